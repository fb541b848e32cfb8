"""GPU replay ring in the reference's layout (src/replay_memory.py) -- SURVEY.md row (f1).

`ReplayBuffer` keeps the reference's constructor, tensor names, shapes and dtypes (`states (M,T,S) f32`, `actions
(M,A) i64`, `rewards (M,A) f32`, `next_states (M,T,S) f32`, `dones (M,1) bool`, `imposters (M,n_imp) i16`), its
`add` / `sample` / `populate` methods and the `Batch` namedtuple, so `DQNTeamTrainer.train_step` (train.py:50-149)
consumes its batches unchanged -- but the tensors live on the env's device and a whole batched step (N transitions,
including the T-deep `np.roll` sequence bookkeeping of replay_memory.py:107-134 / train.py:388-399,440-445) is stored
by ONE kernel launch (`sus_replay_push`).
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _lib as L
from .env import _TORCH_TO_SUS

Batch = namedtuple("Batch", ("states", "actions", "rewards", "next_states", "imposters", "dones"))  # replay_memory.py:6-8


class ReplayBuffer:
    def __init__(self, max_size, state_size, trajectory_size, n_agents, n_imposters, device="cuda"):
        assert max_size > 0, "Replay buffer size must be positive"  # replay_memory.py:21-24
        assert trajectory_size > 0, "Trajectory size must be positive"
        assert state_size > 0, "State size must be positive"
        assert n_agents > 0, "Number of agents must be positive"
        self.max_size, self.trajectory_size, self.state_size = max_size, trajectory_size, state_size
        self.n_agents, self.n_imposters = n_agents, n_imposters
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        M, T, S, A, dev = max_size, trajectory_size, state_size, n_agents, self.device
        self.states = torch.empty((M, T, S), device=dev)  # replay_memory.py:33-44
        self.actions = torch.empty((M, A), dtype=torch.long, device=dev)
        self.rewards = torch.empty((M, A), device=dev)
        self.next_states = torch.empty((M, T, S), device=dev)
        self.dones = torch.empty((M, 1), dtype=torch.bool, device=dev)
        self.imposters = torch.empty((M, n_imposters), dtype=torch.int16, device=dev)
        self.idx = 0
        self.size = 0
        self._env = None
        self._seq = None
        # device copies of idx / size: a CUDA-graph replay of collect_step() / sample() cannot take new host integers, so
        # with `static_buffers=True` (attach) the push kernel reads the slot from `_idx_dev`, sample() scales its draws by
        # `_size_dev`, and both are advanced on the device
        self._idx_dev = self._size_dev = None
        self._static = False

    # ---------------------------------------------------------------- reference API
    def add(self, state, action, reward, next_state, done, imposters):
        """One transition (replay_memory.py:50-72)."""
        i, dev = self.idx, self.device
        self.states[i] = torch.as_tensor(np.asarray(state), dtype=torch.float32).to(dev)
        self.actions[i] = torch.as_tensor(np.asarray(action)).to(dev)
        self.rewards[i] = torch.as_tensor(np.asarray(reward), dtype=torch.float32).to(dev)
        self.next_states[i] = torch.as_tensor(np.asarray(next_state), dtype=torch.float32).to(dev)
        self.dones[i] = bool(done)
        self.imposters[i] = torch.as_tensor(np.asarray(imposters)).to(dev)
        self.idx = (self.idx + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def sample(self, batch_size, generator=None):
        """replay_memory.py:74-94: uniform with replacement over the filled part."""
        assert self.size > 0, "Replay buffer is empty, can't sample"
        if self._static:  # graph-safe: the filled size lives on the device (float64 draws: exact for any ring size)
            u = torch.rand(batch_size, device=self.device, dtype=torch.float64, generator=generator)
            sample_idx = (u * self._size_dev).long().clamp_(max=self.max_size - 1)
        else:
            sample_idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=generator)
        return Batch(states=self.states[sample_idx], actions=self.actions[sample_idx], rewards=self.rewards[sample_idx],
                     imposters=self.imposters[sample_idx], next_states=self.next_states[sample_idx],
                     dones=self.dones[sample_idx])

    def populate(self, env, num_steps):
        """Fill with (at least) `num_steps` random-policy transitions (replay_memory.py:96-143).  With a batched env
        every launch adds `env.num_envs` transitions; a reference-mode env is driven one step at a time exactly
        like the reference does."""
        if not env.batched:
            return self._populate_single(env, num_steps)
        self.attach(env)
        launches = -(-int(num_steps) // env.num_envs)
        for _ in range(launches):
            self.collect_step(None)
        return launches * env.num_envs

    def _populate_single(self, env, num_steps):
        """Reference mode (one env, one transition per step): whole random-policy episodes until `num_steps` transitions are
        stored; the last episode is cut where the count is reached (replay_memory.py:103-143)."""
        added = 0
        while added < num_steps:
            for transition in self._random_episode(env):
                self.add(*transition)
                added += 1
                if added >= num_steps:
                    break
        return added

    def _random_episode(self, env):
        """The transitions of one random-policy episode as `add` takes them.  The T-deep window starts as T copies of the reset
        state and slides by one state per step (replay_memory.py:107-134); the episode's imposter ids are read before the step."""
        start = np.asarray(env.flatten_state(env.reset()[0]), dtype=np.float64)
        window = np.tile(start, (self.trajectory_size, 1))
        ended = False
        while not ended:
            who = env.imposter_idxs
            action = env.sample_actions()
            obs, reward, done, truncated, _ = env.step(action)
            newest = np.asarray(env.flatten_state(obs), dtype=np.float64)
            slid = np.vstack([window[1:], newest[None, :]])
            yield window, action, reward, slid, done, who
            window, ended = slid, bool(done or truncated)

    # ---------------------------------------------------------------- batched collection
    def attach(self, env, static_buffers=False):
        """Bind a batched env (already reset, or reset here) and start every env's sequence from T copies of its
        current state (replay_memory.py:107-112, train.py:318-322).  static_buffers: keep every pointer and integer a
        collect_step() / sample() launch uses fixed or on the device, so the calls can be captured in a CUDA graph."""
        assert env.batched and env.device == self.device
        assert env.flattened_state_size == self.state_size and env.n_agents == self.n_agents
        assert env.num_envs <= self.max_size, "the ring must hold at least one batched step"
        if not env._was_reset:
            env.reset()
        env.emit_next_states = True
        env.emit_imposters = True
        self._env = env
        cur = env.flat_states()
        seq = cur[:, None, :].repeat(1, self.trajectory_size, 1)  # a copy even for T == 1 (cur is overwritten every step)
        self._seq = [seq, torch.empty_like(seq)]
        self._cur_flat = cur
        self._static = bool(static_buffers)
        if self._static:
            self._idx_dev = torch.tensor([self.idx], dtype=torch.int64, device=self.device)
            self._size_dev = torch.tensor([self.size], dtype=torch.int64, device=self.device)

    @property
    def state_sequence(self):
        """(N, T, S) f32: the sequence the next action of every env is taken from (train.py:346-348 featurizes it)."""
        return self._seq[0]

    def collect_step(self, actions=None, featurizer=None):
        """env.step(actions) + ReplayBuffer.add for all N envs (one step launch, one export, one push launch).
        Returns env.step's tuple."""
        env = self._env
        out = env.step(actions, featurizer=featurizer)
        next_flat, rewards, dones, truncated, _ = out
        applied = env._applied_actions  # the device tensor the step consumed (given, converted, or drawn by the policy)
        env.flat_states(out=self._cur_flat)
        N = env.num_envs
        # fixed pointers (graph replays) with T == 1: the kernel advances the sequence in place, no copy back
        in_place = self._static and self.trajectory_size == 1
        p = L.SusReplayPush(
            N=N, M=self.max_size, idx=self.idx, T=self.trajectory_size, S=self.state_size, A=self.n_agents,
            n_imposters=self.n_imposters, seq_in=self._seq[0].data_ptr(), seq_out=self._seq[0 if in_place else 1].data_ptr(),
            next_flat=next_flat.data_ptr(), cur_flat=self._cur_flat.data_ptr(), actions=applied.data_ptr(),
            actions_dtype=_TORCH_TO_SUS[applied.dtype], rewards=rewards.data_ptr(), done=dones.data_ptr(),
            truncated=truncated.data_ptr(), imposters=env._imposters_buf.data_ptr(), states=self.states.data_ptr(),
            r_actions=self.actions.data_ptr(), r_rewards=self.rewards.data_ptr(), next_states=self.next_states.data_ptr(),
            r_dones=self.dones.data_ptr(), r_imposters=self.imposters.data_ptr(),
            idx_dev=self._idx_dev.data_ptr() if self._static else None)
        L.check(env.lib.sus_replay_push(C.byref(p), self.device.index, env._stream()))
        if self._static:  # same buffers every step (a graph replays fixed pointers): copy the rolled sequence back
            if not in_place:
                self._seq[0].copy_(self._seq[1])
            self._idx_dev.add_(N).remainder_(self.max_size)
            self._size_dev.add_(N).clamp_(max=self.max_size)
        else:
            self._seq.reverse()
        self.idx = (self.idx + N) % self.max_size
        self.size = min(self.size + N, self.max_size)
        return out

    def sync_host_counters(self):
        """After CUDA-graph replays (which advance only the device copies): read idx / size back (synchronises)."""
        if self._static:
            self.idx, self.size = int(self._idx_dev.item()), int(self._size_dev.item())

