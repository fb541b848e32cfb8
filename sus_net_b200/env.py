"""Host-side mirror of Sus-Net's environment classes over the CUDA library.

`BatchedFourRoomEnv`, `BatchedFourRoomEnvWithTagging` and `BatchedImposterTrainingGround` keep the constructor
arguments, attribute names and method names of `FourRoomEnv` (src/environment/base.py:102-582),
`FourRoomEnvWithTagging` (tagging.py:9-249) and `ImposterTrainingGround` (pred_prey.py:20-99), and add
`num_envs`, `device`, `seed`, `env_id_base` and `auto_reset`.

Two calling conventions:

* ``num_envs == 1`` (default): *reference mode*.  `reset()` / `step()` return exactly what the reference
  returns -- a tuple of numpy arrays ``(agent_positions (A,2) int64, alive_agents (A,) bool
  [, job_positions (J,2) int64, completed_jobs (J,) bool] [, used_tag_actions, tag_counts, time_left])``,
  ``rewards`` as float64 ``(A,)``, python bools and the metrics dict -- so `src/train.py`'s loop,
  `ReplayBuffer.populate` and `AmongUsVisualizer` can drive it unchanged.  No auto-reset (the caller resets,
  train.py:419-445).
* ``num_envs > 1`` (or ``batched=True``): *batched mode*.  Everything is a torch tensor on the env's device and
  nothing synchronises: ``step(actions (N,A)) -> (next_states (N,S) f32, rewards (N,A) f32, dones (N,) bool,
  truncated (N,) bool, info)``; finished envs are reset inside the same kernel launch (``auto_reset=True``),
  `next_states` holds the terminal state of those envs (what the reference stores in replay,
  train.py:388-399) and `flat_states()` the state the next action is taken from.  The returned tensors are
  the env's own output buffers and are overwritten by the next step (the reference returns aliases of its
  internals too); clone what must be kept.

Random draws come from counter-based Philox keyed by (seed, global env id, launch tick) instead of numpy's
global Mersenne Twister; `random_state` / `seed` is the Philox key.
"""
import ctypes as C
from enum import Enum

import numpy as np
import torch

from . import _lib as L

try:  # inside a Sus-Net checkout reuse its enums so that dict keys hash-equal the caller's (train.py:353)
    from src.environment.base import Action, StateFields  # type: ignore
except Exception:  # noqa: BLE001 - any import problem means "not inside Sus-Net"

    class StateFields(Enum):  # base.py:36-43
        AGENT_POSITIONS = 0
        ALIVE_AGENTS = 1
        JOB_POSITIONS = 2
        JOB_STATUS = 3
        USED_TAGS = 4
        TAG_COUNTS = 5
        TAG_RESET_COUNT = 6

    class Action(Enum):  # base.py:46-66
        STAY = 0
        UP = 1
        DOWN = 2
        LEFT = 3
        RIGHT = 4
        KILL = 5
        FIX = 6
        SABOTAGE = 7

        @property
        def is_move_action(self):
            return self in (Action.UP, Action.DOWN, Action.LEFT, Action.RIGHT, Action.STAY)

        @property
        def is_job_action(self):
            return self in (Action.KILL, Action.FIX, Action.SABOTAGE)


from .metrics import EnvMetricView, SusMetrics, METRIC_ORDER, STAT_KEYS  # noqa: E402

CREW_ACTIONS = [Action.STAY, Action.UP, Action.DOWN, Action.LEFT, Action.RIGHT, Action.FIX]  # base.py:82-89
IMPOSTER_ACTIONS = [Action.STAY, Action.UP, Action.DOWN, Action.LEFT, Action.RIGHT, Action.SABOTAGE, Action.KILL]
CREW_ACTIONS_SIMPLE = CREW_ACTIONS[:5]  # pred_prey.py:4-19
IMPOSTER_ACTIONS_SIMPLE = CREW_ACTIONS[:5] + [Action.KILL]

_WALLS = np.array([[0, 4], [2, 4], [3, 4], [4, 4], [5, 4], [6, 4], [8, 4],
                   [4, 0], [4, 2], [4, 3], [4, 5], [4, 6], [4, 8]])  # base.py:172-188

_TORCH_TO_SUS = {torch.uint8: L.U8, torch.int32: L.I32, torch.int64: L.I64, torch.float32: L.F32, torch.float64: L.F64}


class _FieldMap(dict):
    """`env.state_fields` (base.py:36-43,130-135): keyed by this package's `StateFields`, and ALSO by any other enum with
    the same member names -- a caller that imported the reference's own `StateFields` after this package was loaded
    (train.py:353 indexes `env.state_fields[StateFields.ALIVE_AGENTS]`) must find its keys."""

    def __missing__(self, key):
        name = getattr(key, "name", key)
        for k, v in self.items():
            if k.name == name:
                return v
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or any(k.name == getattr(key, "name", key) for k in self)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class _Discrete:
    """The two attributes of gymnasium.spaces.Discrete the reference's callers read."""

    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedFourRoomEnv:
    """GPU `FourRoomEnv` (src/environment/base.py:102)."""

    _VARIANT = L.VARIANT_BASE

    def __init__(self, n_imposters, n_crew, n_jobs, is_action_order_random=True, random_state=None, kill_reward=-5,
                 complete_job_reward=3, sabotage_reward=3, time_step_reward=0, game_end_reward=10, dead_penalty=-2,
                 shuffle_imposter_index=True, debug=False, max_time_steps=1000, include_walls=True, *,
                 tag_reset_interval=50, vote_reward=3, num_envs=1, device=None, seed=None, env_id_base=0,
                 auto_reset=None, batched=None):
        self.lib = L.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("sus_net_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda", )
        if self.device.type != "cuda":
            raise RuntimeError("sus_net_b200 envs live on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.batched = bool(batched) if batched is not None else self.num_envs != 1
        if not self.batched and self.num_envs != 1:
            raise AssertionError("reference mode (batched=False) needs num_envs == 1")
        self.auto_reset = bool(auto_reset) if auto_reset is not None else self.batched
        self.debug = debug

        self.is_action_order_random = is_action_order_random
        self.n_imposters, self.n_crew, self.n_jobs = n_imposters, n_crew, n_jobs
        self.n_agents = n_imposters + n_crew
        self.kill_reward, self.complete_job_reward, self.sabotage_reward = kill_reward, complete_job_reward, sabotage_reward
        self.time_step_reward, self.game_end_reward, self.dead_penalty = time_step_reward, game_end_reward, dead_penalty
        self.shuffle_imposter_index = shuffle_imposter_index
        self.max_time_steps = max_time_steps
        self.tag_reset_interval, self.vote_reward = tag_reset_interval, vote_reward
        self.seed = int(seed if seed is not None else (random_state if random_state is not None else 0))
        self.env_id_base = int(env_id_base)

        cfg = L.SusConfig(
            variant=self._VARIANT, n_imposters=n_imposters, n_crew=n_crew, n_jobs=n_jobs,
            include_walls=int(bool(include_walls)), is_action_order_random=int(bool(is_action_order_random)),
            shuffle_imposter_index=int(bool(shuffle_imposter_index)), max_time_steps=int(max_time_steps),
            tag_reset_interval=int(tag_reset_interval), auto_reset=int(self.auto_reset),
            kill_reward=float(kill_reward), complete_job_reward=float(complete_job_reward),
            sabotage_reward=float(sabotage_reward), time_step_reward=float(time_step_reward),
            game_end_reward=float(game_end_reward), dead_penalty=float(dead_penalty), vote_reward=float(vote_reward),
            num_envs=self.num_envs, seed=self.seed & 0xFFFFFFFFFFFFFFFF, env_id_base=self.env_id_base, reserved=0,
        )
        self._cfg = cfg
        self._h = C.c_void_p()
        L.check(self.lib.sus_env_create(C.byref(cfg), self.device.index, C.byref(self._h)))

        # geometry (base.py:171-207)
        self.walls = _WALLS.copy() if include_walls else np.array([])
        self.grid = np.ones((9, 9), dtype=bool)
        if len(self.walls) != 0:
            self.grid[self.walls[:, 0], self.walls[:, 1]] = 0
        self.valid_positions = np.argwhere(self.grid)
        self.n_rows = self.n_cols = 9

        self._init_action_lists()
        self.state_fields = self._state_fields()
        self._S = L.check(self.lib.sus_flat_state_size(C.byref(cfg)))
        self.action_space = _Discrete(len(Action) + (self.n_agents if self._VARIANT == L.VARIANT_TAGGING else 0))
        self.metrics = EnvMetricView(self)

        N, A, S, dev = self.num_envs, self.n_agents, self._S, self.device
        self._actions = torch.zeros((N, A), dtype=torch.int32, device=dev)
        self._rewards = torch.zeros((N, A), dtype=torch.float32 if self.batched else torch.float64, device=dev)
        self._done = torch.zeros(N, dtype=torch.bool, device=dev)  # the kernel writes 0/1 bytes
        self._trunc = torch.zeros(N, dtype=torch.bool, device=dev)
        self._next_flat = torch.zeros((N, S), dtype=torch.float32, device=dev)
        self._metrics_buf = None
        if not self.batched:
            # reference mode: every per-step output lives in ONE device buffer so a step costs one launch + one D2H copy
            nb_r, nb_m, nb_f = 8 * A, 8 * L.N_METRICS, 4 * S
            self._ref_dev = torch.zeros(nb_r + nb_m + ((nb_f + 7) // 8) * 8 + 8, dtype=torch.uint8, device=dev)
            self._ref_host = torch.zeros_like(self._ref_dev, device="cpu").pin_memory()
            self._rewards = self._ref_dev[:nb_r].view(torch.float64).view(1, A)
            self._metrics_buf = self._ref_dev[nb_r:nb_r + nb_m].view(torch.int64).view(1, L.N_METRICS)
            self._next_flat = self._ref_dev[nb_r + nb_m:nb_r + nb_m + nb_f].view(torch.float32).view(1, S)
            flags = self._ref_dev[nb_r + nb_m + ((nb_f + 7) // 8) * 8:]
            self._done, self._trunc = flags[0:1].view(torch.bool), flags[1:2].view(torch.bool)
            self._ref_layout = (nb_r, nb_m, nb_f)
        self.emit_next_states = True  # batched mode: write the (N, S) replay-layout next-state rows every step
        self.emit_imposters = False   # batched mode: also write the (N, n_imposters) int16 replay column
        self._imposters_buf = None
        self._last_actions = None     # (N, A) int32: the actions the last fused-random-policy step applied
        self._applied_actions = None  # the (N, A) device tensor the last step consumed (what a replay ring stores)
        self._host_state = None  # reference mode: numpy mirror of the single env
        self._imp_cache = None
        self._was_reset = False
        self.t = None

    # ------------------------------------------------------------------ construction helpers
    def _init_action_lists(self):
        self.imposter_actions, self.crew_actions = IMPOSTER_ACTIONS, CREW_ACTIONS
        self.n_imposter_actions, self.n_crew_actions = len(IMPOSTER_ACTIONS), len(CREW_ACTIONS)

    def _state_fields(self):
        return _FieldMap({f: i for i, f in enumerate([StateFields.AGENT_POSITIONS, StateFields.ALIVE_AGENTS,
                                                      StateFields.JOB_POSITIONS, StateFields.JOB_STATUS])})

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.sus_env_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._h = None

    def close(self):
        self.__del__()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ flatten / unflatten (base.py:230-241)
    @property
    def flattened_state_size(self):
        return self._S

    def _field_slices(self):
        A, J = self.n_agents, self.n_jobs
        out, k = [], 0
        out.append(("agent_positions", k, k + 2 * A, (A, 2), np.int64)); k += 2 * A
        out.append(("alive_agents", k, k + A, (A,), np.int8)); k += A
        if J > 0 or self._VARIANT == L.VARIANT_TAGGING:
            out.append(("job_positions", k, k + 2 * J, (J, 2), np.int64)); k += 2 * J
            out.append(("completed_jobs", k, k + J, (J,), np.int8)); k += J
        if self._VARIANT == L.VARIANT_TAGGING:
            out.append(("used_tag_actions", k, k + A, (A,), np.int8)); k += A
            out.append(("tag_counts", k, k + A, (A,), np.int64)); k += A
            out.append(("time_left", k, k + 1, (1,), np.int64)); k += 1
        return out

    def flatten_state(self, state):
        """spaces.flatten(observation_space, state): concat of the fields in tuple order, int64 (base.py:234-235).
        Flat tensors / arrays (batched mode states) are returned unchanged."""
        if isinstance(state, torch.Tensor):
            return state
        if isinstance(state, np.ndarray):
            return state
        return np.concatenate([np.asarray(x, dtype=np.int64).reshape(-1) for x in state])

    def unflatten_state(self, state):
        """spaces.unflatten: split an (S,) row into the state tuple (Box -> int64, MultiBinary -> int8)."""
        if isinstance(state, torch.Tensor):
            state = state.detach().cpu().numpy()
        state = np.asarray(state)
        return tuple(np.asarray(state[a:b], dtype=dt).reshape(shape) for _n, a, b, shape, dt in self._field_slices())

    def compute_state_dims(self, state_field):
        """base.py:565-579 (indexes the observation space by enum VALUE, i.e. tuple order)."""
        A, J = self.n_agents, self.n_jobs
        v = state_field.value
        if v in (0, 2):
            return torch.tensor([9, 9])
        if v == 1:
            return torch.tensor([A])
        if v == 3:
            return torch.tensor([J])
        if self._VARIANT == L.VARIANT_TAGGING:
            if v == 4:
                return torch.tensor([A])
            if v == 5:
                return torch.tensor([A])
            if v == 6:
                return torch.tensor([self.tag_reset_interval - 1])
        raise ValueError(f"Invalid state field: {state_field}")

    # ------------------------------------------------------------------ state access
    def flat_states(self, dtype=torch.float32, out=None):
        """flatten_state of every env's CURRENT state -> (N, S) tensor on the device."""
        if out is None:
            out = torch.empty((self.num_envs, self._S), dtype=dtype, device=self.device)
        L.check(self.lib.sus_env_export_flat(self._h, _TORCH_TO_SUS[out.dtype], _ptr(out), self._stream()))
        return out

    @property
    def imposter_mask_batch(self):
        """(N, A) bool tensor (env.imposter_mask for every env)."""
        out = torch.empty((self.num_envs, self.n_agents), dtype=torch.uint8, device=self.device)
        L.check(self.lib.sus_env_export_imposter_mask(self._h, _ptr(out), self._stream()))
        return out.bool()

    @property
    def imposter_idxs_batch(self):
        """(N, n_imposters) int64 tensor: ascending agent ids of the imposters (env.imposter_idxs)."""
        m = self.imposter_mask_batch
        return m.nonzero()[:, 1].reshape(self.num_envs, self.n_imposters)

    def metrics_batch(self):
        """(N, 8) int64 per-episode counters of every env in METRIC_ORDER."""
        out = torch.empty((self.num_envs, L.N_METRICS), dtype=torch.int64, device=self.device)
        L.check(self.lib.sus_env_export_metrics(self._h, _ptr(out), self._stream()))
        return out

    def episode_stats(self, clear=False):
        """Finished-episode accumulators of this shard as an int64 (10,) device tensor (STAT_KEYS order).
        Multi-GPU drivers all-reduce this tensor over NCCL (`sus_net_b200.distributed.reduce_episode_stats`)."""
        out = torch.empty(L.N_STATS, dtype=torch.int64, device=self.device)
        L.check(self.lib.sus_env_stats(self._h, _ptr(out), self._stream()))
        if clear:
            L.check(self.lib.sus_env_clear_stats(self._h, self._stream()))
        return out

    def track_returns(self, gamma):
        """Keep train()'s running returns `G = reward + gamma * G` (train.py:386) for every agent of every env on the
        device; finished episodes add G[imposter_mask].mean() / G[~imposter_mask].mean() to `return_sums()`."""
        L.check(self.lib.sus_env_track_returns(self._h, float(gamma), self._stream()))
        self._gamma = float(gamma)

    def return_sums(self):
        """(2,) float64 device tensor: summed imposter / crew returns of the episodes counted in episode_stats()[0]."""
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        L.check(self.lib.sus_env_return_sums(self._h, _ptr(out), self._stream()))
        return out

    def episode_stats_dict(self):
        return dict(zip(STAT_KEYS, self.episode_stats().tolist()))

    # reference-mode attribute mirrors (base.py:158-161, tagging.py:29-31; read by src/visualize.py)
    def _need_host(self):
        if self.batched:
            raise AttributeError("per-env numpy attributes exist in reference mode (num_envs == 1) only; "
                                 "use flat_states() / imposter_mask_batch in batched mode")
        if self._host_state is None:
            raise AttributeError("env has not been reset yet")
        return self._host_state

    @property
    def agent_positions(self):
        return self._need_host()["agent_positions"]

    @property
    def alive_agents(self):
        return self._need_host()["alive_agents"]

    @property
    def job_positions(self):
        return self._need_host()["job_positions"]

    @property
    def completed_jobs(self):
        return self._need_host()["completed_jobs"]

    @property
    def imposter_mask(self):
        if self.batched:
            return self.imposter_mask_batch
        self._need_host()
        return self._imp_cache

    @property
    def crew_mask(self):
        return ~self.imposter_mask

    @property
    def imposter_idxs(self):
        if self.batched:
            return self.imposter_idxs_batch
        return np.where(self.imposter_mask)[0]

    @property
    def crew_idxs(self):
        return np.where(self.crew_mask)[0]

    @property
    def agent_action_map(self):
        """base.py:304-312 (tagging.py:69-75 appends the tag targets)."""
        m = {}
        mask = self.imposter_mask
        for i in range(self.n_agents):
            acts = list(self.imposter_actions if mask[i] else self.crew_actions)
            if self._VARIANT == L.VARIANT_TAGGING:
                acts = acts + [j for j in range(self.n_agents) if j != i]
            m[i] = acts
        return m

    def compute_action(self, agent_idx, action_idx):
        return str(Action(self.agent_action_map[agent_idx][action_idx]))  # base.py:581-582

    # ------------------------------------------------------------------ reset (base.py:251-324)
    def reset(self, seed=None, mask=None, **kwargs):
        """Reference mode: `-> (state_tuple, info)`.  Batched mode: `-> (flat_states (N,S) f32, {})`; `mask` (N,)
        restricts the reset to a subset.  `seed` re-keys nothing (the Philox key is fixed at construction) but is
        accepted for API compatibility when None."""
        if seed is not None:
            raise NotImplementedError("per-reset reseeding is not supported: pass `seed=` to the constructor")
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        L.check(self.lib.sus_env_reset(self._h, _ptr(m), self._stream()))
        self._was_reset = True
        if self.batched:
            return self.flat_states(), {}
        self._sync_host()
        return self._state_tuple(), self._reset_info()

    def _reset_info(self):
        return self.metrics.get_metrics()

    def _sync_host(self):
        self._set_host_state(self.flat_states(dtype=torch.int64)[0].cpu().numpy())
        self._imp_cache = self.imposter_mask_batch[0].cpu().numpy()
        mt = self.metrics_batch()[0].cpu().numpy()
        self._host_metrics = mt
        self.t = int(min(mt[0], self.max_time_steps - 1))

    def _set_host_state(self, flat):
        hs = {}
        for name, a, b, shape, dt in self._field_slices():
            arr = np.asarray(flat[a:b]).reshape(shape)
            hs[name] = arr.astype(bool) if dt == np.int8 else arr.astype(np.int64)
        if "job_positions" not in hs:
            hs["job_positions"] = np.zeros((0, 2), dtype=np.int64)
            hs["completed_jobs"] = np.zeros((0,), dtype=bool)
        self._host_state = hs

    def _state_tuple(self):
        hs = self._host_state
        out = [hs["agent_positions"], hs["alive_agents"]]
        if self.n_jobs > 0:
            out += [hs["job_positions"], hs["completed_jobs"]]
        return tuple(out)

    # ------------------------------------------------------------------ actions (base.py:326-330)
    def sample_actions(self):
        """Uniform over each agent's role list, dead agents included.  Reference mode: (A,) int numpy array;
        batched mode: (N, A) int32 tensor (a live buffer)."""
        L.check(self.lib.sus_env_sample_actions(self._h, _ptr(self._actions), self._stream()))
        if self.batched:
            return self._actions
        return self._actions[0].cpu().numpy().astype(int)

    # ------------------------------------------------------------------ step (base.py:332-407)
    @property
    def compact(self):
        """The compact host protocol of this env (`sus_net_b200.compact.CompactProtocol`: record geometry, packers and
        the bit-exact float64 decode table)."""
        if getattr(self, "_compact", None) is None:
            from .compact import CompactProtocol

            self._compact = CompactProtocol(self)
        return self._compact

    def step(self, agent_actions=None, featurizer=None, check=None, out=None, packed_actions=False, packed_out=None):
        """One env step for every env.

        agent_actions: (N, A) (batched) or (A,) (reference mode) role-list indices; None in batched mode means the
        fused random policy (`env.step(env.sample_actions())` in one launch).
        featurizer: optional sus_net_b200 featurizer; its tensors are written by the same kernel launch from the
        state the next action is taken from (batched mode, T = 1).
        check: validate action indices on the device and raise IndexError like the reference (synchronises);
        defaults to True in reference mode, False in batched mode.
        out: batched mode only -- `(rewards (N,A) f32|f64, dones (N,) bool, truncated (N,) bool)` device tensors to
        write instead of the env's own output buffers (double-buffering by `HostStepper`).
        packed_actions / packed_out: batched mode, the compact host protocol (`env.compact`): `agent_actions` is an
        (N, action_bytes) uint8 device tensor of bit-packed role-list indices; `packed_out` an (N, result_bytes) uint8
        device tensor that receives reward codes + done / truncated bits INSTEAD of rewards / dones / truncated (the
        step then returns `(next_states, packed_out, None, None, {})`)."""
        if not self._was_reset:
            raise AssertionError("reset() must be called before step()")
        N, A = self.num_envs, self.n_agents
        io = L.SusStepIO()
        keep = None
        if agent_actions is None:
            if not self.batched:
                raise AssertionError(f"Expected {A} actions, got none")
            io.actions = None
            if self.emit_imposters:  # a replay ring is attached: it needs the actions the random policy drew
                if self._last_actions is None:
                    self._last_actions = torch.zeros((N, A), dtype=torch.int32, device=self.device)
                io.actions_out = self._last_actions.data_ptr()
                self._applied_actions = self._last_actions
        else:
            if not self.batched:
                assert len(agent_actions) == A, f"Expected {A} actions, got {len(agent_actions)}"  # base.py:357-359
                assert all(a < self.action_space.n for a in agent_actions), f"Invalid action(s) {agent_actions}"
                n_tag = A - 1 if self._VARIANT == L.VARIANT_TAGGING else 0
                for i, a in enumerate(agent_actions):  # agent_action_map[i][a] (base.py:381): IndexError past the list
                    n_role = (len(self.imposter_actions) if self._imp_cache[i] else len(self.crew_actions)) + n_tag
                    if not 0 <= int(a) < n_role:
                        raise IndexError("list index out of range")
                agent_actions = np.asarray(agent_actions).reshape(1, A)
            if packed_actions:
                assert self.batched and isinstance(agent_actions, torch.Tensor) and agent_actions.dtype == torch.uint8 \
                    and agent_actions.device == self.device and agent_actions.is_contiguous() \
                    and tuple(agent_actions.shape) == (N, self.compact.action_bytes), \
                    f"packed actions must be a contiguous ({N}, {self.compact.action_bytes}) uint8 tensor on {self.device}"
                keep = agent_actions
            elif isinstance(agent_actions, torch.Tensor) and agent_actions.device == self.device and \
                    agent_actions.dtype in (torch.uint8, torch.int32, torch.int64) and agent_actions.is_contiguous():
                keep = agent_actions
            else:
                keep = torch.as_tensor(np.asarray(agent_actions) if not isinstance(agent_actions, torch.Tensor)
                                       else agent_actions).to(device=self.device, dtype=torch.int32).contiguous()
            if not packed_actions:
                assert tuple(keep.shape) == (N, A), f"Expected actions of shape {(N, A)}, got {tuple(keep.shape)}"
            io.actions = keep.data_ptr()
            io.actions_dtype = L.PACKED if packed_actions else _TORCH_TO_SUS[keep.dtype]
            self._applied_actions = keep
        if packed_out is not None:
            assert self.batched and packed_out.dtype == torch.uint8 and packed_out.is_contiguous() and \
                tuple(packed_out.shape) == (N, self.compact.result_bytes) and packed_out.device == self.device
            io.packed_out = packed_out.data_ptr()
            rewards, done, trunc = packed_out, None, None
        else:
            rewards, done, trunc = (self._rewards, self._done, self._trunc) if out is None else out
            io.rewards = rewards.data_ptr()
            io.rewards_dtype = _TORCH_TO_SUS[rewards.dtype]
            io.done = done.data_ptr()
            io.truncated = trunc.data_ptr()
        if self.emit_next_states:
            io.next_flat = self._next_flat.data_ptr()
        if self._metrics_buf is not None:
            io.metrics = self._metrics_buf.data_ptr()
        if self.emit_imposters:
            if self._imposters_buf is None:
                self._imposters_buf = torch.zeros((N, self.n_imposters), dtype=torch.int16, device=self.device)
            io.imposters = self._imposters_buf.data_ptr()
        spec = None
        if featurizer is not None:
            spec = featurizer._bind_for_fused_step(self)
            io.encode = C.pointer(spec)
            io.spatial = featurizer._sp_buf.data_ptr() if featurizer._sp_buf is not None else None
            io.non_spatial = featurizer._ns_buf.data_ptr()
        L.check(self.lib.sus_env_step(self._h, C.byref(io), self._stream()))
        if check if check is not None else False:
            L.check(self.lib.sus_env_check_actions(self._h, self._stream()))
        if self.batched:
            return (self._next_flat if self.emit_next_states else None), rewards, done, trunc, {}
        # reference mode: one D2H copy of the packed outputs, then numpy views
        self._ref_host.copy_(self._ref_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        nb_r, nb_m, nb_f = self._ref_layout
        h = self._ref_host.numpy()
        rewards = h[:nb_r].view(np.float64).copy()
        self._host_metrics = h[nb_r:nb_r + nb_m].view(np.int64).copy()
        flat = h[nb_r + nb_m:nb_r + nb_m + nb_f].view(np.float32).astype(np.int64)
        flags = h[nb_r + nb_m + ((nb_f + 7) // 8) * 8:]
        self._set_host_state(flat)
        self.t = int(min(self._host_metrics[0], self.max_time_steps - 1))
        return (self._full_state_tuple(), rewards, bool(flags[0]), bool(flags[1]), self.metrics.get_metrics())

    def _full_state_tuple(self):
        return self._state_tuple()

    def rollout(self, n_steps, reward_sums=False):
        """`n_steps` random-policy steps of every env in ONE kernel launch (batched mode): the same trajectory, draws,
        auto-resets and episode statistics as `n_steps` calls of `step(None)`, without any per-step output.  Returns
        the (N, A) float64 per-agent reward sums if `reward_sums` else None."""
        if not self._was_reset:
            raise AssertionError("reset() must be called before rollout()")
        out = torch.zeros((self.num_envs, self.n_agents), dtype=torch.float64, device=self.device) if reward_sums else None
        L.check(self.lib.sus_env_rollout(self._h, int(n_steps), _ptr(out), self._stream()))
        return out

    def device_ticks(self, enable=True):
        """Keep the Philox launch ticks in device memory so that `reset` / `sample_actions` / `step` / `rollout` launches
        can be captured in a CUDA graph (`torch.cuda.graph`) and replayed: every replay advances the ticks on the device
        and gives the same results as the individual calls would.  With host ticks (the default) a replayed launch would
        repeat its draws.  Costs one extra one-thread launch per call."""
        L.check(self.lib.sus_env_device_ticks(self._h, int(bool(enable)), self._stream()))

    def check_actions(self):
        """Raise IndexError if any step since the last check saw an index outside an agent's role list."""
        L.check(self.lib.sus_env_check_actions(self._h, self._stream()))

    # ------------------------------------------------------------------ checkpoint / resume (no reference analogue)
    def state_dict(self):
        ptrs = (C.c_void_p * 4)()
        sizes = (C.c_int32 * 4)()
        L.check(self.lib.sus_env_state_arrays(self._h, ptrs, sizes))
        torch.cuda.current_stream(self.device).synchronize()
        arrays = []
        for p, s in zip(ptrs, sizes):
            arrays.append(_as_device_bytes(p, self.num_envs * s, self.device).cpu())
        ticks = [C.c_uint64(), C.c_uint64(), C.c_uint64()]
        L.check(self.lib.sus_env_get_ticks(self._h, *[C.byref(t) for t in ticks]))
        # the accumulators a resumed run must continue from: finished-episode statistics, the pending invalid-action
        # counter, and (if tracked) train()'s running returns G of every agent + the two return sums, with their gamma
        aux = [None if not p or not n else _as_device_bytes(p, n, self.device).cpu() for p, n in zip(*self._aux_arrays())]
        return {"arrays": arrays, "ticks": [t.value for t in ticks], "stats": self.episode_stats().cpu(),
                "aux": aux, "gamma": getattr(self, "_gamma", None)}

    def _aux_arrays(self):
        ptrs = (C.c_void_p * 3)()
        sizes = (C.c_int64 * 3)()
        L.check(self.lib.sus_env_aux_arrays(self._h, ptrs, sizes))
        return list(ptrs), list(sizes)

    def load_state_dict(self, sd):
        ptrs = (C.c_void_p * 4)()
        sizes = (C.c_int32 * 4)()
        L.check(self.lib.sus_env_state_arrays(self._h, ptrs, sizes))
        for p, s, a in zip(ptrs, sizes, sd["arrays"]):
            _as_device_bytes(p, self.num_envs * s, self.device).copy_(a.to(self.device))
        L.check(self.lib.sus_env_set_ticks(self._h, *[int(t) for t in sd["ticks"]]))
        aux = sd.get("aux")
        if aux is not None:
            if aux[2] is not None:  # returns were tracked: allocate them here too, then restore G, the sums and gamma
                self.track_returns(sd["gamma"])
            for (p, n), a in zip(zip(*self._aux_arrays()), aux):
                if a is not None and p:
                    assert n == a.numel(), "checkpoint was made for a different batch size"
                    _as_device_bytes(p, n, self.device).copy_(a.to(self.device))
        elif "stats" in sd:  # (older checkpoints carry the statistics only)
            p, n = self._aux_arrays()
            _as_device_bytes(p[0], n[0], self.device).copy_(sd["stats"].to(self.device).view(torch.uint8))
        self._was_reset = True
        if not self.batched:
            self._sync_host()

    # ------------------------------------------------------------------ parity mode
    def debug_inject_words(self, step_words=None, reset_words=None, act_words=None):
        """Raw 32-bit words replacing the Philox output of the next launch (tests only)."""
        keep = []
        ptrs = []
        for w in (step_words, reset_words, act_words):
            if w is None:
                ptrs.append(None)
            else:
                t = torch.as_tensor(np.ascontiguousarray(w, dtype=np.uint32).view(np.int32)).to(self.device).contiguous()
                keep.append(t)
                ptrs.append(C.c_void_p(t.data_ptr()))
        self._inject_keep = keep
        L.check(self.lib.sus_env_debug_inject_words(self._h, *ptrs))

    def import_flat(self, flat, imposter_mask, t=None):
        """Load env states from flatten-order rows (N, S), role masks (N, A) and optional time steps (N,)."""
        f = torch.as_tensor(np.asarray(flat)).to(device=self.device, dtype=torch.int64).contiguous()
        m = torch.as_tensor(np.asarray(imposter_mask)).to(device=self.device, dtype=torch.uint8).contiguous()
        tt = None if t is None else torch.as_tensor(np.asarray(t)).to(device=self.device, dtype=torch.int32).contiguous()
        L.check(self.lib.sus_env_import_flat(self._h, _ptr(f), _ptr(m), _ptr(tt), self._stream()))
        self._was_reset = True
        if not self.batched:
            self._sync_host()


def _as_device_bytes(ptr, nbytes, device):
    """uint8 tensor aliasing `nbytes` of device memory at `ptr` (via __cuda_array_interface__)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=device)


class BatchedFourRoomEnvWithTagging(BatchedFourRoomEnv):
    """GPU `FourRoomEnvWithTagging` (src/environment/tagging.py:9)."""

    _VARIANT = L.VARIANT_TAGGING

    def __init__(self, *args, tag_reset_interval=50, vote_reward=3, **kwargs):
        super().__init__(*args, tag_reset_interval=tag_reset_interval, vote_reward=vote_reward, **kwargs)
        # AmongUsVisualizer decides whether to draw the voting panel with `env.__dict__.get("tag_counts")`
        # (visualize.py:153); the value itself comes from the property below (a data descriptor wins the lookup)
        self.__dict__["tag_counts"] = "property"

    def _init_action_lists(self):
        super()._init_action_lists()
        self.n_imposter_actions += self.n_agents - 1  # tagging.py:35-36
        self.n_crew_actions += self.n_agents - 1

    def _state_fields(self):
        # TUPLE order (tagging.py:221-230).  The reference's own map (tagging.py:15-28) disagrees with its state
        # tuple and breaks every featurizer on this env; documented deviation.
        return _FieldMap({f: i for i, f in enumerate([StateFields.AGENT_POSITIONS, StateFields.ALIVE_AGENTS,
                                                      StateFields.JOB_POSITIONS, StateFields.JOB_STATUS, StateFields.USED_TAGS,
                                                      StateFields.TAG_COUNTS, StateFields.TAG_RESET_COUNT])})

    @property
    def used_tag_actions(self):
        return self._need_host()["used_tag_actions"]

    @property
    def tag_counts(self):
        return self._need_host()["tag_counts"]

    @property
    def tag_reset_timer(self):
        return self.tag_reset_interval - int(self._need_host()["time_left"][0])

    def _reset_info(self):
        return {}  # tagging.py:101

    def _state_tuple(self):
        hs = self._host_state  # tagging.py:94-99
        return (hs["agent_positions"], hs["alive_agents"], hs["job_positions"], hs["completed_jobs"],
                hs["used_tag_actions"], hs["tag_counts"], int(hs["time_left"][0]))

    def compute_action(self, agent_idx, action_idx):  # tagging.py:243-249
        if action_idx < len(Action):
            return str(Action(action_idx))
        players = np.arange(self.n_agents)
        return f"Vote Player {players[players != agent_idx][action_idx - len(Action)]}"


class BatchedImposterTrainingGround(BatchedFourRoomEnv):
    """GPU `ImposterTrainingGround` (src/environment/pred_prey.py:20)."""

    _VARIANT = L.VARIANT_TRAINING_GROUND

    def __init__(self, n_crew, n_jobs, time_step_reward, kill_reward, sabotage_reward, end_of_game_reward,
                 random_state=None, debug=False, shuffle_imposter_index=False, include_walls=True, **kwargs):
        super().__init__(  # pred_prey.py:52-66
            n_imposters=1, n_crew=n_crew, n_jobs=n_jobs, time_step_reward=time_step_reward, kill_reward=kill_reward,
            sabotage_reward=sabotage_reward, debug=debug, dead_penalty=0, game_end_reward=end_of_game_reward,
            random_state=random_state, is_action_order_random=False, shuffle_imposter_index=shuffle_imposter_index,
            include_walls=include_walls, **kwargs)

    def _init_action_lists(self):  # pred_prey.py:68-73
        self.imposter_actions, self.crew_actions = IMPOSTER_ACTIONS_SIMPLE, CREW_ACTIONS_SIMPLE
        self.n_imposter_actions, self.n_crew_actions = len(IMPOSTER_ACTIONS_SIMPLE), len(CREW_ACTIONS_SIMPLE)
