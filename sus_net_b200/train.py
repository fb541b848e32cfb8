"""Batched restatement of the callers of the hot path: the ε-greedy acting part of
`train()` (src/train.py:349-381), `ExponentialSchedule` (src/scheduler.py:4-40), `DQNTeamTrainer.train_step`
(src/train.py:40-149) and a batched training loop with the reference's cadence (target sync, train interval,
auto-reset inside the env kernel, replay writes on the GPU).  Q-networks are out of scope: any `torch.nn.Module`
with the reference's `forward(spatial, non_spatial) -> (B, n_actions)` signature works (e.g. the reference's own
`MLP` / `SpatialDQN` moved to the GPU); `None` stands for the reference's `RandomEquiprobable` model.

Nothing here synchronises the host with the device:
* acting reads the feature tensors the FUSED step kernel wrote (T = 1) or a T-deep device ring of them
  (`FeatureSequence`, rolled by `sus_seq_roll`), runs the networks and hands their Q-values to ONE kernel
  (`sus_env_select_actions`: ε-greedy, role-aware ranges, dead agents keep action 0, Philox draws);
* `DQNTeamTrainer.train_step` uses masked losses over the whole batch instead of boolean gathers, keeps the losses
  on the device and applies Adam through a device-side gate (an empty team subset leaves weights AND optimizer state
  untouched, like the reference's `continue`);
* with `use_graphs=True` an iteration (act + fused step + replay push) and a train step are two CUDA graphs.

Multi-GPU: one process per GPU, each with its own env shard and replay ring; the per-view gradient SUMS, sample counts
and loss sums of all ranks are added with one NCCL all-reduce per optimizer step, so every sample of the global subset
carries the same weight (`_FlatAdam.apply`); episode statistics with `reduce_episode_stats`.
"""
import ctypes as C
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib as L


class ExponentialSchedule:
    """scheduler.py:4-40: value(t) = a * exp(b * t) clipped to [value_from, value_to]."""

    def __init__(self, value_from, value_to, num_steps):
        self.value_from, self.value_to, self.num_steps = value_from, value_to, num_steps
        self.a = value_from
        self.b = math.log(value_to / value_from) / (num_steps - 1)

    def value(self, step):
        if step < 1:
            return self.value_from
        if step >= self.num_steps:
            return self.value_to
        return self.a * math.exp(self.b * step)


def _distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ------------------------------------------------------------------------------------------------ feature sequences
class FeatureSequence:
    """T-deep device ring of ENCODED features: what `featurizer.fit(state_sequence)` would produce for every env's running
    sequence (train.py:346-348), maintained by rolling in the newest time step -- the tensors the fused step kernel just
    wrote -- instead of re-encoding all T states every iteration (`sus_seq_roll`; an env whose episode ended restarts from
    T copies of its reset state's features, train.py:440-445).  For T == 1 it is the featurizer's own output, no copy."""

    def __init__(self, env, featurizer, T):
        self.env, self.featurizer, self.T = env, featurizer, int(T)
        self.sp = self.ns = None
        self._alt = None

    def start(self):
        """Encode the envs' current states and start every sequence from T copies of them (train.py:318-322)."""
        f, T, N = self.featurizer, self.T, self.env.num_envs
        f.encode_env(self.env)
        if T == 1:
            return
        sp, ns = f._sp_buf, f._ns_buf  # (Vs, N, R) or None, (Vn, N, F)
        rep = lambda t: None if t is None else t[:, :, None, :].repeat(1, 1, T, 1).contiguous()  # noqa: E731
        self.sp, self.ns = rep(sp), rep(ns)
        self._alt = (None if sp is None else torch.empty_like(self.sp), torch.empty_like(self.ns))

    def push(self, done, truncated):
        """Roll the newest features (the featurizer's buffers, just written by the fused step) into the sequences."""
        if self.T == 1:
            return
        f, env = self.featurizer, self.env
        for cur, alt, new in ((self.sp, self._alt[0], f._sp_buf), (self.ns, self._alt[1], f._ns_buf)):
            if cur is None:
                continue
            V, N, T, R = cur.shape
            L.check(env.lib.sus_seq_roll(C.c_void_p(cur.data_ptr()), C.c_void_p(alt.data_ptr()), C.c_void_p(new.data_ptr()),
                                         C.c_void_p(done.data_ptr()), C.c_void_p(truncated.data_ptr()), V * N, N, T, R,
                                         env.device.index, env._stream()))
            cur.copy_(alt)  # fixed pointers (CUDA graphs); the ring is T x the size of one time step

    def views(self):
        """(spatial (Vs, N, T, C, 9, 9) or None, non_spatial (Vn, N, T, F)): `featurizer.stacked_views()` of the sequences."""
        f, N, A = self.featurizer, self.env.num_envs, self.env.n_agents
        if self.T == 1:
            f.B, f.T = N, 1
            return f.stacked_views()
        sp = None if self.sp is None else self.sp.view(self.sp.shape[0], N, self.T, A + 2, 9, 9)
        return sp, self.ns


# ------------------------------------------------------------------------------------------------ acting
class BatchedActor:
    """train.py:349-381 for every env at once: per agent view, alive imposters act with `imposter_model`, alive crew
    with `crew_model`, each ε-greedy over its role's action count; dead agents keep action 0.  A model that is `None`
    acts uniformly at random (the reference's `RandomEquiprobable`).

    `act_kernel` is the production path (no host sync, one selection kernel, graph-capturable); `act` / `act_grouped`
    are the torch-op restatements kept for comparison (same distribution; identical actions for eps == 0)."""

    def __init__(self, env, imposter_model, crew_model, generator=None, dense=False, fused_mlp=True):
        """fused_mlp: evaluate Linear + PReLU / ReLU stacks (the reference's MLP estimator) with the one-launch inference kernel
        (`sus_net_b200.mlp.FusedMLP`) in `act_kernel`; other modules, and everything in `act` / `act_grouped`, run as they are."""
        self.env, self.imposter_model, self.crew_model = env, imposter_model, crew_model
        self.generator, self.dense = generator, dense
        self._q_imp = self._q_crew = self._actions = self._eps = None
        from .mlp import FusedMLP

        self._infer = {}
        for name, m in (("imp", imposter_model), ("crew", crew_model)):
            self._infer[name] = FusedMLP(m) if (fused_mlp and FusedMLP.supports(m)) else m

    # ---- production path
    @torch.no_grad()
    def act_kernel(self, sp, ns, eps, imposter_index=None, out=None):
        """sp / ns: stacked views ((Vs, N, T, C, 9, 9) or None, (Vn, N, T, F)) of the states the actions are taken from;
        eps: python float or 0-d / 1-element float32 DEVICE tensor (use a tensor under CUDA graphs);
        imposter_index: (N,) int64 agent id of every env's imposter (n_imposters == 1) -- default: read from the env;
        pass a constant tensor when `shuffle_imposter_index=False`.  Returns the (N, A) int32 action tensor (a live buffer)."""
        env = self.env
        N, A, dev = env.num_envs, env.n_agents, env.device
        io = L.SusPolicyIO()
        keep = []
        T = ns.shape[2]

        def view(t, k):
            if t is None:
                return torch.zeros(N, T, 1, device=dev)  # FlatFeaturizer's spatial placeholder (model_ready.py:362)
            return t[0] if t.shape[0] == 1 else t[k]

        imp_net, crew_net = self._infer["imp"], self._infer["crew"]
        if self.imposter_model is not None:
            if env.n_imposters == 1:
                if ns.shape[0] == 1 and (sp is None or sp.shape[0] == 1):
                    q = imp_net(view(sp, 0), ns[0])  # every view is identical (Flat)
                else:
                    if imposter_index is None:
                        imposter_index = torch.argmax(env.imposter_mask_batch.to(torch.uint8), dim=1)
                    rows = torch.arange(N, device=dev)
                    spv = view(sp, 0) if (sp is None or sp.shape[0] == 1) else sp[imposter_index, rows]
                    q = imp_net(spv, ns[imposter_index, rows])
                io.imposter_per_view = 0
            else:
                q = torch.stack([imp_net(view(sp, k), view(ns, k)) for k in range(A)])
                io.imposter_per_view = 1
            q = q.float().contiguous()
            assert q.shape[-1] == env.n_imposter_actions
            keep.append(q)
            io.q_imposter = q.data_ptr()
        if self.crew_model is not None:
            q = torch.stack([crew_net(view(sp, k), view(ns, k)) for k in range(A)]).float().contiguous()
            assert q.shape[-1] == env.n_crew_actions
            keep.append(q)
            io.q_crew = q.data_ptr()
        if isinstance(eps, torch.Tensor):
            assert eps.dtype == torch.float32 and eps.device == dev
            io.eps = eps.data_ptr()
        else:
            io.eps_value = float(eps)
        if out is None:
            if self._actions is None:
                self._actions = torch.zeros((N, A), dtype=torch.int32, device=dev)
            out = self._actions
        io.actions = out.data_ptr()
        io.actions_dtype = L.U8 if out.dtype == torch.uint8 else L.I32
        L.check(env.lib.sus_env_select_actions(env._h, C.byref(io), env._stream()))
        self._keep = keep  # the Q tensors must outlive the (asynchronous) launch
        return out

    # ---- torch-op restatements
    @torch.no_grad()
    def act_grouped(self, featurizer, eps, flat_states):
        """Sync-free acting for n_imposters == 1 on the featurizer's stacked views with torch ops: the imposter network runs
        ONCE on N rows (row e = the view of env e's imposter), the crew network once per agent view on all N rows, and the
        per-env choice is made with `torch.where`.  Same distribution as `act`; identical actions for eps == 0."""
        env = self.env
        assert env.n_imposters == 1
        sp, ns = featurizer.stacked_views()
        N, A, dev = flat_states.shape[0], env.n_agents, flat_states.device
        alive = flat_states[:, 2 * A:3 * A] != 0
        imp = env.imposter_mask_batch
        imp_idx = torch.argmax(imp.to(torch.uint8), dim=1)  # the single imposter of every env, no host sync
        rows = torch.arange(N, device=dev)

        def view(t, k):  # tensor of agent view k (k: int or per-env index tensor)
            if t is None:
                return torch.zeros(N, featurizer.T, 1, device=dev)  # FlatFeaturizer's spatial placeholder
            if t.shape[0] == 1:
                return t[0]
            return t[k, rows] if isinstance(k, torch.Tensor) else t[k]

        def choose(model, n_act, spatial, non_spatial):
            explore = torch.rand(N, device=dev, generator=self.generator) <= eps
            rand_a = torch.randint(0, n_act, (N,), device=dev, generator=self.generator)
            return torch.where(explore, rand_a, torch.argmax(model(spatial, non_spatial), dim=1))

        a_imp = choose(self.imposter_model, env.n_imposter_actions, view(sp, imp_idx), view(ns, imp_idx))
        actions = torch.zeros((N, A), dtype=torch.int32, device=dev)
        for k in range(A):
            a_crew = choose(self.crew_model, env.n_crew_actions, view(sp, k), view(ns, k))
            a = torch.where(imp[:, k], a_imp, a_crew)
            actions[:, k] = torch.where(alive[:, k], a, torch.zeros_like(a)).to(torch.int32)
        return actions

    @torch.no_grad()
    def act(self, views, eps, flat_states, imposter_mask=None):
        """views: featurizer.generate_featurized_states(); flat_states (N, S): the states the views were made from
        (alive flags live at [2A, 3A)); returns (N, A) int32 role-list indices."""
        env = self.env
        N, A, dev = flat_states.shape[0], env.n_agents, flat_states.device
        alive = flat_states[:, 2 * A:3 * A] != 0
        imp = env.imposter_mask_batch if imposter_mask is None else imposter_mask
        actions = torch.zeros((N, A), dtype=torch.int32, device=dev)
        roles = ((self.imposter_model, env.n_imposter_actions), (self.crew_model, env.n_crew_actions))
        for k, (spatial, non_spatial) in enumerate(views):
            if self.dense:
                picks = []
                for model, n_act in roles:
                    explore = torch.rand(N, device=dev, generator=self.generator) <= eps  # train.py:363,374
                    rand_a = torch.randint(0, n_act, (N,), device=dev, generator=self.generator)
                    greedy = torch.argmax(model(spatial, non_spatial), dim=1)  # train.py:368-370,379-381
                    picks.append(torch.where(explore, rand_a, greedy))
                a = torch.where(imp[:, k], picks[0], picks[1])
                actions[:, k] = torch.where(alive[:, k], a, torch.zeros_like(a)).to(torch.int32)
                continue
            for mask, (model, n_act) in zip((imp[:, k] & alive[:, k], ~imp[:, k] & alive[:, k]), roles):
                idx = mask.nonzero(as_tuple=True)[0]
                if idx.numel() == 0:
                    continue
                explore = torch.rand(idx.numel(), device=dev, generator=self.generator) <= eps
                rand_a = torch.randint(0, n_act, (idx.numel(),), device=dev, generator=self.generator)
                greedy = torch.argmax(model(spatial[idx], non_spatial[idx]), dim=1)
                actions[idx, k] = torch.where(explore, rand_a, greedy).to(torch.int32)
        return actions


# ------------------------------------------------------------------------------------------------ training
def allreduce_grads(model, group=None):
    """Average gradients over ranks (one flattened NCCL all-reduce); a no-op outside torch.distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


class _FlatAdam:
    """torch.optim.Adam's update (defaults: no weight decay, no amsgrad) on ONE flat buffer holding all parameters of the
    optimizer, with a device-side gate: `apply()` does nothing -- weights, moments and step count all untouched -- when the
    (global) sample count of the subset is zero, which is what the reference's `continue` does for an empty team subset
    (train.py:88-92), without the host ever reading the count.

    The parameters become views of `flat`, their `.grad` views of `grad[:-2]`; `grad[-2]` carries the subset's sample count
    and `grad[-1]` its summed squared error, so a multi-GPU step is ONE all-reduce of `grad` (gradient sums, counts and loss
    sums add up; every sample of the global subset then weighs the same)."""

    def __init__(self, opt):
        g = opt.param_groups[0]
        assert isinstance(opt, torch.optim.Adam) and len(opt.param_groups) == 1 and not g.get("amsgrad", False) and \
            g.get("weight_decay", 0) == 0 and not g.get("maximize", False), "only plain Adam is restated here"
        self.opt = opt
        self.params = [p for p in g["params"] if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat = torch.empty(n, device=dev)
        self.grad = torch.zeros(n + 2, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k
        self.acc = torch.zeros(n, device=dev)  # gradients accumulated over the agent views (train.py:65-68: zero_grad once per train_step)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.step = torch.zeros((), device=dev, dtype=torch.float64)
        self.lr, (self.b1, self.b2), self.eps = float(g["lr"]), g["betas"], float(g["eps"])

    def state_dict(self):
        """Moments and step count (the wrapped torch optimizer's own state stays empty: its `step()` is never called)."""
        return {"m": self.m.clone(), "v": self.v.clone(), "step": self.step.clone(), "lr": self.lr, "betas": (self.b1, self.b2), "eps": self.eps}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.step.copy_(sd["step"])

    def begin_train_step(self):
        self.acc.zero_()

    def begin_view(self):
        self.grad.zero_()

    def apply(self, count, loss_sum):
        """count / loss_sum: 0-d device tensors of the LOCAL subset; returns the (global) mean loss (0 for an empty subset)."""
        self.grad[-2] = count
        self.grad[-1] = loss_sum
        if _distributed():
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
        cnt = self.grad[-2]
        on = cnt > 0
        denom = cnt.clamp(min=1.0)
        self.acc.add_(self.grad[:-2] / denom)  # the mean-MSE gradient of this view joins the accumulated one
        g = self.acc
        self.step.add_(on.to(self.step.dtype))
        m2 = torch.lerp(self.m, g, 1.0 - self.b1)
        v2 = self.v * self.b2 + (1.0 - self.b2) * g * g
        bc1 = (1.0 - self.b1 ** self.step).clamp(min=1e-30).float()
        bc2 = (1.0 - self.b2 ** self.step).clamp(min=1e-30).float()
        upd = (self.lr / bc1) * m2 / (v2.sqrt() / bc2.sqrt() + self.eps)
        self.m.copy_(torch.where(on, m2, self.m))
        self.v.copy_(torch.where(on, v2, self.v))
        self.flat.sub_(torch.where(on, upd, torch.zeros_like(upd)))
        return self.grad[-1] / denom


class DQNTeamTrainer:
    """`DQNTeamTrainer` (train.py:40-149): for every agent view, one optimizer step per team on the samples where
    that agent plays for the team; MSE between Q(s, a) and r + γ max_a' Q_target(s', a') (r alone where done).  Like the
    reference, gradients are zeroed once per train_step and ACCUMULATE over the agent views ("training via gradient
    accumulation", train.py:65-68,86-143).

    Sync-free restatement: instead of gathering each subset (`batch.imposters == agent_idx`, a boolean gather whose size
    the host has to learn) the loss of a (view, team) pair is the masked sum over the whole batch, and the optimizer step
    is gated on the device by the subset's sample count (`_FlatAdam`).  `static_imposters=True` (envs built with
    `shuffle_imposter_index=False`: the imposters are agents 0..n_imposters-1 in every sample) skips the (view, team) pairs
    that are empty by construction.  `train_step` returns the two summed losses as a (2,) DEVICE tensor."""

    def __init__(self, imposter_optimizer, crew_optimizer, gamma, static_imposters=None):
        self.imposter_optimizer, self.crew_optimizer, self.gamma = imposter_optimizer, crew_optimizer, gamma
        self.train = imposter_optimizer is not None or crew_optimizer is not None
        self.static_imposters = static_imposters  # None: unknown (use masks); int n: imposters are agents [0, n)
        self._adam = {}

    def _flat(self, opt):
        if id(opt) not in self._adam:
            self._adam[id(opt)] = _FlatAdam(opt)
        return self._adam[id(opt)]

    def optimizer_state_dict(self):
        """{"imposter": ..., "crew": ...}: Adam moments / step counts of the device-gated optimizer steps, for checkpoints."""
        return {name: self._flat(opt).state_dict() for name, opt in (("imposter", self.imposter_optimizer), ("crew", self.crew_optimizer))
                if opt is not None}

    def load_optimizer_state_dict(self, sd):
        for name, opt in (("imposter", self.imposter_optimizer), ("crew", self.crew_optimizer)):
            if opt is not None and name in sd:
                self._flat(opt).load_state_dict(sd[name])

    def train_step(self, batch, featurizer, imposter_model, imposter_target_model, crew_model, crew_target_model):
        dev = batch.states.device
        losses = torch.zeros(2, device=dev)
        if not self.train:
            return losses
        B = batch.states.shape[0]
        # one encode launch for states and next_states together (train.py:70-74)
        featurizer.fit(torch.cat([batch.states, batch.next_states], dim=0))
        sp, ns = featurizer.stacked_views()
        A = batch.actions.shape[1]
        T = ns.shape[2]
        placeholder = torch.zeros(B, T, 1, device=dev) if sp is None else None

        def view(t, k, lo):
            if t is None:
                return placeholder
            return (t[0] if t.shape[0] == 1 else t[k])[lo:lo + B]

        done = batch.dones.view(-1)
        teams = ((0, self.imposter_optimizer, imposter_model, imposter_target_model),
                 (1, self.crew_optimizer, crew_model, crew_target_model))
        for _idx, opt, _m, _t in teams:
            if opt is not None:
                self._flat(opt).begin_train_step()
        for agent_idx in range(A):
            is_imp = (batch.imposters == agent_idx).any(dim=1)  # train.py:81-82 (`.any` generalises n_imposters == 1)
            for loss_idx, opt, model, target in teams:
                if opt is None:
                    continue
                if self.static_imposters is not None and (agent_idx < self.static_imposters) != (loss_idx == 0):
                    continue  # empty by construction: the reference skips it too (train.py:88-92)
                adam = self._flat(opt)
                mask = (is_imp if loss_idx == 0 else ~is_imp).float()
                model.train()
                adam.begin_view()
                q = model(view(sp, agent_idx, 0), view(ns, agent_idx, 0))  # train.py:107-110
                # (rows of the OTHER team are masked out below; their action index may exceed this team's action count)
                a_idx = batch.actions[:, agent_idx].clamp(max=q.shape[1] - 1).view(-1, 1)
                values = torch.gather(q, 1, a_idx).view(-1)
                with torch.no_grad():
                    rewards = batch.rewards[:, agent_idx].view(-1)
                    tq = torch.max(target(view(sp, agent_idx, B), view(ns, agent_idx, B)), dim=1)[0]
                    target_values = torch.where(done, rewards, rewards + self.gamma * tq)  # train.py:124-137
                sq_sum = (((values - target_values) ** 2) * mask).sum()  # = count x F.mse_loss over the subset (train.py:139-143)
                sq_sum.backward()
                losses[loss_idx] += adam.apply(mask.sum(), sq_sum.detach())
        return losses


# ------------------------------------------------------------------------------------------------ the loop
class BatchedTrainingLoop:
    """The loop of train() (train.py:284-471) over a batched env.  Every iteration advances ALL envs one step: the networks
    read the feature tensors the previous fused step wrote (a `FeatureSequence` for T > 1), one kernel picks the ε-greedy
    actions, one kernel steps + encodes, one kernel stores the N transitions in the replay ring; every
    `train_step_interval` iterations a batch is sampled and `trainer.train_step` runs; every `target_update_interval` the
    target networks are synced.  Episode resets happen inside the step kernel; episode statistics accumulate on the device.

    use_graphs: capture [act + step + push] and [sample + train_step] as two CUDA graphs after a few eager iterations and
    replay them (device-resident env ticks, replay index / size and ε make the replays advance like eager calls)."""

    def __init__(self, env, replay_buffer, featurizer, imposter_model, crew_model, trainer, scheduler, batch_size=1024,
                 train_step_interval=5, target_update_interval=1000, use_graphs=False, imposter_index=None, fused_mlp=True):
        self.env, self.buf, self.feat, self.trainer, self.sched = env, replay_buffer, featurizer, trainer, scheduler
        self.imposter_model, self.crew_model = imposter_model, crew_model
        self.batch_size, self.train_every, self.target_every = batch_size, train_step_interval, target_update_interval
        self.use_graphs = use_graphs
        dev = env.device
        self.imposter_target = _copy_model(imposter_model, dev)
        self.crew_target = _copy_model(crew_model, dev)
        self.actor = BatchedActor(env, imposter_model, crew_model, fused_mlp=fused_mlp)
        if use_graphs:
            env.device_ticks(True)
        replay_buffer.attach(env, static_buffers=use_graphs)
        self.seq = FeatureSequence(env, featurizer, replay_buffer.trajectory_size)
        self.seq.start()
        self.eps = torch.zeros(1, dtype=torch.float32, device=dev)
        if imposter_index is None and env.n_imposters == 1 and not env.shuffle_imposter_index:
            imposter_index = torch.zeros(env.num_envs, dtype=torch.int64, device=dev)  # base.py:278: agent 0
        self.imposter_index = imposter_index
        self.it = 0
        self.losses = []
        self._g_iter = self._g_train = None
        self._loss_buf = torch.zeros(2, device=dev)
        for opt in (trainer.imposter_optimizer, trainer.crew_optimizer):
            if opt is not None:
                trainer._flat(opt)  # flatten the parameters NOW: a graph captured later must see their final addresses
        if trainer.static_imposters is None and not env.shuffle_imposter_index:
            trainer.static_imposters = env.n_imposters  # base.py:278: the imposters are agents 0..n_imposters-1

    # one iteration = train.py:346-399 for all envs
    def _iteration(self):
        sp, ns = self.seq.views()
        actions = self.actor.act_kernel(sp, ns, self.eps, imposter_index=self.imposter_index)
        _nf, _r, done, trunc, _ = self.buf.collect_step(actions, featurizer=self.feat)
        self.seq.push(done, trunc)

    def _train(self):
        batch = self.buf.sample(self.batch_size)
        self._loss_buf.copy_(self.trainer.train_step(batch, self.feat_train, self.imposter_model, self.imposter_target,
                                                     self.crew_model, self.crew_target))

    def _sync_targets(self):  # train.py:341-343
        for m, t in ((self.imposter_model, self.imposter_target), (self.crew_model, self.crew_target)):
            if m is not None and t is not None:
                with torch.no_grad():
                    for p, q in zip(m.parameters(), t.parameters()):
                        q.copy_(p)
                    for p, q in zip(m.buffers(), t.buffers()):
                        q.copy_(p)

    @property
    def feat_train(self):
        """The train step featurizes replay batches with its OWN featurizer instance, so that its output buffers never alias
        the acting features the fused step kernel writes."""
        if getattr(self, "_feat_train", None) is None:
            self._feat_train = self.feat.clone_for(self.env)
        return self._feat_train

    def run(self, num_iterations, on_iteration=None):
        warmup_eager = self.train_every + 1  # at least one eager iteration AND one eager train step before any capture
        for _ in range(num_iterations):
            it = self.it
            if it % self.target_every == 0:
                self._sync_targets()
            self.eps.fill_(self.sched.value(it))
            train_now = it % self.train_every == 0 and self.trainer.train
            if not self.use_graphs or it < warmup_eager:
                self._iteration()
                if train_now:
                    self._train()
                    self.losses.append(self._loss_buf.clone())
            else:
                if self._g_iter is None:
                    self._g_iter = self._capture(self._iteration)
                else:
                    self._g_iter.replay()
                if train_now:
                    if self._g_train is None:
                        self._g_train = self._capture(self._train)
                    else:
                        self._g_train.replay()
                    self.losses.append(self._loss_buf.clone())
            self.it += 1
            if on_iteration is not None:
                on_iteration(it)
        return self.losses

    def _capture(self, fn):
        """Capture `fn` (already run eagerly at least once: kernel attributes, buffers, cuBLAS workspaces exist) into a graph.
        The capture itself executes nothing, so the graph is replayed once right away to perform this iteration's work."""
        dev = self.env.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        g.replay()
        return g

    def finish(self):
        """Host-visible bookkeeping after graph replays; returns the losses as a list of [imposter_loss, crew_loss]."""
        self.buf.sync_host_counters()
        self.env.check_actions()  # every action the actor produced was inside its agent's role list
        return [l.tolist() for l in self.losses]


def train_batched(env, replay_buffer, featurizer, imposter_model, crew_model, trainer, scheduler, num_iterations,
                  batch_size=1024, train_step_interval=5, target_update_interval=1000, generator=None,
                  on_iteration=None, use_graphs=False):
    """Convenience wrapper: build a `BatchedTrainingLoop`, run `num_iterations`, return the list of
    [imposter_loss, crew_loss] (one host read at the very end)."""
    loop = BatchedTrainingLoop(env, replay_buffer, featurizer, imposter_model, crew_model, trainer, scheduler,
                               batch_size=batch_size, train_step_interval=train_step_interval,
                               target_update_interval=target_update_interval, use_graphs=use_graphs)
    loop.run(num_iterations, on_iteration=on_iteration)
    return loop.finish()


def _copy_model(model, device):
    if model is None:
        return None
    import copy

    m = model.create_copy() if hasattr(model, "create_copy") else copy.deepcopy(model)
    return m.to(device)
