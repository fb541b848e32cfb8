"""Batched restatement of the callers of the hot path: the ε-greedy acting part of
`train()` (src/train.py:351-381), `ExponentialSchedule` (src/scheduler.py:4-40), `DQNTeamTrainer.train_step`
(src/train.py:40-149) and a batched training loop with the reference's cadence (target sync, train interval,
auto-reset inside the env kernel, replay writes on the GPU).  Q-networks are out of scope: any `torch.nn.Module`
with the reference's `forward(spatial, non_spatial) -> (B, n_actions)` signature works (e.g. the reference's own
`MLP` / `SpatialDQN` moved to the GPU).

Multi-GPU: one process per GPU, each with its own env shard and replay ring; gradients are averaged with one
NCCL all-reduce per optimizer step (`allreduce_grads`), episode statistics with `reduce_episode_stats`.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F


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


class BatchedActor:
    """train.py:351-381 for every env at once: per agent view, alive imposters act with `imposter_model`, alive crew
    with `crew_model`, each ε-greedy over its role's action count; dead agents keep action 0.

    dense=False (default) gathers the rows of each role first and evaluates each network only on its own rows (one
    `nonzero` sync per role and view), like the reference evaluates one row per agent; dense=True evaluates both networks
    on all N rows and selects with `torch.where` -- no host synchronisation, but A x more network rows (measured on the
    cfg5 shape: 2.2e7 vs 4.6e7 env-steps/s in the training loop, so gather is the default)."""

    def __init__(self, env, imposter_model, crew_model, generator=None, dense=False):
        self.env, self.imposter_model, self.crew_model = env, imposter_model, crew_model
        self.generator, self.dense = generator, dense

    @torch.no_grad()
    def act_grouped(self, featurizer, eps, flat_states):
        """Sync-free acting for n_imposters == 1 on the featurizer's stacked views: the imposter network runs ONCE on N
        rows (row e = the view of env e's imposter), the crew network once per agent view on all N rows, and the per-env
        choice is made with `torch.where`.  Same distribution as `act`; identical actions for eps == 0."""
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


class DQNTeamTrainer:
    """`DQNTeamTrainer` (train.py:40-149): for every agent view, one optimizer step per team on the samples where
    that agent plays for the team; MSE between Q(s, a) and r + γ max_a' Q_target(s', a') (r alone where done)."""

    def __init__(self, imposter_optimizer, crew_optimizer, gamma):
        self.imposter_optimizer, self.crew_optimizer, self.gamma = imposter_optimizer, crew_optimizer, gamma
        self.train = imposter_optimizer is not None or crew_optimizer is not None

    def train_step(self, batch, featurizer, imposter_model, imposter_target_model, crew_model, crew_target_model):
        losses = [0, 0]
        if not self.train:
            return losses
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        for opt in (self.imposter_optimizer, self.crew_optimizer):
            if opt is not None:
                opt.zero_grad()
        featurizer.fit(batch.states)  # train.py:70-74
        feat_state = featurizer.generate_featurized_states()
        feat_state = [(sp.detach().clone(), ns.detach().clone()) for sp, ns in feat_state]  # fit() reuses its buffers
        featurizer.fit(batch.next_states)
        feat_next = featurizer.generate_featurized_states()
        for agent_idx, (state_feat, next_feat) in enumerate(zip(feat_state, feat_next)):
            # samples in which this agent is an imposter / crew member (train.py:81-82; `.any` generalises the
            # reference's n_imposters == 1 comparison)
            imposter_samples = (batch.imposters == agent_idx).any(dim=1)
            crew_samples = ~imposter_samples
            for loss_idx, (opt, samples, model, target) in enumerate((
                    (self.imposter_optimizer, imposter_samples, imposter_model, imposter_target_model),
                    (self.crew_optimizer, crew_samples, crew_model, crew_target_model))):
                if opt is None or (not distributed and samples.sum() == 0):
                    continue
                model.train()
                # NOTE: like the reference, gradients are zeroed once per train_step and ACCUMULATE over the agent views
                # ("training via gradient accumulation", train.py:65-68,86-143)
                if samples.sum() > 0:  # (distributed: an empty subset still takes part in the all-reduce below)
                    q = model(state_feat[0][samples], state_feat[1][samples])  # train.py:107-110
                    actions = batch.actions[samples, agent_idx]
                    values = torch.gather(q, 1, actions.view(-1, 1)).view(-1)
                    with torch.no_grad():
                        done_mask = batch.dones[samples].view(-1)
                        rewards = batch.rewards[samples, agent_idx].view(-1)
                        target_values = rewards + self.gamma * torch.max(
                            target(next_feat[0][samples].detach(), next_feat[1][samples].detach()), dim=1)[0]
                        target_values[done_mask] = rewards[done_mask]
                    loss = F.mse_loss(values, target_values)  # train.py:139-143
                    loss.backward()
                    losses[loss_idx] += loss.item()
                allreduce_grads(model)
                opt.step()
        return losses


def train_batched(env, replay_buffer, featurizer, imposter_model, crew_model, trainer, scheduler, num_iterations,
                  batch_size=1024, train_step_interval=5, target_update_interval=1000, generator=None,
                  on_iteration=None):
    """The loop of train() (train.py:284-471) over a batched env: every iteration advances ALL envs one step
    (ε-greedy acting on the GPU, fused step, replay push), trains every `train_step_interval` iterations and syncs
    the target networks every `target_update_interval`.  Episode resets happen inside the step kernel; episode
    statistics accumulate on the device (env.episode_stats()).  Returns the list of [imposter_loss, crew_loss]."""
    imposter_target = imposter_model.create_copy() if hasattr(imposter_model, "create_copy") else _copy(imposter_model)
    crew_target = crew_model.create_copy() if hasattr(crew_model, "create_copy") else _copy(crew_model)
    imposter_target.to(env.device); crew_target.to(env.device)
    actor = BatchedActor(env, imposter_model, crew_model, generator=generator)
    replay_buffer.attach(env)
    losses = []
    for it in range(num_iterations):
        if it % target_update_interval == 0:  # train.py:341-343
            imposter_target.load_state_dict(imposter_model.state_dict())
            crew_target.load_state_dict(crew_model.state_dict())
        seq = replay_buffer.state_sequence
        featurizer.fit(seq)  # train.py:346-348
        if env.n_imposters == 1:
            actions = actor.act_grouped(featurizer, scheduler.value(it), seq[:, -1])
        else:
            actions = actor.act(featurizer.generate_featurized_states(), scheduler.value(it), seq[:, -1])
        replay_buffer.collect_step(actions)  # env.step + replay add (train.py:383-399)
        if it % train_step_interval == 0:  # train.py:402-416
            batch = replay_buffer.sample(batch_size, generator=generator)
            losses.append(trainer.train_step(batch, featurizer, imposter_model, imposter_target, crew_model, crew_target))
        if on_iteration is not None:
            on_iteration(it)
    return losses


def _copy(model):
    import copy

    return copy.deepcopy(model)
