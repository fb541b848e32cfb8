#!/usr/bin/env python
"""Mint the golden fixtures under tests/golden/ by running the UNMODIFIED reference (through the gymnasium
stand-in and the draw-injection harness).  The reference holds no golden vectors of its own, so
these files are how its behaviour travels to the GPU box.

    python tools/make_golden.py            # writes tests/golden/*.npz

Two trajectory fixtures per case:
  <case>.philox.npz  draws follow the susnet Philox spec for the stored seed (pins Philox + draw derivation)
  <case>.words.npz   draws are arbitrary injected 32-bit words, incl. 0 and 0xffffffff (pins the derivation's
                     edge cases independently of Philox)
and one feature fixture per case with the reference featurizers' outputs on states from the trajectory.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle import rng_spec as R  # noqa: E402
from tests.cases import CASES, FLAT_COMPONENT_SETS, GLOBAL_CASES  # noqa: E402
from tools.check_oracle_vs_reference import reference_features  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_ENVS, N_STEPS, SEED, ENV_ID_BASE = 12, 160, 0xB200, 1000


def edge_words(rng, shape):
    w = rng.integers(0, 2**32, size=shape, dtype=np.uint64).astype(np.uint32)
    m = rng.random(shape)
    w[m < 0.05] = 0
    w[m > 0.95] = 0xFFFFFFFF
    return w


def trajectory(name, injected):
    cfg = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
    ref = H.ReferenceBatch(cfg, N_ENVS, SEED, env_id_base=ENV_ID_BASE)
    A, J, nI = ref.A, ref.J, ref.cfg["n_imposters"]
    n_rs, n_ss = R.n_reset_slots(nI, A, J), R.n_step_slots(A)
    rec = dict(actions=[], next_flat=[], rewards=[], done=[], trunc=[], metrics=[], cur_flat=[], imp=[],
               step_words=[], reset_words=[], act_words=[])
    w0 = edge_words(rng, (N_ENVS, n_rs)) if injected else None
    reset_flat = ref.reset(reset_words=w0)

    def imp_mask():
        return np.stack([np.asarray(e.imposter_mask, dtype=np.uint8) for e in ref.envs])

    reset_imp = imp_mask()
    for t in range(N_STEPS):
        sw = edge_words(rng, (N_ENVS, n_ss)) if injected else None
        rw = edge_words(rng, (N_ENVS, n_rs)) if injected else None
        aw = edge_words(rng, (N_ENVS, A)) if injected else None
        a = ref.sample_actions(act_words=aw)
        if t % 7 == 3:  # also drive some steps with non-random actions: everybody uses the last role action
            a = np.stack([[len(e.agent_action_map[i]) - 1 - (t // 7 + i) % 2 for i in range(A)] for e in ref.envs])
        o = ref.step(a, step_words=sw, reset_words=rw)
        rec["actions"].append(a); rec["next_flat"].append(o["next_flat"]); rec["rewards"].append(o["rewards"])
        rec["done"].append(o["done"]); rec["trunc"].append(o["trunc"]); rec["metrics"].append(o["metrics"])
        rec["cur_flat"].append(ref.flat_states()); rec["imp"].append(imp_mask())
        if injected:
            rec["step_words"].append(sw); rec["reset_words"].append(rw); rec["act_words"].append(aw)
    out = dict(
        case=name, seed=SEED, env_id_base=ENV_ID_BASE, injected=injected, reset_flat=reset_flat.astype(np.int16),
        reset_imp=reset_imp, actions=np.array(rec["actions"], dtype=np.int8),
        next_flat=np.array(rec["next_flat"], dtype=np.int16), rewards=np.array(rec["rewards"], dtype=np.float64),
        done=np.array(rec["done"], dtype=np.uint8), trunc=np.array(rec["trunc"], dtype=np.uint8),
        metrics=np.array(rec["metrics"], dtype=np.int32), cur_flat=np.array(rec["cur_flat"], dtype=np.int16),
        imp=np.array(rec["imp"], dtype=np.uint8),
    )
    if injected:
        out.update(reset_words0=w0, step_words=np.array(rec["step_words"]), reset_words=np.array(rec["reset_words"]),
                   act_words=np.array(rec["act_words"]))
    return out, ref.envs[0]


def features(name, env, flat):
    cfg = CASES[name]
    out = dict(case=name, flat=flat.astype(np.int16))
    if name in GLOBAL_CASES:
        sp, ns = reference_features(env, "global", flat)
        for k in range(1, sp.shape[0]):
            assert np.array_equal(sp[0], sp[k])
        out["global_spatial"] = sp[0].astype(np.uint8)  # exact 0/1 planes
        assert np.array_equal(out["global_spatial"].astype(np.float32), sp[0])
        out["global_non_spatial"] = ns
        sp, ns = reference_features(env, "perspective", flat)
        out["perspective_spatial"] = sp.astype(np.uint8)
        assert np.array_equal(out["perspective_spatial"].astype(np.float32), sp)
        out["perspective_non_spatial"] = ns
    for i, comps in enumerate(FLAT_COMPONENT_SETS.get(name, [])):
        _, ns = reference_features(env, "flat", flat, comps)
        out[f"flat{i}_components"] = np.array(comps)
        out[f"flat{i}"] = ns[0]
    del cfg
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    replay_fixture("cfg4_base_1v4", T=1, M=100)
    replay_fixture("base_2v3_j3", T=3, M=64)
    replay_fixture("tagging_2v5_short", T=2, M=50)
    train_step_fixture()
    acting_fixture("acting.cfg4_global_spatial_T2", dict(CASES["cfg4_base_1v4"], max_time_steps=40), "global", T=2, steps=260)
    acting_fixture("acting.cfg4alt_flat98_mlp_T1", dict(CASES["cfg4alt_itg_1v4"], shuffle_imposter_index=True), "flat", T=1,
                   steps=1010)
    if "--replay-only" in sys.argv or "--acting-only" in sys.argv:
        return
    for name in CASES:
        for injected in (False, True):
            tr, env = trajectory(name, injected)
            path = os.path.join(OUT, f"{name}.{'words' if injected else 'philox'}.npz")
            np.savez_compressed(path, **tr)
            eps = int(((tr["done"] | tr["trunc"]) != 0).sum())
            print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, {eps} finished episodes")
            if not injected:
                flat = np.concatenate([tr["next_flat"][::8].reshape(-1, tr["next_flat"].shape[-1]),
                                       tr["reset_flat"]]).astype(np.int64)
                ft = features(name, env, flat)
                if len(ft) > 2:
                    fpath = os.path.join(OUT, f"{name}.features.npz")
                    np.savez_compressed(fpath, **ft)
                    print(f"{fpath}: {os.path.getsize(fpath) / 1024:.1f} KiB, {flat.shape[0]} states")




def replay_fixture(name, T, M, n_envs=6, n_steps=45):
    """The reference's own ReplayBuffer (src/replay_memory.py) fed the way populate()/train() feed it, N envs in lock
    step (ring order: step-major, env-minor).  Pins the layout, dtypes and the T-deep np.roll sequence bookkeeping."""
    import torch

    H.import_reference()
    from src.replay_memory import ReplayBuffer as RefReplay

    cfg = CASES[name]
    ref = H.ReferenceBatch(cfg, n_envs, SEED, env_id_base=ENV_ID_BASE)
    flat = ref.reset()
    S = flat.shape[1]
    buf = RefReplay(max_size=M, state_size=S, trajectory_size=T, n_agents=ref.A, n_imposters=ref.cfg["n_imposters"])
    for t_ in (buf.states, buf.next_states, buf.rewards):
        t_.zero_()
    buf.actions.zero_(); buf.dones.zero_(); buf.imposters.zero_()
    seqs = [np.repeat(flat[i][None].astype(np.float64), T, axis=0) for i in range(n_envs)]  # replay_memory.py:107-112
    actions = []
    for t in range(n_steps):
        imps = [np.asarray(e.imposter_idxs).copy() for e in ref.envs]
        a = ref.sample_actions()
        o = ref.step(a)
        cur = ref.flat_states()
        actions.append(a)
        for i in range(n_envs):
            nxt = np.roll(seqs[i].copy(), -1, axis=0)  # replay_memory.py:121-126
            nxt[-1] = o["next_flat"][i]
            buf.add(state=seqs[i], action=a[i], reward=o["rewards"][i], next_state=nxt, done=bool(o["done"][i]),
                    imposters=imps[i])
            if o["done"][i] or o["trunc"][i]:
                seqs[i] = np.repeat(cur[i][None].astype(np.float64), T, axis=0)  # train.py:440-445
            else:
                seqs[i] = nxt
    out = dict(case=name, T=T, M=M, n_envs=n_envs, seed=SEED, env_id_base=ENV_ID_BASE,
               actions=np.array(actions, dtype=np.int8), states=buf.states.numpy(), r_actions=buf.actions.numpy(),
               rewards=buf.rewards.numpy(), next_states=buf.next_states.numpy(), dones=buf.dones.numpy(),
               imposters=buf.imposters.numpy(), idx=buf.idx, size=buf.size,
               final_seq=np.stack(seqs).astype(np.float32))
    path = os.path.join(OUT, f"{name}.replay_T{T}.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB")
    del torch


def train_step_fixture():
    """The reference's DQNTeamTrainer.train_step (src/train.py:50-149) on a batch from its own ReplayBuffer with its
    own MLP Q-networks and FlatFeaturizer: losses and updated weights are the known answers for the GPU trainer."""
    import torch

    H.import_reference()
    from src.features import (AliveCrewFeaturizer, ClosestAliveCrewFeaturizer, CompositeFeaturizer, FlatFeaturizer,
                              OneHotAgentPositionFeaturizer)
    from src.models.dqn import MLP
    from src.replay_memory import ReplayBuffer as RefReplay
    from src.train import DQNTeamTrainer

    name, T, M, n_envs, n_steps = "cfg4alt_itg_1v4", 2, 96, 8, 12
    cfg = dict(CASES[name]); cfg["shuffle_imposter_index"] = True
    ref = H.ReferenceBatch(cfg, n_envs, SEED, env_id_base=ENV_ID_BASE)
    flat = ref.reset()
    S = flat.shape[1]
    buf = RefReplay(max_size=M, state_size=S, trajectory_size=T, n_agents=ref.A, n_imposters=1)
    seqs = [np.repeat(flat[i][None].astype(np.float64), T, axis=0) for i in range(n_envs)]
    for t in range(n_steps):
        imps = [np.asarray(e.imposter_idxs).copy() for e in ref.envs]
        a = ref.sample_actions()
        o = ref.step(a)
        cur = ref.flat_states()
        for i in range(n_envs):
            nxt = np.roll(seqs[i].copy(), -1, axis=0); nxt[-1] = o["next_flat"][i]
            buf.add(state=seqs[i], action=a[i], reward=o["rewards"][i], next_state=nxt, done=bool(o["done"][i]), imposters=imps[i])
            seqs[i] = np.repeat(cur[i][None].astype(np.float64), T, axis=0) if (o["done"][i] or o["trunc"][i]) else nxt
    torch.manual_seed(7)
    batch = buf.sample(48)
    env = ref.envs[0]
    feat = FlatFeaturizer(env, CompositeFeaturizer([OneHotAgentPositionFeaturizer(env), AliveCrewFeaturizer(env),
                                                    ClosestAliveCrewFeaturizer(env)]))
    F_ = 98 * T
    torch.manual_seed(11)
    imp_model, crew_model = MLP([F_, 32, 16, 6]), MLP([F_, 24, 5])
    imp_target, crew_target = imp_model.create_copy(), crew_model.create_copy()
    with torch.no_grad():  # make the targets differ from the online nets
        for p_ in list(imp_target.parameters()) + list(crew_target.parameters()):
            p_.add_(0.01 * torch.randn_like(p_))
    before = {f"{n}.{k}": v.detach().clone().numpy() for n, m in (("imp", imp_model), ("crew", crew_model),
              ("imp_target", imp_target), ("crew_target", crew_target)) for k, v in m.state_dict().items()}
    trainer = DQNTeamTrainer(torch.optim.Adam(imp_model.parameters(), lr=1e-3),
                             torch.optim.Adam(crew_model.parameters(), lr=1e-3), gamma=0.9)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses = trainer.train_step(batch, feat, imp_model, imp_target, crew_model, crew_target)
    after = {f"{n}.{k}": v.detach().clone().numpy() for n, m in (("imp", imp_model), ("crew", crew_model))
             for k, v in m.state_dict().items()}
    out = dict(T=T, losses=np.array(losses, dtype=np.float64), states=batch.states.numpy(), actions=batch.actions.numpy(),
               rewards=batch.rewards.numpy(), next_states=batch.next_states.numpy(), dones=batch.dones.numpy(),
               imposters=batch.imposters.numpy())
    out.update({f"before.{k}": v for k, v in before.items()})
    out.update({f"after.{k}": v for k, v in after.items()})
    path = os.path.join(OUT, "train_step.cfg4alt_flat98_T2.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, losses {losses}")


def acting_fixture(tag, cfg, kind, T, steps):
    """Row f2: the reference's OWN greedy acting (src/train.py:349-381, unmodified `train()` with epsilon ~ 0 and a trainer
    without optimizers, so nothing is trained) on its own env, featurizer and randomly initialised networks.  The replay
    ring it fills holds, per transition, the state sequence the actions were taken from, the imposter ids and the actions:
    known answers for `BatchedActor` / `sus_env_select_actions` at eps = 0."""
    import contextlib
    import io
    import pathlib
    import tempfile

    import torch

    train_mod, replay_mod, dqn_mod, metrics_mod, sched_mod = H.import_reference_training()
    _, feat_mod = H.import_reference()

    class RandomlyDriven(H.DrawDrivenReferenceEnv):
        """The env advances on RANDOM actions (diverse states, kills, dead agents) while train() records the greedy actions
        its unmodified acting code chose for each state: the recorded (state, action) pairs are what the fixture pins."""

        def step(self, agent_actions):
            return super().step(self.sample_actions())

    env = RandomlyDriven(cfg, SEED, env_id=ENV_ID_BASE)
    e = env._env
    if kind == "global":
        feat = feat_mod.GlobalFeaturizer(e)
        ns = int(feat.featurized_shape[1][0])
        spatial = dict(input_image_size=9, non_spatial_input_size=ns, n_channels=[e.n_agents + 2, 6, 6], strides=[1, 1],
                       paddings=[1, 1], kernel_size=[3, 3], dilations=[1, 1], rnn_layers=1, rnn_hidden_dim=24,
                       rnn_dropout=0.0, mlp_hidden_layer_dims=[16])
        torch.manual_seed(5)
        imp = dqn_mod.SpatialDQN(n_actions=e.n_imposter_actions, **spatial)
        crew = dqn_mod.SpatialDQN(n_actions=e.n_crew_actions, **spatial)
    else:
        feat = feat_mod.FlatFeaturizer(e, feat_mod.CompositeFeaturizer([feat_mod.OneHotAgentPositionFeaturizer(e),
                                                                        feat_mod.AliveCrewFeaturizer(e),
                                                                        feat_mod.ClosestAliveCrewFeaturizer(e)]))
        torch.manual_seed(5)
        imp = dqn_mod.MLP([98 * T, 48, 24, e.n_imposter_actions])
        crew = dqn_mod.MLP([98 * T, 32, e.n_crew_actions])
    with torch.no_grad():  # larger weights: the argmax depends on the input instead of on the output biases
        for m in (imp, crew):
            for p_ in m.parameters():
                if p_.dim() > 1:
                    p_.mul_(4.0)
    np.random.seed(99)
    rb = replay_mod.ReplayBuffer(max_size=steps, trajectory_size=T, state_size=e.flattened_state_size,
                                 n_imposters=e.n_imposters, n_agents=e.n_agents)
    trainer = train_mod.DQNTeamTrainer(imposter_optimizer=None, crew_optimizer=None, gamma=0.9)
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        train_mod.train(env=env, metrics=metrics_mod.EpisodicMetricHandler(), num_steps=steps, replay_buffer=rb, featurizer=feat,
                        imposter_model=imp, crew_model=crew, scheduler=sched_mod.ExponentialSchedule(1e-12, 1e-12, 2),
                        save_directory_path=pathlib.Path(d), trainer=trainer, train_step_interval=5, batch_size=8, num_saves=2,
                        target_update_interval=10_000)
    assert rb.size == steps
    out = dict(kind=kind, T=T, states=rb.states.numpy(), actions=rb.actions.numpy().astype(np.int8),
               imposters=rb.imposters.numpy(), cfg_json=np.frombuffer(__import__("json").dumps(cfg).encode(), dtype=np.uint8))
    out.update({f"imp.{k}": v.detach().numpy() for k, v in imp.state_dict().items()})
    out.update({f"crew.{k}": v.detach().numpy() for k, v in crew.state_dict().items()})
    path = os.path.join(OUT, f"{tag}.npz")
    np.savez_compressed(path, **out)
    alive = rb.states.numpy()[:, -1, 2 * e.n_agents:3 * e.n_agents] != 0
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, {steps} greedy steps, {int((~alive).sum())} dead-agent slots, "
          f"action histogram {np.bincount(out['actions'].reshape(-1).astype(np.int64)).tolist()}")


if __name__ == "__main__":
    if "--acting-only" in sys.argv:
        acting_fixture("acting.cfg4_global_spatial_T2", dict(CASES["cfg4_base_1v4"], max_time_steps=40), "global", T=2, steps=260)
        acting_fixture("acting.cfg4alt_flat98_mlp_T1", dict(CASES["cfg4alt_itg_1v4"], shuffle_imposter_index=True), "flat", T=1,
                       steps=1010)
    else:
        main()
