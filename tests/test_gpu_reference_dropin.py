"""GPU + the UNMODIFIED reference side by side (needs the reference's sources: `/root/reference` in the build container,
the staged `baseline/_ref` of tools/stage_reference.py on the GPU box; skipped when neither exists).

1. Drop-in proof under the reference's OWN callers: `src.replay_memory.ReplayBuffer.populate` (replay_memory.py:96-143) and
   `src.train.train` (train.py:284-471) -- unmodified -- drive (a) the reference's env + featurizer fed with the Philox draws
   and (b) `sus_net_b200`'s reference-mode env + featurizer (`output_device="cpu"`), from the same numpy / torch seeds.
   Everything the callers produce must be IDENTICAL: replay tensors, per-step losses, episode metrics, final weights.
2. Direct reference <-> CUDA trajectories (no oracle in between): N reference envs in lock step with a batched CUDA env on
   the same draws, every case of tests.cases.CASES.
"""
import numpy as np
import pytest
import torch

from oracle import ref_harness as H
from tests.cases import CASES, GLOBAL_CASES
from tests.util import make_cuda_env, reward_bits

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.reference_available(), reason="reference sources not staged")]


def cpu(t):
    return t.detach().cpu().numpy()


SCENARIOS = {
    # BASELINE configs[3]'s env with the reference's spatial recipe (notebooks/experiment.ipynb): both teams learn, T = 2
    # (max_time_steps = 60 so that episodes end inside a short run: train() indexes an empty list otherwise, train.py:462)
    "cfg4_global_spatial_dqn_T2": dict(case="cfg4_base_1v4", featurizer="global", T=2, crew="spatial_dqn", steps=300,
                                       override=dict(max_time_steps=60)),
    # the reference's 1-imposter recipe scaled to 4 crew (notebooks/experiment_1v1.ipynb, BASELINE configs[4]): T = 1
    # (ImposterTrainingGround fixes max_time_steps = 1000: 1100 steps reach the first truncation)
    "cfg5_flat98_mlp_vs_random_T1": dict(case="cfg4alt_itg_1v4", featurizer="flat", T=1, crew="random", steps=1100, override={}),
}


def build_models(dqn_mod, env, kind, crew, ns_size):
    spatial = dict(input_image_size=9, non_spatial_input_size=ns_size, n_channels=[env.n_agents + 2, 4, 4], strides=[1, 1],
                   paddings=[1, 1], kernel_size=[3, 3], dilations=[1, 1], rnn_layers=1, rnn_hidden_dim=16, rnn_dropout=0.0,
                   mlp_hidden_layer_dims=[16])
    if kind == "global":
        imp = dqn_mod.ModelType.build(dqn_mod.ModelType.SPATIAL_DQN, n_actions=env.n_imposter_actions, **spatial)
    else:
        imp = dqn_mod.ModelType.build(dqn_mod.ModelType.MLP, layer_dims=[ns_size, 32, 16, env.n_imposter_actions])
    if crew == "random":
        crw = dqn_mod.ModelType.build(dqn_mod.ModelType.RANDOM, n_actions=env.n_crew_actions)
    else:
        crw = dqn_mod.ModelType.build(dqn_mod.ModelType.SPATIAL_DQN, n_actions=env.n_crew_actions, **spatial)
    return imp, crw


def run_reference_callers(env, featurizer, sc, tmp_path, tag):
    """The reference's own populate() + train() on `env` / `featurizer` (whatever implements them)."""
    train_mod, replay_mod, dqn_mod, metrics_mod, sched_mod = H.import_reference_training()
    np.random.seed(1234)
    torch.manual_seed(1234)
    T = sc["T"]
    ns_size = int(featurizer.featurized_shape[1][0]) * (T if sc["featurizer"] == "flat" else 1)
    imp, crw = build_models(dqn_mod, env, sc["featurizer"], sc["crew"], ns_size)
    rb = replay_mod.ReplayBuffer(max_size=2000, trajectory_size=T, state_size=env.flattened_state_size,
                                 n_imposters=env.n_imposters, n_agents=env.n_agents)
    rb.populate(env=env, num_steps=300)  # replay_memory.py:96-143, unmodified
    after_populate = {k: getattr(rb, k).clone() for k in ("states", "actions", "rewards", "next_states", "dones", "imposters")}
    size_after_populate = rb.size
    trainer = train_mod.DQNTeamTrainer(
        imposter_optimizer=train_mod.OptimizerType.build(train_mod.OptimizerType.ADAM, imp, 1e-3),
        crew_optimizer=train_mod.OptimizerType.build(train_mod.OptimizerType.ADAM, crw, 1e-3), gamma=0.9)
    metrics = metrics_mod.EpisodicMetricHandler()
    out_dir = tmp_path / tag
    out_dir.mkdir()
    train_mod.train(env=env, metrics=metrics, num_steps=sc["steps"], replay_buffer=rb, featurizer=featurizer, imposter_model=imp,
                    crew_model=crw, scheduler=sched_mod.ExponentialSchedule(1.0, 0.05, 200), save_directory_path=out_dir,
                    trainer=trainer, train_step_interval=5, batch_size=16, gamma=0.9, num_saves=3,
                    target_update_interval=50)  # train.py:284-471, unmodified
    return dict(populate=after_populate, size_after_populate=size_after_populate, size=rb.size, idx=rb.idx,
                replay={k: getattr(rb, k).clone() for k in after_populate}, metrics=metrics.metrics,
                weights=[p.detach().clone() for m in (imp, crw) for p in m.parameters()],
                files=sorted(p.name for p in out_dir.iterdir()))


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_reference_populate_and_train_run_identically_on_the_dropin(cuda_lib, tmp_path, name):
    import sus_net_b200 as S

    sc = SCENARIOS[name]
    cfg = dict(CASES[sc["case"]], **sc["override"])
    seed = 77
    env_mod, feat_mod = H.import_reference()
    # (a) the reference's env + featurizer, fed with the draws the GPU env makes
    ref_env = H.DrawDrivenReferenceEnv(cfg, seed)
    # (b) the drop-in: reference-mode GPU env + GPU featurizer handing CPU tensors to the reference's CPU models
    gpu_env = make_cuda_env(cfg, 1, seed=seed, auto_reset=False, batched=False)
    if sc["featurizer"] == "global":
        ref_feat = feat_mod.GlobalFeaturizer(ref_env._env)
        gpu_feat = S.GlobalFeaturizer(gpu_env, output_device="cpu")
    else:
        parts = (feat_mod.OneHotAgentPositionFeaturizer, feat_mod.AliveCrewFeaturizer, feat_mod.ClosestAliveCrewFeaturizer)
        ref_feat = feat_mod.FlatFeaturizer(ref_env._env, feat_mod.CompositeFeaturizer([p(ref_env._env) for p in parts]))
        gparts = (S.OneHotAgentPositionFeaturizer, S.AliveCrewFeaturizer, S.ClosestAliveCrewFeaturizer)
        gpu_feat = S.FlatFeaturizer(gpu_env, S.CompositeFeaturizer([p(gpu_env) for p in gparts]), output_device="cpu")
    a = run_reference_callers(ref_env, ref_feat, sc, tmp_path, "ref")
    b = run_reference_callers(gpu_env, gpu_feat, sc, tmp_path, "gpu")
    assert a["size_after_populate"] == b["size_after_populate"] >= 300
    n = a["size_after_populate"]
    for k in a["populate"]:
        assert torch.equal(a["populate"][k][:n], b["populate"][k][:n]), f"{name}: populate() filled `{k}` differently"
    assert (a["size"], a["idx"]) == (b["size"], b["idx"]) and a["size"] >= min(2000, 300 + sc["steps"])
    for k in a["replay"]:
        assert torch.equal(a["replay"][k][:a["size"]], b["replay"][k][:a["size"]]), f"{name}: train() stored `{k}` differently"
    assert set(map(str, a["metrics"])) == set(map(str, b["metrics"]))
    for k, v in a["metrics"].items():
        w = b["metrics"][k]
        assert np.array_equal(np.asarray(v, dtype=np.float64), np.asarray(w, dtype=np.float64)), f"{name}: metric {k} differs"
    assert len(a["metrics"]["imposter_loss"]) == sc["steps"] // 5 and len(a["metrics"]["total_time_steps"]) >= 1
    assert all(torch.equal(p, q) for p, q in zip(a["weights"], b["weights"])), f"{name}: trained weights differ"
    assert a["files"] == b["files"] and any(f.endswith("100%.pt") for f in a["files"])


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_the_reference_directly(cuda_lib, name):
    """N unmodified reference envs against the batched CUDA env on the same draws: flat states, float64 reward bit patterns,
    dones, truncations, per-episode metrics, post-reset states, role masks and sampled actions; Global features too."""
    import sus_net_b200 as S

    cfg = CASES[name]
    N, T, seed, base = 48, 130, 4242, 900
    ref = H.ReferenceBatch(cfg, N, seed, env_id_base=base)
    env = make_cuda_env(cfg, N, seed=seed, env_id_base=base)
    env._rewards = torch.zeros((N, env.n_agents), dtype=torch.float64, device=env.device)
    env._metrics_buf = torch.zeros((N, 8), dtype=torch.int64, device=env.device)
    feat = S.GlobalFeaturizer(env) if name in GLOBAL_CASES else None
    flat, _ = env.reset()
    assert np.array_equal(cpu(flat).astype(np.int64), ref.reset())
    episodes = 0
    for t in range(T):
        a = ref.sample_actions()
        assert np.array_equal(cpu(env.sample_actions()), a), f"{name}: sample_actions differs at step {t}"
        nf, r, d, tr, _ = env.step(torch.as_tensor(a.astype(np.int32)), featurizer=feat, check=True)
        o = ref.step(a)
        assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"{name}: state differs at step {t}"
        assert np.array_equal(reward_bits(cpu(r)), reward_bits(o["rewards"])), f"{name}: rewards differ at step {t}"
        assert np.array_equal(cpu(d), o["done"] != 0) and np.array_equal(cpu(tr), o["trunc"] != 0)
        assert np.array_equal(cpu(env._metrics_buf), o["metrics"]), f"{name}: metrics differ at step {t}"
        cur = ref.flat_states()
        assert np.array_equal(cpu(env.flat_states(torch.int64)), cur), f"{name}: post-reset state differs at step {t}"
        imp = np.stack([np.asarray(e.imposter_mask, dtype=np.uint8) for e in ref.envs])
        assert np.array_equal(cpu(env.imposter_mask_batch).astype(np.uint8), imp)
        episodes += int(((o["done"] | o["trunc"]) != 0).sum())
        if feat is not None and t % 16 == 0:  # the reference's own GlobalFeaturizer on the same states
            _, feat_mod = H.import_reference()
            rf = feat_mod.GlobalFeaturizer(ref.envs[0])
            rf.fit(torch.tensor(cur, dtype=torch.float32).unsqueeze(1))
            for (rsp, rns), (gsp, gns) in zip(rf.generate_featurized_states(), feat.generate_featurized_states()):
                assert np.array_equal(rsp.detach().numpy(), cpu(gsp)) and np.array_equal(rns.detach().numpy(), cpu(gns))
    assert int(cpu(env.episode_stats())[0]) == episodes


@pytest.mark.parametrize("T", [1, 3])
def test_reference_mode_populate_fills_the_ring_like_the_reference_populate(cuda_lib, T):
    """`sus_net_b200.ReplayBuffer.populate` on a reference-mode env (one env, one transition per step) against the reference's
    own `ReplayBuffer.populate` (replay_memory.py:96-143) on a twin env with the same seed: same ring contents, index and size,
    incl. the episode that is cut where the requested count is reached."""
    import sus_net_b200 as S

    _, replay_mod, _, _, _ = H.import_reference_training()
    kw = dict(n_crew=2, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0, end_of_game_reward=0, random_state=7)
    rings = []
    for cls in (replay_mod.ReplayBuffer, S.ReplayBuffer):
        env = S.ImposterTrainingGround(**kw)
        buf = cls(max_size=400, trajectory_size=T, state_size=env.flattened_state_size, n_imposters=env.n_imposters,
                  n_agents=env.n_agents)
        for t_ in (buf.states, buf.next_states, buf.rewards, buf.actions, buf.dones, buf.imposters):
            t_.zero_()
        buf.populate(env, 250)
        rings.append(buf)
    ref, own = rings
    assert (ref.idx, ref.size) == (own.idx, own.size) == (250, 250)
    for k in ("states", "actions", "rewards", "next_states", "dones", "imposters"):
        assert np.array_equal(cpu(getattr(ref, k)), cpu(getattr(own, k))), k
