"""GPU: the CUDA path (through the C ABI + the host mirror) against the golden fixtures minted from the reference
and against the oracle on seeded inputs.  Integer state, dones, truncations, metrics: bit-exact.  Rewards: exact
(float64 path compares bit patterns incl. the sign of zero; the float32 replay-layout path compares values).
Features: bit-exact float32 (tolerance stated by the north star is 1e-6; exact construction meets it with 0)."""
import numpy as np
import pytest
import torch

import oracle
from tests.cases import FLAT_COMPONENT_SETS, FLAT_COMPONENT_SETS_TAGGING, GLOBAL_CASES
from tests.util import CASES, case_of, flat_featurizer, golden_files, load, make_cuda_env, reward_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["ws", "tma", "direct", "staged"], autouse=True)
def store_path(request, monkeypatch):
    """Every GPU test runs on all output paths of the kernels: warp-specialised emitter + TMA bulk stores (default where
    planes are written), per-warp staging + TMA bulk stores, direct register stores (the fallback), and byte-staged
    rows expanded with coalesced stores (default for Flat encodes; other encodes take the per-warp TMA path there)."""
    monkeypatch.setenv("SUSNET_PATH", request.param)
    return request.param


def cpu(t):
    return t.detach().cpu().numpy()


def replay_cuda(g, cfg):
    T, N, A = g["actions"].shape
    injected = bool(g["injected"])
    env = make_cuda_env(cfg, N, seed=int(g["seed"]), env_id_base=int(g["env_id_base"]))
    env._rewards = torch.zeros((N, A), dtype=torch.float64, device=env.device)  # numpy-identical reward path
    env._metrics_buf = torch.zeros((N, 8), dtype=torch.int64, device=env.device)
    if injected:
        env.debug_inject_words(reset_words=g["reset_words0"])
    flat, _ = env.reset()
    assert np.array_equal(cpu(flat).astype(np.int64), g["reset_flat"])
    assert np.array_equal(cpu(env.imposter_mask_batch).astype(np.uint8), g["reset_imp"])
    for t in range(T):
        if injected:
            env.debug_inject_words(act_words=g["act_words"][t])
        a = cpu(env.sample_actions())
        if t % 7 != 3:
            assert np.array_equal(a, g["actions"][t]), f"sample_actions differs at step {t}"
        if injected:
            env.debug_inject_words(step_words=g["step_words"][t], reset_words=g["reset_words"][t])
        nf, r, d, tr, _ = env.step(torch.as_tensor(g["actions"][t].astype(np.int32)), check=True)
        assert np.array_equal(cpu(nf).astype(np.int64), g["next_flat"][t]), f"state differs at step {t}"
        assert np.array_equal(reward_bits(cpu(r)), reward_bits(g["rewards"][t])), f"rewards differ at step {t}"
        assert np.array_equal(cpu(d), g["done"][t] != 0) and np.array_equal(cpu(tr), g["trunc"][t] != 0)
        assert np.array_equal(cpu(env._metrics_buf), g["metrics"][t]), f"metrics differ at step {t}"
        assert np.array_equal(cpu(env.flat_states(torch.int64)), g["cur_flat"][t]), f"post-reset state differs at {t}"
        assert np.array_equal(cpu(env.imposter_mask_batch).astype(np.uint8), g["imp"][t])
    fin = (g["done"] | g["trunc"]) != 0
    st = cpu(env.episode_stats())
    assert st[0] == fin.sum() and st[9] == (g["trunc"] != 0).sum() and st[8] == g["metrics"][..., 0][fin].sum()


@pytest.mark.parametrize("path", golden_files("philox"), ids=case_of)
def test_cuda_matches_reference_philox_draws(cuda_lib, path):
    replay_cuda(load(path), CASES[case_of(path)])


@pytest.mark.parametrize("path", golden_files("words"), ids=case_of)
def test_cuda_matches_reference_injected_words(cuda_lib, path):
    replay_cuda(load(path), CASES[case_of(path)])


@pytest.mark.parametrize("path", golden_files("features"), ids=case_of)
def test_cuda_features_match_reference(cuda_lib, path):
    import sus_net_b200 as S

    g = load(path)
    name = case_of(path)
    env = make_cuda_env(CASES[name], 4, seed=1)
    flat = g["flat"].astype(np.int64)
    n = flat.shape[0]
    for dtype in (torch.float32, torch.float64, torch.int64):
        seq = torch.as_tensor(flat).to(dtype).reshape(n // 4, 4, -1)  # (B, T, S) with T = 4
        if "global_spatial" in g:
            f = S.GlobalFeaturizer(env)
            f.fit(seq)
            views = f.generate_featurized_states()
            assert len(views) == env.n_agents
            for k, (sp, ns) in enumerate(views):
                assert tuple(sp.shape) == (n // 4, 4, env.n_agents + 2, 9, 9) and sp.requires_grad
                assert np.array_equal(cpu(sp).reshape(n, -1, 9, 9), g["global_spatial"].astype(np.float32))
                assert np.array_equal(cpu(ns).reshape(n, -1), g["global_non_spatial"][k])
            f = S.PerspectiveFeaturizer(env)
            f.fit(seq)
            for k, (sp, ns) in enumerate(f.generate_featurized_states()):
                assert np.array_equal(cpu(sp).reshape(n, -1, 9, 9), g["perspective_spatial"][k].astype(np.float32))
                assert np.array_equal(cpu(ns).reshape(n, -1), g["perspective_non_spatial"][k])
        i = 0
        while f"flat{i}" in g:
            f = flat_featurizer(env, [str(c) for c in g[f"flat{i}_components"]])
            f.fit(seq)
            views = f.generate_featurized_states()
            for sp, ns in views:
                assert tuple(sp.shape) == (n // 4, 4, 1) and not cpu(sp).any()
                assert np.array_equal(cpu(ns).reshape(n, -1).view(np.int32), g[f"flat{i}"].view(np.int32))
            i += 1


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_oracle_at_scale(cuda_lib, name):
    """Seeded Philox draws, fused random policy, ragged N (not a multiple of the warp or CTA size)."""
    cfg = CASES[name]
    N, T, seed, base = 4099, 260, 777, 123456
    env = make_cuda_env(cfg, N, seed=seed, env_id_base=base)
    orc = oracle.OracleEnv(cfg, N, seed=seed, env_id_base=base)
    flat, _ = env.reset()
    assert np.array_equal(cpu(flat).astype(np.int64), orc.reset())
    acts = torch.zeros((N, env.n_agents), dtype=torch.int32, device=env.device)
    env._metrics_buf = torch.zeros((N, 8), dtype=torch.int64, device=env.device)
    for t in range(T):
        if t % 3 == 0:  # fused random policy
            io_actions = None
        else:
            io_actions = env.sample_actions().clone()
            assert np.array_equal(cpu(io_actions), orc.sample_actions())
        nf, r, d, tr, _ = env.step(io_actions)
        o = orc.step(None if io_actions is None else cpu(io_actions))
        assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"{name}: state differs at step {t}"
        assert np.array_equal(cpu(r), o["rewards"].astype(np.float32)), f"{name}: rewards differ at step {t}"
        assert np.array_equal(cpu(d), o["done"] != 0) and np.array_equal(cpu(tr), o["trunc"] != 0)
        assert np.array_equal(cpu(env._metrics_buf), o["metrics"])
    assert np.array_equal(cpu(env.flat_states(torch.int64)), orc.flat_states())
    assert np.array_equal(cpu(env.metrics_batch()), orc.metrics())
    assert np.array_equal(cpu(env.episode_stats()), orc.stats())
    env.check_actions()
    del acts


@pytest.mark.parametrize("name", GLOBAL_CASES + list(FLAT_COMPONENT_SETS) + list(FLAT_COMPONENT_SETS_TAGGING))
def test_fused_step_encode_matches_oracle(cuda_lib, name):
    """The fused K1+K2 launch must write the features of the state the next action is taken from (post auto-reset)."""
    import sus_net_b200 as S

    cfg = CASES[name]
    N, T = 1000, 120
    env = make_cuda_env(cfg, N, seed=5, env_id_base=7)
    orc = oracle.OracleEnv(cfg, N, seed=5, env_id_base=7)
    env.reset(); orc.reset()
    feats = []
    if name in GLOBAL_CASES:
        feats += [("global", S.GlobalFeaturizer(env)), ("perspective", S.PerspectiveFeaturizer(env))]
    for comps in FLAT_COMPONENT_SETS.get(name, []) + FLAT_COMPONENT_SETS_TAGGING.get(name, []):
        feats.append((comps, flat_featurizer(env, comps)))
    for t in range(T):
        kind, f = feats[t % len(feats)]
        env.step(None, featurizer=f)
        orc.step(None)
        cur = orc.flat_states()
        views = f.generate_featurized_states()
        if kind == "global":
            sp, ns = oracle.encode_global(cfg, cur)
            for k, (vsp, vns) in enumerate(views):
                assert np.array_equal(cpu(vsp)[:, 0], sp) and np.array_equal(cpu(vns)[:, 0], ns[k])
        elif kind == "perspective":
            sp, ns = oracle.encode_perspective(cfg, cur)
            for k, (vsp, vns) in enumerate(views):
                assert np.array_equal(cpu(vsp)[:, 0], sp[k]) and np.array_equal(cpu(vns)[:, 0], ns[k])
        else:
            want = oracle.encode_flat(cfg, kind, cur)
            assert np.array_equal(cpu(views[0][1])[:, 0].view(np.int32), want.view(np.int32))
        # the standalone K2 launch on the live state must agree with the fused one
        views2 = f.encode_env()
        assert all(torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]) for a, b in zip(views, views2))
    assert np.array_equal(cpu(env.flat_states(torch.int64)), orc.flat_states())


def test_reference_mode_api(cuda_lib):
    """num_envs == 1: reset/step return what the reference returns (types, shapes, dtypes, error behaviour)."""
    import sus_net_b200 as S

    env = S.FourRoomEnv(n_imposters=1, n_crew=4, n_jobs=5, random_state=3)
    assert env.flattened_state_size == 30 and env.n_agents == 5 and env.action_space.n == 8
    assert env.n_imposter_actions == 7 and env.n_crew_actions == 6
    assert env.grid.shape == (9, 9) and len(env.valid_positions) == 68 and env.valid_positions[4].tolist() == [0, 5]
    state, info = env.reset()
    pos, alive, jpos, jdone = state
    assert pos.shape == (5, 2) and pos.dtype == np.int64 and alive.dtype == bool and alive.all()
    assert jpos.shape == (5, 2) and jdone.dtype == bool and not jdone.any()
    assert set(k.value if hasattr(k, "value") else k for k in info) >= {"total_time_steps", "crew_won", "imposter_won"}
    assert env.imposter_mask.sum() == 1 and env.imposter_idxs.tolist() == np.where(env.imposter_mask)[0].tolist()
    flat = env.flatten_state(state)
    assert flat.shape == (30,) and flat.dtype == np.int64
    un = env.unflatten_state(flat)
    assert all(np.array_equal(a, b) for a, b in zip(un, state))
    a = env.sample_actions()
    assert a.shape == (5,) and all(a[i] < (7 if env.imposter_mask[i] else 6) for i in range(5))
    nstate, r, done, trunc, info = env.step(a)
    assert r.dtype == np.float64 and r.shape == (5,) and isinstance(done, bool) and isinstance(trunc, bool)
    assert info[S.SusMetrics.TOTAL_TIME_STEPS] == 1
    with pytest.raises(AssertionError):
        env.step([0, 0, 0])  # wrong length (base.py:357-359)
    with pytest.raises(AssertionError):
        env.step([8, 0, 0, 0, 0])  # >= action_space.n (base.py:360-362)
    crew = int(np.where(~env.imposter_mask)[0][0])
    bad = [0] * 5
    bad[crew] = 6  # crew list has 6 entries: IndexError in the reference (base.py:381)
    with pytest.raises(IndexError):
        env.step(bad)
    with pytest.raises(AssertionError):
        S.FourRoomEnv(n_imposters=2, n_crew=2, n_jobs=1)  # base.py:247-249
    with pytest.raises(AssertionError):
        S.ImposterTrainingGround(n_crew=0, n_jobs=0, time_step_reward=0, kill_reward=0, sabotage_reward=0,
                                 end_of_game_reward=0)
    # tagging env: 7-field state tuple, {} reset info (tagging.py:94-101)
    tenv = S.FourRoomEnvWithTagging(1, 2, 5)
    st, info = tenv.reset()
    assert len(st) == 7 and st[6] == 50 and info == {} and tenv.flattened_state_size == 31
    assert tenv.action_space.n == 11 and tenv.n_imposter_actions == 9 and tenv.n_crew_actions == 8
    st, r, d, tr, info = tenv.step(tenv.sample_actions())
    assert st[6] == 49 and len(info) == 13


def test_reference_mode_matches_oracle_episode(cuda_lib):
    """Drive the single-env API like train.py does (manual reset on done/trunc) against the oracle."""
    import sus_net_b200 as S

    cfg = CASES["cfg4_base_1v4"]
    env = S.FourRoomEnv(1, 4, 5, random_state=11)
    orc = oracle.OracleEnv(cfg, 1, seed=11, auto_reset=False)
    state, _ = env.reset()
    assert np.array_equal(env.flatten_state(state), orc.reset()[0])
    for t in range(300):
        a = env.sample_actions()
        assert np.array_equal(a, orc.sample_actions()[0])
        state, r, d, tr, info = env.step(a)
        o = orc.step(a[None])
        assert np.array_equal(env.flatten_state(state), o["next_flat"][0])
        assert np.array_equal(reward_bits(r), reward_bits(o["rewards"][0])) and d == bool(o["done"][0])
        assert [info[k] for k in S.METRIC_ORDER] == o["metrics"][0].tolist()
        if d or tr:
            state, _ = env.reset()
            assert np.array_equal(env.flatten_state(state), orc.reset()[0])


def test_state_dict_roundtrip_and_sharding_invariance(cuda_lib):
    """Results depend on (seed, global env id, tick) only: two half-size shards reproduce one full-size env, and a
    restored checkpoint continues identically."""
    cfg = CASES["cfg3_tagging_1v2"]
    full = make_cuda_env(cfg, 512, seed=9, env_id_base=0)
    lo = make_cuda_env(cfg, 256, seed=9, env_id_base=0)
    hi = make_cuda_env(cfg, 256, seed=9, env_id_base=256)
    for e in (full, lo, hi):
        e.reset()
    for t in range(80):
        nf, r, d, tr, _ = full.step(None)
        a = lo.step(None)
        b = hi.step(None)
        assert torch.equal(nf, torch.cat([a[0], b[0]])) and torch.equal(r, torch.cat([a[1], b[1]]))
        assert torch.equal(d, torch.cat([a[2], b[2]]))
        if t == 40:
            sd = full.state_dict()
    assert torch.equal(full.episode_stats(), lo.episode_stats() + hi.episode_stats())
    again = make_cuda_env(cfg, 512, seed=9, env_id_base=0)
    again.load_state_dict(sd)
    ref = make_cuda_env(cfg, 512, seed=9, env_id_base=0)
    ref.reset()
    for t in range(41):
        ref.step(None)
    for t in range(20):
        assert torch.equal(again.step(None)[0], ref.step(None)[0])


def test_empty_batch_and_survey_anchors(cuda_lib):
    import sus_net_b200 as S

    env = make_cuda_env(CASES["cfg4_base_1v4"], 0, seed=1)
    env.reset()
    nf, r, d, tr, _ = env.step(None)
    assert nf.shape == (0, 30) and r.shape == (0, 5)
    # SURVEY.md 8(c) anchors through the CUDA path
    e = S.FourRoomEnv(1, 2, 0, shuffle_imposter_index=False)
    e.reset()
    _, r, d, _, _ = e.step([0, 0, 0])
    assert r.tolist() == [-10.0, 10.0, 10.0] and d
    e = S.ImposterTrainingGround(n_crew=2, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                 end_of_game_reward=0)
    e.import_flat(np.array([[3, 3, 3, 3, 3, 3, 1, 1, 1]]), np.array([[1, 0, 0]]))
    st, r, d, _, _ = e.step([5, 0, 0])
    assert r.tolist() == [3.0, 0.0, 0.0] and st[1].sum() == 2 and not d
    e = S.ImposterTrainingGround(n_crew=1, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                 end_of_game_reward=0)
    e.reset()
    for k in range(1000):
        _, _, d, tr, _ = e.step([0, 0])
        assert tr == (k == 999)


@pytest.mark.parametrize("k", range(16))
def test_random_constructor_arguments_match_oracle(cuda_lib, k):
    """Random variants / sizes / NON-INTEGER reward constants (the oracle is pinned against the reference on the same
    generator: profiles/r01_oracle_pin_random_configs.log).  Rewards are compared as float64 bit patterns."""
    import sus_net_b200 as S
    from tests.cases import random_case

    cfg = random_case(np.random.default_rng(100000 + k))
    N, T = 777, 90
    env = make_cuda_env(cfg, N, seed=k, env_id_base=50 * k)
    if cfg["variant"] == "training_ground":
        assert cfg["max_time_steps"] == 1000
    env._rewards = torch.zeros((N, env.n_agents), dtype=torch.float64, device=env.device)
    env._metrics_buf = torch.zeros((N, 8), dtype=torch.int64, device=env.device)
    orc = oracle.OracleEnv(cfg, N, seed=k, env_id_base=50 * k)
    env.reset(); orc.reset()
    feat = None
    if cfg["n_jobs"] > 0:
        feat = S.PerspectiveFeaturizer(env) if k % 2 else S.GlobalFeaturizer(env)
    for t in range(T):
        nf, r, d, tr, _ = env.step(None, featurizer=feat)
        o = orc.step(None)
        assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"state differs at step {t}: {cfg}"
        assert np.array_equal(reward_bits(cpu(r)), reward_bits(o["rewards"])), f"rewards differ at step {t}: {cfg}"
        assert np.array_equal(cpu(d), o["done"] != 0) and np.array_equal(cpu(tr), o["trunc"] != 0)
        assert np.array_equal(cpu(env._metrics_buf), o["metrics"])
    cur = orc.flat_states()
    assert np.array_equal(cpu(env.flat_states(torch.int64)), cur)
    assert np.array_equal(cpu(env.episode_stats()), orc.stats())
    if feat is not None:
        views = feat.generate_featurized_states()
        if k % 2:
            sp, ns = oracle.encode_perspective(cfg, cur)
            assert all(np.array_equal(cpu(v[0])[:, 0], sp[i]) and np.array_equal(cpu(v[1])[:, 0], ns[i]) for i, v in enumerate(views))
        else:
            sp, ns = oracle.encode_global(cfg, cur)
            assert all(np.array_equal(cpu(v[0])[:, 0], sp) and np.array_equal(cpu(v[1])[:, 0], ns[i]) for i, v in enumerate(views))


@pytest.mark.parametrize("k", range(10))
def test_random_flat_component_sets_match_oracle(cuda_lib, k):
    """Random env configurations (any variant, up to 8 agents / 8 jobs) x random orderings of flat components, with and
    without the float-valued scent component (i.e. byte-staged and float-row kernels): fused step + encode, the
    standalone encode of the live state and fit() from float32 / float64 / int64 rows all equal the oracle's rows."""
    from tests.cases import random_case

    rng = np.random.default_rng(424200 + k)
    cfg = random_case(rng)
    names = ["onehot_pos", "coords", "alive_crew", "dist_to_imposter", "walls", "rooms", "state_alive"]
    if cfg["n_imposters"] == 1:  # the per-crew components index crew by agent_idx - 1 (component.py:440,467)
        names += ["closest_crew", "l1_crew"]
    if cfg["n_jobs"] > 0:
        names.append("state_job_status")
    if cfg["variant"] == "tagging":
        names += ["state_used_tags", "state_tag_counts"]
    if k % 3 == 0:
        names.append("scent")
    comps = [names[i] for i in rng.permutation(len(names))[:int(rng.integers(1, len(names) + 1))]]
    if k % 4 == 1:
        comps.append("onehot_pos")  # a second one-hot segment in the same row
    N, T = 1003, 40
    env = make_cuda_env(cfg, N, seed=k, env_id_base=11 * k)
    orc = oracle.OracleEnv(cfg, N, seed=k, env_id_base=11 * k)
    env.reset(); orc.reset()
    feat = flat_featurizer(env, comps)
    for t in range(T):
        env.step(None, featurizer=feat)
        orc.step(None)
        if t % 13 == 0 or t == T - 1:
            want = oracle.encode_flat(cfg, comps, orc.flat_states())
            got = cpu(feat.generate_featurized_states()[0][1])[:, 0]
            assert np.array_equal(got.view(np.int32), want.view(np.int32)), f"fused rows differ at step {t}: {cfg} {comps}"
    cur = orc.flat_states()
    want = oracle.encode_flat(cfg, comps, cur)
    assert np.array_equal(cpu(feat.encode_env()[0][1])[:, 0].view(np.int32), want.view(np.int32))
    for dtype in (torch.float32, torch.float64, torch.int64):
        feat.fit(torch.as_tensor(cur).to(dtype).reshape(N, 1, -1))
        assert np.array_equal(cpu(feat.generate_featurized_states()[0][1])[:, 0].view(np.int32), want.view(np.int32)), str(dtype)
    assert np.array_equal(cpu(env.flat_states(torch.int64)), cur)


@pytest.mark.parametrize("name", ["cfg2_itg_1v1_wall", "cfg3_tagging_1v2", "cfg4_base_1v4", "base_2v3_j3"])
def test_rollout_kernel_equals_repeated_steps(cuda_lib, name):
    """rollout(T) == T x step(None): final states, per-episode counters, episode statistics, reward sums."""
    cfg = CASES[name]
    N, T = 1500, 137
    a = make_cuda_env(cfg, N, seed=3, env_id_base=9)
    b = make_cuda_env(cfg, N, seed=3, env_id_base=9)
    orc = oracle.OracleEnv(cfg, N, seed=3, env_id_base=9)
    a.reset(); b.reset(); orc.reset()
    b._rewards = torch.zeros((N, b.n_agents), dtype=torch.float64, device=b.device)
    sums = a.rollout(T, reward_sums=True)
    want = torch.zeros_like(sums)
    for _ in range(T):
        want += b.step(None)[1]
        orc.step(None)
    assert torch.equal(a.flat_states(torch.int64), b.flat_states(torch.int64))
    assert np.array_equal(cpu(a.flat_states(torch.int64)), orc.flat_states())
    assert torch.equal(a.metrics_batch(), b.metrics_batch()) and torch.equal(a.episode_stats(), b.episode_stats())
    assert np.array_equal(cpu(a.episode_stats()), orc.stats())
    assert torch.equal(sums, want)  # same float64 additions in the same order
    a.rollout(5); [b.step(None) for _ in range(5)]  # ticks stay aligned afterwards
    assert torch.equal(a.flat_states(torch.int64), b.flat_states(torch.int64))


def test_unaligned_output_pointers_through_the_c_abi(cuda_lib):
    """Feature tensors that are only 4-byte aligned (sliced views): the TMA paths must fall back to ordinary stores for
    every block whose address or size is not a 16-byte multiple, with identical results."""
    import ctypes as C

    import sus_net_b200 as S
    from sus_net_b200 import _lib as L

    cfg = CASES["cfg4_base_1v4"]
    for N in (1003, 64):
        env = make_cuda_env(cfg, N, seed=2)
        env.reset()
        for _ in range(3):
            env.step(None)
        cur = cpu(env.flat_states(torch.int64))
        for kind, enc in ((L.ENCODE_GLOBAL, oracle.encode_global), (L.ENCODE_PERSPECTIVE, oracle.encode_perspective)):
            sh = L.SusEncodeShape()
            spec = L.SusEncodeSpec(kind=kind)
            L.check(env.lib.sus_encode_shape(C.byref(env._cfg), C.byref(spec), C.byref(sh)))
            sp_n = sh.spatial_views * N * sh.spatial_floats
            ns_n = sh.non_spatial_views * N * sh.non_spatial_floats
            for off in (1, 2, 3):  # floats: 4, 8, 12 bytes past a 16-byte boundary
                sp_raw = torch.full((sp_n + 8,), -7.0, device=env.device)
                ns_raw = torch.full((ns_n + 8,), -7.0, device=env.device)
                sp, ns = sp_raw[off:off + sp_n], ns_raw[off:off + ns_n]
                L.check(env.lib.sus_env_encode(env._h, C.byref(spec), C.c_void_p(sp.data_ptr()), C.c_void_p(ns.data_ptr()),
                                               env._stream()))
                want_sp, want_ns = enc(cfg, cur)
                assert np.array_equal(cpu(sp).reshape(want_sp.shape), want_sp)
                assert np.array_equal(cpu(ns).reshape(want_ns.shape), want_ns)
                # nothing outside the tensors was touched
                assert (cpu(sp_raw[:off]) == -7).all() and (cpu(sp_raw[off + sp_n:]) == -7).all()
                assert (cpu(ns_raw[:off]) == -7).all() and (cpu(ns_raw[off + ns_n:]) == -7).all()


@pytest.mark.parametrize("name", ["cfg4_base_1v4", "cfg3_tagging_1v2", "cfg2_itg_1v1_wall"])
def test_cuda_graph_replay_matches_oracle(cuda_lib, name, store_path):
    """Device-resident ticks: a captured graph of [sample_actions, step (+ fused encode), 3-step rollout] replayed K times
    walks the same trajectory as the oracle doing the same calls one by one; afterwards the env keeps stepping in
    lock-step with host calls, and the checkpointed ticks are the advanced ones."""
    import sus_net_b200 as S

    cfg = CASES[name]
    N, K = 777, 25
    env = make_cuda_env(cfg, N, seed=21)
    orc = oracle.OracleEnv(cfg, N, seed=21)
    feat = S.GlobalFeaturizer(env) if name in GLOBAL_CASES else None
    env.reset(); orc.reset()
    env.device_ticks(True)
    env.step(env.sample_actions(), featurizer=feat)  # eager warm-up call (loads the kernels' attributes outside the capture)
    orc.step(orc.sample_actions())
    side = torch.cuda.Stream(env.device)
    side.wait_stream(torch.cuda.current_stream(env.device))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        env.rollout(3)  # warm-up of the rollout kernel as well
        orc_r = [orc.step(None) for _ in range(3)]
        with torch.cuda.graph(graph, stream=side):
            acts = env.sample_actions()
            nf, r, d, tr, _ = env.step(acts, featurizer=feat)
            env.rollout(3)
    torch.cuda.current_stream(env.device).wait_stream(side)
    del orc_r
    for k in range(K):
        graph.replay()
        o = orc.step(orc.sample_actions())
        want_cur = orc.flat_states()
        for _ in range(3):
            orc.step(None)
        if k % 6 == 0 or k == K - 1:
            assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"{name}: replay {k}"
            assert np.array_equal(cpu(r), o["rewards"].astype(np.float32)) and np.array_equal(cpu(d), o["done"] != 0)
            if feat is not None:
                sp, ns = oracle.encode_global(cfg, want_cur)
                views = feat.generate_featurized_states()
                assert np.array_equal(cpu(views[0][0])[:, 0], sp) and np.array_equal(cpu(views[2][1])[:, 0], ns[2])
            assert np.array_equal(cpu(env.flat_states(torch.int64)), orc.flat_states()), f"{name}: replay {k}"
    assert np.array_equal(cpu(env.episode_stats()), orc.stats())
    sd = env.state_dict()
    assert sd["ticks"][0] == 1 + 3 + 4 * K and sd["ticks"][2] == 1 + K  # step ticks / act epochs advanced on the device
    env.device_ticks(False)  # back to host ticks: they continue from the device's values
    nf2, *_ = env.step(None)
    assert np.array_equal(cpu(nf2).astype(np.int64), orc.step(None)["next_flat"])


def test_viewer_surface_of_reference_mode(cuda_lib):
    """The statements AmongUsVisualizer / run_game execute on the env (visualize.py:40-43,92,142,153-160,214-261,555-560),
    without pygame: a reference-mode env exposes the same attributes with the shapes the viewer indexes."""
    import sus_net_b200 as S
    from sus_net_b200.metrics import SusMetrics

    env = S.BatchedFourRoomEnvWithTagging(1, 3, 2, seed=3)
    env.reset()
    for _ in range(3):
        env.step(env.sample_actions())
    assert env.n_rows == 9 and env.n_cols == 9 and env.grid.shape == (9, 9)
    assert env.__dict__.get("tag_counts") is not None  # the voting panel is drawn
    vote_counts = env.tag_counts.flatten()
    vote_counts[env.alive_agents == 0] = -1
    voted = env.used_tag_actions.flatten() + 0
    voted[env.alive_agents == 0] = -1
    assert len(vote_counts) == 4 and len(voted) == 4
    assert 0 < env.tag_reset_interval - env.tag_reset_timer <= env.tag_reset_interval
    for i, (pos, alive) in enumerate(zip(env.agent_positions, env.alive_agents)):
        x, y = pos
        assert 0 <= x < 9 and 0 <= y < 9 and env.grid[x, y] and alive in (0, 1) and env.imposter_mask[i] in (True, False)
    assert len(env.job_positions) == env.n_jobs == len(env.completed_jobs)
    assert env.metrics.metrics[SusMetrics.IMPOSTER_WON] in (0, 1)
    assert env.compute_action(0, 1) == str(S.Action.UP) and env.compute_action(1, 8) == "Vote Player 0"
    assert any(i in env.imposter_idxs for i in range(env.n_agents))
    assert env.flatten_state(env.reset()[0]).shape == (env.flattened_state_size,)
    base = S.BatchedFourRoomEnv(1, 2, 1)
    assert base.__dict__.get("tag_counts") is None  # no voting panel for the plain env


def test_compressible_feature_memory(cuda_lib, monkeypatch):
    """sus_alloc_compressible: the featurizers' big output buffers live in L2-compressible memory (where the device has
    it), results are identical to the ones written into torch-allocated memory, small buffers stay with torch, blocks are
    freed with their last tensor and a foreign / repeated free is refused."""
    import ctypes as C
    import gc

    import sus_net_b200 as S
    from sus_net_b200 import _lib as L
    from sus_net_b200 import memory as M

    lib = L.lib()
    ptr, got = C.c_void_p(), C.c_uint64()
    rc = lib.sus_alloc_compressible(0, 5 << 20, C.byref(ptr), C.byref(got))
    if rc == L.SUS_ERR_UNSUPPORTED:
        pytest.skip("device without generic compression")
    L.check(rc)
    assert ptr.value and got.value >= 5 << 20 and got.value % (2 << 20) == 0
    L.check(lib.sus_free_compressible(ptr))
    with pytest.raises(AssertionError):
        L.check(lib.sus_free_compressible(ptr))  # not a live block any more
    with pytest.raises(AssertionError):
        L.check(lib.sus_alloc_compressible(0, 0, C.byref(ptr), None))

    cfg = CASES["cfg4_base_1v4"]
    N = 8192  # planes: 8192 x 567 x 4 B = 18.6 MB (compressible); non-spatial views: 2.5 MB (torch)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("SUSNET_COMPRESSIBLE", flag)
        env = make_cuda_env(cfg, N, seed=4)
        feat = S.GlobalFeaturizer(env)
        env.reset()
        for _ in range(5):
            env.step(None, featurizer=feat)
        assert M.is_compressible(feat._sp_buf) == (flag == "1") and not M.is_compressible(feat._ns_buf)
        views = feat.generate_featurized_states()
        outs.append((cpu(views[0][0]).copy(), cpu(views[3][1]).copy(), cpu(env.flat_states(torch.int64))))
        t = feat._sp_buf
        del feat, views, env
        gc.collect()
        assert float(t.sum()) > 0  # the tensor alone keeps its block mapped
        del t
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    small = M.empty_f32((16, 16), torch.device("cuda", 0))
    assert not M.is_compressible(small)


def test_flat_rows_unaligned_and_ragged_through_the_c_abi(cuda_lib):
    """Flat rows with an odd float count (F = 39) on ragged env counts, written to 4-byte-aligned tensors by the standalone
    encode AND by the fused step (rewards / replay row / features all unaligned): the staged paths must fall back to scalar
    stores where a block is not 16-byte aligned or a multiple of 16 bytes, and touch nothing outside the tensors."""
    import ctypes as C

    from sus_net_b200 import _lib as L

    cfg = CASES["cfg4alt_itg_1v4"]
    comps = ["alive_crew", "l1_crew", "dist_to_imposter", "walls", "closest_crew", "coords"]  # 4 + 4 + 8 + 9 + 4 + 10
    ids = (C.c_int32 * L.MAX_FLAT_COMPONENTS)(*[oracle.FLAT_COMPONENTS[c] for c in comps])
    for N in (1003, 37, 5):
        env = make_cuda_env(cfg, N, seed=3)
        orc = oracle.OracleEnv(cfg, N, seed=3)
        env.reset(); orc.reset()
        spec = L.SusEncodeSpec(kind=L.ENCODE_FLAT, n_components=len(comps), components=ids)
        sh = L.SusEncodeShape()
        L.check(env.lib.sus_encode_shape(C.byref(env._cfg), C.byref(spec), C.byref(sh)))
        F, A, S = sh.non_spatial_floats, env.n_agents, env.flattened_state_size
        assert F == 39
        for off in (0, 1, 2, 3):
            ns_raw = torch.full((N * F + 8,), -7.0, device=env.device)
            ns = ns_raw[off:off + N * F]
            L.check(env.lib.sus_env_encode(env._h, C.byref(spec), None, C.c_void_p(ns.data_ptr()), env._stream()))
            assert np.array_equal(cpu(ns).reshape(N, F), oracle.encode_flat(cfg, comps, orc.flat_states()))
            assert (cpu(ns_raw[:off]) == -7).all() and (cpu(ns_raw[off + N * F:]) == -7).all()
            # fused step: random policy in the kernel, every dense output 4-byte aligned only
            rew_raw = torch.full((N * A + 8,), -7.0, device=env.device)
            nf_raw = torch.full((N * S + 8,), -7.0, device=env.device)
            ns_raw.fill_(-7.0)
            rew, nf = rew_raw[off:off + N * A], nf_raw[off:off + N * S]
            done, trunc = torch.zeros(N, dtype=torch.uint8, device=env.device), torch.zeros(N, dtype=torch.uint8, device=env.device)
            io = L.SusStepIO(actions=None, actions_dtype=L.I32, rewards=rew.data_ptr(), rewards_dtype=L.F32,
                             done=done.data_ptr(), truncated=trunc.data_ptr(), next_flat=nf.data_ptr(),
                             encode=C.pointer(spec), non_spatial=ns.data_ptr())
            L.check(env.lib.sus_env_step(env._h, C.byref(io), env._stream()))
            o = orc.step(None)
            assert np.array_equal(cpu(rew).reshape(N, A), o["rewards"].astype(np.float32))
            assert np.array_equal(cpu(nf).reshape(N, S).astype(np.int64), o["next_flat"])
            assert np.array_equal(cpu(done), o["done"]) and np.array_equal(cpu(trunc), o["trunc"])
            assert np.array_equal(cpu(ns).reshape(N, F), oracle.encode_flat(cfg, comps, orc.flat_states()))
            for raw, n in ((rew_raw, N * A), (nf_raw, N * S), (ns_raw, N * F)):
                assert (cpu(raw[:off]) == -7).all() and (cpu(raw[off + n:]) == -7).all()


@pytest.mark.parametrize("name", ["cfg4_base_1v4", "tagging_2v5_short", "base_fixed_order_tsr"])
def test_tracked_returns_match_oracle(cuda_lib, name):
    """train()'s running returns G = r + gamma * G per agent (train.py:386) kept on the device: per-episode means summed
    over finished episodes.  The per-env recurrences are bit-exact; the cross-env sums are order-dependent float adds,
    compared to 1e-9 relative."""
    import sus_net_b200 as S

    cfg = CASES[name]
    N, T, gamma = 3001, 150, 0.9
    env = make_cuda_env(cfg, N, seed=8)
    orc = oracle.OracleEnv(cfg, N, seed=8)
    env.reset(); orc.reset()
    env.track_returns(gamma); orc.track_returns(gamma)
    feat = S.GlobalFeaturizer(env)
    for t in range(T):
        env.step(None, featurizer=feat if t % 2 else None)
        orc.step(None)
    got, want = cpu(env.return_sums()), orc.return_sums()
    assert int(env.episode_stats()[0]) == int(orc.stats()[0]) > 0
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (got, want)
    with pytest.raises(NotImplementedError):
        env.rollout(3)  # the rollout kernel does not maintain tracked returns


def test_component_extract_features_and_cpu_output(cuda_lib):
    """Single-state `extract_features` of the component classes and `output_device="cpu"` (what the reference's unmodified
    CPU models need) against the oracle."""
    import sus_net_b200 as S

    cfg = CASES["cfg4_base_1v4"]
    env = S.FourRoomEnv(1, 4, 5, random_state=5)
    state, _ = env.reset()
    for _ in range(7):
        state, *_ = env.step(env.sample_actions())
    flat = env.flatten_state(state)[None]
    sp, ns = oracle.encode_global(cfg, flat)
    assert np.array_equal(cpu(S.AgentPositionsFeaturizer(env).extract_features(state)), sp[0, :5])
    assert np.array_equal(cpu(S.JobFeaturizer(env).extract_features(state)), sp[0, 5:])
    assert S.AgentPositionsFeaturizer(env).shape.tolist() == [5, 9, 9] and S.JobFeaturizer(env).shape.tolist() == [2, 9, 9]
    one = S.OneHotAgentPositionFeaturizer(env).extract_features(state)
    assert np.array_equal(cpu(one), oracle.encode_flat(cfg, ["onehot_pos"], flat)[0])
    f = S.GlobalFeaturizer(env, output_device="cpu")
    f.fit(torch.tensor(flat, dtype=torch.float64).unsqueeze(0))  # what train.py:346-348 passes
    views = f.generate_featurized_states()
    assert all(v[0].device.type == "cpu" and v[1].device.type == "cpu" and v[0].requires_grad for v in views)
    assert np.array_equal(views[2][0].detach().numpy()[0, 0], sp[0]) and np.array_equal(views[2][1].detach().numpy()[0, 0], ns[2, 0])


def _edge_names():
    from tests.cases import EDGE_CASES

    return list(EDGE_CASES)


@pytest.mark.parametrize("name", _edge_names())
def test_edge_case_configs_match_oracle(cuda_lib, name):
    """Extremes of the constructor-argument space (oracle pinned vs the reference: profiles/r01_oracle_pin_edge_cases.log):
    1-step episodes, 1-step vote windows, 8 agents x 8 jobs, all-zero rewards (-0.0 in the tagging env)."""
    import sus_net_b200 as S
    from tests.cases import EDGE_CASES

    cfg = EDGE_CASES[name]
    N, T = 515, 120
    env = make_cuda_env(cfg, N, seed=12)
    env._rewards = torch.zeros((N, env.n_agents), dtype=torch.float64, device=env.device)
    env._metrics_buf = torch.zeros((N, 8), dtype=torch.int64, device=env.device)
    orc = oracle.OracleEnv(cfg, N, seed=12)
    assert np.array_equal(cpu(env.reset()[0]).astype(np.int64), orc.reset())
    feat = S.PerspectiveFeaturizer(env) if cfg["n_jobs"] > 0 else None
    for t in range(T):
        acts = env.sample_actions().clone() if t % 2 else None
        if acts is not None:
            assert np.array_equal(cpu(acts), orc.sample_actions())
        nf, r, d, tr, _ = env.step(acts, featurizer=feat)
        o = orc.step(None if acts is None else cpu(acts))
        assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"state differs at step {t}"
        assert np.array_equal(reward_bits(cpu(r)), reward_bits(o["rewards"])), f"reward bits differ at step {t}"
        assert np.array_equal(cpu(d), o["done"] != 0) and np.array_equal(cpu(tr), o["trunc"] != 0)
        assert np.array_equal(cpu(env._metrics_buf), o["metrics"])
    cur = orc.flat_states()
    assert np.array_equal(cpu(env.flat_states(torch.int64)), cur) and np.array_equal(cpu(env.episode_stats()), orc.stats())
    if feat is not None:
        sp, ns = oracle.encode_perspective(cfg, cur)
        views = feat.generate_featurized_states()
        assert all(np.array_equal(cpu(v[0])[:, 0], sp[i]) and np.array_equal(cpu(v[1])[:, 0], ns[i]) for i, v in enumerate(views))


def test_cfg2_4096_envs_10k_trajectories_bit_exact(cuda_lib):
    """BASELINE.json configs[1]: ImposterTrainingGround 1v1 on the walled grid, 4096 batched envs, every trajectory run to
    done / truncation, > 10 000 finished trajectories, every step compared (state, rewards, done, truncated).  The same
    configuration is pinned oracle-vs-reference on > 10 000 trajectories in profiles/r01_oracle_pin_cfg2_10k_trajectories.log."""
    cfg = CASES["cfg2_itg_1v1_wall"]
    N, T = 4096, 1500
    env = make_cuda_env(cfg, N, seed=2, env_id_base=0)
    orc = oracle.OracleEnv(cfg, N, seed=2, env_id_base=0)
    assert np.array_equal(cpu(env.reset()[0]).astype(np.int64), orc.reset())
    out = None
    for t in range(T):
        nf, r, d, tr, _ = env.step(None)
        out = orc.step(None, out=out)
        assert np.array_equal(cpu(nf).astype(np.int64), out["next_flat"]), f"state differs at step {t}"
        assert np.array_equal(cpu(r), out["rewards"].astype(np.float32)), f"rewards differ at step {t}"
        assert np.array_equal(cpu(d), out["done"] != 0) and np.array_equal(cpu(tr), out["trunc"] != 0)
    stats = cpu(env.episode_stats())
    assert np.array_equal(stats, orc.stats()) and stats[0] > 10_000, stats


def test_env_on_a_non_current_device_and_side_stream(cuda_lib):
    """The handle remembers its device (every C-ABI call switches to it and back) and all launches go to the caller's
    current stream: an env on cuda:1 driven while cuda:0 is current, and an env driven on a side stream, reproduce the
    default-stream env on cuda:0."""
    import sus_net_b200 as S

    cfg = CASES["cfg4_base_1v4"]
    N, T = 4096, 25
    ref = make_cuda_env(cfg, N, seed=6, device="cuda:0")
    ref.reset()
    fr = S.GlobalFeaturizer(ref)
    for _ in range(T):
        ref.step(None, featurizer=fr)
    torch.cuda.synchronize()
    want_state, want_sp = ref.flat_states(torch.int64).cpu(), fr.spatial.detach().cpu()
    # side stream on the same device
    side = torch.cuda.Stream("cuda:0")
    with torch.cuda.stream(side):
        e2 = make_cuda_env(cfg, N, seed=6, device="cuda:0")
        e2.reset()
        f2 = S.GlobalFeaturizer(e2)
        for _ in range(T):
            e2.step(None, featurizer=f2)
        got_state, got_sp = e2.flat_states(torch.int64), f2.spatial.detach()
    side.synchronize()
    assert torch.equal(got_state.cpu(), want_state) and torch.equal(got_sp.cpu(), want_sp)
    if torch.cuda.device_count() >= 2:
        assert torch.cuda.current_device() == 0
        e3 = make_cuda_env(cfg, N, seed=6, device="cuda:1")
        e3.reset()
        f3 = S.GlobalFeaturizer(e3)
        for _ in range(T):
            e3.step(None, featurizer=f3)
        assert torch.cuda.current_device() == 0
        assert torch.equal(e3.flat_states(torch.int64).cpu(), want_state) and torch.equal(f3.spatial.detach().cpu(), want_sp)


def test_action_dtypes_and_invalid_indices(cuda_lib):
    """uint8 / int32 / int64 action tensors give identical steps; an index outside an agent's role list (or negative, or
    >= 256) leaves that env untouched and is reported by check_actions() as the reference's IndexError."""
    cfg = CASES["cfg3_tagging_1v2"]
    N = 2000
    envs = [make_cuda_env(cfg, N, seed=4) for _ in range(3)]
    for e in envs:
        e.reset()
    for t in range(30):
        a = envs[0].sample_actions().clone()
        outs = [e.step(a.to(dt)) for e, dt in zip(envs, (torch.int32, torch.int64, torch.uint8))]
        assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2]) for o in outs[1:])
        for e in envs[1:]:
            e.sample_actions()  # keep the act epochs aligned
    env = envs[0]
    before = env.flat_states(torch.int64).clone()
    a = env.sample_actions().clone().to(torch.int64)
    bad_rows = torch.tensor([3, 77, 1999], device=env.device)
    a[bad_rows[0], 1] = 9          # past every role list of a 3-agent tagging env (8 or 9 entries)
    a[bad_rows[1], 0] = -1         # negative
    a[bad_rows[2], 2] = 300        # >= 256
    env.step(a)
    after = env.flat_states(torch.int64)
    assert torch.equal(after[bad_rows], before[bad_rows])  # rejected envs did not move
    changed = (after != before).any(dim=1)
    assert changed.sum() > N // 2
    with pytest.raises(IndexError):
        env.check_actions()
    env.check_actions()  # the counter was cleared
