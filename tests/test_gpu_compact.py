"""GPU: the compact host protocol (include/susnet_b200.h `SusCompactLayout`, sus_net_b200/compact.py).  Bit-packed
actions in, reward codes + done / truncated bits out; the host decodes the codes through the float64 table of
`sus_reward_lut`.  The bar: decoded rewards BIT-EQUAL (float64 bit patterns incl. the sign of zero) to the reference's
rewards on every golden case, and to the oracle's on random constructor arguments with non-integer reward constants."""
import numpy as np
import pytest
import torch

import oracle
from tests.cases import EDGE_CASES, FLAT_COMPONENT_SETS, GLOBAL_CASES, random_case
from tests.util import CASES, case_of, flat_featurizer, golden_files, load, make_cuda_env, reward_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["ws", "tma", "direct", "staged"], autouse=True)
def store_path(request, monkeypatch):
    monkeypatch.setenv("SUSNET_PATH", request.param)
    return request.param


def cpu(t):
    return t.detach().cpu().numpy()


def packed_step(env, actions, featurizer=None):
    """One step through the compact protocol; returns decoded (rewards f64, done, trunc) and the raw records."""
    cp = env.compact
    pa = cp.pack_actions(torch.as_tensor(np.asarray(actions)).to(env.device))
    assert tuple(pa.shape) == (env.num_envs, cp.action_bytes)
    out = torch.full((env.num_envs, cp.result_bytes), 0xA5, dtype=torch.uint8, device=env.device)
    nf, rec, d, tr, _ = env.step(pa.contiguous(), featurizer=featurizer, packed_actions=True, packed_out=out)
    assert rec is out and d is None and tr is None
    return cp.decode(cpu(out)), nf


@pytest.mark.parametrize("path", golden_files("philox") + golden_files("words"), ids=lambda p: case_of(p) + ("-words" if ".words." in p else ""))
def test_compact_rewards_bit_equal_reference_on_golden_cases(cuda_lib, path):
    """Replay every golden trajectory (minted from the unmodified reference) through packed actions / packed results."""
    g = load(path)
    cfg = CASES[case_of(path)]
    T, N, A = g["actions"].shape
    injected = bool(g["injected"])
    env = make_cuda_env(cfg, N, seed=int(g["seed"]), env_id_base=int(g["env_id_base"]))
    cp = env.compact
    assert np.array_equal(cp.unpack_actions(cp.pack_actions(g["actions"][0])), g["actions"][0])
    if injected:
        env.debug_inject_words(reset_words=g["reset_words0"])
    env.reset()
    for t in range(T):
        if injected:
            env.debug_inject_words(step_words=g["step_words"][t], reset_words=g["reset_words"][t])
        (r, d, tr), nf = packed_step(env, g["actions"][t])
        assert np.array_equal(cpu(nf).astype(np.int64), g["next_flat"][t]), f"state differs at step {t}"
        assert np.array_equal(reward_bits(r), reward_bits(g["rewards"][t])), f"decoded rewards differ at step {t}"
        assert np.array_equal(d, g["done"][t] != 0) and np.array_equal(tr, g["trunc"][t] != 0)
    env.check_actions()


@pytest.mark.parametrize("name", list(CASES))
def test_compact_matches_oracle_at_scale(cuda_lib, name):
    """Ragged N (byte tails of the 3-byte records), fused encode where the case has one, against the oracle's float64 rewards."""
    import sus_net_b200 as S

    cfg = CASES[name]
    N, T, seed = 4099, 150, 31
    env = make_cuda_env(cfg, N, seed=seed, env_id_base=17)
    orc = oracle.OracleEnv(cfg, N, seed=seed, env_id_base=17)
    env.reset(); orc.reset()
    feat = S.GlobalFeaturizer(env) if name in GLOBAL_CASES else (
        flat_featurizer(env, FLAT_COMPONENT_SETS[name][0]) if name in FLAT_COMPONENT_SETS else None)
    for t in range(T):
        a = orc.sample_actions()
        assert np.array_equal(cpu(env.sample_actions()), a)
        (r, d, tr), nf = packed_step(env, a, featurizer=feat if t % 2 else None)
        o = orc.step(a)
        assert np.array_equal(cpu(nf).astype(np.int64), o["next_flat"]), f"{name}: state differs at step {t}"
        assert np.array_equal(reward_bits(r), reward_bits(o["rewards"])), f"{name}: rewards differ at step {t}"
        assert np.array_equal(d, o["done"] != 0) and np.array_equal(tr, o["trunc"] != 0)
    assert np.array_equal(cpu(env.episode_stats()), orc.stats())
    if feat is not None and name in GLOBAL_CASES:
        env.step(None, featurizer=feat); orc.step(None)
        sp, ns = oracle.encode_global(cfg, orc.flat_states())
        views = feat.generate_featurized_states()
        assert np.array_equal(cpu(views[0][0])[:, 0], sp) and np.array_equal(cpu(views[1][1])[:, 0], ns[1])


@pytest.mark.parametrize("k", range(12))
def test_compact_random_constructor_arguments(cuda_lib, k):
    """Random variants / sizes / NON-INTEGER reward constants: decoded rewards equal the oracle's float64 bit patterns."""
    cfg = random_case(np.random.default_rng(770000 + k))
    N, T = 515, 70
    env = make_cuda_env(cfg, N, seed=k, env_id_base=3 * k)
    orc = oracle.OracleEnv(cfg, N, seed=k, env_id_base=3 * k)
    env.reset(); orc.reset()
    for t in range(T):
        a = orc.sample_actions()
        env.sample_actions()
        (r, d, tr), _ = packed_step(env, a)
        o = orc.step(a)
        assert np.array_equal(reward_bits(r), reward_bits(o["rewards"])), f"rewards differ at step {t}: {cfg}"
        assert np.array_equal(d, o["done"] != 0) and np.array_equal(tr, o["trunc"] != 0)
    assert np.array_equal(cpu(env.flat_states(torch.int64)), orc.flat_states())


@pytest.mark.parametrize("name", list(EDGE_CASES))
def test_compact_edge_cases(cuda_lib, name):
    cfg = EDGE_CASES[name]
    N, T = 300, 45
    env = make_cuda_env(cfg, N, seed=4)
    orc = oracle.OracleEnv(cfg, N, seed=4)
    env.reset(); orc.reset()
    for t in range(T):
        a = orc.sample_actions()
        env.sample_actions()
        (r, d, tr), _ = packed_step(env, a)
        o = orc.step(a)
        assert np.array_equal(reward_bits(r), reward_bits(o["rewards"])), f"{name}: rewards differ at step {t}"
        assert np.array_equal(d, o["done"] != 0) and np.array_equal(tr, o["trunc"] != 0)


def test_compact_rejected_actions_have_defined_outputs(cuda_lib):
    """An index outside the role list: the env is untouched, the record carries the all-ones code (NaN) with done = trunc = 0,
    the dense outputs are NaN / 0 / the unchanged state, and check_actions() raises IndexError like the reference."""
    cfg = CASES["cfg4_base_1v4"]
    N = 100
    env = make_cuda_env(cfg, N, seed=8)
    env.reset()
    before = cpu(env.flat_states(torch.int64)).copy()
    imp = cpu(env.imposter_mask_batch)
    a = np.zeros((N, 5), dtype=np.int64)
    bad = np.arange(0, N, 7)
    for e in bad:
        crew = int(np.where(~imp[e])[0][0])
        a[e, crew] = 6  # the crew list has 6 entries
    (r, d, tr), nf = packed_step(env, a)
    assert np.isnan(r[bad]).all() and not d[bad].any() and not tr[bad].any()
    good = np.setdiff1d(np.arange(N), bad)
    assert not np.isnan(r[good]).any()
    assert np.array_equal(cpu(nf).astype(np.int64)[bad], before[bad])
    assert np.array_equal(cpu(env.flat_states(torch.int64))[bad], before[bad])
    with pytest.raises(IndexError):
        env.check_actions()
    # dense protocol: NaN rewards, flags 0, next_flat = the unchanged state
    before = cpu(env.flat_states(torch.int64)).copy()
    imp = cpu(env.imposter_mask_batch)
    a[:] = 0
    for e in bad:
        a[e, int(np.where(~imp[e])[0][0])] = 6
    nf, r, d, tr, _ = env.step(torch.as_tensor(a.astype(np.int32)))
    assert torch.isnan(r[bad]).all() and not d[bad].any() and not tr[bad].any() and not torch.isnan(r[good]).any()
    assert np.array_equal(cpu(nf).astype(np.int64)[bad], before[bad])
    with pytest.raises(IndexError):
        env.check_actions()


def test_compact_unaligned_record_pointers(cuda_lib):
    """Packed buffers at odd byte offsets (sliced views): the staged paths fall back to byte copies; nothing outside is touched."""
    import ctypes as C

    from sus_net_b200 import _lib as L

    cfg = CASES["cfg4_base_1v4"]
    for N in (1003, 64):
        env = make_cuda_env(cfg, N, seed=2)
        ref = make_cuda_env(cfg, N, seed=2)
        env.reset(); ref.reset()
        cp = env.compact
        for off in (1, 2, 3, 5):
            a = ref.sample_actions().clone()
            env.sample_actions()
            _, want_r, want_d, want_t, _ = ref.step(a)
            raw_a = torch.zeros(N * cp.action_bytes + 16, dtype=torch.uint8, device=env.device)
            raw_o = torch.full((N * cp.result_bytes + 16,), 0x5A, dtype=torch.uint8, device=env.device)
            pa = raw_a[off:off + N * cp.action_bytes]
            pa.copy_(cp.pack_actions(a).reshape(-1))
            po = raw_o[off:off + N * cp.result_bytes]
            io = L.SusStepIO()
            io.actions = pa.data_ptr(); io.actions_dtype = L.PACKED; io.packed_out = po.data_ptr()
            L.check(env.lib.sus_env_step(env._h, C.byref(io), env._stream()))
            r, d, tr = cp.decode(cpu(po).reshape(N, cp.result_bytes))
            assert np.array_equal(r.astype(np.float32), cpu(want_r)) and np.array_equal(d, cpu(want_d)) and np.array_equal(tr, cpu(want_t))
            assert (cpu(raw_o[:off]) == 0x5A).all() and (cpu(raw_o[off + N * cp.result_bytes:]) == 0x5A).all()
        assert torch.equal(env.flat_states(torch.int64), ref.flat_states(torch.int64))
        # packed_out together with dense outputs is refused
        io = L.SusStepIO()
        io.packed_out = po.data_ptr(); io.done = po.data_ptr()
        assert env.lib.sus_env_step(env._h, C.byref(io), env._stream()) == L.SUS_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("protocol", ["compact", "dense"])
def test_host_stepper_matches_direct_stepping(cuda_lib, protocol):
    """The pipelined host-buffer driver (3 streams, 2 slots, per-slot feature tensors) returns what plain stepping returns."""
    import sus_net_b200 as S

    cfg = CASES["cfg4_base_1v4"]
    N, T = 3000, 40
    ref = make_cuda_env(cfg, N, seed=21)
    env = make_cuda_env(cfg, N, seed=21)
    ref.reset(); env.reset()
    ref._rewards = torch.zeros((N, 5), dtype=torch.float64, device=ref.device)
    cp = env.compact
    acts, want, want_flat = [], [], []
    for t in range(T):  # record a valid action stream and the expected results
        a = ref.sample_actions().clone()
        acts.append((cp.pack_actions(a) if protocol == "compact" else a.to(torch.uint8)).cpu().pin_memory())
        _, r, d, tr, _ = ref.step(a)
        want.append((cpu(r).copy(), cpu(d).copy(), cpu(tr).copy()))
        want_flat.append(cpu(ref.flat_states(torch.int64)).copy())
    feat = S.GlobalFeaturizer(env)
    torch.cuda.synchronize()
    stepper = S.HostStepper(env, featurizer=feat, protocol=protocol)
    assert stepper.d2h_bytes_per_step == (N * cp.result_bytes if protocol == "compact" else N * 22)
    prev = None
    for t in range(T):
        slot = stepper.step(acts[t])
        if prev is not None:  # results and features of step t-1 are read while step t is in flight
            pslot, pt = prev
            r, d, tr = stepper.wait(pslot)
            r, d, tr = (np.asarray(x) if not isinstance(x, torch.Tensor) else x.numpy() for x in (r, d, tr))
            if protocol == "compact":
                assert np.array_equal(reward_bits(r), reward_bits(want[pt][0]))
            else:
                assert np.array_equal(r, want[pt][0].astype(np.float32))
            assert np.array_equal(d, want[pt][1]) and np.array_equal(tr, want[pt][2])
            views = stepper.features(pslot)
            sp, ns = oracle.encode_global(cfg, want_flat[pt])
            assert np.array_equal(cpu(views[0][0])[:, 0], sp) and np.array_equal(cpu(views[3][1])[:, 0], ns[3])
            stepper.release_features(pslot)
        prev = (slot, t)
    stepper.drain()
    torch.cuda.synchronize()
    env.check_actions()
    assert torch.equal(env.flat_states(torch.int64), ref.flat_states(torch.int64))
