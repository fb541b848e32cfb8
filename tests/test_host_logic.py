"""CPU: host-side logic that needs no GPU -- the C-ABI library loads and exports every symbol the header
declares, config validation mirrors the reference's asserts, shape queries, sharding, the gloo stat reduce."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sus_net_b200 import build as B

    B.build()
    from sus_net_b200 import _lib as L

    return L


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "susnet_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*\*?\s*(sus_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 28, declared
    assert declared == set(lib.EXPORTED_SYMBOLS)
    L = lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.sus_abi_version() == lib.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), f"{name} is not an exported text symbol"


def test_struct_layout_matches_header(lib):
    # sizes the C compiler gives the header's structs (computed by compiling a probe with gcc)
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "susnet_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(SusConfig), sizeof(SusEncodeSpec), sizeof(SusEncodeShape),
             sizeof(SusStepIO), offsetof(SusConfig, kill_reward), offsetof(SusConfig, num_envs),
             sizeof(SusCompactLayout), offsetof(SusStepIO, packed_out));
      return 0;
    }'''
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write(probe)
        exe = os.path.join(d, "p")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    want = [C.sizeof(lib.SusConfig), C.sizeof(lib.SusEncodeSpec), C.sizeof(lib.SusEncodeShape), C.sizeof(lib.SusStepIO),
            lib.SusConfig.kill_reward.offset, lib.SusConfig.num_envs.offset, C.sizeof(lib.SusCompactLayout),
            lib.SusStepIO.packed_out.offset]
    assert [int(x) for x in out] == want


def cfg(lib, **kw):
    base = dict(variant=0, n_imposters=1, n_crew=4, n_jobs=5, include_walls=1, is_action_order_random=1,
                shuffle_imposter_index=1, max_time_steps=1000, tag_reset_interval=50, auto_reset=1, num_envs=16)
    base.update(kw)
    return lib.SusConfig(**base)


def test_compact_protocol_layout_and_decode_table(lib):
    """Host calls of the compact host protocol: record geometry per config and the float64 decode table, recomputed here
    from the reference's reward rules (base.py:511-532,553-563,389-390; tagging.py:162,196)."""
    L = lib.lib()

    def layout(c):
        out = lib.SusCompactLayout()
        assert L.sus_compact_layout(C.byref(c), C.byref(out)) == 0
        return out

    l4 = layout(cfg(lib))  # cfg4: 7 imposter actions -> 3 bits x 5 = 2 bytes; 13 codes -> 4 bits x 5 + 2 flags = 3 bytes
    assert (l4.action_bits, l4.action_bytes, l4.reward_bits, l4.result_bytes, l4.n_codes, l4.invalid_code) == (3, 2, 4, 3, 13, 15)
    l3 = layout(cfg(lib, variant=1, n_crew=2))  # cfg3 tagging 1v2: 9 actions -> 4 bits; 37 codes -> 6 bits x 3 + 2 = 20 bits
    assert (l3.action_bits, l3.action_bytes, l3.reward_bits, l3.result_bytes, l3.n_codes, l3.invalid_code) == (4, 2, 6, 3, 37, 63)
    l8 = layout(cfg(lib, variant=1, n_imposters=3, n_crew=5, n_jobs=2))  # 8 agents: 14 actions, 8 x 6 + 2 = 50 bits
    assert (l8.action_bits, l8.action_bytes, l8.reward_bits, l8.result_bytes) == (4, 4, 6, 7)
    # decode table of a config with awkward constants
    c = cfg(lib, variant=1, n_imposters=2, n_crew=3, n_jobs=2, kill_reward=-5.1, complete_job_reward=1.0 / 3.0,
            sabotage_reward=0.7, time_step_reward=-0.25, game_end_reward=10.3, dead_penalty=-2.2, vote_reward=0.1)
    lay = layout(c)
    lut = np.zeros((5, lay.invalid_code + 1))
    assert L.sus_reward_lut(C.byref(c), lut.ctypes.data_as(C.POINTER(C.c_double))) == 0
    for i in range(5):
        for code in range(lay.n_codes):
            if code == lay.n_codes - 1:
                want = -2.2
            else:
                ev, team = code & 3, code >> 2
                win, vote = team % 3, team // 3
                v = [-0.25, -5.1, 1.0 / 3.0, -0.7][ev]
                t = 0.0
                if vote:
                    t += 0.1 * (-1.0 if vote == 2 else 1.0)
                t += [0.0, 10.3, -10.3][win]
                v += t
                want = -v if i < 2 else v
            assert lut[i, code] == want and np.signbit(lut[i, code]) == np.signbit(want), (i, code)
        assert np.isnan(lut[i, lay.n_codes:]).all()
    # base env: zeros become time_step_reward (base.py:389-390), also for a dead agent with dead_penalty == 0
    c = cfg(lib, time_step_reward=-1.5, dead_penalty=0.0, kill_reward=0.0, complete_job_reward=3.0)
    lay = layout(c)
    lut = np.zeros((5, lay.invalid_code + 1))
    assert L.sus_reward_lut(C.byref(c), lut.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert lut[3, 0] == -1.5 and lut[3, 1] == -1.5 and lut[3, lay.n_codes - 1] == -1.5 and lut[0, 2] == -3.0


def test_flat_size_and_action_counts(lib):
    L = lib.lib()
    assert L.sus_flat_state_size(C.byref(cfg(lib))) == 30                       # cfg4
    assert L.sus_flat_state_size(C.byref(cfg(lib, variant=1, n_crew=2))) == 31  # cfg3
    assert L.sus_flat_state_size(C.byref(cfg(lib, variant=2, n_crew=1, n_jobs=0))) == 6   # cfg1/2
    assert L.sus_flat_state_size(C.byref(cfg(lib, variant=2, n_crew=4, n_jobs=0))) == 15  # cfg4-alt
    assert L.sus_n_role_actions(C.byref(cfg(lib)), 1) == 7 and L.sus_n_role_actions(C.byref(cfg(lib)), 0) == 6
    t = cfg(lib, variant=1, n_crew=2)
    assert L.sus_n_role_actions(C.byref(t), 1) == 9 and L.sus_n_role_actions(C.byref(t), 0) == 8
    g = cfg(lib, variant=2, n_crew=4, n_jobs=0)
    assert L.sus_n_role_actions(C.byref(g), 1) == 6 and L.sus_n_role_actions(C.byref(g), 0) == 5


def test_config_validation_mirrors_reference_asserts(lib):
    L = lib.lib()
    for bad in (dict(n_imposters=0), dict(n_crew=0), dict(n_jobs=-1), dict(n_imposters=2, n_crew=2)):
        with pytest.raises(AssertionError):
            lib.check(L.sus_flat_state_size(C.byref(cfg(lib, **bad))))
    for unsupported in (dict(n_crew=8), dict(n_jobs=9), dict(variant=1, n_jobs=0)):
        with pytest.raises(NotImplementedError):
            lib.check(L.sus_flat_state_size(C.byref(cfg(lib, **unsupported))))
    # the training ground only needs n_crew > 0 (pred_prey.py:75-76)
    assert L.sus_flat_state_size(C.byref(cfg(lib, variant=2, n_crew=1, n_jobs=0))) == 6


def test_encode_shapes(lib):
    L = lib.lib()
    sh = lib.SusEncodeShape()
    spec = lib.SusEncodeSpec(kind=lib.ENCODE_GLOBAL)
    lib.check(L.sus_encode_shape(C.byref(cfg(lib)), C.byref(spec), C.byref(sh)))
    assert (sh.spatial_floats, sh.non_spatial_floats, sh.spatial_views, sh.non_spatial_views) == (567, 15, 1, 5)
    spec = lib.SusEncodeSpec(kind=lib.ENCODE_PERSPECTIVE)
    lib.check(L.sus_encode_shape(C.byref(cfg(lib, variant=1)), C.byref(spec), C.byref(sh)))
    assert (sh.spatial_floats, sh.non_spatial_floats, sh.spatial_views, sh.non_spatial_views) == (567, 15, 5, 5)
    spec = lib.SusEncodeSpec(kind=lib.ENCODE_FLAT, n_components=3)
    spec.components[0], spec.components[1], spec.components[2] = 0, 2, 3  # README 1v4 recipe: F = 98
    lib.check(L.sus_encode_shape(C.byref(cfg(lib, variant=2, n_jobs=0)), C.byref(spec), C.byref(sh)))
    assert (sh.spatial_floats, sh.non_spatial_floats, sh.spatial_views, sh.non_spatial_views) == (0, 98, 0, 1)
    with pytest.raises(AssertionError):  # Global needs jobs
        lib.check(L.sus_encode_shape(C.byref(cfg(lib, variant=2, n_jobs=0)), C.byref(lib.SusEncodeSpec(kind=1)), C.byref(sh)))


def test_package_fails_loudly_without_cuda(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import sus_net_b200 as S

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.FourRoomEnv(1, 4, 5)


def test_shard_ranges_cover_and_partition():
    from sus_net_b200.distributed import shard_range

    for total in (0, 1, 7, 8, 65536, 1_000_003):
        for ws in (1, 2, 3, 8):
            parts = [shard_range(total, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sus_net_b200.distributed import reduce_episode_stats, shard_range, max_over_ranks
import numpy as np
import oracle
from tests.cases import CASES
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
cfg = CASES["cfg3_tagging_1v2"]
TOTAL, T = 600, 120
lo, hi = shard_range(TOTAL, rank, 2)
env = oracle.OracleEnv(cfg, hi - lo, seed=4, env_id_base=lo)   # stand-in for the per-rank GPU env
env.reset()
for _ in range(T):
    env.step(None)
total = reduce_episode_stats(torch.as_tensor(env.stats()))
full = oracle.OracleEnv(cfg, TOTAL, seed=4)
full.reset()
for _ in range(T):
    full.step(None)
assert np.array_equal(total.numpy(), full.stats()), (total, full.stats())
assert max_over_ranks(float(rank)) == 1.0
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_stat_reduce_matches_single_shard():
    """world_size 2 on CPU: sharded envs + the one collective of the path (episode-stat all-reduce) reproduce the
    unsharded run.  The per-rank env here is the oracle (the CUDA env needs a GPU); the plumbing is the product's."""
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, "-c", _GLOO_WORKER, ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_missing_cuda_library_is_an_import_error(lib, monkeypatch, tmp_path):
    """No silent fallback: without the built .so the binding raises (the product never routes through the oracle)."""
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "libsusnet_b200.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        lib.lib()
    src = "".join(open(os.path.join(ROOT, "sus_net_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "sus_net_b200"))
                  if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_compact_protocol_host_codec_matches_the_numpy_spec(lib):
    """sus_host_pack_actions / sus_host_decode_results (threaded C over host buffers) against the numpy statement of the record
    format in sus_net_b200/compact.py, for every variant's geometry, ragged sizes, both reward dtypes, and out-of-field indices."""
    from sus_net_b200.compact import CompactProtocol

    L = lib.lib()

    class Env:  # the three attributes CompactProtocol reads
        pass

    rng = np.random.default_rng(0)
    for kw in (dict(), dict(variant=1, n_crew=2), dict(variant=1, n_imposters=3, n_crew=5, n_jobs=2), dict(variant=2, n_crew=1, n_jobs=0)):
        e = Env()
        e.lib, e._cfg = L, cfg(lib, kill_reward=-5.1, complete_job_reward=1 / 3, sabotage_reward=3.0, game_end_reward=10.0,
                               dead_penalty=-2.0, vote_reward=0.7, time_step_reward=-0.5, **kw)
        e.n_agents = e._cfg.n_imposters + e._cfg.n_crew
        cp = CompactProtocol(e)
        for n in (0, 1, 7, 70001):
            a = rng.integers(0, 1 << cp.action_bits, (n, e.n_agents))
            for dt in (np.uint8, np.int32, np.int64):
                p = cp.pack_actions(a.astype(dt))
                assert p.shape == (n, cp.action_bytes) and np.array_equal(cp.unpack_actions(p), a)
            if n:
                bad = a.astype(np.int64).copy()
                bad[0, 0] = -3; bad[n - 1, e.n_agents - 1] = 1 << 20
                q = cp.unpack_actions(cp.pack_actions(bad))
                assert q[0, 0] == (1 << cp.action_bits) - 1 and q[n - 1, e.n_agents - 1] == (1 << cp.action_bits) - 1
            codes = rng.integers(0, cp.invalid_code + 1, (n, e.n_agents)).astype(np.uint64)
            rec = np.zeros(n, dtype=np.uint64)
            for i in range(e.n_agents):
                rec |= codes[:, i] << np.uint64(i * cp.reward_bits)
            rec |= rng.integers(0, 4, n).astype(np.uint64) << np.uint64(e.n_agents * cp.reward_bits)
            r = np.stack([(rec >> np.uint64(8 * b)) & np.uint64(255) for b in range(cp.result_bytes)], 1).astype(np.uint8).reshape(n, cp.result_bytes)
            want = cp.decode_numpy(r)
            got = cp.decode(r)
            assert np.array_equal(got[0].view(np.int64), want[0].view(np.int64)) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
            got32 = cp.decode(r, np.float32)
            assert np.array_equal(got32[0].view(np.int32), want[0].astype(np.float32).view(np.int32))


def test_mlp_workspace_size_is_a_host_computation(lib):
    """sus_mlp_workspace_bytes needs no device: per layer [column blocks of 16 * CT columns] x [chunks of 16 k] floats, zero
    padded, CT = 8 / 4 / 1 for layers with > 64 / > 16 / <= 16 outputs (the rule the forward kernel stages by); 0 = invalid."""
    L = lib.lib()

    def spec_of(dims, weights=True):
        s = lib.SusMlpSpec(n_layers=len(dims) - 1, activation=lib.ACT_PRELU)
        for i, d in enumerate(dims):
            s.dims[i] = d
        for l in range(len(dims) - 1):
            s.weight[l] = 0x1000 if weights else None  # never dereferenced on the host
        return s

    def want(dims):
        total = 0
        for k, m in zip(dims[:-1], dims[1:]):
            cb = 16 * (8 if m > 64 else 4 if m > 16 else 1)
            total += -(-m // cb) * cb * -(-k // 16) * 16
        return 4 * total

    for dims in ([98, 256, 128, 64, 16, 6], [4, 6], [200, 150, 17, 129, 3], [36, 200, 100, 7], [16, 400, 400, 2]):
        assert L.sus_mlp_workspace_bytes(C.byref(spec_of(dims))) == want(dims), dims
    assert want([98, 256, 128, 64, 16, 6]) == 283648
    assert L.sus_mlp_workspace_bytes(C.byref(spec_of([8, 0, 2]))) == 0          # a layer width < 1
    assert L.sus_mlp_workspace_bytes(C.byref(spec_of([8, 4, 2], weights=False))) == 0  # NULL weight
    assert L.sus_mlp_workspace_bytes(None) == 0
