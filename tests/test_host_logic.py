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
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(SusConfig), sizeof(SusEncodeSpec), sizeof(SusEncodeShape),
             sizeof(SusStepIO), offsetof(SusConfig, kill_reward), offsetof(SusConfig, num_envs));
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
            lib.SusConfig.kill_reward.offset, lib.SusConfig.num_envs.offset]
    assert [int(x) for x in out] == want


def cfg(lib, **kw):
    base = dict(variant=0, n_imposters=1, n_crew=4, n_jobs=5, include_walls=1, is_action_order_random=1,
                shuffle_imposter_index=1, max_time_steps=1000, tag_reset_interval=50, auto_reset=1, num_envs=16)
    base.update(kw)
    return lib.SusConfig(**base)


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
