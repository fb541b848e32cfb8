"""GPU: opt-in byte planes (SUS_ENCODE_PLANES_U8; `GlobalFeaturizer(env, plane_dtype=torch.uint8)`): the spatial tensor holds
one byte per cell instead of one float32.  Same oracle as the float planes: the bytes, cast to float, must EQUAL the
reference-pinned planes -- fused with the step, standalone on the live state, and from (B, T, S) rows -- on every store path."""
import numpy as np
import pytest
import torch

import oracle
from tests.cases import GLOBAL_CASES
from tests.util import CASES, make_cuda_env

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["ws", "tma", "direct", "staged"], autouse=True)
def store_path(request, monkeypatch):
    monkeypatch.setenv("SUSNET_PATH", request.param)
    return request.param


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", GLOBAL_CASES)
@pytest.mark.parametrize("N", [1000, 4099, 7])
def test_byte_planes_equal_the_float_planes(cuda_lib, name, N):
    import sus_net_b200 as S

    cfg = CASES[name]
    env = make_cuda_env(cfg, N, seed=5, env_id_base=7)
    orc = oracle.OracleEnv(cfg, N, seed=5, env_id_base=7)
    env.reset(); orc.reset()
    feats = [("global", S.GlobalFeaturizer(env, plane_dtype=torch.uint8)), ("perspective", S.PerspectiveFeaturizer(env, plane_dtype=torch.uint8))]
    for t in range(24):
        kind, f = feats[t % 2]
        env.step(None, featurizer=f)
        orc.step(None)
        cur = orc.flat_states()
        sp, ns = (oracle.encode_global if kind == "global" else oracle.encode_perspective)(cfg, cur)
        for views in (f.generate_featurized_states(), f.encode_env()):  # fused with the step / standalone on the live state
            for k, (vsp, vns) in enumerate(views):
                assert vsp.dtype == torch.uint8 and not vsp.requires_grad and vns.dtype == torch.float32
                want = sp if kind == "global" else sp[k]
                assert np.array_equal(cpu(vsp)[:, 0].astype(np.float32), want), f"{name}: {kind} planes differ at step {t} (view {k})"
                assert np.array_equal(cpu(vns)[:, 0], ns[k])
    cur = orc.flat_states()
    for kind, f in feats:  # fit() on (B, T, S) rows
        for dtype in (torch.float32, torch.int64):
            f.fit(torch.as_tensor(cur).to(dtype).reshape(N, 1, -1))
            sp, ns = (oracle.encode_global if kind == "global" else oracle.encode_perspective)(cfg, cur)
            for k, (vsp, vns) in enumerate(f.generate_featurized_states()):
                assert np.array_equal(cpu(vsp)[:, 0].astype(np.float32), sp if kind == "global" else sp[k])
                assert np.array_equal(cpu(vns)[:, 0], ns[k])
    assert np.array_equal(cpu(env.flat_states(torch.int64)), cur)


def test_byte_planes_unaligned_and_untouched_surroundings(cuda_lib):
    """Plane tensors at odd byte offsets: bulk stores fall back to byte copies; nothing outside the tensors is written."""
    import ctypes as C

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
            spec = L.SusEncodeSpec(kind=kind, flags=L.ENCODE_PLANES_U8)
            L.check(env.lib.sus_encode_shape(C.byref(env._cfg), C.byref(spec), C.byref(sh)))
            sp_n = sh.spatial_views * N * sh.spatial_floats
            ns_n = sh.non_spatial_views * N * sh.non_spatial_floats
            for off in (0, 1, 3, 8):
                sp_raw = torch.full((sp_n + 32,), 0x77, dtype=torch.uint8, device=env.device)
                ns = torch.empty(ns_n, device=env.device)
                sp = sp_raw[off:off + sp_n]
                L.check(env.lib.sus_env_encode(env._h, C.byref(spec), C.c_void_p(sp.data_ptr()), C.c_void_p(ns.data_ptr()), env._stream()))
                want_sp, want_ns = enc(cfg, cur)
                assert np.array_equal(cpu(sp).reshape(want_sp.shape).astype(np.float32), want_sp)
                assert np.array_equal(cpu(ns).reshape(want_ns.shape), want_ns)
                assert (cpu(sp_raw[:off]) == 0x77).all() and (cpu(sp_raw[off + sp_n:]) == 0x77).all()
