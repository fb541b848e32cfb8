"""CPU: WHOLE kernels of the library on the host with real barriers and warp collectives (tests/emu/simt.h: every CUDA thread of
a block is a fiber), against the oracle.  Where tests/test_host_kernel_emulation.py runs the per-env device functions, this runs
the kernels' own data flow through shared memory: k_reset, k_sample_actions (actions staged per warp, written lane-contiguously),
the direct-store fused kernel k_step<V, ENCODE> (planes zero-filled by all lanes of a warp, ones scattered per lane, ragged last
warp) and the byte-staged Flat kernel k_step_flat (rows prefilled and built as bytes in shared memory, expanded to floats by all
lanes; rewards staged and copied) incl. its compile-time-shape instantiation.  The TMA / mbarrier kernels have no host
counterpart and stay with the GPU tests."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
from tests.cases import CASES, EDGE_CASES, FLAT_COMPONENT_SETS, FLAT_COMPONENT_SETS_TAGGING, GLOBAL_CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sus_net_b200", "csrc")
ALL = {**CASES, **EDGE_CASES}


@pytest.fixture(scope="module")
def simt(tmp_path_factory):
    from sus_net_b200 import _lib as L

    d = tmp_path_factory.mktemp("simt")
    for name in ("susnet_tile.cuh", "susnet_ws.cuh"):  # the PTX of the TMA path has no host counterpart: compile it away
        text = open(os.path.join(CSRC, name)).read()
        (d / name).write_text(re.sub(r"asm volatile\(.*?\);", ";", text, flags=re.S))
    api = open(os.path.join(CSRC, "susnet_api.cu")).read()
    consts = "\n".join(re.findall(r"^constexpr (?:int|unsigned) k\w+ = [^;]+;[^\n]*$", api[:api.index("struct StepParams {")], flags=re.M))
    kernels = api[api.index("struct StepParams {"):api.index("// ------------------------------------------------------------------------------------------ host side")]
    kernels = re.sub(r"extern __shared__ __align__\(\d+\)", "extern", kernels)  # the harness defines the dynamic shared array
    host = api[api.index("int flat_size(const SusConfig& c) {"):api.index("inline int32_t align128")]
    stage = re.search(r"bool make_flat_stage\(.*?\n\}\n", api, re.S).group(0)
    inc = d / "kernels.inc"
    inc.write_text("\n".join([consts, kernels, host, stage]))
    so = str(d / "kernels_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", str(d), "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "include"), "-I", CSRC, f'-DKERNEL_SOURCE="{inc}"',
                    os.path.join(ROOT, "tests", "emu", "kernels_emu.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.emu_k_reset.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp, vp, vp]
    lib.emu_k_sample_actions.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp]
    lib.emu_k_step.argtypes = [C.POINTER(L.SusConfig), C.c_int, C.c_uint64, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp,
                               C.POINTER(L.SusEncodeSpec), vp, vp, vp, vp]
    lib.emu_k_rollout.argtypes = [C.POINTER(L.SusConfig), C.c_int, C.c_uint64, C.c_int32, vp, vp, vp, vp, vp, vp]
    lib.emu_k_encode_rows.argtypes = [C.POINTER(L.SusConfig), C.POINTER(L.SusEncodeSpec), C.c_int, C.c_int, vp, C.c_int64, vp, vp]
    return lib


class KernelEnv:
    """Env state in numpy; every call is one emulated kernel launch (what BatchedFourRoomEnv does on the GPU)."""

    def __init__(self, lib, cfg, N, seed, env_id_base=0):
        from sus_net_b200 import _lib as L

        self.L, self.lib, self.N, self.A = L, lib, N, cfg["n_imposters"] + cfg["n_crew"]
        kw = {k: (oracle.VARIANT_IDS[v] if k == "variant" else (int(v) if isinstance(v, bool) else v)) for k, v in cfg.items()}
        self.cfg = L.SusConfig(num_envs=N, seed=seed, env_id_base=env_id_base, auto_reset=1, **kw)
        self.S = oracle.flat_size(cfg)
        self.pos, self.jobpos = np.zeros(N, np.uint64), np.zeros(N, np.uint64)
        self.aux, self.met = np.zeros((N, 4), np.uint32), np.zeros((N, 4), np.uint32)
        self.stats, self.err = np.zeros(10, np.uint64), np.zeros(1, np.uint32)
        self.ticks = [0, 0, 0]  # step, reset, act

    def _state(self):
        return [a.ctypes.data for a in (self.pos, self.jobpos, self.aux, self.met)]

    def reset(self):
        self.lib.emu_k_reset(C.byref(self.cfg), self.ticks[1], *self._state())
        self.ticks[1] += 1

    def sample_actions(self):
        out = np.full((self.N, self.A), -1, np.int32)
        self.lib.emu_k_sample_actions(C.byref(self.cfg), self.ticks[2], self.aux.ctypes.data, out.ctypes.data)
        self.ticks[2] += 1
        return out

    def step(self, actions, path=0, kind=0, components=(), f64=True):
        L = self.L
        spec = L.SusEncodeSpec(kind=kind, n_components=len(components))
        for i, name in enumerate(components):
            spec.components[i] = oracle.FLAT_COMPONENTS[name]
        shape = L.SusEncodeShape()
        assert L.lib().sus_encode_shape(C.byref(self.cfg), C.byref(spec), C.byref(shape)) == 0
        a = None if actions is None else np.ascontiguousarray(actions, np.int32)
        r = np.full((self.N, self.A), np.nan, np.float64 if f64 else np.float32)
        done, trunc = np.full(self.N, 9, np.uint8), np.full(self.N, 9, np.uint8)
        nf = np.full((self.N, self.S), np.nan, np.float32)
        sp = np.full((max(shape.spatial_views, 1), self.N, max(shape.spatial_floats, 1)), np.nan, np.float32)
        ns = np.full((max(shape.non_spatial_views, 1), self.N, max(shape.non_spatial_floats, 1)), np.nan, np.float32)
        rc = self.lib.emu_k_step(C.byref(self.cfg), path, self.ticks[0], *self._state(), None if a is None else a.ctypes.data,
                                 r.ctypes.data, L.F64 if f64 else L.F32, done.ctypes.data, trunc.ctypes.data, nf.ctypes.data,
                                 C.byref(spec) if kind else None, sp.ctypes.data, ns.ctypes.data, self.stats.ctypes.data,
                                 self.err.ctypes.data)
        if rc != 0:
            return None
        self.ticks[0] += 1
        return dict(rewards=r, done=done, trunc=trunc, next_flat=nf.astype(np.int64), spatial=sp, non_spatial=ns)


def _same_step(got, want, where):
    assert np.array_equal(got["next_flat"], want["next_flat"]), where
    assert np.array_equal(got["rewards"].astype(np.float64).view(np.int64), want["rewards"].astype(got["rewards"].dtype).astype(np.float64).view(np.int64)), where
    assert np.array_equal(got["done"], want["done"]) and np.array_equal(got["trunc"], want["trunc"]), where


@pytest.mark.parametrize("case", sorted(ALL))
def test_step_only_kernels_equal_the_oracle(simt, case):
    """k_reset, k_sample_actions and k_step<V, false> on a ragged batch (two full blocks + a partial warp): states, sampled
    actions, float64 reward bit patterns, done / truncated, statistics accumulated through the warp reductions."""
    cfg = ALL[case]
    N, seed, base = 2 * 256 + 37, 21, 1000
    env, orc = KernelEnv(simt, cfg, N, seed, base), oracle.OracleEnv(cfg, N, seed=seed, env_id_base=base)
    env.reset()
    orc.reset()
    for t in range(40):
        a, want_a = env.sample_actions(), orc.sample_actions()
        assert np.array_equal(a, want_a), (case, t)
        use = a if t % 3 else None  # every third step through the fused random policy
        if use is None:
            want = orc.step(None)
        else:
            want = orc.step(use)
        _same_step(env.step(use), want, (case, t))
    assert np.array_equal(env.stats.astype(np.int64), orc.stats()) and int(env.err[0]) == 0


@pytest.mark.parametrize("case", GLOBAL_CASES)
def test_direct_fused_kernel_with_plane_encodes_equals_the_oracle(simt, case):
    """k_step<V, true> + warp_encode: Global and Perspective feature tensors written next to the step's own outputs."""
    cfg = CASES[case]
    A = cfg["n_imposters"] + cfg["n_crew"]
    N, seed = 256 + 70, 4
    for kind, encode in ((1, oracle.encode_global), (2, oracle.encode_perspective)):
        env, orc = KernelEnv(simt, cfg, N, seed), oracle.OracleEnv(cfg, N, seed=seed)
        env.reset()
        orc.reset()
        for t in range(25):
            want = orc.step(None)
            got = env.step(None, path=0, kind=kind, f64=False)
            _same_step(got, want, (case, kind, t))
            sp, ns = encode(cfg, orc.flat_states())
            assert np.array_equal(got["non_spatial"], ns), (case, kind, t)
            got_sp = got["spatial"][0].reshape(-1, A + 2, 9, 9) if kind == 1 else got["spatial"].reshape(A, -1, A + 2, 9, 9)
            assert np.array_equal(got_sp, sp), (case, kind, t)


@pytest.mark.parametrize("case", sorted({**FLAT_COMPONENT_SETS, **FLAT_COMPONENT_SETS_TAGGING}))
def test_flat_kernels_equal_the_oracle(simt, case):
    """Flat rows through the direct kernel (float rows), the byte-staged kernel k_step_flat and -- for the training shape -- its
    compile-time instantiation: rows, float32 rewards, done / truncated, states."""
    cfg = CASES[case]
    N, seed = 256 + 45, 8
    for comps in {**FLAT_COMPONENT_SETS, **FLAT_COMPONENT_SETS_TAGGING}[case]:
        ran = []
        for path in (0, 1, 2):
            env, orc = KernelEnv(simt, cfg, N, seed), oracle.OracleEnv(cfg, N, seed=seed)
            env.reset()
            orc.reset()
            for t in range(20):
                got = env.step(None, path=path, kind=3, components=comps, f64=False)
                if got is None:
                    break  # the path does not apply (float-valued scent rows are never byte-staged; other shapes for path 2)
                want = orc.step(None)
                _same_step(got, want, (case, comps, path, t))
                assert np.array_equal(got["non_spatial"][0], oracle.encode_flat(cfg, comps, orc.flat_states())), (case, comps, path, t)
            else:
                ran.append(path)
        F = oracle.encode_flat(cfg, comps, orc.flat_states()).shape[1]
        # (rows wider than ~170 floats do not fit four byte-staged CTAs per SM: the library takes the TMA path for them)
        assert 0 in ran and (1 in ran or "scent" in comps or F > 170)
        if case == "cfg4alt_itg_1v4" and "scent" not in comps:
            assert 2 in ran


@pytest.mark.parametrize("case", ["cfg2_itg_1v1_wall", "cfg3_tagging_1v2", "cfg4_base_1v4", "cfg4alt_itg_1v4", "tagging_2v5_short",
                                  "base_fixed_order_tsr", "edge_max_time_steps_1"])
def test_rollout_kernel_equals_n_oracle_steps(simt, case):
    """k_rollout: n random-policy steps per launch with the env state in registers == n x step(None): states, statistics and the
    per-agent reward sums, for the generic instantiation and for the compile-time shapes the library picks."""
    cfg = ALL[case]
    N, seed, n = 256 + 33, 6, 37
    for shape in (0, 1):
        env, orc = KernelEnv(simt, cfg, N, seed), oracle.OracleEnv(cfg, N, seed=seed)
        env.reset()
        orc.reset()
        sums = np.zeros((N, env.A), np.float64)
        want_sums = np.zeros((N, env.A), np.float64)
        ok = True
        for launch in range(2):
            rc = simt.emu_k_rollout(C.byref(env.cfg), shape, env.ticks[0], n, *env._state(), env.stats.ctypes.data, sums.ctypes.data)
            if rc != 0:
                ok = False
                break
            env.ticks[0] += n
            part = np.zeros_like(want_sums)
            for _ in range(n):
                part += orc.step(None)["rewards"]
            assert np.array_equal(sums, part), (case, shape, launch)  # same summation order per env: bit-identical
            assert np.array_equal(_flat(env), orc.flat_states()), (case, shape, launch)
        if ok:
            assert np.array_equal(env.stats.astype(np.int64), orc.stats()), (case, shape)
        else:
            assert shape == 1  # no compile-time instantiation for this shape


def _flat(env):
    """Flattened states of a KernelEnv through the per-env device code of the other harness' twin: decode the records here."""
    A = env.A
    N = env.N
    cfgd = env.cfg
    J = cfgd.n_jobs
    out = []
    pos = env.pos.view(np.uint8).reshape(N, 8)[:, :A]
    out += [np.stack([pos >> 4, pos & 15], axis=2).reshape(N, 2 * A).astype(np.int64)]
    alive = env.aux[:, 0] & 0xff
    out += [((alive[:, None] >> np.arange(A)) & 1).astype(np.int64)]
    if J > 0 or cfgd.variant == 1:
        jp = env.jobpos.view(np.uint8).reshape(N, 8)[:, :J]
        out += [np.stack([jp >> 4, jp & 15], axis=2).reshape(N, 2 * J).astype(np.int64)]
        jd = (env.aux[:, 0] >> 16) & 0xff
        out += [((jd[:, None] >> np.arange(J)) & 1).astype(np.int64)]
    if cfgd.variant == 1:
        used = env.aux[:, 0] >> 24
        out += [((used[:, None] >> np.arange(A)) & 1).astype(np.int64)]
        out += [((env.aux[:, 3][:, None] >> (4 * np.arange(A))) & 15).astype(np.int64)]
        out += [(cfgd.tag_reset_interval - env.aux[:, 2].astype(np.int64))[:, None]]
    return np.concatenate(out, axis=1)


def test_fit_kernels_on_replay_rows_equal_the_oracle(simt):
    """SequenceStateFeaturizer.fit: k_encode_rows<T> for float32 / float64 / int64 rows (every encode kind, direct stores) and the
    byte-staged k_encode_flat<T, true>, on a ragged batch of states taken along oracle trajectories."""
    from sus_net_b200 import _lib as L

    def rows_of(cfg, n):
        orc = oracle.OracleEnv(cfg, n, seed=2)
        orc.reset()
        for _ in range(30):
            orc.step(None)
        return orc.flat_states()

    def run(cfg, kind, comps, staged, rows, dtype):
        env = KernelEnv(simt, cfg, 1, 0)
        spec = L.SusEncodeSpec(kind=kind, n_components=len(comps))
        for i, name in enumerate(comps):
            spec.components[i] = oracle.FLAT_COMPONENTS[name]
        shape = L.SusEncodeShape()
        assert L.lib().sus_encode_shape(C.byref(env.cfg), C.byref(spec), C.byref(shape)) == 0
        n = rows.shape[0]
        sp = np.full((max(shape.spatial_views, 1), n, max(shape.spatial_floats, 1)), np.nan, np.float32)
        ns = np.full((max(shape.non_spatial_views, 1), n, max(shape.non_spatial_floats, 1)), np.nan, np.float32)
        r = np.ascontiguousarray(rows, dtype)
        code = {np.float32: L.F32, np.float64: L.F64, np.int64: L.I64}[dtype]
        rc = simt.emu_k_encode_rows(C.byref(env.cfg), C.byref(spec), staged, code, r.ctypes.data, n, sp.ctypes.data, ns.ctypes.data)
        return None if rc != 0 else (sp, ns)

    n = 256 + 19
    for case in GLOBAL_CASES[:3]:
        cfg = CASES[case]
        A = cfg["n_imposters"] + cfg["n_crew"]
        rows = rows_of(cfg, n)
        for dtype in (np.float32, np.float64, np.int64):
            sp, ns = run(cfg, 1, (), 0, rows, dtype)
            want_sp, want_ns = oracle.encode_global(cfg, rows)
            assert np.array_equal(sp[0].reshape(-1, A + 2, 9, 9), want_sp) and np.array_equal(ns, want_ns), (case, dtype)
        sp, ns = run(cfg, 2, (), 0, rows, np.float32)
        want_sp, want_ns = oracle.encode_perspective(cfg, rows)
        assert np.array_equal(sp.reshape(A, -1, A + 2, 9, 9), want_sp) and np.array_equal(ns, want_ns), case
    for case, sets in FLAT_COMPONENT_SETS.items():
        cfg = CASES[case]
        rows = rows_of(cfg, n)
        for comps in sets:
            want = oracle.encode_flat(cfg, comps, rows)
            for staged in (0, 1):
                for dtype in (np.float32, np.int64):
                    got = run(cfg, 3, comps, staged, rows, dtype)
                    if got is None:
                        assert staged == 1 and ("scent" in comps or want.shape[1] > 170)
                        continue
                    assert np.array_equal(got[1][0].view(np.int32), want.view(np.int32)), (case, comps, staged, dtype)
