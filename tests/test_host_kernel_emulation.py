"""CPU: the index logic of the replay-push kernels, run on the host thread by thread (tests/emu/): the vectorised
k_replay_push_v must store exactly what the one-thread-per-float k_replay_push stores -- and what the reference's bookkeeping
(replay_memory.py:103-143: ring slots, np.roll of the state sequence, restart from T copies of the reset state) says -- for
ragged sizes, every ring position incl. the wrap point inside a vector, T = 1 and T > 1, and unaligned buffers."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    from sus_net_b200 import _lib as L

    d = tmp_path_factory.mktemp("emu")
    src = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_replay.cu")).read()
    body = re.search(r"namespace \{\n(.*)\n\}  // namespace", src, re.S).group(1)
    with open(d / "kernels.inc", "w") as f:
        f.write(body)
    so = str(d / "replay_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "include"), f'-DKERNEL_SOURCE="{d / "kernels.inc"}"',
                    os.path.join(ROOT, "tests", "emu", "replay_emu.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    for fn in (lib.emu_replay_push_v1, lib.emu_replay_push_v2):
        fn.argtypes = [C.POINTER(L.SusReplayPush)]
        fn.restype = None
    return lib


def _off(arr, off):
    """A view of `arr`'s storage that starts `off` elements into a larger buffer (off = 1: 4-byte but not 16-byte aligned)."""
    big = np.zeros(arr.size + off + 8, arr.dtype)
    view = big[off:off + arr.size].reshape(arr.shape)
    view[...] = arr
    return view


def _case(rng, N, M, idx, T, S, A, n_imp, dtype, off):
    seq = rng.integers(0, 9, (N, T, S)).astype(np.float32)
    nf = rng.integers(0, 9, (N, S)).astype(np.float32)
    cf = rng.integers(10, 19, (N, S)).astype(np.float32)
    acts = rng.integers(0, 7, (N, A)).astype(dtype)
    rew = rng.standard_normal((N, A)).astype(np.float32)
    done = (rng.random(N) < 0.2).astype(np.uint8)
    trunc = (rng.random(N) < 0.2).astype(np.uint8)
    imps = rng.integers(0, A, (N, n_imp)).astype(np.int16)
    ins = dict(seq=_off(seq, off), nf=_off(nf, off), cf=_off(cf, off), acts=acts, rew=rew, done=done, trunc=trunc, imps=imps)
    # reference bookkeeping in numpy
    slots = (idx + np.arange(N)) % M
    ring = dict(states=np.full((M, T, S), -1, np.float32), next_states=np.full((M, T, S), -1, np.float32),
                actions=np.full((M, A), -1, np.int64), rewards=np.full((M, A), -1, np.float32),
                dones=np.full((M, 1), 7, np.uint8), imposters=np.full((M, n_imp), -1, np.int16))
    want = {k: v.copy() for k, v in ring.items()}
    nxt = np.roll(seq, -1, axis=1)
    nxt[:, -1] = nf
    want["states"][slots] = seq
    want["next_states"][slots] = nxt
    want["actions"][slots] = acts
    want["rewards"][slots] = rew
    want["dones"][slots, 0] = done
    want["imposters"][slots] = imps
    fin = (done | trunc).astype(bool)
    want_seq = np.where(fin[:, None, None], np.broadcast_to(cf[:, None, :], (N, T, S)), nxt)
    return ins, ring, want, want_seq


def _run(fn, L, ins, ring, N, M, idx, T, S, A, n_imp, dtype, off, idx_dev, in_place=False):
    out = {k: _off(v, off if v.dtype == np.float32 else 0) for k, v in ring.items()}
    if in_place:  # T == 1 only: the sequence is advanced where it stands
        ins = dict(ins, seq=_off(ins["seq"], off))
        seq_out = ins["seq"]
    else:
        seq_out = _off(np.full((N, T, S), -5, np.float32), off)
    code = {np.uint8: L.U8, np.int32: L.I32, np.int64: L.I64}[dtype]
    dev_idx = np.array([idx], np.int64)
    p = L.SusReplayPush(N=N, M=M, idx=0 if idx_dev else idx, T=T, S=S, A=A, n_imposters=n_imp, seq_in=ins["seq"].ctypes.data,
                        seq_out=seq_out.ctypes.data, next_flat=ins["nf"].ctypes.data, cur_flat=ins["cf"].ctypes.data,
                        actions=ins["acts"].ctypes.data, actions_dtype=code, rewards=ins["rew"].ctypes.data,
                        done=ins["done"].ctypes.data, truncated=ins["trunc"].ctypes.data, imposters=ins["imps"].ctypes.data,
                        states=out["states"].ctypes.data, r_actions=out["actions"].ctypes.data,
                        r_rewards=out["rewards"].ctypes.data, next_states=out["next_states"].ctypes.data,
                        r_dones=out["dones"].ctypes.data, r_imposters=out["imposters"].ctypes.data,
                        idx_dev=dev_idx.ctypes.data if idx_dev else None)
    fn(C.byref(p))
    return out, seq_out


SHAPES = [  # N, M, T, S, A, n_imp
    (1, 1, 1, 6, 2, 1), (7, 9, 1, 6, 2, 1), (33, 40, 1, 30, 5, 1), (64, 64, 1, 15, 5, 1), (50, 130, 2, 30, 5, 2),
    (19, 19, 3, 31, 3, 1), (257, 300, 1, 5, 5, 2), (40, 41, 4, 8, 8, 3), (3, 1000, 1, 3, 3, 1), (100, 128, 1, 16, 4, 1),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "N%d_M%d_T%d_S%d_A%d_I%d" % s)
def test_vector_replay_push_equals_scalar_kernel_and_numpy(emu, shape):
    from sus_net_b200 import _lib as L

    N, M, T, S, A, n_imp = shape
    rng = np.random.default_rng(hash(shape) & 0xffff)
    ring_positions = sorted({0, 1, 2, 3, M - 1, max(M - N, 0), max(M - N + 1, 0), M // 2, max(M - N // 2, 0)} & set(range(M)))
    for idx in ring_positions:
        for dtype in (np.int32, np.uint8, np.int64):
            for off in (0, 1):  # 16-byte aligned float buffers / 4-byte aligned only
                for idx_dev in (False, True):
                    ins, ring, want, want_seq = _case(rng, N, M, idx, T, S, A, n_imp, dtype, off)
                    got = {}
                    for name, fn in (("v1", emu.emu_replay_push_v1), ("v2", emu.emu_replay_push_v2)):
                        if name == "v1" and T * S < max(A, n_imp):
                            continue  # the scalar kernel needs a sequence block at least as long as the action row
                        out, seq_out = _run(fn, L, ins, ring, N, M, idx, T, S, A, n_imp, dtype, off, idx_dev)
                        for k in want:
                            assert np.array_equal(out[k], want[k]), (name, k, idx, dtype, off)
                        assert np.array_equal(seq_out, want_seq), (name, "seq_out", idx, dtype, off)
                        got[name] = out
                        if T == 1:
                            out, seq_out = _run(fn, L, ins, ring, N, M, idx, T, S, A, n_imp, dtype, off, idx_dev, in_place=True)
                            assert all(np.array_equal(out[k], want[k]) for k in want) and np.array_equal(seq_out, want_seq)


def test_mlp_weight_packer_writes_the_chunk_images_the_forward_kernel_copies(tmp_path):
    """k_mlp_pack on the host: for every layer the image is [column block][16-k chunk][kk][column] with zeros outside the
    weight matrix -- exactly the [kk][column] tiles the forward kernel's staging buffer holds per chunk (susnet_mlp.cu, `put`)."""
    from sus_net_b200 import _lib as L

    src = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_mlp.cu")).read()
    pieces = [re.search(r"constexpr int kKc = 16;[^\n]*\n", src).group(0),
              re.search(r"struct MlpParams \{.*?\n\};\n", src, re.S).group(0),
              re.search(r"__host__ __device__ inline int mlp_ct[^\n]*\n", src).group(0),
              re.search(r"__host__ __device__ inline int64_t mlp_packed_floats.*?\n\}\n", src, re.S).group(0),
              re.search(r"__global__ void __launch_bounds__\(256\) k_mlp_pack.*?\n\}\n", src, re.S).group(0)]
    inc = tmp_path / "mlp_pack.inc"
    inc.write_text("\n".join(pieces))
    so = str(tmp_path / "mlp_pack_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "include"), f'-DKERNEL_SOURCE="{inc}"',
                    os.path.join(ROOT, "tests", "emu", "mlp_pack_emu.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.emu_mlp_pack.argtypes = [C.POINTER(L.SusMlpSpec), C.c_void_p]
    lib.emu_mlp_pack.restype = None
    rng = np.random.default_rng(3)
    for dims in ([98, 256, 128, 64, 16, 6], [4, 6], [200, 150, 17, 129, 3], [5, 65, 1]):
        ws = [rng.standard_normal((m, k)).astype(np.float32) for k, m in zip(dims[:-1], dims[1:])]
        spec = L.SusMlpSpec(n_layers=len(ws), activation=L.ACT_RELU)
        for i, d in enumerate(dims):
            spec.dims[i] = d
        for l, w in enumerate(ws):
            spec.weight[l] = w.ctypes.data
        want = []
        for w in ws:
            m, k = w.shape
            cb = 16 * (8 if m > 64 else 4 if m > 16 else 1)
            n_blocks, n_chunks = -(-m // cb), -(-k // 16)
            padded = np.zeros((n_blocks * cb, n_chunks * 16), np.float32)
            padded[:m, :k] = w
            # [block][col][chunk][kk] -> [block][chunk][kk][col]
            want.append(padded.reshape(n_blocks, cb, n_chunks, 16).transpose(0, 2, 3, 1).ravel())
        want = np.concatenate(want)
        got = np.full(want.size, np.nan, np.float32)
        lib.emu_mlp_pack(C.byref(spec), got.ctypes.data)
        assert np.array_equal(got, want), dims


# ---------------------------------------------------------------------------------------------------------------------------
# The DEVICE code of reset / step / flatten / feature rows, compiled for the host (tests/emu/step_emu.cpp) and held against the
# oracle: the CPU suite then covers the very functions the kernels run -- reset_env, step_env, agent_reward, finish_one's
# auto-reset and statistics, write_flat, flat_component (float rows, byte-staged rows, compile-time agent count), the Global /
# Perspective rows and plane offsets -- not only their restatement.
@pytest.fixture(scope="module")
def step_emu(tmp_path_factory):
    from sus_net_b200 import _lib as L

    d = tmp_path_factory.mktemp("step_emu")
    src = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_api.cu")).read()
    pieces = [re.search(r"struct StepParams \{.*?\n\};\n", src, re.S).group(0),
              src[src.index("__device__ __forceinline__ int reward_row_bytes"):src.index("// K1 (+K2), direct-store path")],
              src[src.index("// unflatten one row (gymnasium.spaces.unflatten"):src.index("template <typename T>\n__global__ void __launch_bounds__(kThreads) k_encode_rows")],
              src[src.index("int flat_size(const SusConfig& c) {"):src.index("inline int32_t align128")]]
    inc = d / "step_glue.inc"
    inc.write_text("\n".join(pieces))
    so = str(d / "step_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "sus_net_b200", "csrc"),
                    f'-DKERNEL_SOURCE="{inc}"', os.path.join(ROOT, "tests", "emu", "step_emu.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.emu_reset.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp, vp, vp]
    lib.emu_step.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.emu_step_compact.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.emu_export_flat.argtypes = [C.POINTER(L.SusConfig), vp, vp, vp, vp, vp]
    lib.emu_encode.argtypes = [C.POINTER(L.SusConfig), C.POINTER(L.SusEncodeSpec), C.c_int, vp, vp, vp, vp, vp, vp]
    lib.emu_encode_rows.argtypes = [C.POINTER(L.SusConfig), C.POINTER(L.SusEncodeSpec), C.c_int, vp, C.c_int64, vp, vp]
    return lib


class _EmuEnv:
    """Structure-of-arrays env state in numpy + the emulated device functions (what BatchedFourRoomEnv does with kernels)."""

    def __init__(self, lib, cfg, N, seed, env_id_base=0):
        import oracle
        from sus_net_b200 import _lib as L

        self.lib, self.N, self.A = lib, N, cfg["n_imposters"] + cfg["n_crew"]
        kw = {k: (oracle.VARIANT_IDS[v] if k == "variant" else (int(v) if isinstance(v, bool) else v)) for k, v in cfg.items()}
        self.cfg = L.SusConfig(num_envs=N, seed=seed, env_id_base=env_id_base, auto_reset=1, **kw)
        self.S = oracle.flat_size(cfg)
        self.pos, self.jobpos = np.zeros(N, np.uint64), np.zeros(N, np.uint64)
        self.aux, self.met = np.zeros((N, 4), np.uint32), np.zeros((N, 4), np.uint32)
        self.stats = np.zeros(10, np.uint64)
        self.err = np.zeros(1, np.uint32)
        self.step_tick = self.reset_epoch = 0

    def _state(self):
        return [a.ctypes.data for a in (self.pos, self.jobpos, self.aux, self.met)]

    def reset(self):
        self.lib.emu_reset(C.byref(self.cfg), self.reset_epoch, *self._state())
        self.reset_epoch += 1
        return self.flat_states()

    def imposter_mask(self):
        return ((self.aux[:, 0:1] >> 8 >> np.arange(self.A, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)

    def flat_states(self):
        out = np.zeros((self.N, self.S), np.int64)
        self.lib.emu_export_flat(C.byref(self.cfg), *self._state(), out.ctypes.data)
        return out

    def step(self, actions=None):
        a = None if actions is None else np.ascontiguousarray(actions, np.int32)
        r = np.zeros((self.N, self.A), np.float64)
        done, trunc = np.zeros(self.N, np.uint8), np.zeros(self.N, np.uint8)
        nf = np.zeros((self.N, self.S), np.float32)
        a_out = np.zeros((self.N, self.A), np.int32)
        self.lib.emu_step(C.byref(self.cfg), self.step_tick, *self._state(), None if a is None else a.ctypes.data, r.ctypes.data,
                          done.ctypes.data, trunc.ctypes.data, nf.ctypes.data, a_out.ctypes.data, self.stats.ctypes.data,
                          self.err.ctypes.data)
        self.step_tick += 1
        return dict(rewards=r, done=done, trunc=trunc, next_flat=nf.astype(np.int64), actions=a_out)

    def step_compact(self, proto, actions):
        """One step through the compact host protocol: the library's host packer, the device code, the library's host decoder."""
        packed = proto.pack_actions(np.ascontiguousarray(actions, np.int32))
        rec = np.zeros((self.N, proto.result_bytes), np.uint8)
        self.lib.emu_step_compact(C.byref(self.cfg), self.step_tick, *self._state(), packed.ctypes.data, rec.ctypes.data,
                                  self.stats.ctypes.data, self.err.ctypes.data)
        self.step_tick += 1
        return proto.decode(rec)

    def encode(self, kind, components=(), mode=0, rows=None):
        """Feature tensors of the live envs, or (rows = (n, S) int64 flattened states) of replay rows like fit()."""
        import oracle
        from sus_net_b200 import _lib as L

        spec = L.SusEncodeSpec(kind=kind, n_components=len(components))
        for i, name in enumerate(components):
            spec.components[i] = oracle.FLAT_COMPONENTS[name]
        shape = L.SusEncodeShape()
        assert L.lib().sus_encode_shape(C.byref(self.cfg), C.byref(spec), C.byref(shape)) == 0
        n = self.N if rows is None else rows.shape[0]
        sp = np.full((max(shape.spatial_views, 1), n, max(shape.spatial_floats, 1)), np.nan, np.float32)
        ns = np.full((shape.non_spatial_views, n, shape.non_spatial_floats), np.nan, np.float32)
        if rows is None:
            rc = self.lib.emu_encode(C.byref(self.cfg), C.byref(spec), mode, *self._state(), sp.ctypes.data, ns.ctypes.data)
        else:
            rows = np.ascontiguousarray(rows, np.int64)
            rc = self.lib.emu_encode_rows(C.byref(self.cfg), C.byref(spec), mode, rows.ctypes.data, n, sp.ctypes.data, ns.ctypes.data)
        assert rc == 0
        return sp, ns


def _all_cases():
    from tests.cases import CASES, EDGE_CASES

    return {**CASES, **EDGE_CASES}


def _case_ids():
    return sorted(_all_cases())


@pytest.mark.parametrize("case", _case_ids())
def test_device_step_code_on_the_host_equals_the_oracle(step_emu, case):
    """The 12 canonical and 8 edge configurations (1-step episodes, 1-step vote windows, 8 agents x 8 jobs, all-zero rewards incl.
    -0.0): reset + 220 fused random-policy steps (every 7th with explicit sampled actions) of 96 envs: flat states, float64 reward
    bit patterns, done / truncated, the actions the fused policy drew, auto-reset states and the episode statistics."""
    import oracle

    cfg = _all_cases()[case]
    N, seed = 96, 11
    emu, orc = _EmuEnv(step_emu, cfg, N, seed), oracle.OracleEnv(cfg, N, seed=seed)
    assert np.array_equal(emu.reset(), orc.reset())
    for t in range(220):
        acts = orc.sample_actions() if t % 7 == 3 else None
        want, got = orc.step(acts), emu.step(acts)
        assert np.array_equal(got["next_flat"], want["next_flat"]), (case, t)
        assert np.array_equal(got["rewards"].view(np.int64), want["rewards"].view(np.int64)), (case, t)
        assert np.array_equal(got["done"], want["done"]) and np.array_equal(got["trunc"], want["trunc"]), (case, t)
        assert np.array_equal(got["actions"], want["actions"]), (case, t)
        assert np.array_equal(emu.flat_states(), orc.flat_states()), (case, t)
    assert np.array_equal(emu.stats.astype(np.int64), orc.stats()) and int(emu.err[0]) == 0
    # (random play rarely ends an ImposterTrainingGround episode with 3-4 crew inside 220 steps; every other case finishes some)
    assert int(emu.stats[0]) > 0 or cfg["variant"] == "training_ground"


def test_device_feature_rows_on_the_host_equal_the_oracle(step_emu):
    """Flat rows (float rows, byte-staged rows, byte-staged rows built with the compile-time agent count), Global and Perspective
    non-spatial rows and planes, on states taken along oracle trajectories."""
    import oracle
    from sus_net_b200 import _lib as L
    from tests.cases import CASES, FLAT_COMPONENT_SETS, FLAT_COMPONENT_SETS_TAGGING, GLOBAL_CASES

    def walk(case):
        cfg = CASES[case]
        emu = _EmuEnv(step_emu, cfg, 64, 5)
        emu.reset()
        for t in range(60):
            emu.step(None)
            if t % 20 == 19:
                yield cfg, emu, emu.flat_states()

    for case, sets in {**FLAT_COMPONENT_SETS, **FLAT_COMPONENT_SETS_TAGGING}.items():
        for cfg, emu, flat in walk(case):
            for comps in sets:
                want = oracle.encode_flat(cfg, comps, flat)
                for mode in ((0,) if "scent" in comps else (0, 1, 2)):  # the float-valued scent row is never byte-staged
                    got = emu.encode(L.ENCODE_FLAT, comps, mode)[1][0]
                    assert np.array_equal(got, want), (case, comps, mode)
    for case in GLOBAL_CASES:
        for cfg, emu, flat in walk(case):
            A = cfg["n_imposters"] + cfg["n_crew"]
            sp, ns = oracle.encode_global(cfg, flat)
            got_sp, got_ns = emu.encode(L.ENCODE_GLOBAL)
            assert np.array_equal(got_sp[0].reshape(-1, A + 2, 9, 9), sp) and np.array_equal(got_ns, ns), case
            sp, ns = oracle.encode_perspective(cfg, flat)
            got_sp, got_ns = emu.encode(L.ENCODE_PERSPECTIVE)
            assert np.array_equal(got_sp.reshape(A, -1, A + 2, 9, 9), sp) and np.array_equal(got_ns, ns), case


def _philox_fixtures():
    from tests.util import golden_files

    return golden_files("philox")


@pytest.mark.parametrize("path", _philox_fixtures(), ids=lambda p: os.path.basename(p).split(".")[0])
def test_device_step_code_on_the_host_equals_the_reference_fixtures(step_emu, path):
    """The same device code against outputs of the UNMODIFIED reference (tests/golden/*.philox.npz, minted by
    tools/make_golden.py on the Philox draws): reset states and roles, then per step the state, the float64 reward bit patterns,
    done / truncated, the post-auto-reset state and roles."""
    from tests.cases import CASES
    from tests.util import case_of, load, reward_bits

    g = load(path)
    cfg = CASES[case_of(path)]
    T, N, A = g["actions"].shape
    emu = _EmuEnv(step_emu, cfg, N, int(g["seed"]), env_id_base=int(g["env_id_base"]))
    assert np.array_equal(emu.reset(), g["reset_flat"])
    assert np.array_equal(emu.imposter_mask(), g["reset_imp"].astype(np.uint8))
    for t in range(T):
        o = emu.step(g["actions"][t])
        assert np.array_equal(o["next_flat"], g["next_flat"][t]), f"state differs at step {t}"
        assert np.array_equal(reward_bits(o["rewards"]), reward_bits(g["rewards"][t])), f"rewards differ at step {t}"
        assert np.array_equal(o["done"], g["done"][t]) and np.array_equal(o["trunc"], g["trunc"][t])
        assert np.array_equal(emu.flat_states(), g["cur_flat"][t]), f"post-reset state differs at step {t}"
        assert np.array_equal(emu.imposter_mask(), g["imp"][t].astype(np.uint8))
    fin = (g["done"] | g["trunc"]) != 0
    assert int(emu.stats[0]) == fin.sum() and int(emu.stats[9]) == (g["trunc"] != 0).sum() and int(emu.err[0]) == 0


def _feature_fixtures():
    from tests.util import golden_files

    return golden_files("features")


@pytest.mark.parametrize("path", _feature_fixtures(), ids=lambda p: os.path.basename(p).split(".")[0])
def test_device_feature_rows_on_the_host_equal_the_reference_fixtures(step_emu, path):
    """The device featurizer code on replay rows (parse_row, the fit() path) against the UNMODIFIED reference's featurizers
    (tests/golden/*.features.npz): Global and Perspective planes and rows, every flat component list incl. the float scent one
    (bit patterns), byte-staged rows where the list is integer-valued."""
    from sus_net_b200 import _lib as L
    from tests.cases import CASES
    from tests.util import case_of, load

    g = load(path)
    cfg = CASES[case_of(path)]
    flat = g["flat"].astype(np.int64)
    A = cfg["n_imposters"] + cfg["n_crew"]
    emu = _EmuEnv(step_emu, cfg, 1, 0)
    if "global_spatial" in g:
        sp, ns = emu.encode(L.ENCODE_GLOBAL, rows=flat)
        assert np.array_equal(sp[0].reshape(-1, A + 2, 9, 9), g["global_spatial"].astype(np.float32))
        assert np.array_equal(ns, g["global_non_spatial"])
        sp, ns = emu.encode(L.ENCODE_PERSPECTIVE, rows=flat)
        assert np.array_equal(sp.reshape(A, -1, A + 2, 9, 9), g["perspective_spatial"].astype(np.float32))
        assert np.array_equal(ns, g["perspective_non_spatial"])
    i = 0
    while f"flat{i}" in g:
        comps = [str(c) for c in g[f"flat{i}_components"]]
        for mode in ((0,) if "scent" in comps else (0, 1, 2)):
            out = emu.encode(L.ENCODE_FLAT, comps, mode, rows=flat)[1][0]
            assert np.array_equal(out.view(np.int32), g[f"flat{i}"].view(np.int32)), (comps, mode)
        i += 1


def test_device_step_code_on_the_host_equals_the_oracle_on_random_configurations(step_emu):
    """24 random constructor-argument sets (every variant, up to 8 agents / 8 jobs, non-integer reward constants, short episodes
    and vote windows): 150 steps of 40 envs each, everything the step returns and the episode statistics."""
    import oracle
    from tests.cases import random_case

    rng = np.random.default_rng(2026)
    episodes = 0
    for k in range(24):
        cfg = random_case(rng)
        emu, orc = _EmuEnv(step_emu, cfg, 40, 100 + k, env_id_base=7 * k), oracle.OracleEnv(cfg, 40, seed=100 + k, env_id_base=7 * k)
        assert np.array_equal(emu.reset(), orc.reset()), cfg
        for t in range(150):
            acts = orc.sample_actions() if t % 5 == 2 else None
            want, got = orc.step(acts), emu.step(acts)
            assert np.array_equal(got["next_flat"], want["next_flat"]), (cfg, t)
            assert np.array_equal(got["rewards"].view(np.int64), want["rewards"].view(np.int64)), (cfg, t)
            assert np.array_equal(got["done"], want["done"]) and np.array_equal(got["trunc"], want["trunc"]), (cfg, t)
            assert np.array_equal(emu.flat_states(), orc.flat_states()), (cfg, t)
        assert np.array_equal(emu.stats.astype(np.int64), orc.stats()), cfg
        episodes += int(emu.stats[0])
    assert episodes > 1000


@pytest.mark.parametrize("case", _case_ids())
def test_device_step_code_through_the_compact_host_protocol_equals_the_oracle(step_emu, case):
    """Bit-packed action records in, reward codes + done / truncated bits out (SUS_PACKED / SusStepIO.packed_out): packed by
    sus_host_pack_actions, stepped by the device code, decoded by sus_host_decode_results -- the float64 rewards must carry the
    oracle's bit patterns (incl. -0.0 and non-integer constants), an env whose action is rejected reports NaN and stays put."""
    import types

    import oracle
    from sus_net_b200 import _lib as L
    from sus_net_b200.compact import CompactProtocol

    cfg = _all_cases()[case]
    N, seed = 64, 3
    emu, orc = _EmuEnv(step_emu, cfg, N, seed), oracle.OracleEnv(cfg, N, seed=seed)
    proto = CompactProtocol(types.SimpleNamespace(lib=L.lib(), _cfg=emu.cfg, n_agents=emu.A))
    assert np.array_equal(emu.reset(), orc.reset())
    for t in range(120):
        acts = orc.sample_actions()
        want = orc.step(acts)
        r, d, tr = emu.step_compact(proto, acts)
        assert np.array_equal(r.view(np.int64), want["rewards"].view(np.int64)), (case, t)
        assert np.array_equal(d, want["done"] != 0) and np.array_equal(tr, want["trunc"] != 0), (case, t)
        assert np.array_equal(emu.flat_states(), orc.flat_states()), (case, t)
    assert np.array_equal(emu.stats.astype(np.int64), orc.stats())
    # a role-list index past the end of the crew list (the imposter list is the longer one): rejected, counted, NaN, state kept
    before = emu.flat_states()
    bad = orc.sample_actions()
    crew = np.argwhere(orc.imposter_mask()[0] == 0)[0, 0]
    bad[0, crew] = oracle.n_role_actions(cfg, False)
    r, d, tr = emu.step_compact(proto, bad)
    assert np.isnan(r[0]).all() and not d[0] and not tr[0] and int(emu.err[0]) == 1
    assert np.array_equal(emu.flat_states()[0], before[0])


@pytest.fixture(scope="module")
def policy_emu(tmp_path_factory):
    from sus_net_b200 import _lib as L

    d = tmp_path_factory.mktemp("policy_emu")
    pol = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_policy.cu")).read()
    api = open(os.path.join(ROOT, "sus_net_b200", "csrc", "susnet_api.cu")).read()
    pieces = [pol[pol.index("constexpr uint32_t P_POLICY"):pol.index("}  // namespace")],
              api[api.index("int flat_size(const SusConfig& c) {"):api.index("int component_size(const SusConfig& c, int comp)")]]
    inc = d / "policy.inc"
    inc.write_text("\n".join(pieces))
    so = str(d / "policy_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "sus_net_b200", "csrc"),
                    f'-DKERNEL_SOURCE="{inc}"', os.path.join(ROOT, "tests", "emu", "policy_emu.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.emu_select_actions.argtypes = [C.POINTER(L.SusConfig), C.c_uint64, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp]
    lib.emu_seq_roll.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int]
    return lib


@pytest.mark.parametrize("case", ["cfg4_base_1v4", "cfg4alt_itg_1v4", "cfg3_tagging_1v2", "base_2v3_j3", "tagging_2v5_short"])
def test_selection_kernel_on_the_host_equals_the_numpy_restatement(step_emu, policy_emu, case):
    """k_select_actions (the acting part of train(), train.py:349-381): epsilon-greedy with Philox draws, role-aware ranges, first
    argmax on ties, dead agents keep 0, both imposter Q layouts, eps as a scalar or from memory, int32 / uint8 actions."""
    import oracle
    from sus_net_b200 import _lib as L

    cfg = _all_cases()[case]
    N, seed, base = 300, 17, 555
    env = _EmuEnv(step_emu, cfg, N, seed, env_id_base=base)
    env.reset()
    for _ in range(25):
        env.step(None)  # some agents die, some episodes restart
    A, nI = env.A, cfg["n_imposters"]
    nia, nca = oracle.n_role_actions(cfg, True), oracle.n_role_actions(cfg, False)
    rng = np.random.default_rng(3)
    ids = np.arange(base, base + N)
    flat = env.flat_states()
    for trial, (eps, use_imp, use_crew, per_view, dtype) in enumerate([
            (0.0, True, True, nI != 1, np.int32), (0.3, True, True, True, np.uint8), (1.0, True, True, nI != 1, np.int32),
            (0.2, True, False, nI != 1, np.int32), (0.0, False, True, True, np.uint8), (0.5, False, False, True, np.int32)]):
        q_imp = rng.standard_normal((A, N, nia) if per_view else (N, nia)).astype(np.float32) if use_imp else None
        q_crew = rng.standard_normal((A, N, nca)).astype(np.float32) if use_crew else None
        if q_imp is not None:
            q_imp[..., 3] = q_imp[..., 1]  # ties: the FIRST maximum wins (torch.argmax, train.py:368-370)
        eps_arr = np.array([eps], np.float32)
        out = np.full((N, A), 77, dtype)
        policy_emu.emu_select_actions(C.byref(env.cfg), trial, env.aux.ctypes.data, None if q_imp is None else q_imp.ctypes.data,
                                      None if q_crew is None else q_crew.ctypes.data, eps_arr.ctypes.data if trial % 2 else None,
                                      eps, int(per_view), L.U8 if dtype == np.uint8 else L.I32, out.ctypes.data)
        want = oracle.select_actions(cfg, seed, ids, trial, flat[:, 2 * A:3 * A], env.imposter_mask(), q_imp, q_crew, eps,
                                     imposter_per_view=per_view)
        assert np.array_equal(out.astype(np.int32), want), (case, trial)


def test_feature_sequence_roll_on_the_host_equals_numpy(policy_emu):
    """k_seq_roll (np.roll of T-deep ENCODED feature sequences, restart from T copies where the episode ended,
    train.py:388-389,440-445): the 128-bit and the scalar variant against numpy, view-major rows."""
    rng = np.random.default_rng(0)
    for views, n_envs, T, R in ((1, 37, 1, 8), (5, 40, 3, 16), (2, 19, 4, 7), (3, 64, 2, 567)):
        rows = views * n_envs
        seq = rng.standard_normal((rows, T, R)).astype(np.float32)
        newest = rng.standard_normal((rows, R)).astype(np.float32)
        done = (rng.random(n_envs) < 0.3).astype(np.uint8)
        trunc = (rng.random(n_envs) < 0.2).astype(np.uint8)
        fin = np.tile((done | trunc).astype(bool), views)
        want = np.roll(seq, -1, axis=1)
        want[:, -1] = newest
        want[fin] = newest[fin][:, None, :]
        for vec in ((0, 1) if R % 4 == 0 else (0,)):
            out = np.full_like(seq, np.nan)
            policy_emu.emu_seq_roll(seq.ctypes.data, out.ctypes.data, newest.ctypes.data, done.ctypes.data, trunc.ctypes.data, rows,
                                    n_envs, T, R, vec)
            assert np.array_equal(out, want), (views, n_envs, T, R, vec)
