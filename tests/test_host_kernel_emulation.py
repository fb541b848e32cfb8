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
