"""In-tree build of the CUDA library (sm_100a only).  `python -m sus_net_b200.build` or __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsusnet_b200.so")
SOURCES = ["susnet_api.cu", "susnet_replay.cu", "susnet_alloc.cu", "susnet_policy.cu", "susnet_mlp.cu", "susnet_host.cu"]
HEADERS = ["susnet_device.cuh", "susnet_encode.cuh", "susnet_tile.cuh", "susnet_ws.cuh", os.path.join("..", "..", "include", "susnet_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


HASH_FILE = LIB + ".buildhash"


def _source_hash():
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for rel in SOURCES + HEADERS:
        with open(os.path.join(CSRC, rel), "rb") as f:
            h.update(rel.encode())
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    """Content hash of the sources + flags, not mtimes: a copied tree (gpurun snapshot) must not trigger a rebuild."""
    if not os.path.exists(LIB) or not os.path.exists(HASH_FILE):
        return True
    with open(HASH_FILE) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    with open(HASH_FILE, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
