#!/usr/bin/env python
"""The acting loop's Q-network evaluation (cfg5: MLP [98, 256, 128, 64, 16, 6] on 131 072 rows): torch module (cuBLAS SGEMMs +
elementwise PReLU passes) against the one-launch inference kernel `sus_mlp_forward` (CUDA events, median of --reps).

    python tools/bench_mlp.py [--rows 131072] [--dims 98,256,128,64,16,6]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tools.train_demo import MLPQ  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=131072)
    ap.add_argument("--dims", default="98,256,128,64,16,6")
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    dims = [int(x) for x in a.dims.split(",")]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = MLPQ(dims).to(dev)
    x = (torch.rand(a.rows, 1, dims[0], device=dev) < 0.15).float()
    sp = torch.zeros(a.rows, 1, 1, device=dev)
    f = S.FusedMLP(m)
    flops = 2.0 * a.rows * sum(p * q for p, q in zip(dims[:-1], dims[1:]))

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            ts.append((s, e))
        torch.cuda.synchronize()
        ms = sorted(s.elapsed_time(e) for s, e in ts)
        return ms[len(ms) // 2]

    out = {"rows": a.rows, "dims": dims, "gflop": flops / 1e9}
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        t = timed(lambda: m(sp, x))
        out["torch_fp32"] = {"ms": t, "tflops": flops / t / 1e9}
        torch.backends.cuda.matmul.allow_tf32 = True
        t = timed(lambda: m(sp, x))
        out["torch_tf32"] = {"ms": t, "tflops": flops / t / 1e9}
        torch.backends.cuda.matmul.allow_tf32 = False
    os.environ["SUSNET_MLP_VERBOSE"] = "1"  # every geometry reports its threads and CTAs per SM on stderr
    # rows per tile (128: one CTA per SM, 64: two), k-parts per CTA, weights staged from the repacked workspace image or not
    for rows, split, packed in (("128", "1", "1"), ("128", "1", "0"), ("128", "2", "0"), ("64", "1", "1"), ("64", "1", "0"), ("64", "2", "0")):
        os.environ["SUSNET_MLP_ROWS"], os.environ["SUSNET_MLP_SPLIT"], os.environ["SUSNET_MLP_PACKED"] = rows, split, packed
        t = timed(lambda: f(sp, x))
        err = float(((f(sp, x) - m(sp, x)).abs().max()).item())
        out[f"sus_mlp_forward_fp32_rows{rows}_split{split}_packed{packed}"] = {"ms": t, "tflops": flops / t / 1e9, "max_abs_diff_vs_torch_fp32": err}
    for k in ("SUSNET_MLP_ROWS", "SUSNET_MLP_SPLIT", "SUSNET_MLP_PACKED"):
        os.environ.pop(k)
    t = timed(lambda: f(sp, x))
    out["sus_mlp_forward_fp32_default"] = {"ms": t, "tflops": flops / t / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
