#!/usr/bin/env python
"""What fp32 FFMA rate does a library SGEMM reach on this part?  (torch.matmul, TF32 off: cuBLAS SGEMM.)  The practical ceiling
the one-launch MLP kernel (k_mlp_forward, 8 x 8 register micro-tiles, fp32 FFMA) is to be read against."""
import json

import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a @ b; e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    out[f"sgemm_{n}"] = {"ms_best": min(ts), "tflops_best": 2 * n**3 / min(ts) / 1e9, "tflops_median": 2 * n**3 / sorted(ts)[4] / 1e9}
# the MLP's own shapes as plain GEMMs (131072 x 98 x 256 and 131072 x 256 x 128)
for m, k, n in ((131072, 98, 256), (131072, 256, 128)):
    a = torch.randn(m, k, device="cuda"); b = torch.randn(k, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a @ b; e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    out[f"sgemm_{m}x{k}x{n}"] = {"ms_best": min(ts), "tflops_best": 2 * m * k * n / min(ts) / 1e9}
print(json.dumps(out))
