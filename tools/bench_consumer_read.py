#!/usr/bin/env python
"""What the consumer of the feature tensors sees: reductions / a first conv layer reading the Global planes of 1 Mi envs,
with the planes in L2-compressible memory (default) and in ordinary device memory (SUSNET_COMPRESSIBLE=0).

    python tools/bench_consumer_read.py; SUSNET_COMPRESSIBLE=0 python tools/bench_consumer_read.py
"""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import sus_net_b200 as S
from sus_net_b200.memory import is_compressible
N = 1 << 20
env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=1, device="cuda:0")
feat = S.GlobalFeaturizer(env); env.reset()
for _ in range(3): env.step(None, featurizer=feat)
sp = feat.generate_featurized_states()[0][0].detach()   # (N, 1, 7, 9, 9)
w = torch.randn(16, 7, 3, 3, device="cuda:0")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
x = sp[:, 0]
out = {"compressible": is_compressible(feat._sp_buf), "bytes": x.numel()*4,
       "sum_ms": timed(lambda: x.sum()), "amax_ms": timed(lambda: x.amax()),
       "conv3x3_16ch_first_256k_ms": timed(lambda: torch.nn.functional.conv2d(x[:262144], w, padding=1))}
out["sum_gbs"] = out["bytes"]/out["sum_ms"]/1e6
print(json.dumps(out))
