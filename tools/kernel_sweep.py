#!/usr/bin/env python
"""GPU tuning helper: times the fused step+encode kernel (headline config) under the layout overrides of
make_layout() (SUSNET_TILE_G / SUSNET_TILE_WARPS) and measures the write-only HBM ceiling (memset) beside it.

    python tools/kernel_sweep.py [--envs 1048576] [--steps 30]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402


def time_fused(N, steps, policy_fused=False):
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=1234, device="cuda:0")
    env.emit_next_states = False
    feat = S.GlobalFeaturizer(env)
    env.reset()
    for _ in range(5):
        env.step(None if policy_fused else env.sample_actions(), featurizer=feat)
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        a = None if policy_fused else env.sample_actions()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        env.step(a, featurizer=feat)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    N = a.envs
    out = {}
    # write-only ceiling: memset of the same number of bytes the kernel writes
    buf = torch.empty(N * 2568 // 4, dtype=torch.float32, device="cuda:0")
    for _ in range(3):
        buf.zero_()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); buf.zero_(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    out["memset_gbs_best"] = buf.numel() * 4 / (min(ts) * 1e-3) / 1e9
    out["memset_gbs_median"] = buf.numel() * 4 / (sorted(ts)[5] * 1e-3) / 1e9
    del buf
    os.environ["SUSNET_PATH"] = "tma"
    for g, w in [(8, 8), (4, 8)]:
        os.environ["SUSNET_TILE_G"], os.environ["SUSNET_TILE_WARPS"] = str(g), str(w)
        med, best = time_fused(N, a.steps)
        out[f"tma_G{g}_W{w}"] = {"median_ms": med, "best_ms": best, "gbs_algorithmic": 2650 * N / (med * 1e-3) / 1e9}
    os.environ.pop("SUSNET_TILE_G"); os.environ.pop("SUSNET_TILE_WARPS")
    os.environ["SUSNET_PATH"] = "direct"
    med, best = time_fused(N, a.steps)
    out["direct"] = {"median_ms": med, "best_ms": best, "gbs_algorithmic": 2650 * N / (med * 1e-3) / 1e9}
    os.environ["SUSNET_PATH"] = "ws"
    for te in (16, 8):
        os.environ["SUSNET_WS_TILE"] = str(te)
        for cw in (8, 7, 6, 5, 4):
            os.environ["SUSNET_WS_WARPS"] = str(cw)
            med, best = time_fused(N, a.steps)
            out[f"ws_T{te}_CW{cw}"] = {"median_ms": med, "best_ms": best, "gbs_algorithmic": 2650 * N / (med * 1e-3) / 1e9}
    os.environ.pop("SUSNET_WS_TILE")
    os.environ.pop("SUSNET_WS_WARPS")
    os.environ.pop("SUSNET_PATH")
    med, best = time_fused(N, a.steps, policy_fused=True)
    out["default_fused_policy"] = {"median_ms": med, "best_ms": best, "gbs_algorithmic": 2650 * N / (med * 1e-3) / 1e9}
    for n in (1 << 16, 1 << 18, 1 << 22):
        med, best = time_fused(n, a.steps)
        out[f"default_N{n}"] = {"median_ms": med, "gbs_algorithmic": 2650 * n / (med * 1e-3) / 1e9, "env_steps_per_s": n / (med * 1e-3)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
