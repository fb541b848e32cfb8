#!/usr/bin/env python
"""Small batches of the headline kernel (cfg4 + Global, fused step + encode): kernel time at 2^16 and 2^17 envs under the
warp-specialised kernel's geometry overrides (SUSNET_WS_TILE = envs per plane tile, SUSNET_WS_WARPS = compute warps), to see
whether a geometry other than the 1 Mi-env optimum (8-env tiles, 7 compute warps) shortens the fixed cost of a launch.

    python tools/small_batch_sweep.py [--steps 40]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kernel_sweep import time_fused  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    rows = []
    for n in (1 << 16, 1 << 17):
        med, best = time_fused(n, a.steps)
        rows.append({"envs": n, "geometry": "default", "median_ms": med, "best_ms": best, "gbs_algorithmic": 2650 * n / med / 1e6})
        for te in (8, 16):
            for cw in (4, 5, 6, 7, 8):  # 8 is what fits beside two 8-env tiles at cfg4
                os.environ["SUSNET_WS_TILE"], os.environ["SUSNET_WS_WARPS"] = str(te), str(cw)
                try:
                    med, best = time_fused(n, a.steps)
                    rows.append({"envs": n, "geometry": f"T{te}_CW{cw}", "median_ms": med, "best_ms": best,
                                 "gbs_algorithmic": 2650 * n / med / 1e6})
                finally:
                    os.environ.pop("SUSNET_WS_TILE"); os.environ.pop("SUSNET_WS_WARPS")
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
