#!/usr/bin/env python
"""Headline kernel (cfg4 + Global, fused step + encode) at 65 536 / 131 072 / 1 Mi envs with the library SUSNET_B200_LIB names
(default: the in-tree build): one line per env count, for A/B runs of two builds in one GPU session.

    SUSNET_B200_LIB=_ab/libother.so python tools/micro/small_ab.py; python tools/micro/small_ab.py
"""
import sys, os, json
sys.path.insert(0, os.getcwd())
from tools.kernel_sweep import time_fused
for n in (1 << 16, 1 << 17, 1 << 20):
    med, best = time_fused(n, 60)
    print(json.dumps({"lib": os.environ.get("SUSNET_B200_LIB", "in-tree"), "envs": n, "median_us": med * 1e3, "best_us": best * 1e3, "frac": 2650 * n / med / 1e6 / 6540.2}))
