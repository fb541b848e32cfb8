import sys, os, json
sys.path.insert(0, os.getcwd())
from tools.kernel_sweep import time_fused
for n in (1 << 16, 1 << 17, 1 << 20):
    med, best = time_fused(n, 60)
    print(json.dumps({"lib": os.environ.get("SUSNET_B200_LIB", "in-tree"), "envs": n, "median_us": med * 1e3, "best_us": best * 1e3, "frac": 2650 * n / med / 1e6 / 6540.2}))
