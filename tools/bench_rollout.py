#!/usr/bin/env python
"""Random-policy rollouts in one launch (sus_env_rollout): env-steps/s for the step-only BASELINE configs at their
nominal env counts -- no launch per step, state in registers."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.cases import CASES  # noqa: E402
from tests.util import make_cuda_env  # noqa: E402

rows = []
for label, case, N, T in [("cfg2 ITG 1v1 walled", "cfg2_itg_1v1_wall", 4096, 1000), ("cfg2 ITG 1v1 walled", "cfg2_itg_1v1_wall", 1 << 20, 100),
                          ("cfg3 tagging 1v2 J=5", "cfg3_tagging_1v2", 65536, 500), ("cfg3 tagging 1v2 J=5", "cfg3_tagging_1v2", 1 << 20, 100),
                          ("cfg4 base 1v4 J=5", "cfg4_base_1v4", 1 << 20, 100), ("cfg4-alt ITG 1v4", "cfg4alt_itg_1v4", 131072, 500)]:
    env = make_cuda_env(CASES[case], N, seed=1)
    env.reset()
    env.rollout(10)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.rollout(T); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[2]
    rows.append({"config": label, "envs": N, "steps_per_launch": T, "ms": ms, "env_steps_per_s": N * T / (ms * 1e-3)})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open("gpurun_out/rollout_bench.json", "w"), indent=1)
