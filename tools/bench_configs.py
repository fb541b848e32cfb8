#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration (they are parity-test cases, not bench lines; this table goes
to profiles/ for context).  Random policy from HBM (K3 + step launch), CUDA events, median of `--steps` launches.

    python tools/bench_configs.py [--steps 30]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tests.cases import CASES  # noqa: E402
from tests.util import flat_featurizer, make_cuda_env  # noqa: E402

# (label, case, envs, featurizer, algorithmic bytes per env-step from SURVEY.md 8d)
RUNS = [
    ("cfg2 ITG 1v1 walled, step only", "cfg2_itg_1v1_wall", 4096, None, 28),
    ("cfg2 ITG 1v1 walled, step only", "cfg2_itg_1v1_wall", 1 << 20, None, 28),
    ("cfg3 tagging 1v2 J=5, step only", "cfg3_tagging_1v2", 65536, None, 74),
    ("cfg3 tagging 1v2 J=5, step only", "cfg3_tagging_1v2", 1 << 20, None, 74),
    ("cfg4 base 1v4 J=5, step only", "cfg4_base_1v4", 1 << 20, None, 82),
    ("cfg4 base 1v4 J=5 + Global", "cfg4_base_1v4", 65536, "global", 2650),
    ("cfg4 base 1v4 J=5 + Global", "cfg4_base_1v4", 1 << 20, "global", 2650),
    ("cfg4 base 1v4 J=5 + Perspective", "cfg4_base_1v4", 1 << 18, "perspective", 82 + 5 * 2268 + 5 * 40),
    ("cfg4-alt ITG 1v4 + Flat-98", "cfg4alt_itg_1v4", 1 << 20, ["onehot_pos", "alive_crew", "closest_crew"], 453),
    ("cfg5 env side: ITG 1v4 + Flat-98, 131072 envs", "cfg4alt_itg_1v4", 131072, ["onehot_pos", "alive_crew", "closest_crew"], 453),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    rows = []
    for label, case, N, fkind, nbytes in RUNS:
        env = make_cuda_env(CASES[case], N, seed=1234)
        env.emit_next_states = False
        env.reset()
        feat = None
        if fkind == "global":
            feat = S.GlobalFeaturizer(env)
        elif fkind == "perspective":
            feat = S.PerspectiveFeaturizer(env)
        elif fkind is not None:
            feat = flat_featurizer(env, fkind)
        for _ in range(5):
            env.step(env.sample_actions(), featurizer=feat)
        torch.cuda.synchronize()
        evs = []
        for _ in range(a.steps):
            s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            s.record(); acts = env.sample_actions(); m.record(); env.step(acts, featurizer=feat); e.record()
            evs.append((s, m, e))
        torch.cuda.synchronize()
        step_ms = sorted(m.elapsed_time(e) for _, m, e in evs)[len(evs) // 2]
        both_ms = sorted(s.elapsed_time(e) for s, _, e in evs)[len(evs) // 2]
        env.check_actions()
        rows.append({"config": label, "envs": N, "step_kernel_ms": step_ms, "sample+step_ms": both_ms,
                     "env_steps_per_s": N / (both_ms * 1e-3), "algorithmic_bytes_per_env_step": nbytes,
                     "algorithmic_gbs_step_kernel": nbytes * N / (step_ms * 1e-3) / 1e9})
        print(json.dumps(rows[-1]), flush=True)
        del env, feat
    json.dump(rows, open(os.path.join("gpurun_out", "configs_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
