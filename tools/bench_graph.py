#!/usr/bin/env python
"""Launch-bound configurations with a per-step action tensor (i.e. NOT the one-launch random rollout): K x [sample_actions,
step (+ fused encode)] issued call by call against the same K steps captured once in a CUDA graph and replayed
(device-resident ticks, env.device_ticks()).  CUDA events around `reps` repetitions, median.

    python tools/bench_graph.py [--k 64] [--reps 20]
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

RUNS = [
    ("cfg2 ITG 1v1 walled, step only", "cfg2_itg_1v1_wall", 4096, None),
    ("cfg3 tagging 1v2 J=5, step only", "cfg3_tagging_1v2", 65536, None),
    ("cfg4 base 1v4 J=5 + Global", "cfg4_base_1v4", 65536, "global"),
    ("cfg5 env side: ITG 1v4 + Flat-98", "cfg4alt_itg_1v4", 131072, "flat98"),
]


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    return sorted(ms)[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    rows = []
    for label, case, N, fkind in RUNS:
        env = make_cuda_env(CASES[case], N, seed=1234)
        env.emit_next_states = False
        env.reset()
        feat = S.GlobalFeaturizer(env) if fkind == "global" else (
            flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"]) if fkind == "flat98" else None)

        def k_steps():
            for _ in range(a.k):
                env.step(env.sample_actions(), featurizer=feat)

        eager_ms = timed(k_steps, a.reps)
        env.device_ticks(True)
        k_steps()
        side = torch.cuda.Stream(env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side), torch.cuda.graph(graph, stream=side):
            k_steps()
        torch.cuda.current_stream(env.device).wait_stream(side)
        graph_ms = timed(graph.replay, a.reps)
        rows.append({"config": label, "envs": N, "steps_per_graph": a.k, "eager_ms_per_step": eager_ms / a.k,
                     "graph_ms_per_step": graph_ms / a.k, "eager_env_steps_per_s": N * a.k / eager_ms * 1e3,
                     "graph_env_steps_per_s": N * a.k / graph_ms * 1e3, "speedup": eager_ms / graph_ms})
        print(json.dumps(rows[-1]))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/graph_bench.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
