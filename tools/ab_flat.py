#!/usr/bin/env python
"""A/B timing of one fused step+encode configuration (run once per library build, SUSNET_B200_LIB selects it).

    SUSNET_B200_LIB=_ab/libhead.so python tools/ab_flat.py [--case cfg4alt_itg_1v4] [--envs 1048576] [--steps 200]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="cfg4alt_itg_1v4")
    ap.add_argument("--feat", default="onehot_pos,alive_crew,closest_crew")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    env = make_cuda_env(CASES[a.case], a.envs, seed=1234)
    env.emit_next_states = False
    env.reset()
    if a.feat == "global":
        feat = S.GlobalFeaturizer(env)
    elif a.feat == "perspective":
        feat = S.PerspectiveFeaturizer(env)
    elif a.feat == "none":
        feat = None
    else:
        feat = flat_featurizer(env, a.feat.split(","))
    for _ in range(10):
        env.step(env.sample_actions(), featurizer=feat)
    torch.cuda.synchronize()
    evs = []
    for _ in range(a.steps):
        acts = env.sample_actions()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.step(acts, featurizer=feat); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    print(json.dumps({"lib": os.environ.get("SUSNET_B200_LIB", "in-tree"), "case": a.case, "feat": a.feat, "envs": a.envs,
                      "median_ms": ms[len(ms) // 2], "min_ms": ms[0], "p90_ms": ms[int(len(ms) * 0.9)]}))


if __name__ == "__main__":
    main()
