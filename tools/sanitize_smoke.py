#!/usr/bin/env python
"""Small run of every kernel and both output paths, meant to be wrapped in compute-sanitizer (one tool per call):

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tests.cases import CASES, FLAT_COMPONENT_SETS, GLOBAL_CASES  # noqa: E402
from tests.util import flat_featurizer, make_cuda_env  # noqa: E402

for path in ("tma", "direct"):
    os.environ["SUSNET_PATH"] = path
    for name, cfg in CASES.items():
        for N in (37, 1000):
            env = make_cuda_env(cfg, N, seed=3)
            env.reset()
            feats = [None]
            if name in GLOBAL_CASES:
                feats += [S.GlobalFeaturizer(env), S.PerspectiveFeaturizer(env)]
            for comps in FLAT_COMPONENT_SETS.get(name, []):
                feats.append(flat_featurizer(env, comps))
            for t in range(12):
                f = feats[t % len(feats)]
                a = env.sample_actions() if t % 2 else None
                env.step(a, featurizer=f)
                if f is not None:
                    f.encode_env()
                    f.fit(env.flat_states().reshape(1, N, -1))
            env.flat_states(torch.int64); env.metrics_batch(); env.episode_stats(); env.imposter_mask_batch
            torch.cuda.synchronize()
    print("path", path, "ok")
print("sanitize smoke done")
