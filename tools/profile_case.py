#!/usr/bin/env python
"""A few fused step+encode launches of one configuration, for ncu:
    python tools/profile_case.py cfg4alt_itg_1v4 flat98 [envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tests.cases import CASES  # noqa: E402
from tests.util import flat_featurizer, make_cuda_env  # noqa: E402

case, kind = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
env = make_cuda_env(CASES[case], N, seed=1)
env.emit_next_states = False
env.reset()
feat = {"global": S.GlobalFeaturizer, "perspective": S.PerspectiveFeaturizer}.get(kind)
feat = feat(env) if feat else (flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"]) if kind == "flat98" else None)
for _ in range(8):
    env.step(env.sample_actions(), featurizer=feat)
torch.cuda.synchronize()
print("done")
