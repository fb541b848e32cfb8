#!/usr/bin/env python
"""Reference mode (num_envs = 1, the literal drop-in for src/train.py's loop): steps/s of
`a = env.sample_actions(); env.step(a)` and of the featurizer on a (1, T, S) sequence, beside the reference's own
CPU numbers (4-14 k steps/s step-only, 0.8-2.4 k/s with encode)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402

env = S.FourRoomEnv(1, 4, 5, random_state=0)
feat = S.GlobalFeaturizer(env)
state, _ = env.reset()
seq = np.zeros((1, env.flattened_state_size))
for name, n, with_feat, with_sample in (("step(fixed actions)", 3000, False, False), ("sample_actions+step", 3000, False, True),
                                         ("sample_actions+step+fit+views", 2000, True, True)):
    acts = np.zeros(5, dtype=np.int32)
    t0 = time.perf_counter()
    for _ in range(n):
        if with_sample:
            acts = env.sample_actions()
        state, r, d, tr, info = env.step(acts)
        if with_feat:
            seq[0] = env.flatten_state(state)
            feat.fit(torch.tensor(seq).unsqueeze(0))
            feat.generate_featurized_states()
        if d or tr:
            state, _ = env.reset()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"loop": name, "steps_per_s": n / dt, "us_per_step": 1e6 * dt / n}), flush=True)
