#!/usr/bin/env python
"""Throughput of SequenceStateFeaturizer.fit on (B, T, S) replay batches (kernel K2 from flattened rows)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tests.cases import CASES  # noqa: E402
from tests.util import flat_featurizer, make_cuda_env  # noqa: E402

rows = []
for case, kind, B, T in [("cfg4_base_1v4", "global", 1 << 20, 1), ("cfg4_base_1v4", "global", 1 << 18, 4),
                         ("cfg4_base_1v4", "global", 4096, 2), ("cfg4_base_1v4", "perspective", 1 << 18, 1),
                         ("cfg4alt_itg_1v4", "flat98", 1 << 20, 1), ("cfg4alt_itg_1v4", "flat98", 4096, 2)]:
    env = make_cuda_env(CASES[case], B, seed=1)
    env.reset()
    seq = env.flat_states()[:, None, :].repeat(1, T, 1).contiguous()
    f = {"global": S.GlobalFeaturizer, "perspective": S.PerspectiveFeaturizer}.get(kind, None)
    feat = f(env) if f else flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])
    for _ in range(3):
        feat.fit(seq)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); feat.fit(seq); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[10]
    sh = feat._shape
    out_bytes = B * T * 4 * (sh.spatial_views * sh.spatial_floats + sh.non_spatial_views * sh.non_spatial_floats)
    in_bytes = B * T * env.flattened_state_size * 4
    rows.append({"case": case, "kind": kind, "B": B, "T": T, "fit_ms": ms, "states_per_s": B * T / (ms * 1e-3),
                 "gbs": (out_bytes + in_bytes) / (ms * 1e-3) / 1e9})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open("gpurun_out/fit_bench.json", "w"), indent=1)
