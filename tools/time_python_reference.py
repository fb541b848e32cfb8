#!/usr/bin/env python
"""Context number: the UNMODIFIED Python reference's own loop (a = env.sample_actions(); env.step(a); GlobalFeaturizer
fit + views; reset on done/trunc) timed in the build container (needs /root/reference; cannot run on the GPU box).

    python tools/time_python_reference.py [--steps 3000] [--procs 8]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(args):
    n_steps, seed, with_features = args
    import numpy as np
    import torch

    from oracle import ref_harness as H

    env_mod, feat_mod = H.import_reference()
    torch.set_num_threads(1)
    np.random.seed(seed)
    env = env_mod.FourRoomEnv(n_imposters=1, n_crew=4, n_jobs=5)
    feat = feat_mod.GlobalFeaturizer(env)
    state, _ = env.reset()
    seq = np.zeros((1, env.flattened_state_size))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        a = env.sample_actions()
        state, r, d, tr, info = env.step(a)
        if with_features:
            seq[0] = env.flatten_state(state)
            feat.fit(torch.tensor(seq).unsqueeze(0))
            feat.generate_featurized_states()
        if d or tr:
            state, _ = env.reset()
    return n_steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    out = {"host": "build container", "cores": a.procs, "config": "FourRoomEnv(1, 4, 5) defaults, random policy"}
    with mp.Pool(a.procs) as pool:
        for name, wf in (("sample_actions+step", False), ("sample_actions+step+Global fit+views", True)):
            rates = pool.map(worker, [(a.steps, 100 + i, wf) for i in range(a.procs)])
            out[name] = {"per_core_steps_per_s": sum(rates) / len(rates), "aggregate_steps_per_s": sum(rates)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
