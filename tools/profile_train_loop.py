#!/usr/bin/env python
"""Where the time of the cfg5 training loop goes (one GPU): wall time per phase with a device synchronize after each,
plus the torch profiler's top kernels.

    python tools/profile_train_loop.py [--envs 131072] [--iters 50]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tools.train_demo import MLPQ, RandomQ  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--kernels", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    N = a.envs
    env = S.BatchedImposterTrainingGround(n_crew=4, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                          end_of_game_reward=0, num_envs=N, seed=7, device=dev)
    feat = S.FlatFeaturizer(env, S.CompositeFeaturizer([S.OneHotAgentPositionFeaturizer(env), S.AliveCrewFeaturizer(env),
                                                        S.ClosestAliveCrewFeaturizer(env)]))
    torch.manual_seed(0)
    imp, crew = MLPQ([98, 256, 128, 64, 16, 6]).to(dev), RandomQ(5).to(dev)
    imp_t, crew_t = imp.create_copy().to(dev), crew.create_copy().to(dev)
    trainer = S.DQNTeamTrainer(torch.optim.Adam(imp.parameters(), lr=1e-3), None, gamma=0.9)
    buf = S.ReplayBuffer(8 * N, env.flattened_state_size, 1, env.n_agents, 1, device=dev)
    actor = S.BatchedActor(env, imp, crew)
    buf.attach(env)
    phases = {"fit": 0.0, "act": 0.0, "collect": 0.0, "sample": 0.0, "train": 0.0}

    def timed(name, fn):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(dev)
        phases[name] += time.perf_counter() - t0
        return out

    def iteration(it, record):
        t = timed if record else (lambda _n, fn: fn())
        seq = buf.state_sequence
        t("fit", lambda: feat.fit(seq))
        actions = t("act", lambda: actor.act_grouped(feat, 0.3, seq[:, -1]))
        t("collect", lambda: buf.collect_step(actions))
        if it % 5 == 0:
            batch = t("sample", lambda: buf.sample(a.batch))
            t("train", lambda: trainer.train_step(batch, feat, imp, imp_t, crew, crew_t))

    for it in range(10):
        iteration(it, False)
    for it in range(a.iters):
        iteration(it, True)
    print(json.dumps({"envs": N, "iters": a.iters, "ms_per_iteration": {k: 1e3 * v / a.iters for k, v in phases.items()},
                      "total_ms_per_iteration_synced": 1e3 * sum(phases.values()) / a.iters}))
    if a.kernels:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for it in range(10):
                iteration(it, False)
            torch.cuda.synchronize(dev)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
