#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end batched DQN loop -- ImposterTrainingGround 1v4 (walled), FlatFeaturizer(OneHot +
AliveCrew + ClosestAliveCrew) = 98 features, MLP imposter Q-net [98,256,128,64,16,6] vs a random crew
(notebooks/experiment_1v1.ipynb cell 1 model args), T = 1, gamma = 0.9, lr = 1e-3, 131 072 envs per GPU, replay
writes on the GPU.  One process per GPU (torchrun) or a single process.  Every iteration advances ALL envs one step:
imposter network on the features the fused step kernel wrote -> epsilon-greedy selection kernel -> fused step + Flat-98
encode -> replay push; a train step (batch 4096) every 5 iterations; the whole iteration and the train step are CUDA
graphs (--no-graphs: eager launches).  Timed over >= --seconds of wall clock after warm-up.

    python tools/train_demo.py [--envs-per-gpu 131072] [--seconds 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_demo.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch import nn  # noqa: E402

import sus_net_b200 as S  # noqa: E402


class MLPQ(nn.Module):
    """The reference's MLP (src/models/dqn.py:72-108): Linear + PReLU stack on the flattened non-spatial features."""

    def __init__(self, dims):
        super().__init__()
        layers = []
        for i, d in enumerate(dims[:-1]):
            layers += [nn.Linear(d, dims[i + 1]), nn.PReLU()]
        self.model = nn.Sequential(*layers[:-1])
        self.dims = dims

    def forward(self, spatial_x, non_spatial_x):
        return self.model(non_spatial_x.reshape(spatial_x.size(0), -1))

    def create_copy(self):
        m = MLPQ(self.dims)
        m.load_state_dict(self.state_dict())
        return m


def build(N, dev, rank, batch, graphs, iters_for_schedule=2000, fused_mlp=True):
    env = S.BatchedImposterTrainingGround(n_crew=4, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                          end_of_game_reward=0, num_envs=N, seed=7, env_id_base=rank * N, device=dev)
    feat = S.FlatFeaturizer(env, S.CompositeFeaturizer([S.OneHotAgentPositionFeaturizer(env), S.AliveCrewFeaturizer(env),
                                                        S.ClosestAliveCrewFeaturizer(env)]))
    torch.manual_seed(0)  # identical initial weights on every rank
    imp = MLPQ([98, 256, 128, 64, 16, 6]).to(dev)
    trainer = S.DQNTeamTrainer(torch.optim.Adam(imp.parameters(), lr=1e-3), None, gamma=0.9)
    buf = S.ReplayBuffer(8 * N, env.flattened_state_size, 1, env.n_agents, 1, device=dev)
    sched = S.ExponentialSchedule(1.0, 0.05, iters_for_schedule)
    # crew_model=None: the reference's RandomEquiprobable crew (uniform over the crew's role list)
    loop = S.BatchedTrainingLoop(env, buf, feat, imp, None, trainer, sched, batch_size=batch, train_step_interval=5,
                                 target_update_interval=1000, use_graphs=graphs, fused_mlp=fused_mlp)
    return env, loop


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--tf32", action="store_true", help="allow TF32 tensor-core GEMMs in the Q-network")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--torch-mlp", action="store_true", help="evaluate the acting Q-network with the torch module instead of sus_mlp_forward")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = a.envs_per_gpu
    env, loop = build(N, dev, rank, a.batch, not a.no_graphs, fused_mlp=not a.torch_mlp)
    loop.run(40)  # warm-up: eager iterations, graph captures, first replays
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    chunk, iters = 500, 0
    t0 = time.perf_counter()
    while True:
        loop.run(chunk)
        iters += chunk
        torch.cuda.synchronize(dev)
        flag = torch.tensor([1.0 if time.perf_counter() - t0 >= a.seconds else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank runs the same number of iterations
        if flag.item() > 0:
            break
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    losses = loop.finish()
    stats = S.reduce_episode_stats(env.episode_stats())
    if rank == 0:
        print(json.dumps({"config": "cfg5 batched DQN loop (imposter MLP on fused Flat-98 features -> selection kernel -> fused step "
                                    "+ encode -> replay push; train step every 5 iterations)",
                          "n_gpus": world, "envs_per_gpu": N, "global_envs": world * N, "iterations": iters, "wall_s": dt,
                          "cuda_graphs": not a.no_graphs, "tf32_q_network": a.tf32,
                          "acting_q_network": "torch module" if a.torch_mlp else "sus_mlp_forward (one launch, fp32 FFMA)", "batch_size": a.batch,
                          "ms_per_iteration": 1e3 * dt / iters,
                          "env_steps_per_s_in_training_loop": world * N * iters / dt,
                          "train_steps": len(losses), "last_losses": losses[-1],
                          "episodes": int(stats[0]), "imposter_win_rate": float(stats[2]) / max(int(stats[0]), 1)}))
    if world > 1:
        # The captured graphs hold NCCL work; tearing the communicator down with them alive hung the 8-GPU run at exit (the
        # measurement had finished).  Drop the graphs, drain the device, meet at a barrier and leave without the destructor.
        loop._g_iter = loop._g_train = None
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
