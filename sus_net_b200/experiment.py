"""`run_experiment` for batched envs (src/train.py:152-282): the artefacts of a reference experiment -- an experiment directory
with `config.json` (train.py:185-211), model checkpoints at the reference's cadence (`np.linspace(0, num_steps, num_saves - 1,
endpoint=False)` plus the final `*_100%.pt`, train.py:310,331-338,453-459) and `metrics.json` in the reference's schema
(metrics.py:88-95) -- around `BatchedTrainingLoop`.  The Q-networks are the caller's (dense layers are outside this package):
pass built modules instead of the reference's `ModelType` + args; a module with the reference's `dump_to_checkpoint` /
`model_type` / `config` attributes is saved through them, anything else as `{"state_dict", "config"}`.

Per logging interval the finished-episode statistics (and, with `track_returns`, the return sums) are reduced over ranks with
one all-reduce each and appended to the metric lists; that is the only host synchronisation of the run.
"""
import json
import pathlib
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from .distributed import reduce_episode_stats, reduce_return_sums
from .metrics import EpisodicMetricHandler, SusMetrics
from .replay_memory import ReplayBuffer
from .train import BatchedTrainingLoop, DQNTeamTrainer, ExponentialSchedule


class _Encoder(json.JSONEncoder):  # utils.py:14-21 (GeneralEncoder)
    def default(self, obj):
        if isinstance(obj, pathlib.Path):
            return str(obj)
        return str(obj)


def _model_type(model):
    if model is None:
        return "random"
    return str(getattr(model, "model_type", type(model).__name__))


def _dump(model, path):
    if model is None:
        return  # RandomEquiprobable.dump_to_checkpoint is a no-op (dqn.py:130-131)
    if hasattr(model, "dump_to_checkpoint"):
        model.dump_to_checkpoint(path)
    else:
        torch.save({"state_dict": model.state_dict(), "config": getattr(model, "config", {})}, path)


def run_experiment(env, num_steps, imposter_model, crew_model, featurizer, sequence_length=2, replay_buffer_size=100_000,
                   replay_prepopulate_steps=1000, batch_size=32, gamma=0.99, scheduler_start_eps=1.0, scheduler_end_eps=0.05,
                   scheduler_time_steps=1_000_000, train_imposter=True, train_crew=True, experiment_base_dir=None,
                   learning_rate=0.0001, train_step_interval=5, num_checkpoint_saves=5, target_update_interval=10_000,
                   use_graphs=False, log_interval=1000):
    """train.py:152-282 with a batched env: `num_steps` is the number of loop iterations (every iteration advances all
    `env.num_envs` envs one step).  Returns the `EpisodicMetricHandler`."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    base = pathlib.Path(experiment_base_dir) if experiment_base_dir is not None else pathlib.Path("model_registry") / "experiments"
    experiment_dir = base / datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
    if rank == 0:
        experiment_dir.mkdir(parents=True, exist_ok=True)
        config = {  # the reference's keys (train.py:185-207); model args -> the built modules' own config
            "num_steps": num_steps, "imposter_model_args": getattr(imposter_model, "config", {}),
            "crew_model_args": getattr(crew_model, "config", {}), "imposter_model_type": _model_type(imposter_model),
            "crew_model_type": _model_type(crew_model), "featurizer_type": str(featurizer), "sequence_length": sequence_length,
            "replay_buffer_size": replay_buffer_size, "replay_prepopulate_steps": replay_prepopulate_steps,
            "batch_size": batch_size, "gamma": gamma, "scheduler_start_eps": scheduler_start_eps,
            "scheduler_end_eps": scheduler_end_eps, "scheduler_time_steps": scheduler_time_steps,
            "train_imposter": train_imposter, "train_crew": train_crew, "experiment_base_dir": base,
            "optimizer_type": "adam", "learning_rate": learning_rate, "train_step_interval": train_step_interval,
            "target_update_interval": target_update_interval,
            # batched-run extras
            "num_envs": env.num_envs, "env_id_base": env.env_id_base, "seed": env.seed, "use_graphs": use_graphs,
        }
        with open(experiment_dir / "config.json", "w") as f:
            json.dump(config, f, cls=_Encoder, indent=4)
    imp_opt = torch.optim.Adam(imposter_model.parameters(), lr=learning_rate) if (train_imposter and imposter_model is not None) else None
    crew_opt = torch.optim.Adam(crew_model.parameters(), lr=learning_rate) if (train_crew and crew_model is not None) else None
    trainer = DQNTeamTrainer(imposter_optimizer=imp_opt, crew_optimizer=crew_opt, gamma=gamma)
    scheduler = ExponentialSchedule(scheduler_start_eps, scheduler_end_eps, scheduler_time_steps)
    metrics = EpisodicMetricHandler()
    replay_buffer = ReplayBuffer(max_size=max(replay_buffer_size, env.num_envs), trajectory_size=sequence_length,
                                 state_size=env.flattened_state_size, n_agents=env.n_agents, n_imposters=env.n_imposters,
                                 device=env.device)
    if not env._was_reset:
        env.reset()
    replay_buffer.populate(env=env, num_steps=replay_prepopulate_steps)  # train.py:254 (random policy, N transitions per launch)
    env.episode_stats(clear=True)  # the reference's metrics start with train(), not with the prepopulation
    env.track_returns(gamma)  # G = reward + gamma * G per agent, averaged per team at episode ends (train.py:386,421-424)
    loop = BatchedTrainingLoop(env, replay_buffer, featurizer, imposter_model, crew_model, trainer, scheduler,
                               batch_size=batch_size, train_step_interval=train_step_interval,
                               target_update_interval=target_update_interval, use_graphs=use_graphs)
    t_saves = set(np.linspace(0, num_steps, num_checkpoint_saves - 1, endpoint=False, dtype=int).tolist())  # train.py:310

    def on_iteration(it):
        if it in t_saves and trainer.train and rank == 0:  # train.py:331-338
            pct = f"{int(it * 100 / num_steps)}"
            _dump(imposter_model, experiment_dir / f"imposter_{_model_type(imposter_model)}_{pct}.pt")
            _dump(crew_model, experiment_dir / f"crew_{_model_type(crew_model)}_{pct}.pt")
        if (it + 1) % log_interval == 0 or it + 1 == num_steps:
            metrics.log_interval(reduce_episode_stats(env.episode_stats()), reduce_return_sums(env.return_sums()))

    loop.run(num_steps, on_iteration=on_iteration)
    losses = loop.finish()
    if rank == 0:
        _dump(imposter_model, experiment_dir / f"imposter_{_model_type(imposter_model)}_100%.pt")  # train.py:453-459
        _dump(crew_model, experiment_dir / f"crew_{_model_type(crew_model)}_100%.pt")
        metrics.set({SusMetrics.IMPOSTER_LOSS: [l[0] for l in losses], SusMetrics.CREW_LOSS: [l[1] for l in losses]})  # train.py:468-471
        metrics.extra.update(env_steps=num_steps * env.num_envs, iterations=num_steps)
        metrics.save_metrics(experiment_dir / "metrics.json")  # train.py:280
    metrics.experiment_dir = experiment_dir
    return metrics
