"""GPU: rows f2/f3 -- the batched acting loop and DQNTeamTrainer.train_step against known answers produced by the
REFERENCE's trainer, MLP and FlatFeaturizer (tools/make_golden.py train_step_fixture)."""
import os

import numpy as np
import pytest
import torch
from torch import nn

from tests.util import CASES, GOLDEN, flat_featurizer, load, make_cuda_env

pytestmark = pytest.mark.gpu


class MLPQ(nn.Module):
    """Same network as the reference's MLP (src/models/dqn.py:72-108, make_mlp :316-324): Linear + PReLU stack on
    the flattened non-spatial features; parameter names match so the fixture's state_dict loads."""

    def __init__(self, layer_dims):
        super().__init__()
        layers = []
        for i, d in enumerate(layer_dims[:-1]):
            layers += [nn.Linear(d, layer_dims[i + 1]), nn.PReLU()]
        self.model = nn.Sequential(*layers[:-1])
        self.layer_dims = layer_dims

    def forward(self, spatial_x, non_spatial_x):
        return self.model(non_spatial_x.reshape(spatial_x.size(0), -1))

    def create_copy(self):
        m = MLPQ(self.layer_dims)
        m.load_state_dict(self.state_dict())
        return m


def test_train_step_matches_reference_trainer(cuda_lib):
    import sus_net_b200 as S

    g = load(os.path.join(GOLDEN, "train_step.cfg4alt_flat98_T2.npz"))
    T = int(g["T"])
    cfg = dict(CASES["cfg4alt_itg_1v4"]); cfg["shuffle_imposter_index"] = True
    env = make_cuda_env(cfg, 4, seed=1)
    dev = env.device
    feat = flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])

    def model(name, dims):
        m = MLPQ(dims)
        m.load_state_dict({k[len(f"before.{name}."):]: torch.as_tensor(v) for k, v in g.items() if k.startswith(f"before.{name}.")})
        return m.to(dev)

    F_ = 98 * T
    imp, crew = model("imp", [F_, 32, 16, 6]), model("crew", [F_, 24, 5])
    imp_t, crew_t = model("imp_target", [F_, 32, 16, 6]), model("crew_target", [F_, 24, 5])
    batch = S.Batch(states=torch.as_tensor(g["states"]).to(dev), actions=torch.as_tensor(g["actions"]).to(dev),
                    rewards=torch.as_tensor(g["rewards"]).to(dev), next_states=torch.as_tensor(g["next_states"]).to(dev),
                    imposters=torch.as_tensor(g["imposters"]).to(dev), dones=torch.as_tensor(g["dones"]).to(dev))
    torch.backends.cuda.matmul.allow_tf32 = False
    trainer = S.DQNTeamTrainer(torch.optim.Adam(imp.parameters(), lr=1e-3), torch.optim.Adam(crew.parameters(), lr=1e-3), gamma=0.9)
    losses = trainer.train_step(batch, feat, imp, imp_t, crew, crew_t)
    assert isinstance(losses, torch.Tensor) and losses.device == dev  # the losses stay on the device (no host sync)
    losses = losses.tolist()
    # float32 GEMMs on another device: tolerance 1e-4 relative on the losses, 1e-5 absolute on the updated weights
    assert np.allclose(losses, g["losses"], rtol=1e-4), (losses, g["losses"])
    for name, m in (("imp", imp), ("crew", crew)):
        for k, v in m.state_dict().items():
            assert np.allclose(v.cpu().numpy(), g[f"after.{name}.{k}"], atol=2e-5), (name, k)


def test_batched_training_loop_runs_and_acts_within_role_ranges(cuda_lib):
    import sus_net_b200 as S

    cfg = dict(CASES["cfg4alt_itg_1v4"]); cfg["shuffle_imposter_index"] = True
    N, T = 2048, 1
    env = make_cuda_env(cfg, N, seed=5)
    feat = flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    imp, crew = MLPQ([98 * T, 64, 6]).to(env.device), MLPQ([98 * T, 32, 5]).to(env.device)
    buf = S.ReplayBuffer(8 * N, env.flattened_state_size, T, env.n_agents, env.n_imposters, device=env.device)
    trainer = S.DQNTeamTrainer(torch.optim.Adam(imp.parameters(), lr=1e-3), torch.optim.Adam(crew.parameters(), lr=1e-3), 0.9)
    sched = S.ExponentialSchedule(1.0, 0.05, 50)
    assert sched.value(0) == 1.0 and abs(sched.value(49) - 0.05) < 1e-9 and sched.value(1000) == 0.05
    losses = S.train_batched(env, buf, feat, imp, crew, trainer, sched, num_iterations=40, batch_size=512,
                             train_step_interval=5, target_update_interval=10, generator=gen)
    env.check_actions()  # every action the actor produced was inside its agent's role list
    assert len(losses) == 8 and all(np.isfinite(l).all() for l in losses)
    assert buf.size == 8 * N and int(env.metrics_batch()[:, 0].max()) == 40  # envs advanced 40 steps (fewer only where an episode restarted)
    # greedy acting (eps = 0): imposters pick argmax of the imposter net, dead agents keep action 0
    seq = buf.state_sequence
    feat.fit(seq)
    views = feat.generate_featurized_states()
    acts = S.BatchedActor(env, imp, crew).act(views, 0.0, seq[:, -1])
    assert torch.equal(acts, S.BatchedActor(env, imp, crew, dense=True).act(views, 0.0, seq[:, -1]))  # greedy: both modes agree
    alive = seq[:, -1, 10:15] != 0
    mask = env.imposter_mask_batch
    assert (acts[~alive] == 0).all() and (acts[mask].max() <= 5) and (acts[~mask].max() <= 4)
    assert torch.equal(acts, S.BatchedActor(env, imp, crew).act_grouped(feat, 0.0, seq[:, -1]))  # sync-free variant
    assert torch.equal(acts, S.BatchedActor(env, imp, crew).act_kernel(*feat.stacked_views(), 0.0))  # the selection kernel
    k = 2
    want = torch.argmax(imp(views[k][0], views[k][1]), dim=1)
    sel = mask[:, k] & alive[:, k]
    assert torch.equal(acts[sel, k].long(), want[sel])


def test_run_experiment_writes_the_references_artefacts(cuda_lib, tmp_path):
    """Row f4: config.json with the reference's keys (train.py:185-207), checkpoints at the reference's cadence
    (train.py:310,331-338,453-459) and metrics.json in the reference's schema, loadable by ITS EpisodicMetricHandler
    (metrics.py:88-95)."""
    import json

    import sus_net_b200 as S
    from sus_net_b200.experiment import run_experiment

    cfg = dict(CASES["base_fixed_order_tsr"], max_time_steps=12)
    N = 1024
    env = make_cuda_env(cfg, N, seed=5)
    feat = flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])
    F_ = int(feat.featurized_shape[1][0])
    torch.manual_seed(0)
    imp, crew = MLPQ([F_ * 2, 32, env.n_imposter_actions]).to(env.device), MLPQ([F_ * 2, 16, env.n_crew_actions]).to(env.device)
    m = run_experiment(env, 60, imp, crew, feat, sequence_length=2, replay_buffer_size=4 * N, replay_prepopulate_steps=2 * N,
                       batch_size=128, gamma=0.9, scheduler_time_steps=50, experiment_base_dir=tmp_path, learning_rate=1e-3,
                       train_step_interval=5, num_checkpoint_saves=4, target_update_interval=20, use_graphs=True, log_interval=20)
    d = m.experiment_dir
    files = sorted(p.name for p in d.iterdir())
    assert files == sorted(["config.json", "metrics.json", "metrics_totals.json"] + [f"{t}_MLPQ_{p}.pt" for t in ("imposter", "crew")
                                                                                      for p in ("0", "33", "66", "100%")]), files
    conf = json.loads((d / "config.json").read_text())
    ref_keys = {"num_steps", "imposter_model_args", "crew_model_args", "imposter_model_type", "crew_model_type", "featurizer_type",
                "sequence_length", "replay_buffer_size", "replay_prepopulate_steps", "batch_size", "gamma", "scheduler_start_eps",
                "scheduler_end_eps", "scheduler_time_steps", "train_imposter", "train_crew", "experiment_base_dir",
                "optimizer_type", "learning_rate", "train_step_interval", "target_update_interval"}
    assert ref_keys <= set(conf) and conf["num_steps"] == 60 and conf["sequence_length"] == 2
    ck = torch.load(d / "imposter_MLPQ_100%.pt", weights_only=False)
    assert set(ck) == {"state_dict", "config"} and all(torch.equal(v.cpu(), imp.state_dict()[k].cpu()) for k, v in ck["state_dict"].items())
    mj = json.loads((d / "metrics.json").read_text())
    assert set(mj) == {str(x.value) for x in S.SusMetrics} and all(isinstance(v, list) and len(v) >= 1 for v in mj.values())
    assert len(mj["imposter_loss"]) == 12 and len(mj["crew_won"]) == 3 and len(mj["avg_imposter_returns"]) == 3
    tot = json.loads((d / "metrics_totals.json").read_text())
    assert tot["episodes"] > N and tot["iterations"] == 60 and 0 < tot["averages"]["crew_won"] + tot["averages"]["imposter_won"] <= 1
    from oracle import ref_harness as H

    if H.reference_available():  # the reference's own handler loads and averages the file
        H.import_reference()
        from src.metrics import EpisodicMetricHandler as Ref

        r = Ref()
        r.load_metrics(d / "metrics.json")
        avg = r.compute()
        assert len(avg) == 13 and abs(avg["crew_won"] - np.mean(mj["crew_won"])) < 1e-12
