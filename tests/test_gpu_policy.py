"""GPU: rows f2 / f3 -- the acting kernel (`sus_env_select_actions`) against the oracle's numpy restatement of
src/train.py:349-381 on the same Philox draws and against fixtures of the REFERENCE's own greedy acting
(tools/make_golden.py `acting_fixture`: unmodified `train()` with epsilon ~ 0); the T-deep feature ring (`sus_seq_roll`);
the batched training loop eager vs CUDA graphs."""
import json
import os

import numpy as np
import pytest
import torch
from torch import nn

import oracle
from tests.cases import CASES
from tests.util import GOLDEN, flat_featurizer, load, make_cuda_env

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", ["cfg4_base_1v4", "cfg4alt_itg_1v4", "cfg3_tagging_1v2", "base_2v3_j3", "tagging_2v5_short",
                                  "base_1v7_j8_nowall"])
def test_select_actions_matches_oracle(cuda_lib, name):
    import ctypes as C

    from sus_net_b200 import _lib as L

    cfg = CASES[name]
    N, seed, base = 3001, 17, 555
    env = make_cuda_env(cfg, N, seed=seed, env_id_base=base)
    env.reset()
    for _ in range(25):
        env.step(None)  # some agents die, some episodes restart
    A, nI = env.n_agents, env.n_imposters
    nia, nca = env.n_imposter_actions, env.n_crew_actions
    g = torch.Generator(device=env.device); g.manual_seed(3)
    ids = np.arange(base, base + N)
    for trial, (eps, use_imp, use_crew, per_view, dtype) in enumerate([
            (0.0, True, True, nI != 1, torch.int32), (0.3, True, True, True, torch.uint8), (1.0, True, True, nI != 1, torch.int32),
            (0.2, True, False, nI != 1, torch.int32), (0.0, False, True, True, torch.uint8), (0.5, False, False, True, torch.int32)]):
        q_imp = torch.randn((A, N, nia) if per_view else (N, nia), device=env.device, generator=g) if use_imp else None
        q_crew = torch.randn((A, N, nca), device=env.device, generator=g) if use_crew else None
        if q_imp is not None:  # ties: the FIRST maximum wins (torch.argmax, train.py:368-370)
            q_imp[..., 3] = q_imp[..., 1]
        eps_t = torch.tensor([eps], dtype=torch.float32, device=env.device)
        out = torch.full((N, A), 77, dtype=dtype, device=env.device)
        io = L.SusPolicyIO()
        io.q_imposter = None if q_imp is None else q_imp.data_ptr()
        io.q_crew = None if q_crew is None else q_crew.data_ptr()
        io.imposter_per_view = int(per_view)
        if trial % 2:
            io.eps = eps_t.data_ptr()
        else:
            io.eps_value = eps
        io.actions = out.data_ptr()
        io.actions_dtype = L.U8 if dtype == torch.uint8 else L.I32
        epoch = env.state_dict()["ticks"][2]
        L.check(env.lib.sus_env_select_actions(env._h, C.byref(io), env._stream()))
        assert env.state_dict()["ticks"][2] == epoch + 1
        flat = cpu(env.flat_states(torch.int64))
        want = oracle.select_actions(cfg, seed, ids, epoch, flat[:, 2 * A:3 * A], cpu(env.imposter_mask_batch),
                                     None if q_imp is None else cpu(q_imp), None if q_crew is None else cpu(q_crew), eps,
                                     imposter_per_view=per_view)
        assert np.array_equal(cpu(out).astype(np.int32), want), f"{name}: trial {trial}"
        env.step(out)  # the selected actions are valid role-list indices
        env.check_actions()
    if nI != 1:  # the [N][n_actions] imposter layout is refused for several imposters
        io = L.SusPolicyIO()
        dummy = torch.zeros((N, nia), device=env.device)
        io.q_imposter = dummy.data_ptr(); io.actions = out.data_ptr(); io.actions_dtype = L.I32
        assert env.lib.sus_env_select_actions(env._h, C.byref(io), env._stream()) == L.SUS_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("T,R,V", [(2, 567, 1), (3, 15, 5), (4, 98, 1), (2, 7, 3)])
def test_seq_roll_matches_numpy(cuda_lib, T, R, V):
    import ctypes as C

    import sus_net_b200 as S

    N = 777
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(1)
    seq = torch.randn((V, N, T, R), device=dev, generator=g)
    new = torch.randn((V, N, R), device=dev, generator=g)
    done = torch.rand(N, device=dev, generator=g) < 0.2
    trunc = torch.rand(N, device=dev, generator=g) < 0.1
    out = torch.empty_like(seq)
    lib = S.lib()
    rc = lib.sus_seq_roll(C.c_void_p(seq.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(new.data_ptr()),
                          C.c_void_p(done.data_ptr()), C.c_void_p(trunc.data_ptr()), V * N, N, T, R, 0,
                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    want = np.roll(cpu(seq), -1, axis=2)  # np.roll(sequence, -1, axis=0) per env (train.py:388-389)
    want[:, :, -1] = cpu(new)
    fin = cpu(done | trunc)
    want[:, fin] = cpu(new)[:, fin, None, :]  # T copies of the reset state's features (train.py:440-445)
    assert np.array_equal(cpu(out), want)


class MLPQ(nn.Module):
    """The reference's MLP (src/models/dqn.py:72-108, make_mlp :316-324); parameter names match its state_dict."""

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


class SpatialQ(nn.Module):
    """The reference's SpatialDQN (src/models/dqn.py:204-311) restated with the same parameter names: Conv2d+ReLU stack
    (n_channels + repeated last layer), nn.RNN over time on [cnn features | non-spatial], PReLU MLP head."""

    def __init__(self, c_in, ns, n_actions, ch=6, hidden=24, head=16):
        super().__init__()

        class Wrap(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.model = m

        self.cnn = Wrap(nn.Sequential(nn.Conv2d(c_in, ch, 3, 1, 1), nn.ReLU(), nn.Conv2d(ch, ch, 3, 1, 1), nn.ReLU(),
                                      nn.Conv2d(ch, ch, 3, 1, 1), nn.ReLU()))
        self.rnn = Wrap(nn.RNN(input_size=81 * ch + ns, hidden_size=hidden, num_layers=1, batch_first=True))
        self.prediction_head = nn.Sequential(nn.Linear(hidden, head), nn.PReLU(), nn.Linear(head, n_actions))

    def forward(self, spatial_x, non_spatial_x):
        b, t, c, h, w = spatial_x.size()
        x = self.cnn.model(spatial_x.reshape(b * t, c, h, w)).reshape(b, t, -1)
        out, _ = self.rnn.model(torch.cat((x, non_spatial_x), dim=2))
        return self.prediction_head(out[:, -1, :])


@pytest.mark.parametrize("tag", ["acting.cfg4_global_spatial_T2", "acting.cfg4alt_flat98_mlp_T1"])
def test_acting_matches_the_references_greedy_actions(cuda_lib, tag):
    """Every transition the reference's train() stored holds the state sequence its acting code (train.py:349-381) saw and the
    greedy actions it chose: the GPU featurizer + the same networks on the GPU + the selection kernel must choose the same."""
    import sus_net_b200 as S

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load(os.path.join(GOLDEN, tag + ".npz"))
    cfg = json.loads(bytes(g["cfg_json"]).decode())
    states, want, imposters = g["states"], g["actions"].astype(np.int32), g["imposters"]
    M, T, S_ = states.shape
    A = want.shape[1]
    env = make_cuda_env(cfg, M, seed=1)
    mask = np.zeros((M, A), dtype=np.uint8)
    mask[np.arange(M)[:, None], imposters.astype(np.int64)] = 1
    env.import_flat(states[:, -1].astype(np.int64), mask)  # the selection kernel reads liveness and roles from the env state
    dev = env.device
    sd = lambda p: {k[len(p):]: torch.as_tensor(v) for k, v in g.items() if k.startswith(p)}  # noqa: E731
    if str(g["kind"]) == "global":
        feat = S.GlobalFeaturizer(env)
        ns = int(feat.featurized_shape[1][0])
        imp, crew = SpatialQ(A + 2, ns, env.n_imposter_actions), SpatialQ(A + 2, ns, env.n_crew_actions)
    else:
        feat = flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])
        imp, crew = MLPQ([98 * T, 48, 24, env.n_imposter_actions]), MLPQ([98 * T, 32, env.n_crew_actions])
    imp.load_state_dict(sd("imp.")); crew.load_state_dict(sd("crew."))
    imp, crew = imp.to(dev).eval(), crew.to(dev).eval()
    feat.fit(torch.as_tensor(states).to(dev))
    sp, ns_ = feat.stacked_views()
    actor = S.BatchedActor(env, imp, crew)
    got = cpu(actor.act_kernel(sp, ns_, 0.0))
    alive = states[:, -1, 2 * A:3 * A] != 0
    assert (got[~alive] == 0).all() and (want[~alive] == 0).all()
    diff = np.argwhere(got != want)
    # float32 network outputs on another device can flip an argmax only where the two best Q-values nearly tie
    for e, i in diff:
        views = feat.generate_featurized_states()
        model = imp if mask[e, i] else crew
        q = model(views[i][0][e:e + 1], views[i][1][e:e + 1])[0]
        top = torch.topk(q, 2).values
        assert float(top[0] - top[1]) < 1e-4 * max(1.0, float(top[0].abs())), f"{tag}: env {e} agent {i}: {got[e]} vs {want[e]}"
    assert len(diff) <= max(1, M // 200), f"{tag}: {len(diff)} of {M * A} actions differ"
    # the torch-op restatements agree with the kernel at eps = 0
    flat = torch.as_tensor(states[:, -1]).to(dev)
    assert np.array_equal(cpu(actor.act(feat.generate_featurized_states(), 0.0, flat)), got)
    if env.n_imposters == 1:
        assert np.array_equal(cpu(actor.act_grouped(feat, 0.0, flat)), got)


def _loop(cfg, N, T, graphs, train, kind="flat", iters=23):
    import sus_net_b200 as S

    env = make_cuda_env(cfg, N, seed=5)
    dev = env.device
    if kind == "flat":
        feat = flat_featurizer(env, ["onehot_pos", "alive_crew", "closest_crew"])
        F_ = int(feat.featurized_shape[1][0])
        torch.manual_seed(0)
        imp, crew = MLPQ([F_ * T, 64, env.n_imposter_actions]).to(dev), MLPQ([F_ * T, 32, env.n_crew_actions]).to(dev)
    else:
        feat = S.GlobalFeaturizer(env)
        ns = int(feat.featurized_shape[1][0])
        torch.manual_seed(0)
        imp, crew = SpatialQ(env.n_agents + 2, ns, env.n_imposter_actions).to(dev), SpatialQ(env.n_agents + 2, ns, env.n_crew_actions).to(dev)
    buf = S.ReplayBuffer(8 * N, env.flattened_state_size, T, env.n_agents, env.n_imposters, device=dev)
    trainer = S.DQNTeamTrainer(torch.optim.Adam(imp.parameters(), lr=1e-3) if train else None,
                               torch.optim.Adam(crew.parameters(), lr=1e-3) if train else None, 0.9)
    loop = S.BatchedTrainingLoop(env, buf, feat, imp, crew, trainer, S.ExponentialSchedule(1.0, 0.05, 15), batch_size=256,
                                 train_step_interval=5, target_update_interval=10, use_graphs=graphs)
    loop.run(iters)
    losses = loop.finish()
    return env, buf, loop, losses


@pytest.mark.parametrize("kind,T,case", [("flat", 1, "base_fixed_order_tsr"), ("flat", 2, "base_fixed_order_tsr"), ("global", 2, "cfg4_base_1v4")])
def test_training_loop_graph_replay_equals_eager_without_training(cuda_lib, kind, T, case):
    """No optimizers => nothing random besides the Philox draws: the CUDA-graph loop must walk exactly the eager loop's
    trajectory (states, replay ring, statistics), and the feature ring must equal featurizer.fit() of the state sequences."""
    cfg = dict(CASES[case], shuffle_imposter_index=True, max_time_steps=9)  # short episodes: the reset paths are exercised
    N = 1500
    a_env, a_buf, a_loop, _ = _loop(cfg, N, T, False, False, kind)
    b_env, b_buf, b_loop, _ = _loop(cfg, N, T, True, False, kind)
    assert b_loop._g_iter is not None
    assert torch.equal(a_env.flat_states(torch.int64), b_env.flat_states(torch.int64))
    assert torch.equal(a_env.episode_stats(), b_env.episode_stats()) and int(a_env.episode_stats()[0]) > 0
    assert (a_buf.idx, a_buf.size) == (b_buf.idx, b_buf.size) == ((23 * N) % (8 * N), 8 * N)
    for k in ("states", "actions", "rewards", "next_states", "dones", "imposters"):
        assert torch.equal(getattr(a_buf, k), getattr(b_buf, k)), k
    # the rolled feature sequences are what fit() makes of the raw state sequences (train.py:346-348)
    sp, ns = b_loop.seq.views()
    f2 = b_loop.feat.clone_for()
    f2.fit(b_buf.state_sequence)
    sp2, ns2 = f2.stacked_views()
    assert torch.equal(ns, ns2) and (sp is None or torch.equal(sp, sp2))


def test_training_loop_trains_eager_and_graphed(cuda_lib):
    cfg = dict(CASES["cfg4alt_itg_1v4"]); cfg["shuffle_imposter_index"] = True
    for graphs in (False, True):
        env, buf, loop, losses = _loop(cfg, 2048, 1, graphs, True, iters=40)
        assert len(losses) == 8 and np.isfinite(losses).all() and all(l[0] > 0 and l[1] > 0 for l in losses)
        assert buf.size == 8 * 2048 and int(env.metrics_batch()[:, 0].max()) == 40
        assert (loop._g_train is not None) == graphs
