"""CPU: host-side pieces of the "next" rows that need no GPU -- schedule, replay-buffer add/sample semantics on a CPU
device (same code path as on the GPU apart from the push kernel), metrics adaptor."""
import math

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def S():
    from sus_net_b200 import build as B

    B.build()
    import sus_net_b200

    return sus_net_b200


def test_exponential_schedule_matches_reference_formula(S):
    sch = S.ExponentialSchedule(1.0, 0.05, 1_000_000)  # report p.5: eps 1 -> 0.05 over 1M steps
    assert sch.value(0) == 1.0 and sch.value(-3) == 1.0 and sch.value(1_000_000) == 0.05 and sch.value(2_000_000) == 0.05
    b = math.log(0.05) / 999_999
    for t in (1, 10, 999, 500_000, 999_999):
        assert sch.value(t) == pytest.approx(math.exp(b * t), rel=1e-12)
    from oracle import ref_harness as H

    if H.reference_available():
        H.import_reference()
        from src.scheduler import ExponentialSchedule as Ref

        ref = Ref(1.0, 0.05, 1_000_000)
        for t in (0, 1, 17, 123456, 999_999, 1_000_000):
            assert float(ref.value(t)) == pytest.approx(sch.value(t), rel=1e-12)


def test_replay_buffer_add_sample_on_cpu_matches_reference_layout(S):
    buf = S.ReplayBuffer(max_size=5, state_size=6, trajectory_size=2, n_agents=2, n_imposters=1, device="cpu")
    assert buf.states.shape == (5, 2, 6) and buf.states.dtype == torch.float32
    assert buf.actions.shape == (5, 2) and buf.actions.dtype == torch.int64
    assert buf.rewards.shape == (5, 2) and buf.next_states.shape == (5, 2, 6)
    assert buf.dones.shape == (5, 1) and buf.dones.dtype == torch.bool
    assert buf.imposters.shape == (5, 1) and buf.imposters.dtype == torch.int16
    for k in range(7):  # wraps after 5 (replay_memory.py:70-72)
        st = np.full((2, 6), k, dtype=np.float64)
        buf.add(state=st, action=np.array([k, k + 1]), reward=np.array([0.5 * k, -k]), next_state=st + 1, done=k % 2 == 1,
                imposters=np.array([k % 2]))
    assert buf.idx == 2 and buf.size == 5
    assert buf.states[0, 0, 0] == 5 and buf.states[1, 0, 0] == 6 and buf.states[2, 0, 0] == 2
    assert buf.actions[1].tolist() == [6, 7] and buf.rewards[1].tolist() == [3.0, -6.0] and bool(buf.dones[1, 0]) is False
    assert bool(buf.dones[0, 0]) is True and buf.imposters[0, 0] == 1
    g = torch.Generator().manual_seed(0)
    b = buf.sample(64, generator=g)
    assert b._fields == ("states", "actions", "rewards", "next_states", "imposters", "dones")
    assert b.states.shape == (64, 2, 6) and b.dones.shape == (64, 1)
    assert torch.equal(b.next_states, b.states + 1)  # rows stay aligned across the six tensors
    with pytest.raises(AssertionError):
        S.ReplayBuffer(5, 6, 2, 2, 1, device="cpu").sample(1)  # empty (replay_memory.py:83)
    for bad in (dict(max_size=0), dict(trajectory_size=0), dict(state_size=0), dict(n_agents=0)):
        kw = dict(max_size=5, state_size=6, trajectory_size=2, n_agents=2, n_imposters=1)
        kw.update(bad)
        with pytest.raises(AssertionError):
            S.ReplayBuffer(device="cpu", **kw)
    from oracle import ref_harness as H

    if H.reference_available():  # identical tensors to the reference's own buffer fed the same transitions
        H.import_reference()
        from src.replay_memory import ReplayBuffer as Ref

        ref = Ref(max_size=5, state_size=6, trajectory_size=2, n_agents=2, n_imposters=1)
        mine = S.ReplayBuffer(max_size=5, state_size=6, trajectory_size=2, n_agents=2, n_imposters=1, device="cpu")
        rng = np.random.default_rng(0)
        for k in range(9):
            st, nx = rng.integers(0, 9, (2, 6)).astype(np.float64), rng.integers(0, 9, (2, 6)).astype(np.float64)
            a, r = rng.integers(0, 6, 2), rng.normal(size=2)
            for bfr in (ref, mine):
                bfr.add(st, a, r, nx, bool(k % 3 == 0), np.array([k % 2]))
        for name in ("states", "actions", "rewards", "next_states", "dones", "imposters"):
            assert torch.equal(getattr(ref, name), getattr(mine, name)), name
        assert (ref.idx, ref.size) == (mine.idx, mine.size)


def test_episodic_metric_handler_from_stats(S, tmp_path):
    m = S.EpisodicMetricHandler()
    m.update_from_stats(torch.tensor([10, 6, 4, 20, 30, 2, 1, 0, 900, 1]))
    avg = m.compute()
    assert avg[S.SusMetrics.CREW_WON] == 0.6 and avg[S.SusMetrics.IMPOSTER_WON] == 0.4
    assert avg[S.SusMetrics.TOTAL_TIME_STEPS] == 90.0 and avg[S.SusMetrics.IMP_KILLED_CREW] == 2.0
    assert set(avg) == set(S.SusMetrics) and len(avg) == 13
    p = tmp_path / "metrics.json"
    m.save_metrics(p)
    import json

    # metrics.json: the reference's schema, {metric name: list} for all 13 metrics (metrics.py:88-90)
    d = json.loads(p.read_text())
    assert set(d) == {str(x.value) for x in S.SusMetrics} and d["crew_won"] == [0.6] and d["total_time_steps"] == [90.0]
    # the exact totals live in the sidecar
    t = json.loads((tmp_path / "metrics_totals.json").read_text())
    assert t["episodes"] == 10 and t["truncated_episodes"] == 1 and t["totals"]["completed_jobs"] == 30
    # per-interval lists: means over the episodes finished in each interval
    m2 = S.EpisodicMetricHandler()
    assert m2.log_interval(torch.tensor([10, 6, 4, 20, 30, 2, 1, 0, 900, 1]), [5.0, -3.0]) == 10
    assert m2.log_interval(torch.tensor([30, 16, 14, 50, 70, 4, 1, 0, 2900, 3]), [15.0, -13.0]) == 20
    assert m2.log_interval(torch.tensor([30, 16, 14, 50, 70, 4, 1, 0, 2900, 3]), [15.0, -13.0]) == 0  # nothing finished: no entry
    assert m2.metrics[S.SusMetrics.CREW_WON] == [0.6, 0.5] and m2.metrics[S.SusMetrics.AVG_CREW_RETURNS] == [-0.3, -0.5]
    assert m2.compute()[S.SusMetrics.CREW_WON] == 16 / 30 and m2.compute()[S.SusMetrics.AVG_IMPOSTER_RETURNS] == 0.5
    m2.set({"imposter_loss": [1.0, 3.0], "crew_loss": [0.0, 0.0]})
    m2.save_metrics(p)
    from oracle import ref_harness as H

    if H.reference_available():  # the reference's own handler loads the file and averages every key (metrics.py:84-95)
        H.import_reference()
        from src.metrics import EpisodicMetricHandler as Ref

        r = Ref()
        r.load_metrics(p)
        avg = r.compute()
        assert len(avg) == 13 and avg["crew_won"] == 0.55 and avg["imposter_loss"] == 2.0


_GRAD_WORKER = r'''
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sus_net_b200.train import allreduce_grads
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
torch.manual_seed(0)
m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.PReLU(), torch.nn.Linear(3, 2))
x = torch.full((5, 4), float(rank + 1))
m(x).sum().backward()
local = [p.grad.clone() for p in m.parameters()]
allreduce_grads(m)
# the other rank's gradients, recomputed locally
m2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.PReLU(), torch.nn.Linear(3, 2))
m2.load_state_dict(m.state_dict())
m2(torch.full((5, 4), float(2 - rank))).sum().backward()
for p, g1, p2 in zip(m.parameters(), local, m2.parameters()):
    assert torch.allclose(p.grad, (g1 + p2.grad) / 2, atol=1e-6)
dist.destroy_process_group()
'''


def test_gradient_allreduce_two_ranks_gloo(S):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = str(31000 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, "-c", _GRAD_WORKER, root, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_flat_adam_equals_torch_adam_and_gates_empty_subsets(S):
    """`_FlatAdam` (the device-gated optimizer step of the sync-free trainer) == torch.optim.Adam when the subset is not empty;
    with a zero sample count weights, moments and the step count stay untouched (the reference `continue`s, train.py:88-92)."""
    from sus_net_b200.train import _FlatAdam

    torch.manual_seed(0)
    make = lambda: torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.PReLU(), torch.nn.Linear(5, 3))  # noqa: E731
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    ref = torch.optim.Adam(a.parameters(), lr=1e-2)
    fa = _FlatAdam(torch.optim.Adam(b.parameters(), lr=1e-2))
    for step in range(6):
        x, y = torch.randn(9, 6), torch.randn(9, 3)
        ref.zero_grad()
        torch.nn.functional.mse_loss(a(x), y).backward()  # mean over 9 x 3 elements
        ref.step()
        fa.begin_train_step(); fa.begin_view()
        sq = ((b(x) - y) ** 2).sum()
        sq.backward()
        loss = fa.apply(torch.tensor(27.0), sq.detach())
        assert float(loss) == pytest.approx(float(sq) / 27, rel=1e-6)
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, atol=1e-6), step
        if step == 3:  # an empty subset in between: nothing may move
            before = [q.detach().clone() for q in b.parameters()]
            m, v, st = fa.m.clone(), fa.v.clone(), fa.step.clone()
            fa.begin_train_step(); fa.begin_view()
            assert float(fa.apply(torch.tensor(0.0), torch.tensor(0.0))) == 0.0
            assert all(torch.equal(p, q) for p, q in zip(before, b.parameters()))
            assert torch.equal(m, fa.m) and torch.equal(v, fa.v) and torch.equal(st, fa.step)


_FLAT_ADAM_WORKER = r'''
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sus_net_b200.train import _FlatAdam
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
torch.manual_seed(0)
make = lambda: torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.PReLU(), torch.nn.Linear(3, 2))
m, single = make(), make()
single.load_state_dict(m.state_dict())
X, Y = torch.randn(10, 4), torch.randn(10, 2)
rows = [slice(0, 7), slice(7, 10)][rank]          # uneven local subsets: 7 and 3 samples
fa = _FlatAdam(torch.optim.Adam(m.parameters(), lr=1e-2))
ref = torch.optim.Adam(single.parameters(), lr=1e-2)
for step in range(4):
    fa.begin_train_step(); fa.begin_view()
    sq = ((m(X[rows]) - Y[rows]) ** 2).sum(dim=1).sum()
    sq.backward()
    loss = fa.apply(torch.tensor(float(rows.stop - rows.start)), sq.detach())
    # single process on the whole batch: mean over the 10 samples (every sample weighs the same)
    ref.zero_grad()
    full = ((single(X) - Y) ** 2).sum(dim=1).mean()
    full.backward()
    ref.step()
    assert abs(float(loss) - float(full)) < 1e-5
    for p, q in zip(m.parameters(), single.parameters()):
        assert torch.allclose(p, q, atol=1e-6), (step, rank)
# a subset that is empty on EVERY rank: no rank moves
before = [p.detach().clone() for p in m.parameters()]
fa.begin_train_step(); fa.begin_view()
fa.apply(torch.tensor(0.0), torch.tensor(0.0))
assert all(torch.equal(p, q) for p, q in zip(before, m.parameters()))
# empty on this rank only: it still takes the other rank's step
fa.begin_train_step(); fa.begin_view()
if rank == 0:
    sq = ((m(X) - Y) ** 2).sum()
    sq.backward()
    fa.apply(torch.tensor(10.0), sq.detach())
else:
    fa.apply(torch.tensor(0.0), torch.tensor(0.0))
flat = fa.flat.clone()
dist.all_reduce(flat)
assert torch.allclose(flat, 2 * fa.flat, atol=1e-7) and not all(torch.equal(p, q) for p, q in zip(before, m.parameters()))
dist.destroy_process_group()
'''


def test_flat_adam_two_ranks_gloo_weights_samples_evenly(S):
    """World size 2 (gloo): gradient SUMS, sample counts and loss sums ride in one all-reduce, so uneven local subsets give
    exactly the single-process step on the union; a globally empty subset moves nothing on any rank."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = str(33000 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, "-c", _FLAT_ADAM_WORKER, root, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_trainer_optimizer_state_roundtrip(S):
    """The device-gated Adam keeps its own moments: they travel through optimizer_state_dict() / load_optimizer_state_dict()."""
    from sus_net_b200.train import _FlatAdam

    torch.manual_seed(1)
    m1, m2 = torch.nn.Linear(5, 3), torch.nn.Linear(5, 3)
    m2.load_state_dict(m1.state_dict())
    t1 = S.DQNTeamTrainer(torch.optim.Adam(m1.parameters(), lr=1e-2), None, 0.9)
    t2 = S.DQNTeamTrainer(torch.optim.Adam(m2.parameters(), lr=1e-2), None, 0.9)
    x, y = torch.randn(7, 5), torch.randn(7, 3)

    def step(tr, model):
        fa = tr._flat(tr.imposter_optimizer)
        fa.begin_train_step(); fa.begin_view()
        sq = ((model(x) - y) ** 2).sum()
        sq.backward()
        fa.apply(torch.tensor(7.0), sq.detach())

    for _ in range(3):
        step(t1, m1)
    sd = t1.optimizer_state_dict()
    assert set(sd) == {"imposter"} and float(sd["imposter"]["step"]) == 3.0
    with torch.no_grad():
        for p, q in zip(m2.parameters(), m1.parameters()):
            p.copy_(q)
    t2._flat(t2.imposter_optimizer)  # flatten first, then restore
    t2.load_optimizer_state_dict(sd)
    step(t1, m1); step(t2, m2)
    assert all(torch.equal(p, q) for p, q in zip(m1.parameters(), m2.parameters()))
