"""GPU: the replay ring (row f1) against fixtures written by the REFERENCE's own ReplayBuffer (tools/make_golden.py
replay_fixture): layout, dtypes, ring wrap-around and the T-deep np.roll sequence bookkeeping."""
import numpy as np
import pytest
import torch

from tests.util import CASES, case_of, golden_files, load, make_cuda_env

pytestmark = pytest.mark.gpu


def replay_files():
    import glob
    import os

    from tests.util import GOLDEN

    return sorted(glob.glob(os.path.join(GOLDEN, "*.replay_T*.npz")))


@pytest.mark.parametrize("path", replay_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_replay_ring_matches_reference_replay_buffer(cuda_lib, path):
    import sus_net_b200 as S

    g = load(path)
    cfg = CASES[case_of(path)]
    T, M, N = int(g["T"]), int(g["M"]), int(g["n_envs"])
    env = make_cuda_env(cfg, N, seed=int(g["seed"]), env_id_base=int(g["env_id_base"]))
    env.reset()
    buf = S.ReplayBuffer(max_size=M, state_size=env.flattened_state_size, trajectory_size=T, n_agents=env.n_agents,
                         n_imposters=env.n_imposters, device=env.device)
    assert buf.states.dtype == torch.float32 and buf.actions.dtype == torch.int64 and buf.dones.dtype == torch.bool
    assert buf.imposters.dtype == torch.int16 and tuple(buf.dones.shape) == (M, 1)
    for t_ in (buf.states, buf.next_states, buf.rewards, buf.actions, buf.dones, buf.imposters):
        t_.zero_()
    buf.attach(env)
    for t in range(g["actions"].shape[0]):
        sampled = env.sample_actions()
        assert np.array_equal(sampled.cpu().numpy(), g["actions"][t])  # same trajectory as the reference run
        buf.collect_step(torch.as_tensor(g["actions"][t].astype(np.int32)))
    env.check_actions()
    assert buf.idx == int(g["idx"]) and buf.size == int(g["size"])
    assert np.array_equal(buf.states.cpu().numpy(), g["states"])
    assert np.array_equal(buf.next_states.cpu().numpy(), g["next_states"])
    assert np.array_equal(buf.actions.cpu().numpy(), g["r_actions"])
    assert np.array_equal(buf.rewards.cpu().numpy(), g["rewards"])
    assert np.array_equal(buf.dones.cpu().numpy(), g["dones"])
    assert np.array_equal(buf.imposters.cpu().numpy(), g["imposters"])
    assert np.array_equal(buf.state_sequence.cpu().numpy(), g["final_seq"])
    b = buf.sample(16)
    assert tuple(b.states.shape) == (16, T, env.flattened_state_size) and tuple(b.actions.shape) == (16, env.n_agents)
    assert b._fields == ("states", "actions", "rewards", "next_states", "imposters", "dones")


def test_populate_random_policy_and_single_env_add(cuda_lib):
    """populate() with the fused random policy stores exactly the transitions a manual loop stores; the
    reference-mode populate fills one transition per step like the reference."""
    import sus_net_b200 as S

    cfg = CASES["cfg4_base_1v4"]
    N, M, T = 500, 2000, 2
    env = make_cuda_env(cfg, N, seed=4)
    buf = S.ReplayBuffer(M, env.flattened_state_size, T, env.n_agents, env.n_imposters, device=env.device)
    added = buf.populate(env, 1700)
    assert added == 2000 and buf.size == 2000 and buf.idx == 0
    twin = make_cuda_env(cfg, N, seed=4)
    twin.reset()
    seq = twin.flat_states()[:, None, :].repeat(1, T, 1)
    for step in range(4):
        twin.emit_imposters = True
        nf, r, d, tr, _ = twin.step(None)
        sl = slice(step * N, (step + 1) * N)
        nxt = torch.cat([seq[:, 1:], nf[:, None]], dim=1)
        assert torch.equal(buf.states[sl], seq) and torch.equal(buf.next_states[sl], nxt)
        assert torch.equal(buf.rewards[sl], r) and torch.equal(buf.dones[sl, 0], d)
        assert torch.equal(buf.actions[sl], twin._last_actions.long())
        assert torch.equal(buf.imposters[sl], twin._imposters_buf)
        fin = (d | tr)[:, None, None]
        seq = torch.where(fin, twin.flat_states()[:, None, :].expand(-1, T, -1), nxt)
    # reference mode
    single = S.FourRoomEnv(1, 4, 5, random_state=1)
    sbuf = S.ReplayBuffer(64, single.flattened_state_size, 3, 5, 1, device=single.device)
    assert sbuf.populate(single, 40) == 40 and sbuf.size == 40
    assert torch.equal(sbuf.next_states[:39, -1][~sbuf.dones[:39, 0]], sbuf.states[1:40, -1][~sbuf.dones[:39, 0]])
