"""GPU: checkpoint / resume of the FULL env state (advisor finding of round 1): besides the state arrays and the Philox ticks,
`state_dict()` carries the finished-episode statistics, the pending invalid-action counter and the tracked returns
(per-agent G, the two return sums, gamma); a resumed env continues exactly like the original one."""
import numpy as np
import pytest
import torch

from tests.util import CASES, make_cuda_env

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,track", [("cfg4_base_1v4", True), ("cfg3_tagging_1v2", True), ("cfg2_itg_1v1_wall", False)])
def test_state_dict_restores_statistics_and_tracked_returns(cuda_lib, name, track):
    cfg = dict(CASES[name], max_time_steps=min(CASES[name]["max_time_steps"], 37))
    N = 1234
    a = make_cuda_env(cfg, N, seed=3, env_id_base=10)
    a.reset()
    if track:
        a.track_returns(0.9)
    for _ in range(60):
        a.step(None)
    sd = a.state_dict()
    assert int(sd["stats"][0]) > 0 and sd["aux"][0] is not None and (sd["aux"][2] is not None) == track
    b = make_cuda_env(cfg, N, seed=3, env_id_base=10)  # a fresh env: nothing but the checkpoint
    b.load_state_dict(sd)
    assert torch.equal(a.episode_stats(), b.episode_stats())
    if track:
        assert torch.equal(a.return_sums(), b.return_sums())
    for _ in range(45):
        ra, rb = a.step(None), b.step(None)
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2])
    assert torch.equal(a.episode_stats(), b.episode_stats()) and int(b.episode_stats()[0]) > int(sd["stats"][0])
    if track:  # float64 atomics over the same finished episodes in both envs: equal up to summation order
        sa, sb = a.return_sums().cpu().numpy(), b.return_sums().cpu().numpy()
        assert np.allclose(sa, sb, rtol=1e-12, atol=1e-9) and np.abs(sa).sum() > 0
        pa, na = a._aux_arrays()
        from sus_net_b200.env import _as_device_bytes

        ga = _as_device_bytes(pa[2], na[2] - 16, a.device).view(torch.float64)
        pb, nb = b._aux_arrays()
        gb = _as_device_bytes(pb[2], nb[2] - 16, b.device).view(torch.float64)
        assert torch.equal(ga, gb)  # the per-agent running returns G are bit-identical
    # an older checkpoint (statistics only) still restores them
    c = make_cuda_env(cfg, N, seed=3, env_id_base=10)
    c.load_state_dict({k: v for k, v in sd.items() if k not in ("aux", "gamma")})
    assert torch.equal(c.episode_stats().cpu(), sd["stats"])
