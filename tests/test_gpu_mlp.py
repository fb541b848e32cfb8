"""GPU: the one-launch MLP inference kernel (`sus_mlp_forward`, row f2: the Q-network evaluation of the acting loop,
src/train.py:367-370 on src/models/dqn.py:72-108) against the torch module it restates: same float32 arithmetic up to
summation order (tolerance 2e-5 relative to the row's largest |Q|, stated here; the reference's own CPU BLAS differs from
any GPU by as much)."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def make(dims, act):
    layers = []
    for i, d in enumerate(dims[:-1]):
        layers += [nn.Linear(d, dims[i + 1]), act()]
    return nn.Sequential(*layers[:-1])


class RefMLP(nn.Module):  # dqn.py:72-93
    def __init__(self, dims, act=nn.PReLU):
        super().__init__()
        self.model = make(dims, act)

    def forward(self, spatial_x, non_spatial_x):
        return self.model(non_spatial_x.view(spatial_x.size(0), -1))


@pytest.mark.parametrize("dims,act,rows", [
    ([98, 256, 128, 64, 16, 6], nn.PReLU, 131072), ([98, 256, 128, 64, 16, 6], nn.PReLU, 1000), ([196, 32, 16, 6], nn.PReLU, 4099),
    ([98, 24, 5], nn.ReLU, 777), ([36, 200, 100, 7], nn.ReLU, 130), ([4, 6], nn.PReLU, 5), ([78, 64, 6], nn.PReLU, 1),
    ([200, 150, 17, 129, 3], nn.PReLU, 515), ([16, 400, 400, 2], nn.PReLU, 300)])
@pytest.mark.parametrize("tile_rows,k_parts,packed", [("128", "1", "1"), ("128", "1", "0"), ("128", "2", "0"), ("64", "1", "1"),
                                                      ("64", "1", "0"), ("64", "2", "0")])
def test_fused_mlp_matches_the_module(cuda_lib, monkeypatch, dims, act, rows, tile_rows, k_parts, packed):
    """Every geometry of the kernel: 128-row tiles (one CTA per SM) or 64-row tiles (two), one or two k-parts per CTA, weights
    staged from the repacked workspace image (the default) or gathered from the torch layout."""
    import sus_net_b200 as S

    monkeypatch.setenv("SUSNET_MLP_ROWS", tile_rows)  # ([16, 400, 400, 2] only fits as a 64-row tile: chosen automatically)
    monkeypatch.setenv("SUSNET_MLP_SPLIT", k_parts)
    monkeypatch.setenv("SUSNET_MLP_PACKED", packed)

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    torch.manual_seed(len(dims) * 1000 + rows)
    m = RefMLP(dims, act).to(dev)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() > 1:
                p.mul_(2.0)
    T = 2 if dims[0] % 2 == 0 else 1
    ns = (torch.rand(rows, T, dims[0] // T, device=dev) < 0.15).float() * torch.randint(1, 9, (rows, T, dims[0] // T), device=dev)
    sp = torch.zeros(rows, T, 1, device=dev)
    f = S.FusedMLP(m)
    got = f(sp, ns)
    with torch.no_grad():
        want = m.double()(sp.double(), ns.double()).float()
        m.float()
    scale = want.abs().amax(dim=1, keepdim=True).clamp(min=1.0)
    assert got.shape == want.shape and torch.isfinite(got).all()
    assert float(((got - want).abs() / scale).max()) < 2e-5
    # the live parameters are read at every call
    with torch.no_grad():
        m.model[0].weight.mul_(0.5)
        want2 = m(sp, ns)
    got2 = f(sp, ns)
    assert float(((got2 - want2).abs() / want2.abs().amax(dim=1, keepdim=True).clamp(min=1.0)).max()) < 2e-5
    if rows > 1:
        assert not torch.equal(got, got2)


def test_fused_mlp_scope(cuda_lib):
    import sus_net_b200 as S

    assert S.FusedMLP.supports(RefMLP([8, 4, 2]))
    assert not S.FusedMLP.supports(nn.Sequential(nn.Linear(4, 4), nn.Tanh(), nn.Linear(4, 2)))
    assert not S.FusedMLP.supports(nn.Conv2d(1, 1, 1)) and not S.FusedMLP.supports(None)
    with pytest.raises(NotImplementedError):
        S.FusedMLP(nn.Sequential(nn.Linear(4, 4), nn.Tanh(), nn.Linear(4, 2)))
    # the raw entry point refuses a workspace that is too small or misaligned, and runs without one
    import ctypes as C
    from sus_net_b200 import _lib as L
    f = S.FusedMLP(RefMLP([8, 4, 2]).to("cuda"))
    spec = f._spec()
    need = int(L.lib().sus_mlp_workspace_bytes(C.byref(spec)))
    assert need == 4 * (16 * 16 + 16 * 16)  # two layers, each one 16-column block x one 16-k chunk
    x, out = torch.ones(3, 8, device="cuda"), torch.zeros(3, 2, device="cuda")
    ws = torch.empty(need + 16, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (C.byref(spec), C.c_void_p(x.data_ptr()), 3, C.c_void_p(out.data_ptr()))
    assert L.lib().sus_mlp_forward_ws(*args, C.c_void_p(ws.data_ptr()), need - 1, 0, st) == L.SUS_ERR_INVALID_ARGUMENT
    assert L.lib().sus_mlp_forward_ws(*args, C.c_void_p(ws.data_ptr() + 4), need, 0, st) == L.SUS_ERR_INVALID_ARGUMENT
    assert L.lib().sus_mlp_forward_ws(*args, C.c_void_p(ws.data_ptr()), need, 0, st) == 0
    a = out.clone()
    assert L.lib().sus_mlp_forward_ws(*args, None, 0, 0, st) == 0 and torch.equal(a, out)
    assert torch.allclose(out, f.module.model(x), atol=1e-6)
    big = RefMLP([16, 800, 800, 2]).to("cuda")  # two adjacent 800-wide layers: 410 KB of activations even per 64-row tile
    with pytest.raises(NotImplementedError):
        S.FusedMLP(big)(torch.zeros(4, 1, 1, device="cuda"), torch.zeros(4, 1, 16, device="cuda"))
