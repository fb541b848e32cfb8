"""Inference of the reference's MLP Q-estimator (src/models/dqn.py:72-108) for all envs in one kernel launch
(`sus_mlp_forward`, csrc/susnet_mlp.cu) -- the acting loop's network evaluation (train.py:367-370) without five GEMM launches
and four activation passes over HBM.  Training still goes through the module itself (autograd); `FusedMLP` reads the module's
LIVE parameter tensors at every call, so it always evaluates the current weights.

    q = FusedMLP(model)(spatial, non_spatial)      # same signature and result as model(spatial, non_spatial), no autograd
"""
import ctypes as C

import torch
from torch import nn

from . import _lib as L


def _stack_of(module):
    """The Linear / activation stack of `module` (the reference's MLP keeps it in `.model`), or None."""
    seq = getattr(module, "model", module)
    if not isinstance(seq, nn.Sequential):
        return None
    layers = list(seq.children())
    if not layers or len(layers) % 2 == 0:
        return None
    linears, acts = layers[0::2], layers[1::2]
    if not all(isinstance(l, nn.Linear) for l in linears) or len(linears) > L.MLP_MAX_LAYERS:
        return None
    if acts:
        kind = type(acts[0])
        if kind not in (nn.PReLU, nn.ReLU) or not all(type(a) is kind for a in acts):
            return None
        if kind is nn.PReLU and not all(a.num_parameters == 1 for a in acts):
            return None
    for a, b in zip(linears[:-1], linears[1:]):
        if a.out_features != b.in_features:
            return None
    return linears, acts


class FusedMLP:
    def __init__(self, module):
        st = _stack_of(module)
        if st is None:
            raise NotImplementedError("FusedMLP restates Linear + PReLU / ReLU stacks (the reference's MLP) only")
        self.module = module
        self.linears, self.acts = st
        self.activation = L.ACT_NONE if not self.acts else (L.ACT_PRELU if isinstance(self.acts[0], nn.PReLU) else L.ACT_RELU)
        self.in_features, self.out_features = self.linears[0].in_features, self.linears[-1].out_features
        self._workspace = None  # device scratch for the repacked weights (owned here: one network, one stream at a time)

    @staticmethod
    def supports(module):
        return module is not None and _stack_of(module) is not None

    def _spec(self):
        spec = L.SusMlpSpec(n_layers=len(self.linears), activation=self.activation)
        spec.dims[0] = self.in_features
        for l, lin in enumerate(self.linears):
            spec.dims[l + 1] = lin.out_features
            assert lin.weight.dtype == torch.float32 and lin.weight.is_contiguous()
            spec.weight[l] = lin.weight.data_ptr()
            spec.bias[l] = lin.bias.data_ptr() if lin.bias is not None else None
            spec.alpha[l] = self.acts[l].weight.data_ptr() if (self.activation == L.ACT_PRELU and l < len(self.acts)) else None
        return spec

    @torch.no_grad()
    def __call__(self, spatial_x, non_spatial_x, out=None):
        """forward(spatial, non_spatial) of the reference's MLP (the spatial input is ignored there too, dqn.py:88-92)."""
        B = spatial_x.size(0) if spatial_x is not None else non_spatial_x.size(0)
        x = non_spatial_x.reshape(B, -1)
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        assert x.shape[1] == self.in_features, f"expected {self.in_features} features per row, got {x.shape[1]}"
        dev = x.device
        if out is None:
            out = torch.empty((B, self.out_features), dtype=torch.float32, device=dev)
        spec = self._spec()
        lib = L.lib()
        if self._workspace is None or self._workspace.device != dev:
            # allocated at the first call (outside any CUDA-graph capture in the training loop, whose warm-up runs eagerly)
            n = int(lib.sus_mlp_workspace_bytes(C.byref(spec)))
            self._workspace = torch.empty(max(n, 16), dtype=torch.uint8, device=dev)
        L.check(lib.sus_mlp_forward_ws(C.byref(spec), C.c_void_p(x.data_ptr()), B, C.c_void_p(out.data_ptr()),
                                       C.c_void_p(self._workspace.data_ptr()), self._workspace.numel(), dev.index,
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out
