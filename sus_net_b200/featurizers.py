"""Host-side mirror of Sus-Net's observation featurizers (src/features/component.py, model_ready.py).

The classes keep the reference's names and methods -- `GlobalFeaturizer`, `PerspectiveFeaturizer`,
`FlatFeaturizer` with `fit(state_sequence (B, T, S))`, `generate_featurized_states()` and `featurized_shape`;
component featurizers (`OneHotAgentPositionFeaturizer`, ...) and `CompositeFeaturizer` -- but the work is one
CUDA kernel launch over all B*T states instead of Python loops over states, agents and jobs.

Extra (no reference analogue): `encode_env(env)` featurizes the live state of every env of a batched env (T = 1)
and `env.step(actions, featurizer=f)` fuses that into the step kernel.
"""
import ctypes as C
from enum import StrEnum, auto

import numpy as np
import torch

from . import _lib as L
from .env import StateFields, _ptr, _TORCH_TO_SUS
from .memory import empty as empty_dev
from .memory import empty_f32


# ---------------------------------------------------------------------------------------------- components
class ComponentFeaturizer:
    """Declarative stand-in for component.py's per-state featurizers: `code` selects the device routine."""

    code = None

    def __init__(self, env):
        self.env = env

    def _size(self):
        raise NotImplementedError

    @property
    def shape(self):
        return torch.tensor([self._size()], dtype=torch.int)

    def extract_features(self, state):
        """Featurize ONE state tuple (reference signature, component.py:27) -> 1-D float32 tensor (on the device)."""
        flat = torch.as_tensor(self.env.flatten_state(state)).reshape(1, 1, -1)
        f = FlatFeaturizer(self.env, CompositeFeaturizer([self]))
        f.fit(flat)
        return f.featurized_state[0, 0]


def _component(name, code, size_fn, anchor):
    cls = type(name, (ComponentFeaturizer,), {"code": code, "_size": lambda self: size_fn(self.env),
                                              "__doc__": f"GPU `{name}` ({anchor})."})
    return cls


OneHotAgentPositionFeaturizer = _component("OneHotAgentPositionFeaturizer", 0, lambda e: e.n_agents * 18, "component.py:221-247")
CoordinateAgentPositionsFeaturizer = _component("CoordinateAgentPositionsFeaturizer", 1, lambda e: e.n_agents * 2, "component.py:384-403")
AliveCrewFeaturizer = _component("AliveCrewFeaturizer", 2, lambda e: e.n_agents - 1, "component.py:406-425")
ClosestAliveCrewFeaturizer = _component("ClosestAliveCrewFeaturizer", 3, lambda e: e.n_crew, "component.py:455-482")
L1CrewFeaturizer = _component("L1CrewFeaturizer", 4, lambda e: e.n_crew, "component.py:428-452")
DistanceToImposterFeaturizer = _component("DistanceToImposterFeaturizer", 5, lambda e: (e.n_agents - 1) * 2, "component.py:250-278")
WallsFeaturizer = _component("WallsFeaturizer", 6, lambda e: 9, "component.py:281-300")
ImposterVSCrewRoomLocaionFeaturizer = _component("ImposterVSCrewRoomLocaionFeaturizer", 7, lambda e: 8, "component.py:303-334")
ImposterScentFeaturizer = _component("ImposterScentFeaturizer", 8, lambda e: 4, "component.py:339-380")

_STATE_FIELD_CODES = {"ALIVE_AGENTS": 9, "JOB_STATUS": 10, "USED_TAGS": 11, "TAG_COUNTS": 12}


class StateFieldFeaturizer(ComponentFeaturizer):
    """GPU `StateFieldFeaturizer` (component.py:200-218) for the 1-D state fields."""

    def __init__(self, env, state_field):
        super().__init__(env)
        self.state_field = state_field
        if state_field.name not in _STATE_FIELD_CODES:
            raise ValueError(f"state field {state_field} is not a 1-D per-agent / per-job field")
        self.code = _STATE_FIELD_CODES[state_field.name]

    def _size(self):
        return self.env.n_jobs if self.state_field.name == "JOB_STATUS" else self.env.n_agents


class _SpatialComponent(ComponentFeaturizer):
    """The two spatial components of component.py (used by the reference only inside Global / Perspective): featurize ONE
    state tuple with the Global plane kernel and slice the planes."""

    def _planes(self, state):
        f = GlobalFeaturizer(self.env)
        f.fit(torch.as_tensor(self.env.flatten_state(state)).reshape(1, 1, -1))
        return f.spatial[0, 0]  # (A + 2, 9, 9)


class AgentPositionsFeaturizer(_SpatialComponent):
    """GPU `AgentPositionsFeaturizer` (component.py:83-106): one plane per agent, 1 at (x, y) if alive."""

    def extract_features(self, agent_state, alive_only=True):
        return self._planes(agent_state)[: self.env.n_agents].clone()

    @property
    def shape(self):
        return torch.tensor([self.env.n_agents, self.env.n_cols, self.env.n_rows], dtype=torch.int)


class JobFeaturizer(_SpatialComponent):
    """GPU `JobFeaturizer` (component.py:109-131): plane 0 = incomplete jobs, plane 1 = done jobs."""

    def extract_features(self, agent_state):
        return self._planes(agent_state)[self.env.n_agents:].clone()

    @property
    def shape(self):
        return torch.tensor([2, self.env.n_cols, self.env.n_rows], dtype=torch.int)


class CompositeFeaturizer(ComponentFeaturizer):
    """`CompositeFeaturizer` (component.py:134-159): concatenation of flat components in list order."""

    def __init__(self, featurizers):
        assert len(featurizers) > 0, "No featurizers provided."
        self.featurizers = list(featurizers)
        self.env = self.featurizers[0].env

    def _size(self):
        return sum(f._size() for f in self.featurizers)

    def codes(self):
        return [f.code for f in self.featurizers]

    def __repr__(self):
        return str([f.__class__.__name__ for f in self.featurizers])


# ---------------------------------------------------------------------------------------------- model-ready
class FeaturizerType(StrEnum):  # model_ready.py:20-37
    PERPSECTIVE = auto()
    GLOBAL = auto()
    FLAT = auto()

    @staticmethod
    def build(featurizer_type, env, **kwargs):
        assert featurizer_type in [f.value for f in FeaturizerType], f"Invalid featurizer type: {featurizer_type}"
        if featurizer_type == FeaturizerType.PERPSECTIVE:
            return PerspectiveFeaturizer(env=env)
        if featurizer_type == FeaturizerType.GLOBAL:
            return GlobalFeaturizer(env=env)
        featurizers = kwargs.get("featurizers", None)
        assert featurizers is not None, "Need to provide a featurizer for FlatFeaturizer."
        return FlatFeaturizer(env=env, featurizer=featurizers)


class SequenceStateFeaturizer:
    """model_ready.py:40-79."""

    _KIND = L.ENCODE_NONE

    def __init__(self, env, clone_views=False, output_device=None, plane_dtype=torch.float32):
        """plane_dtype=torch.uint8 (opt-in, Global / Perspective only; no reference analogue): the spatial tensors hold one
        byte per cell instead of one float32 -- same shape and order, 4x fewer bytes to write and to read back -- for a
        network that casts in its first layer.  The returned spatial views are then uint8 tensors without `requires_grad`."""
        assert plane_dtype in (torch.float32, torch.uint8)
        self.plane_dtype = plane_dtype
        self.env = env
        self.state_size = env.flattened_state_size
        self.clone_views = clone_views  # True: every agent view owns its memory like the reference's .clone()
        # None: views stay on the env's GPU (for a Q-network on the GPU); "cpu": views are copied to the host, which
        # is what the reference's unmodified CPU models in src/train.py expect
        self.output_device = torch.device(output_device) if output_device is not None else None
        self._spec = self._make_spec()
        shape = L.SusEncodeShape()
        L.check(env.lib.sus_encode_shape(C.byref(env._cfg), C.byref(self._spec), C.byref(shape)))
        self._shape = shape
        self._sp_buf = self._ns_buf = None
        self._buffers = {}  # n_items -> (spatial, non_spatial): acting (N envs) and training (batch) sizes alternate
        self.B = self.T = None

    def _make_spec(self):
        return L.SusEncodeSpec(kind=self._KIND, n_components=0,
                               flags=L.ENCODE_PLANES_U8 if self.plane_dtype == torch.uint8 else 0)

    def _alloc(self, n_items):
        sh, dev = self._shape, self.env.device
        if n_items not in self._buffers:
            if len(self._buffers) >= 4:
                self._buffers.pop(next(iter(self._buffers)))
            # the output tensors are almost all zeros: L2-compressible memory where the device has it (memory.py)
            sp = empty_dev((sh.spatial_views, n_items, sh.spatial_floats), dev, self.plane_dtype) if sh.spatial_views else None
            ns = empty_f32((sh.non_spatial_views, n_items, sh.non_spatial_floats), dev)
            self._buffers[n_items] = (sp, ns)
        self._sp_buf, self._ns_buf = self._buffers[n_items]

    def clone_for(self, env=None):
        """A second featurizer of the same kind with its OWN output buffers (e.g. one for acting, one for replay batches)."""
        import copy

        f = copy.copy(self)
        f.env = env or self.env
        f._buffers = {}
        f._sp_buf = f._ns_buf = None
        f.B = f.T = None
        return f

    def new_buffers(self, n_items):
        """A fresh (spatial, non_spatial) output pair for `n_items` items (callers that double-buffer, e.g. `HostStepper`)."""
        sh, dev = self._shape, self.env.device
        sp = empty_dev((sh.spatial_views, n_items, sh.spatial_floats), dev, self.plane_dtype) if sh.spatial_views else None
        return sp, empty_f32((sh.non_spatial_views, n_items, sh.non_spatial_floats), dev)

    def bind_buffers(self, sp, ns):
        """Make `(sp, ns)` (from `new_buffers`) the output pair the next fit / encode / fused step of that size writes."""
        self._buffers[ns.shape[1]] = (sp, ns)
        self._sp_buf, self._ns_buf = sp, ns

    def fit(self, state_sequence):
        """Featurize a (B, T, S) batch of flattened states (train.py:70-74,346-348) in one kernel launch."""
        if not isinstance(state_sequence, torch.Tensor):
            state_sequence = torch.as_tensor(np.asarray(state_sequence))
        assert state_sequence.dim() == 3, f"Expected 3D tensor. Got: {state_sequence.dim()}"
        self.B, self.T, S = state_sequence.size()
        assert S == self.state_size, f"Expected state size {self.state_size}. Got: {S}"
        x = state_sequence.to(self.env.device)
        if x.dtype not in (torch.float32, torch.float64, torch.int64):
            x = x.to(torch.float32)
        x = x.contiguous()
        n = self.B * self.T
        self._alloc(n)
        env = self.env
        L.check(env.lib.sus_encode_from_flat(C.byref(env._cfg), C.byref(self._spec), _ptr(x), _TORCH_TO_SUS[x.dtype], n,
                                             _ptr(self._sp_buf), _ptr(self._ns_buf), env.device.index, env._stream()))

    def encode_env(self, env=None):
        """Featurize the CURRENT state of every env (B = num_envs, T = 1) straight from the device state."""
        env = env or self.env
        self.B, self.T = env.num_envs, 1
        self._alloc(env.num_envs)
        L.check(env.lib.sus_env_encode(env._h, C.byref(self._spec), _ptr(self._sp_buf), _ptr(self._ns_buf), env._stream()))
        return self.generate_featurized_states()

    def _bind_for_fused_step(self, env):
        assert env is self.env or env._S == self.env._S
        self.B, self.T = env.num_envs, 1
        self._alloc(env.num_envs)
        return self._spec

    def _views(self):
        """(spatial (A or 1, B, T, C, 9, 9) or None, non_spatial (views, B, T, F)) tensor views of the buffers."""
        sh, A = self._shape, self.env.n_agents
        sp = None
        if self._sp_buf is not None:
            sp = self._sp_buf.view(sh.spatial_views, self.B, self.T, A + 2, 9, 9)
        ns = self._ns_buf.view(sh.non_spatial_views, self.B, self.T, sh.non_spatial_floats)
        return sp, ns

    def stacked_views(self):
        """The featurized tensors of the last fit / encode without per-agent slicing: `(spatial, non_spatial)` with
        shapes (Vs, B, T, C, 9, 9) or None and (Vn, B, T, F); Vs / Vn are 1 where all agent views share the tensor."""
        sp, ns = self._views()
        return sp, ns

    def _leaf(self, t):
        t = t.detach()
        if self.output_device is not None and t.device != self.output_device:
            t = t.to(self.output_device)
        elif self.clone_views:
            t = t.clone()
        return t.requires_grad_(True) if t.is_floating_point() else t


class GlobalFeaturizer(SequenceStateFeaturizer):
    """GPU `GlobalFeaturizer` (model_ready.py:219-306)."""

    _KIND = L.ENCODE_GLOBAL

    @property
    def featurized_shape(self):
        A = self.env.n_agents
        return torch.tensor([A + 2, 9, 9], dtype=torch.int), torch.tensor([self._shape.non_spatial_floats], dtype=torch.int)

    @property
    def spatial(self):
        return self._views()[0][0]

    def generate_featurized_states(self):
        sp, ns = self._views()
        return [(self._leaf(sp[0]), self._leaf(ns[k])) for k in range(self.env.n_agents)]


class PerspectiveFeaturizer(SequenceStateFeaturizer):
    """GPU `PerspectiveFeaturizer` (model_ready.py:82-216)."""

    _KIND = L.ENCODE_PERSPECTIVE

    @property
    def featurized_shape(self):
        A = self.env.n_agents
        return torch.tensor([A + 2, 9, 9], dtype=torch.int), torch.tensor([self._shape.non_spatial_floats], dtype=torch.int)

    def generate_featurized_states(self):
        sp, ns = self._views()
        return [(self._leaf(sp[k]), self._leaf(ns[k])) for k in range(self.env.n_agents)]


class FlatFeaturizer(SequenceStateFeaturizer):
    """GPU `FlatFeaturizer` (model_ready.py:309-370) over a `CompositeFeaturizer` of flat components."""

    _KIND = L.ENCODE_FLAT

    def __init__(self, env, featurizer, clone_views=False, output_device=None):
        self.featurizer = featurizer if isinstance(featurizer, CompositeFeaturizer) else CompositeFeaturizer(list(featurizer))
        super().__init__(env, clone_views=clone_views, output_device=output_device)

    def _make_spec(self):
        codes = self.featurizer.codes()
        spec = L.SusEncodeSpec(kind=self._KIND, n_components=len(codes), flags=0)
        for i, c in enumerate(codes):
            spec.components[i] = c
        return spec

    @property
    def featurized_shape(self):
        return (1, self.featurizer.shape)  # model_ready.py:318-323

    @property
    def featurized_state(self):
        return self._views()[1][0]

    def generate_featurized_states(self):
        _, ns = self._views()
        out = []
        for _k in range(self.env.n_agents):  # every view is identical (model_ready.py:356-367)
            dev = self.output_device or self.env.device
            out.append((torch.zeros(self.B, self.T, 1, device=dev).requires_grad_(True), self._leaf(ns[0])))
        return out

    def __repr__(self):
        return f"FlatFeaturizer_{self.featurizer}"
