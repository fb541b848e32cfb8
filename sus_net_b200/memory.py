"""Device buffers for the feature tensors in L2-compressible memory (`sus_alloc_compressible`, csrc/susnet_alloc.cu).

The planes / flat rows are almost all zeros; in a compressible allocation the B200's L2 keeps such lines compressed on
their way to and from HBM, which is worth +18 % on the fused kernel's tile-store stream and +35 % on reading the
tensors back (tools/micro/compressible_bench.cu).  torch's own allocator cannot make such allocations, so the featurizers
get their buffers here; small buffers, devices without generic compression and `SUSNET_COMPRESSIBLE=0` use `torch.empty`.
"""
import ctypes as C
import os
import warnings

import numpy as np
import torch

from . import _lib as L

MIN_BYTES = 4 << 20  # below this a 2 MiB-granular private allocation is not worth it
_unsupported = set()  # device indices that refused a compressible allocation
_deferred = []  # (lib, ptr) of blocks whose last tensor died during a CUDA graph capture; freed at the next allocation


class _Block:
    """Owner of one compressible allocation; tensors made from it keep it alive (torch holds a reference to the object
    that exports __cuda_array_interface__ until the storage dies) and the memory is unmapped when the last one goes."""

    def __init__(self, lib, ptr, shape, typestr="<f4"):
        self._lib, self._ptr, self._pid = lib, ptr, os.getpid()
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}

    def __del__(self):
        if os.getpid() != self._pid:  # a forked child (e.g. a DataLoader worker) must not touch the parent's CUDA context
            return
        try:
            if torch.cuda.is_current_stream_capturing():
                _deferred.append((self._lib, self._ptr))  # freeing synchronises the device: not inside a graph capture
            else:
                self._lib.sus_free_compressible(C.c_void_p(self._ptr))
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def compressible_enabled():
    return os.environ.get("SUSNET_COMPRESSIBLE", "1") != "0"


def empty_f32(shape, device):
    """Uninitialised float32 tensor of `shape` on `device`, in L2-compressible memory where that applies."""
    return empty(shape, device, torch.float32)


def empty(shape, device, dtype=torch.float32):
    """Uninitialised float32 / uint8 tensor of `shape` on `device`, in L2-compressible memory where that applies."""
    assert dtype in (torch.float32, torch.uint8)
    typestr, itemsize = ("<f4", 4) if dtype == torch.float32 else ("|u1", 1)
    device = torch.device(device)
    nbytes = itemsize * int(np.prod(shape))
    if compressible_enabled() and nbytes >= MIN_BYTES and device.index not in _unsupported:
        lib = L.lib()
        while _deferred and not torch.cuda.is_current_stream_capturing():
            dlib, dptr = _deferred.pop()
            dlib.sus_free_compressible(C.c_void_p(dptr))
        ptr = C.c_void_p()
        rc = lib.sus_alloc_compressible(device.index, nbytes, C.byref(ptr), None)
        if rc == L.SUS_OK:
            t = torch.as_tensor(_Block(lib, ptr.value, shape, typestr), device=device)
            t._sus_compressible = True
            return t
        if rc != L.SUS_ERR_UNSUPPORTED:  # e.g. the driver refuses virtual-memory allocations in this container: say so once
            warnings.warn(f"sus_alloc_compressible failed ({lib.sus_last_error().decode()}); feature tensors of cuda:{device.index} "
                          "stay in ordinary device memory", RuntimeWarning, stacklevel=2)
        _unsupported.add(device.index)
    return torch.empty(tuple(shape), dtype=dtype, device=device)


def is_compressible(t):
    return bool(getattr(t, "_sus_compressible", False))
