"""Compact host protocol (include/susnet_b200.h, `SusCompactLayout`): what crosses PCIe per env-step when the policy
or the consumer of the rewards lives on the host.

The reference hands `step()`'s numpy arrays to a caller in the same process (src/train.py:383-399); a GPU env has to move
them over the host link, and at 1 Mi envs the dense form (int32 / uint8 actions in, float32 rewards + done + truncated
out: 27 B per env-step) saturates it.  A step's reward is one of a handful of values, so the kernel can emit it as a
small code next to the done / truncated bits, and actions need 3-4 bits each:

    actions  (N, action_bytes) uint8   agent i's role-list index in bits [i*action_bits, (i+1)*action_bits)
    results  (N, result_bytes) uint8   agent i's reward code in bits [i*reward_bits, (i+1)*reward_bits),
                                       done in bit A*reward_bits, truncated in bit A*reward_bits + 1

(little-endian bit order; cfg4, `FourRoomEnv` 1v4: 2 B in, 3 B out).  `decode()` maps the codes through the float64 table
of `sus_reward_lut`, which the library computes with the same float64 operation sequence as the kernel: decoded rewards
are bit-identical to the dense float64 rewards for arbitrary reward constants (GPU test on every golden case).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


class CompactProtocol:
    def __init__(self, env):
        self._lib, self._cfg = env.lib, env._cfg
        self.n_agents = A = env.n_agents
        lay = L.SusCompactLayout()
        L.check(env.lib.sus_compact_layout(C.byref(env._cfg), C.byref(lay)))
        self.action_bits, self.action_bytes = lay.action_bits, lay.action_bytes
        self.reward_bits, self.result_bytes = lay.reward_bits, lay.result_bytes
        self.n_codes, self.invalid_code = lay.n_codes, lay.invalid_code
        lut = np.empty((A, lay.invalid_code + 1), dtype=np.float64)
        L.check(env.lib.sus_reward_lut(C.byref(env._cfg), lut.ctypes.data_as(C.POINTER(C.c_double))))
        self.lut = lut  # (A, 2**reward_bits) float64; NaN past n_codes

    # ------------------------------------------------------------------ actions
    def pack_actions(self, actions):
        """(N, A) integer role-list indices (numpy / CPU or CUDA tensor) -> (N, action_bytes) uint8 of the same kind."""
        if isinstance(actions, torch.Tensor):
            a = actions.to(torch.int64)
            rec = torch.zeros(a.shape[0], dtype=torch.int64, device=a.device)
            for i in range(self.n_agents):
                rec |= a[:, i] << (i * self.action_bits)
            return torch.stack([(rec >> (8 * b)) & 0xFF for b in range(self.action_bytes)], dim=1).to(torch.uint8)
        a = np.ascontiguousarray(actions)
        if a.dtype in (np.uint8, np.int32, np.int64) and a.ndim == 2:  # the library's host packer (threads over N)
            out = np.empty((a.shape[0], self.action_bytes), dtype=np.uint8)
            dt = {np.dtype(np.uint8): L.U8, np.dtype(np.int32): L.I32, np.dtype(np.int64): L.I64}[a.dtype]
            L.check(self._lib.sus_host_pack_actions(C.byref(self._cfg), a.ctypes.data_as(C.c_void_p), dt, a.shape[0],
                                                    out.ctypes.data_as(C.c_void_p), 0))
            return out
        a = a.astype(np.uint64)
        rec = np.zeros(a.shape[0], dtype=np.uint64)
        for i in range(self.n_agents):
            rec |= a[:, i] << np.uint64(i * self.action_bits)
        return np.stack([(rec >> np.uint64(8 * b)) & np.uint64(0xFF) for b in range(self.action_bytes)], axis=1).astype(np.uint8)

    def unpack_actions(self, packed):
        rec = self._records(packed)
        m = np.uint64((1 << self.action_bits) - 1)
        return np.stack([(rec >> np.uint64(i * self.action_bits)) & m for i in range(self.n_agents)], axis=1).astype(np.int64)

    # ------------------------------------------------------------------ results
    @staticmethod
    def _records(packed):
        if isinstance(packed, torch.Tensor):
            packed = packed.detach().cpu().numpy()
        p = np.asarray(packed, dtype=np.uint8)
        rec = np.zeros(p.shape[0], dtype=np.uint64)
        for b in range(p.shape[1]):
            rec |= p[:, b].astype(np.uint64) << np.uint64(8 * b)
        return rec

    def codes(self, results):
        """(N, result_bytes) uint8 -> (N, A) reward codes."""
        rec = self._records(results)
        m = np.uint64((1 << self.reward_bits) - 1)
        return np.stack([(rec >> np.uint64(i * self.reward_bits)) & m for i in range(self.n_agents)], axis=1).astype(np.int64)

    def decode(self, results, dtype=np.float64):
        """(N, result_bytes) uint8 records -> (rewards (N, A) `dtype`, dones (N,) bool, truncated (N,) bool)."""
        if isinstance(results, torch.Tensor):
            results = results.detach().cpu().numpy()
        results = np.ascontiguousarray(results, dtype=np.uint8)
        if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.float64)):  # the library's host decoder (threads over N)
            n = results.shape[0]
            rewards = np.empty((n, self.n_agents), dtype=dtype)
            done, trunc = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint8)
            L.check(self._lib.sus_host_decode_results(
                C.byref(self._cfg), results.ctypes.data_as(C.c_void_p), n, rewards.ctypes.data_as(C.c_void_p),
                L.F32 if np.dtype(dtype) == np.dtype(np.float32) else L.F64, done.ctypes.data_as(C.c_void_p),
                trunc.ctypes.data_as(C.c_void_p), 0))
            return rewards, done.view(bool), trunc.view(bool)
        return self.decode_numpy(results, dtype)

    def decode_numpy(self, results, dtype=np.float64):
        """The same decode in numpy (the written spec of the record format; tests compare both)."""
        rec = self._records(results)
        A, rb = self.n_agents, self.reward_bits
        m = np.uint64((1 << rb) - 1)
        rewards = np.empty((rec.shape[0], A), dtype=np.float64)
        for i in range(A):
            rewards[:, i] = self.lut[i][((rec >> np.uint64(i * rb)) & m).astype(np.int64)]
        done = ((rec >> np.uint64(A * rb)) & np.uint64(1)).astype(bool)
        trunc = ((rec >> np.uint64(A * rb + 1)) & np.uint64(1)).astype(bool)
        return rewards.astype(dtype, copy=False), done, trunc

    def flags(self, results):
        rec = self._records(results)
        A, rb = self.n_agents, self.reward_bits
        return ((rec >> np.uint64(A * rb)) & np.uint64(1)).astype(bool), ((rec >> np.uint64(A * rb + 1)) & np.uint64(1)).astype(bool)
