"""Host-buffer driver for a batched env: actions arrive in (pinned) host memory, rewards / dones / truncations go
back to pinned host memory, and the three legs of a step -- H2D copy, fused step(+encode) kernel, D2H copy -- run on
three CUDA streams over rotating slots, so consecutive steps overlap (PCIe is full duplex).  The feature tensors
stay on the device for the consumer (the Q-network).  No reference analogue: the reference is a host-only loop
(src/train.py:383-399 hands numpy arrays to a caller in the same process).

Two wire formats:

* ``protocol="compact"`` (default): the compact host protocol of `sus_net_b200.compact` -- bit-packed role-list indices
  in ((N, action_bytes) uint8, 2 B per env at cfg4), reward codes + done / truncated bits out ((N, result_bytes) uint8,
  3 B per env at cfg4).  `wait(slot)` returns a `CompactResult` whose `rewards` / `dones` / `truncated` decode lazily on
  the host through the float64 table (bit-identical to the dense float64 rewards).
* ``protocol="dense"``: `actions_dtype` role-list indices in, `[rewards (N, A) f32 | dones (N,) | truncated (N,)]` out
  (27 B per env-step at cfg4 with uint8 actions).

Feature buffers: with a featurizer every slot owns its own output tensors, so the fused kernel of step k+1 never writes
the tensors step k's consumer is still reading.  Contract for a consumer on another stream:
`features(slot)` makes the CURRENT stream wait for the kernel that filled the slot and returns the featurizer's views;
`release_features(slot)` (called on the consumer's stream when it is done reading) is what the next kernel into that
slot waits for.  A consumer that never calls `release_features` must synchronise itself before the slot comes round
again (`slots` steps later).
"""
import torch


class CompactResult:
    """Results of one step in pinned host memory, still packed; decoding is lazy (numpy, float64 table lookup)."""

    def __init__(self, protocol, records):
        self.protocol, self.records = protocol, records  # records: (N, result_bytes) uint8 pinned CPU tensor
        self._decoded = None

    def _decode(self):
        if self._decoded is None:
            self._decoded = self.protocol.decode(self.records.numpy())
        return self._decoded

    @property
    def rewards(self):
        return self._decode()[0]

    @property
    def dones(self):
        return self._decode()[1]

    @property
    def truncated(self):
        return self._decode()[2]

    def __iter__(self):  # (rewards, dones, truncated) like the dense protocol
        return iter(self._decode())


class HostStepper:
    def __init__(self, env, featurizer=None, slots=2, actions_dtype=torch.uint8, protocol="compact"):
        """actions_dtype (dense protocol only): dtype of the host action rows (uint8 = 1 byte per agent on the PCIe link,
        role-list indices are < 256; int32 / int64 are accepted too, the reference's loop uses np.int32)."""
        assert env.batched, "HostStepper drives batched envs"
        assert protocol in ("compact", "dense")
        self.env, self.featurizer, self.slots, self.protocol = env, featurizer, slots, protocol
        dev, N, A = env.device, env.num_envs, env.n_agents
        self.s_h2d, self.s_run, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
        self.compact = env.compact if protocol == "compact" else None
        if self.compact is not None:
            cp = self.compact
            self.d_actions = [torch.empty((N, cp.action_bytes), dtype=torch.uint8, device=dev) for _ in range(slots)]
            self.d_block = [torch.empty((N, cp.result_bytes), dtype=torch.uint8, device=dev) for _ in range(slots)]
            self.h_block = [torch.empty((N, cp.result_bytes), dtype=torch.uint8).pin_memory() for _ in range(slots)]
            self.h_out = [CompactResult(cp, b) for b in self.h_block]
        else:
            assert actions_dtype in (torch.uint8, torch.int32, torch.int64)
            self.d_actions = [torch.empty((N, A), dtype=actions_dtype, device=dev) for _ in range(slots)]
            # per slot ONE packed block [rewards (N, A) f32 | dones (N,) | truncated (N,)] on the device and in pinned host
            # memory: the results of a step leave the device in a single D2H copy
            nb_r = N * A * 4

            def views(block):
                return (block[:nb_r].view(torch.float32).view(N, A), block[nb_r:nb_r + N].view(torch.bool),
                        block[nb_r + N:nb_r + 2 * N].view(torch.bool))

            self.d_block = [torch.empty(nb_r + 2 * N, dtype=torch.uint8, device=dev) for _ in range(slots)]
            self.h_block = [torch.empty(nb_r + 2 * N, dtype=torch.uint8).pin_memory() for _ in range(slots)]
            self.d_out = [views(b) for b in self.d_block]
            self.h_out = [views(b) for b in self.h_block]
        self.feat_bufs = None
        if featurizer is not None:  # one output pair per slot (see the module docstring)
            self.feat_bufs = [featurizer.new_buffers(N) for _ in range(slots)]
        self.ev_h2d = [torch.cuda.Event() for _ in range(slots)]
        self.ev_run = [torch.cuda.Event() for _ in range(slots)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(slots)]
        self.ev_consumed = [None] * slots
        self.k = 0
        start = torch.cuda.current_stream(dev)
        for s in (self.s_h2d, self.s_run, self.s_d2h):
            s.wait_stream(start)

    @property
    def h2d_bytes_per_step(self):
        return self.d_actions[0].numel() * self.d_actions[0].element_size()

    @property
    def d2h_bytes_per_step(self):
        return self.d_block[0].numel()

    def step(self, host_actions):
        """Enqueue one step on `host_actions` (compact: (N, action_bytes) uint8 from `env.compact.pack_actions`; dense:
        (N, A) of `actions_dtype`; ideally pinned).  Returns the slot whose pinned host buffers will hold the results
        once `wait(slot)` returns."""
        slot = self.k % self.slots
        first_use = self.k < self.slots
        with torch.cuda.stream(self.s_h2d):
            if not first_use:
                self.s_h2d.wait_event(self.ev_run[slot])  # the kernel that read this action buffer has finished
            self.d_actions[slot].copy_(host_actions, non_blocking=True)
            self.ev_h2d[slot].record(self.s_h2d)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_h2d[slot])
            if not first_use:
                self.s_run.wait_event(self.ev_d2h[slot])  # the previous results of this slot have left the device
            if self.ev_consumed[slot] is not None:
                self.s_run.wait_event(self.ev_consumed[slot])  # the consumer is done with this slot's feature tensors
                self.ev_consumed[slot] = None
            if self.feat_bufs is not None:
                self.featurizer.bind_buffers(*self.feat_bufs[slot])
            if self.compact is not None:
                self.env.step(self.d_actions[slot], featurizer=self.featurizer, packed_actions=True,
                              packed_out=self.d_block[slot])
            else:
                self.env.step(self.d_actions[slot], featurizer=self.featurizer, out=self.d_out[slot])
            self.ev_run[slot].record(self.s_run)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_run[slot])
            self.h_block[slot].copy_(self.d_block[slot], non_blocking=True)
            self.ev_d2h[slot].record(self.s_d2h)
        self.k += 1
        return slot

    def wait(self, slot):
        """Block the host until the slot's results are in pinned host memory.  Compact: a `CompactResult` (unpacks to
        `rewards (N, A) f64, dones, truncated` on first use); dense: `(rewards (N, A) f32, dones, truncated)` tensors."""
        self.ev_d2h[slot].synchronize()
        out = self.h_out[slot]
        if self.compact is not None:
            out._decoded = None  # the pinned block was overwritten since the last decode
        return out

    def features(self, slot):
        """Make the current stream wait for the kernel that filled `slot` and return the featurizer's views of it."""
        torch.cuda.current_stream(self.env.device).wait_event(self.ev_run[slot])
        self.featurizer.bind_buffers(*self.feat_bufs[slot])
        self.featurizer.B, self.featurizer.T = self.env.num_envs, 1
        return self.featurizer.generate_featurized_states()

    def release_features(self, slot):
        """Record, on the current stream, that the consumer has finished reading `slot`'s feature tensors."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.env.device))
        self.ev_consumed[slot] = ev

    def drain(self):
        """Make the current stream wait for everything enqueued so far."""
        cur = torch.cuda.current_stream(self.env.device)
        for s in (self.s_h2d, self.s_run, self.s_d2h):
            cur.wait_stream(s)
