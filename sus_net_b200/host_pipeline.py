"""Host-buffer driver for a batched env: actions arrive in (pinned) host memory, rewards / dones / truncations go
back to pinned host memory, and the three legs of a step -- H2D copy, fused step(+encode) kernel, D2H copy -- run on
three CUDA streams over two rotating slots, so consecutive steps overlap (PCIe is full duplex).  The feature tensors
stay on the device for the consumer (the Q-network).  No reference analogue: the reference is a host-only loop."""
import torch


class HostStepper:
    def __init__(self, env, featurizer=None, slots=2, actions_dtype=torch.uint8):
        """actions_dtype: dtype of the host action rows (uint8 = 1 byte per agent on the PCIe link, role-list indices
        are < 256; int32 / int64 are accepted too, the reference's loop uses np.int32)."""
        assert env.batched, "HostStepper drives batched envs"
        self.env, self.featurizer, self.slots = env, featurizer, slots
        dev, N, A = env.device, env.num_envs, env.n_agents
        self.s_h2d, self.s_run, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
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
        self.ev_h2d = [torch.cuda.Event() for _ in range(slots)]
        self.ev_run = [torch.cuda.Event() for _ in range(slots)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(slots)]
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
        """Enqueue one step on `host_actions` ((N, A) of `actions_dtype`, ideally pinned).  Returns the slot whose pinned host
        buffers `(rewards, dones, truncated)` will hold the results once `wait(slot)` returns."""
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
            self.env.step(self.d_actions[slot], featurizer=self.featurizer, out=self.d_out[slot])
            self.ev_run[slot].record(self.s_run)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_run[slot])
            self.h_block[slot].copy_(self.d_block[slot], non_blocking=True)
            self.ev_d2h[slot].record(self.s_d2h)
        self.k += 1
        return slot

    def wait(self, slot):
        self.ev_d2h[slot].synchronize()
        return self.h_out[slot]

    def drain(self):
        """Make the current stream wait for everything enqueued so far."""
        cur = torch.cuda.current_stream(self.env.device)
        for s in (self.s_h2d, self.s_run, self.s_d2h):
            cur.wait_stream(s)
