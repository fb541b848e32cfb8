#!/usr/bin/env python
"""Where the time of the cfg5 training loop goes (one GPU): device time per phase of an EAGER iteration (CUDA events around
each phase, averaged), then the same loop as CUDA graphs for the per-iteration total, plus (optionally) the torch profiler's
top kernels.

    python tools/profile_train_loop.py [--envs 131072] [--iters 200] [--tf32] [--kernels]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from tools.train_demo import build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--kernels", action="store_true")
    ap.add_argument("--torch-mlp", action="store_true", help="evaluate the Q-network with the torch module instead of sus_mlp_forward")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    dev = torch.device("cuda", 0)
    N = a.envs
    env, loop = build(N, dev, 0, a.batch, graphs=False, fused_mlp=not a.torch_mlp)
    loop.run(10)
    actor, buf, feat, seq = loop.actor, loop.buf, loop.feat, loop.seq
    names = ("q_network_forward", "select_actions_kernel", "fused_step_encode_kernel", "export_flat+replay_push", "sample_batch", "train_step")
    acc = {k: 0.0 for k in names}
    launches0 = int(S.lib().sus_launch_count())

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    spans = []
    for it in range(a.iters):
        loop.eps.fill_(0.3)
        sp, ns = seq.views()
        e0 = ev()
        with torch.no_grad():  # the network the actor evaluates: the one-launch MLP kernel (or the torch module with --torch-mlp)
            q = loop.actor._infer["imp"](torch.zeros(N, 1, 1, device=dev), ns[0]).float().contiguous()
        e1 = ev()
        # (act_kernel would run the network again; call the selection kernel directly on the Q-values just computed)
        import ctypes as C
        from sus_net_b200 import _lib as L
        io = L.SusPolicyIO()
        io.q_imposter = q.data_ptr(); io.eps = loop.eps.data_ptr(); io.actions_dtype = L.I32
        if actor._actions is None:
            actor._actions = torch.zeros((N, env.n_agents), dtype=torch.int32, device=dev)
        io.actions = actor._actions.data_ptr()
        L.check(env.lib.sus_env_select_actions(env._h, C.byref(io), env._stream()))
        e2 = ev()
        out = env.step(actor._actions, featurizer=feat)
        e3 = ev()
        # the rest of collect_step(): export of the post-reset state + the push kernel
        next_flat, rewards, dones, truncated, _ = out
        env.flat_states(out=buf._cur_flat)
        p = L.SusReplayPush(N=N, M=buf.max_size, idx=buf.idx, T=1, S=buf.state_size, A=buf.n_agents, n_imposters=1,
                            seq_in=buf._seq[0].data_ptr(), seq_out=buf._seq[1].data_ptr(), next_flat=next_flat.data_ptr(),
                            cur_flat=buf._cur_flat.data_ptr(), actions=actor._actions.data_ptr(), actions_dtype=L.I32,
                            rewards=rewards.data_ptr(), done=dones.data_ptr(), truncated=truncated.data_ptr(),
                            imposters=env._imposters_buf.data_ptr(), states=buf.states.data_ptr(), r_actions=buf.actions.data_ptr(),
                            r_rewards=buf.rewards.data_ptr(), next_states=buf.next_states.data_ptr(), r_dones=buf.dones.data_ptr(),
                            r_imposters=buf.imposters.data_ptr())
        L.check(env.lib.sus_replay_push(C.byref(p), dev.index, env._stream()))
        buf._seq.reverse()
        buf.idx = (buf.idx + N) % buf.max_size
        buf.size = min(buf.size + N, buf.max_size)
        e4 = ev()
        span = [(names[0], e0, e1), (names[1], e1, e2), (names[2], e2, e3), (names[3], e3, e4)]
        if it % 5 == 0:
            batch = buf.sample(a.batch)
            e5 = ev()
            loop.trainer.train_step(batch, loop.feat_train, loop.imposter_model, loop.imposter_target, None, None)
            e6 = ev()
            span += [(names[4], e4, e5), (names[5], e5, e6)]
        spans.append(span)
    torch.cuda.synchronize(dev)
    for span in spans:
        for k, s, e in span:
            acc[k] += s.elapsed_time(e)
    per_iter = {k: v / a.iters for k, v in acc.items()}
    launches = (int(S.lib().sus_launch_count()) - launches0) / a.iters
    # the same loop through BatchedTrainingLoop, eager and as CUDA graphs: wall time per iteration
    totals = {}
    for graphs in (False, True):
        env2, loop2 = build(N, dev, 0, a.batch, graphs=graphs, fused_mlp=not a.torch_mlp)
        loop2.run(40)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        loop2.run(a.iters)
        torch.cuda.synchronize(dev)
        totals["cuda_graphs" if graphs else "eager"] = 1e3 * (time.perf_counter() - t0) / a.iters
        del env2, loop2
    print(json.dumps({"envs": N, "iters": a.iters, "batch": a.batch, "tf32_q_network": a.tf32,
                      "q_network": "torch module" if a.torch_mlp else "sus_mlp_forward (one launch, fp32 FFMA)",
                      "device_ms_per_iteration_by_phase (eager, CUDA events; train phases amortised over 5 iterations)": per_iter,
                      "sum_of_phases_ms": sum(per_iter.values()), "library_kernel_launches_per_iteration": launches,
                      "wall_ms_per_iteration": totals,
                      "env_steps_per_s": {k: N / (v * 1e-3) for k, v in totals.items()}}))
    if a.kernels:
        from torch.profiler import ProfilerActivity, profile

        env2, loop2 = build(N, dev, 0, a.batch, graphs=False)
        loop2.run(10)
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            loop2.run(20)
            torch.cuda.synchronize(dev)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
