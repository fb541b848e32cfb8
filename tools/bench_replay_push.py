#!/usr/bin/env python
"""Row f1's replay push at full size (cfg4, 1 Mi transitions per launch, T = 1 and T = 4): the vectorised stream kernel
(k_replay_push_v, default) against the one-thread-per-float kernel (SUSNET_REPLAY_PUSH=v1), CUDA events, median of --reps, with
the algorithmic bytes per transition (24 T S + 20 A + 7) and the fraction of the measured copy peak.

    python tools/bench_replay_push.py [--envs 1048576] [--reps 30]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sus_net_b200 as S  # noqa: E402
from sus_net_b200 import _lib as L  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_gbps"):
            if k in d:
                return float(d[k])
        for v in d.values():
            if isinstance(v, dict):
                for k in ("hbm_gbs", "copy_gbs", "gbs"):
                    if k in v:
                        return float(v[k])
    except Exception:
        pass
    return 6540.2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    N = a.envs
    peak = peak_gbs()
    out = {"envs": N, "copy_peak_gbs": peak}
    for T in (1, 4):
        env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=1234, device=dev)
        env.reset()
        buf = S.ReplayBuffer(2 * N, env.flattened_state_size, T, env.n_agents, 1, device=dev)
        buf.attach(env)
        for _ in range(2):
            buf.collect_step(env.sample_actions())
        acts = env.sample_actions()
        next_flat, rewards, dones, truncated, _ = env.step(acts)
        env.flat_states(out=buf._cur_flat)
        nbytes = 24 * T * buf.state_size + 20 * buf.n_agents + 7
        res = {}
        keep = {}
        for ver in ("v1", "v2"):
            os.environ["SUSNET_REPLAY_PUSH"] = ver
            evs = []
            for r in range(a.reps + 3):
                idx = (r % 2) * N  # both halves of the ring, 16-byte aligned
                p = L.SusReplayPush(N=N, M=buf.max_size, idx=idx, T=T, S=buf.state_size, A=buf.n_agents, n_imposters=1,
                                    seq_in=buf._seq[0].data_ptr(), seq_out=buf._seq[1].data_ptr(), next_flat=next_flat.data_ptr(),
                                    cur_flat=buf._cur_flat.data_ptr(), actions=acts.data_ptr(), actions_dtype=L.I32,
                                    rewards=rewards.data_ptr(), done=dones.data_ptr(), truncated=truncated.data_ptr(),
                                    imposters=env._imposters_buf.data_ptr(), states=buf.states.data_ptr(),
                                    r_actions=buf.actions.data_ptr(), r_rewards=buf.rewards.data_ptr(),
                                    next_states=buf.next_states.data_ptr(), r_dones=buf.dones.data_ptr(),
                                    r_imposters=buf.imposters.data_ptr())
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); L.check(env.lib.sus_replay_push(C.byref(p), dev.index, env._stream())); e.record()
                if r >= 3:
                    evs.append((s, e))
            torch.cuda.synchronize(dev)
            ms = sorted(s.elapsed_time(e) for s, e in evs)[len(evs) // 2]
            res[ver] = {"ms": ms, "gbs": nbytes * N / ms / 1e6, "frac_of_copy_peak": nbytes * N / ms / 1e6 / peak}
            keep[ver] = [t.clone() for t in (buf.states, buf.next_states, buf.actions, buf.rewards, buf.dones, buf.imposters, buf._seq[1])]
        os.environ.pop("SUSNET_REPLAY_PUSH")
        res["identical_results"] = all(torch.equal(x, y) for x, y in zip(keep["v1"], keep["v2"]))
        res["bytes_per_transition"] = nbytes
        out[f"T{T}"] = res
        del env, buf, keep
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
