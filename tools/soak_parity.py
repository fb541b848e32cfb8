#!/usr/bin/env python
"""Full-size parity soak (BASELINE.json headline size): 1 Mi FourRoomEnv(1v4, 5 jobs) envs, fused step + Global encode,
random policy, compared with the oracle on the same Philox draws:
  * every `--check-every` steps: ALL flat states identical, episode statistics identical;
  * at the end: rewards / dones of the last step identical, the feature tensors of a 65 536-env slice identical, and for
    ALL envs the size-independent property "ones in the planes == alive agents + jobs" and a checksum of the non-spatial
    views against the oracle's.
`--config flat` runs the cfg4-alt shape instead (ImposterTrainingGround 1v4, walled, Flat-98 features through the
byte-staged k_step_flat) and compares the feature rows of ALL envs at every checkpoint.
    python tools/soak_parity.py [--envs 1048576] [--steps 1000] [--check-every 100] [--config global|flat]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
import sus_net_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--check-every", type=int, default=100)
    ap.add_argument("--config", choices=["global", "flat", "tagging", "compact"], default="global")
    a = ap.parse_args()
    N, T = a.envs, a.steps
    if a.config == "flat":
        return soak_flat(N, T, a.check_every)
    if a.config == "tagging":
        return soak_tagging(N, T, a.check_every)
    if a.config == "compact":
        return soak_compact(N, T, a.check_every)
    cfg = oracle.default_config("base", n_crew=4, n_jobs=5)
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=99, device="cuda:0")
    feat = S.GlobalFeaturizer(env)
    orc = oracle.OracleEnv(cfg, N, seed=99)
    oracle.set_threads(os.cpu_count() or 1)
    assert np.array_equal(env.reset()[0].cpu().numpy().astype(np.int64), orc.reset())
    t0 = time.time()
    out = None
    checks = 0
    for t in range(1, T + 1):
        nf, r, d, tr, _ = env.step(None, featurizer=feat)
        out = orc.step(None, want_flat=False, want_metrics=False, out=out)
        if t % a.check_every == 0 or t == T:
            assert np.array_equal(env.flat_states(torch.int64).cpu().numpy(), orc.flat_states()), f"states differ at step {t}"
            assert np.array_equal(env.episode_stats().cpu().numpy(), orc.stats()), f"episode stats differ at step {t}"
            assert np.array_equal(r.cpu().numpy(), out["rewards"].astype(np.float32)), f"rewards differ at step {t}"
            assert np.array_equal(d.cpu().numpy(), out["done"] != 0), f"dones differ at step {t}"
            checks += 1
    cur = orc.flat_states()
    views = feat.generate_featurized_states()
    sp = views[0][0].detach()[:, 0]
    ones = sp.sum(dim=(1, 2, 3)).cpu().numpy()
    alive = cur[:, 10:15].sum(axis=1)
    assert np.array_equal(ones, (alive + 5).astype(np.float32)), "plane population != alive agents + jobs"
    k = min(N, 65536)
    want_sp, want_ns = oracle.encode_global(cfg, cur[:k])
    assert np.array_equal(sp[:k].cpu().numpy(), want_sp)
    for v in range(5):
        assert np.array_equal(views[v][1].detach()[:k, 0].cpu().numpy(), want_ns[v])
    _, ns_all = oracle.encode_global(cfg, cur)
    chk = float(sum(views[v][1].detach().double().sum().item() for v in range(5)))
    assert chk == float(ns_all.astype(np.float64).sum()), "non-spatial checksum differs"
    stats = dict(zip(S.STAT_KEYS, [int(x) for x in orc.stats()]))
    print(json.dumps({"envs": N, "steps": T, "env_steps": N * T, "full_state_comparisons": checks,
                      "finished_trajectories": stats["episodes"], "stats": stats, "wall_s": round(time.time() - t0, 1),
                      "result": "identical"}))


def soak_flat(N, T, check_every):
    from tests.cases import CASES
    from tests.util import flat_featurizer, make_cuda_env

    cfg = CASES["cfg4alt_itg_1v4"]
    comps = ["onehot_pos", "alive_crew", "closest_crew"]
    env = make_cuda_env(cfg, N, seed=99)
    feat = flat_featurizer(env, comps)
    orc = oracle.OracleEnv(cfg, N, seed=99)
    oracle.set_threads(os.cpu_count() or 1)
    assert np.array_equal(env.reset()[0].cpu().numpy().astype(np.int64), orc.reset())
    t0, out, checks = time.time(), None, 0
    for t in range(1, T + 1):
        nf, r, d, tr, _ = env.step(None, featurizer=feat)
        out = orc.step(None, want_flat=False, want_metrics=False, out=out)
        if t % check_every == 0 or t == T:
            cur = orc.flat_states()
            assert np.array_equal(env.flat_states(torch.int64).cpu().numpy(), cur), f"states differ at step {t}"
            assert np.array_equal(env.episode_stats().cpu().numpy(), orc.stats()), f"episode stats differ at step {t}"
            assert np.array_equal(r.cpu().numpy(), out["rewards"].astype(np.float32)), f"rewards differ at step {t}"
            assert np.array_equal(d.cpu().numpy(), out["done"] != 0), f"dones differ at step {t}"
            got = feat.generate_featurized_states()[0][1].detach()[:, 0].cpu().numpy()
            assert np.array_equal(got.view(np.int32), oracle.encode_flat(cfg, comps, cur).view(np.int32)), f"features differ at step {t}"
            checks += 1
    stats = dict(zip(S.STAT_KEYS, [int(x) for x in orc.stats()]))
    print(json.dumps({"config": "cfg4-alt ITG 1v4 walled + Flat-98 (k_step_flat)", "envs": N, "steps": T, "env_steps": N * T,
                      "full_state_and_feature_comparisons": checks, "finished_trajectories": stats["episodes"], "stats": stats,
                      "wall_s": round(time.time() - t0, 1), "result": "identical"}))


def soak_tagging(N, T, check_every):
    """cfg3 (FourRoomEnvWithTagging 1v2, 5 jobs; BASELINE configs[2]), step only, explicit sample_actions + step."""
    from tests.cases import CASES
    from tests.util import make_cuda_env

    cfg = CASES["cfg3_tagging_1v2"]
    env = make_cuda_env(cfg, N, seed=99)
    orc = oracle.OracleEnv(cfg, N, seed=99)
    oracle.set_threads(os.cpu_count() or 1)
    assert np.array_equal(env.reset()[0].cpu().numpy().astype(np.int64), orc.reset())
    t0, out, checks = time.time(), None, 0
    for t in range(1, T + 1):
        acts = env.sample_actions()
        nf, r, d, tr, _ = env.step(acts)
        oa = orc.sample_actions()
        out = orc.step(oa, want_flat=False, want_metrics=False, out=out)
        if t % check_every == 0 or t == T:
            assert np.array_equal(acts.cpu().numpy(), oa), f"sampled actions differ at step {t}"
            assert np.array_equal(env.flat_states(torch.int64).cpu().numpy(), orc.flat_states()), f"states differ at step {t}"
            assert np.array_equal(env.episode_stats().cpu().numpy(), orc.stats()), f"episode stats differ at step {t}"
            assert np.array_equal(r.cpu().numpy(), out["rewards"].astype(np.float32)), f"rewards differ at step {t}"
            assert np.array_equal(d.cpu().numpy(), out["done"] != 0) and np.array_equal(tr.cpu().numpy(), out["trunc"] != 0)
            checks += 1
    env.check_actions()
    stats = dict(zip(S.STAT_KEYS, [int(x) for x in orc.stats()]))
    print(json.dumps({"config": "cfg3 FourRoomEnvWithTagging 1v2, 5 jobs, step only (sample_actions + step)", "envs": N, "steps": T,
                      "env_steps": N * T, "full_state_comparisons": checks, "finished_trajectories": stats["episodes"],
                      "stats": stats, "wall_s": round(time.time() - t0, 1), "result": "identical"}))


def soak_compact(N, T, check_every):
    """The headline config through the COMPACT HOST PROTOCOL at full size: bit-packed actions in, reward codes + done / truncated
    bits out of the fused step + Global encode; at every checkpoint the records of all envs are decoded on the host (threaded C
    decoder AND its numpy statement) and the float64 rewards must equal the oracle's BIT PATTERNS, with flags, states and stats."""
    cfg = oracle.default_config("base", n_crew=4, n_jobs=5)
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=99, device="cuda:0")
    env.emit_next_states = False
    feat = S.GlobalFeaturizer(env)
    cp = env.compact
    orc = oracle.OracleEnv(cfg, N, seed=99)
    oracle.set_threads(os.cpu_count() or 1)
    assert np.array_equal(env.reset()[0].cpu().numpy().astype(np.int64), orc.reset())
    rec = torch.empty((N, cp.result_bytes), dtype=torch.uint8, device=env.device)
    t0, out, checks, oa = time.time(), None, 0, None
    for t in range(1, T + 1):
        acts = env.sample_actions()
        env.step(cp.pack_actions(acts).contiguous(), featurizer=feat, packed_actions=True, packed_out=rec)
        oa = orc.sample_actions(out=oa)
        out = orc.step(oa, want_flat=False, want_metrics=False, out=out)
        if t % check_every == 0 or t == T:
            h = rec.cpu().numpy()
            r, d, tr = cp.decode(h)
            r2, d2, tr2 = cp.decode_numpy(h)
            assert np.array_equal(r.view(np.int64), out["rewards"].view(np.int64)), f"decoded rewards differ at step {t}"
            assert np.array_equal(r2.view(np.int64), r.view(np.int64)) and np.array_equal(d, d2) and np.array_equal(tr, tr2)
            assert np.array_equal(d, out["done"] != 0) and np.array_equal(tr, out["trunc"] != 0), f"flags differ at step {t}"
            assert np.array_equal(acts.cpu().numpy(), oa), f"sampled actions differ at step {t}"
            assert np.array_equal(env.flat_states(torch.int64).cpu().numpy(), orc.flat_states()), f"states differ at step {t}"
            assert np.array_equal(env.episode_stats().cpu().numpy(), orc.stats()), f"episode stats differ at step {t}"
            checks += 1
    env.check_actions()
    cur = orc.flat_states()
    views = feat.generate_featurized_states()
    k = min(N, 65536)
    want_sp, want_ns = oracle.encode_global(cfg, cur[:k])
    assert np.array_equal(views[0][0].detach()[:k, 0].cpu().numpy(), want_sp) and np.array_equal(views[4][1].detach()[:k, 0].cpu().numpy(), want_ns[4])
    stats = dict(zip(S.STAT_KEYS, [int(x) for x in orc.stats()]))
    print(json.dumps({"config": "cfg4 FourRoomEnv 1v4 + Global encode through the compact host protocol (packed actions in, reward codes out)",
                      "envs": N, "steps": T, "env_steps": N * T, "full_comparisons": checks, "record_bytes": [cp.action_bytes, cp.result_bytes],
                      "finished_trajectories": stats["episodes"], "stats": stats, "wall_s": round(time.time() - t0, 1), "result": "identical"}))


if __name__ == "__main__":
    main()
