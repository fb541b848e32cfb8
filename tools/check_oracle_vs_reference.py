#!/usr/bin/env python
"""Pin the C oracle against the UNMODIFIED reference: replay many trajectories on identical draws and
require identical flat states, rewards, dones, truncations and per-episode metrics, then identical features.

    python tools/check_oracle_vs_reference.py --envs 64 --steps 400 [--case NAME] [--procs 8]

Needs /root/reference (build container only)."""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests.cases import CASES  # noqa: E402


def reference_features(env, kind, flat, components=None):
    """Run the reference's model-ready featurizer on (n, S) flat states (as a (n, 1, S) float batch)."""
    import torch
    from oracle import ref_harness as H

    _, feat = H.import_reference()
    if kind == "global":
        f = feat.GlobalFeaturizer(env)
    elif kind == "perspective":
        f = feat.PerspectiveFeaturizer(env)
    else:
        classes = dict(onehot_pos=feat.OneHotAgentPositionFeaturizer, coords=feat.CoordinateAgentPositionsFeaturizer,
                       alive_crew=feat.AliveCrewFeaturizer, closest_crew=feat.ClosestAliveCrewFeaturizer,
                       l1_crew=feat.L1CrewFeaturizer, dist_to_imposter=feat.DistanceToImposterFeaturizer,
                       walls=feat.WallsFeaturizer, rooms=feat.ImposterVSCrewRoomLocaionFeaturizer,
                       scent=feat.ImposterScentFeaturizer)
        from src.environment.base import StateFields
        fields = dict(state_alive=StateFields.ALIVE_AGENTS, state_job_status=StateFields.JOB_STATUS,
                      state_used_tags=StateFields.USED_TAGS, state_tag_counts=StateFields.TAG_COUNTS)
        parts = [classes[c](env) if c in classes else feat.StateFieldFeaturizer(env, fields[c]) for c in components]
        comp = feat.CompositeFeaturizer.__new__(feat.CompositeFeaturizer)  # skip the shape assert (scent has a 0-d shape)
        comp.featurizers = parts
        f = feat.FlatFeaturizer(env, comp)
    f.fit(torch.tensor(flat, dtype=torch.float32).unsqueeze(1))
    views = f.generate_featurized_states()
    sp = np.stack([v[0].detach().numpy()[:, 0] for v in views])
    ns = np.stack([v[1].detach().numpy()[:, 0] for v in views])
    return sp, ns


def check_features(name, cfg, env, flat):
    import oracle
    from tests.cases import GLOBAL_CASES, FLAT_COMPONENT_SETS

    if name in GLOBAL_CASES:
        sp, ns = reference_features(env, "global", flat)
        osp, ons = oracle.encode_global(cfg, flat)
        for k in range(sp.shape[0]):
            assert np.array_equal(sp[k], osp), f"{name}: global spatial mismatch (view {k})"
        assert np.array_equal(ns, ons), f"{name}: global non-spatial mismatch"
        sp, ns = reference_features(env, "perspective", flat)
        osp, ons = oracle.encode_perspective(cfg, flat)
        assert np.array_equal(sp, osp), f"{name}: perspective spatial mismatch"
        assert np.array_equal(ns, ons), f"{name}: perspective non-spatial mismatch"
    for comps in FLAT_COMPONENT_SETS.get(name, []):
        sp, ns = reference_features(env, "flat", flat, comps)
        o = oracle.encode_flat(cfg, comps, flat)
        for k in range(ns.shape[0]):
            assert np.array_equal(ns[k].view(np.int32), o.view(np.int32)), f"{name}: flat {comps} mismatch"
        assert sp.shape[2:] == (1,) and not sp.any()


def run_case(args):
    name, n_envs, n_steps, seed, base = args
    import oracle
    from oracle import ref_harness as H

    oracle.set_threads(1)
    if isinstance(name, tuple) and name[0] == "edge":
        from tests.cases import EDGE_CASES

        cfg, name = EDGE_CASES[name[1]], name[1]
    elif isinstance(name, tuple):  # ("random", k): k-th random constructor-argument set
        from tests.cases import random_case

        cfg = random_case(np.random.default_rng(100000 + name[1]))
        name = f"random{name[1]}:{cfg['variant']}:{cfg['n_imposters']}v{cfg['n_crew']}j{cfg['n_jobs']}"
    else:
        cfg = CASES[name]
    ref = H.ReferenceBatch(cfg, n_envs, seed, env_id_base=base)
    orc = oracle.OracleEnv(cfg, n_envs, seed, env_id_base=base)
    f_ref, f_orc = ref.reset(), orc.reset()
    assert np.array_equal(f_ref, f_orc), f"{name}: reset mismatch"
    gamma = 0.9  # running returns exactly as train() keeps them (train.py:324,386,421-424,436)
    orc.track_returns(gamma)
    G = np.zeros((n_envs, ref.A))
    ret_sums = np.zeros(2)
    assert np.array_equal(ref.imposter_idxs().astype(bool).shape, ref.imposter_idxs().shape)
    episodes = 0
    for t in range(n_steps):
        prev_imp = np.stack([np.asarray(e.imposter_mask, dtype=bool).copy() for e in ref.envs])
        a_ref = ref.sample_actions()
        a_orc = orc.sample_actions()
        assert np.array_equal(a_ref, a_orc), f"{name}: sample_actions mismatch at step {t}"
        o_ref = ref.step(a_ref)
        o_orc = orc.step(a_orc)
        for k in ("next_flat", "done", "trunc", "metrics"):
            if not np.array_equal(o_ref[k], o_orc[k]):
                bad = np.argwhere(o_ref[k] != o_orc[k])[0]
                raise AssertionError(f"{name}: {k} mismatch at step {t}, idx {bad}: ref={o_ref[k][bad[0]]} orc={o_orc[k][bad[0]]}")
        # rewards: bit-exact including the sign of zero
        if not np.array_equal(o_ref["rewards"].view(np.int64), o_orc["rewards"].view(np.int64)):
            raise AssertionError(f"{name}: reward bits mismatch at step {t}")
        for i, e in enumerate(ref.envs):
            G[i] = o_ref["rewards"][i] + gamma * G[i]  # train.py:386
        masks = np.stack([np.zeros(ref.A, dtype=bool)] * n_envs)
        for i in range(n_envs):  # the roles of the episode that just advanced (pre-reset mask = non-zero rows of o_orc)
            masks[i] = prev_imp[i]
        for i in range(n_envs):
            if o_ref["done"][i] or o_ref["trunc"][i]:
                ret_sums[0] += G[i][masks[i]].mean().item()  # train.py:421-422
                ret_sums[1] += G[i][~masks[i]].mean().item()
                G[i] = 0  # train.py:436
        assert np.array_equal(G.view(np.int64), orc.returns().view(np.int64)), f"{name}: running returns differ at step {t}"
        assert np.array_equal(ret_sums.view(np.int64), orc.return_sums().view(np.int64)), f"{name}: return sums differ at step {t}"
        episodes += int(((o_ref["done"] | o_ref["trunc"]) != 0).sum())
        cur_ref, cur_orc = ref.flat_states(), orc.flat_states()
        assert np.array_equal(cur_ref, cur_orc), f"{name}: post-reset state mismatch at step {t}"
        imp_ref = np.zeros((n_envs, ref.A), dtype=np.uint8)
        for i, e in enumerate(ref.envs):
            imp_ref[i] = e.imposter_mask
        assert np.array_equal(imp_ref, orc.imposter_mask()), f"{name}: imposter mask mismatch at step {t}"
        if t % 10 == 0:
            check_features(name, cfg, ref.envs[0], np.concatenate([o_ref["next_flat"], cur_ref]))
    return name, n_envs * n_steps, episodes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--case", default=None)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--seed", type=int, default=20260)
    ap.add_argument("--shards", type=int, default=1, help="independent env-id shards per case")
    ap.add_argument("--edge", action="store_true", help="check tests.cases.EDGE_CASES instead")
    ap.add_argument("--random-configs", type=int, default=0, help="check this many random constructor-argument sets instead")
    a = ap.parse_args()
    names = [a.case] if a.case else list(CASES)
    if a.random_configs:
        names = [("random", k) for k in range(a.random_configs)]
    if a.edge:
        from tests.cases import EDGE_CASES

        names = [("edge", k) for k in EDGE_CASES]
    jobs = [(n, a.envs, a.steps, a.seed, s * a.envs) for n in names for s in range(a.shards)]
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        res = pool.map(run_case, jobs)
    tot = {}
    for name, steps, eps in res:
        s, e = tot.get(name, (0, 0))
        tot[name] = (s + steps, e + eps)
    for name, (steps, eps) in tot.items():
        print(f"OK {name:28s} env-steps={steps:8d} finished-episodes={eps}")
    print(f"all identical; {sum(s for s, _ in tot.values())} env-steps, {sum(e for _, e in tot.values())} trajectories, {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
