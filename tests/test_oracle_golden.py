"""CPU: the C oracle (oracle/susnet_oracle.c) against the golden fixtures minted from the UNMODIFIED reference
(tools/make_golden.py).  This is what pins the oracle on boxes that have no copy of the reference."""
import numpy as np
import pytest

import oracle
from tests.util import CASES, case_of, golden_files, load, reward_bits


def replay_oracle(g, cfg):
    T, N, A = g["actions"].shape
    injected = bool(g["injected"])
    env = oracle.OracleEnv(cfg, N, seed=int(g["seed"]), env_id_base=int(g["env_id_base"]))
    if injected:
        env.inject_words(reset_words=g["reset_words0"])
    flat = env.reset()
    assert np.array_equal(flat, g["reset_flat"])
    assert np.array_equal(env.imposter_mask(), g["reset_imp"])
    for t in range(T):
        if injected:
            env.inject_words(act_words=g["act_words"][t])
        a = env.sample_actions()
        if t % 7 != 3:
            assert np.array_equal(a, g["actions"][t]), f"sample_actions differs at step {t}"
        if injected:
            env.inject_words(step_words=g["step_words"][t], reset_words=g["reset_words"][t])
        o = env.step(g["actions"][t].astype(np.int32))
        assert np.array_equal(o["next_flat"], g["next_flat"][t]), f"state differs at step {t}"
        assert np.array_equal(reward_bits(o["rewards"]), reward_bits(g["rewards"][t])), f"rewards differ at step {t}"
        assert np.array_equal(o["done"], g["done"][t]) and np.array_equal(o["trunc"], g["trunc"][t])
        assert np.array_equal(o["metrics"], g["metrics"][t]), f"metrics differ at step {t}"
        assert np.array_equal(env.flat_states(), g["cur_flat"][t]), f"post-reset state differs at step {t}"
        assert np.array_equal(env.imposter_mask(), g["imp"][t])
    fin = (g["done"] | g["trunc"]) != 0
    st = env.stats()
    assert st[0] == fin.sum() and st[9] == (g["trunc"] != 0).sum()
    assert st[8] == g["metrics"][..., 0][fin].sum()


@pytest.mark.parametrize("path", golden_files("philox"), ids=case_of)
def test_oracle_matches_reference_philox_draws(path):
    g = load(path)
    replay_oracle(g, CASES[case_of(path)])


@pytest.mark.parametrize("path", golden_files("words"), ids=case_of)
def test_oracle_matches_reference_injected_words(path):
    g = load(path)
    replay_oracle(g, CASES[case_of(path)])


@pytest.mark.parametrize("path", golden_files("features"), ids=case_of)
def test_oracle_features_match_reference(path):
    g = load(path)
    cfg = CASES[case_of(path)]
    flat = g["flat"].astype(np.int64)
    if "global_spatial" in g:
        sp, ns = oracle.encode_global(cfg, flat)
        assert np.array_equal(sp, g["global_spatial"].astype(np.float32))
        assert np.array_equal(ns, g["global_non_spatial"])
        sp, ns = oracle.encode_perspective(cfg, flat)
        assert np.array_equal(sp, g["perspective_spatial"].astype(np.float32))
        assert np.array_equal(ns, g["perspective_non_spatial"])
    i = 0
    while f"flat{i}" in g:
        out = oracle.encode_flat(cfg, [str(c) for c in g[f"flat{i}_components"]], flat)
        assert np.array_equal(out.view(np.int32), g[f"flat{i}"].view(np.int32))  # bit-exact, incl. scent
        i += 1


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors)."""
    from oracle.rng_spec import philox4x32_10

    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert philox4x32_10(ctr, key).tolist() == want


def test_survey_anchor_cases():
    """Hand-checkable anchors recorded in SURVEY.md 8(c) (probed on the reference)."""
    # FourRoomEnv(1, 2, n_jobs=0): first step -> rewards [-10, 10, 10], done
    cfg = oracle.default_config("base", n_crew=2, n_jobs=0, shuffle_imposter_index=False)
    env = oracle.OracleEnv(cfg, 1, auto_reset=False)
    env.reset()
    o = env.step(np.zeros((1, 3), dtype=np.int32))
    assert o["rewards"].tolist() == [[-10.0, 10.0, 10.0]] and o["done"][0] == 1
    # ITG(n_crew=2, kill_reward=-3): all on one cell, imposter KILL -> rewards [3, 0, 0], exactly one crew dead
    cfg = oracle.default_config("training_ground", n_crew=2, n_jobs=0, kill_reward=-3.0, sabotage_reward=0.0,
                                game_end_reward=0.0)
    env = oracle.OracleEnv(cfg, 1, auto_reset=False)
    env.import_flat(np.array([[3, 3, 3, 3, 3, 3, 1, 1, 1]]), np.array([[1, 0, 0]]))
    o = env.step(np.array([[5, 0, 0]], dtype=np.int32))
    assert o["rewards"].tolist() == [[3.0, 0.0, 0.0]] and o["next_flat"][0, 6:].sum() == 2 and o["done"][0] == 0
    # FourRoomEnv(1,2,1) with the imposter at index 2 killing one of two co-located crew -> [10, -2, -15] or
    # [-2, 10, -15] up to the victim pick; done (index-negation quirk C-1)
    cfg = oracle.default_config("base", n_crew=2, n_jobs=1, is_action_order_random=False)
    env = oracle.OracleEnv(cfg, 1, auto_reset=False)
    env.import_flat(np.array([[3, 3, 3, 3, 3, 3, 1, 1, 1, 0, 0, 0]]), np.array([[0, 0, 1]]))
    env.inject_words(step_words=np.full((1, 5), 0xFFFFFFFF, dtype=np.uint32))  # victim pick 1 of 2 -> agent 1
    o = env.step(np.array([[0, 0, 6]], dtype=np.int32))
    assert o["rewards"].tolist() == [[10.0, -2.0, -15.0]] and o["done"][0] == 1
    env.import_flat(np.array([[3, 3, 3, 3, 3, 3, 1, 1, 1, 0, 0, 0]]), np.array([[0, 0, 1]]))
    env.inject_words(step_words=np.zeros((1, 5), dtype=np.uint32))  # victim pick 0 -> agent 0
    o = env.step(np.array([[0, 0, 6]], dtype=np.int32))
    assert o["rewards"].tolist() == [[-2.0, -10.0, -15.0]] and o["done"][0] == 1
    # ITG STAY-only episode truncates on step call #1000
    cfg = oracle.default_config("training_ground", n_crew=1, n_jobs=0)
    env = oracle.OracleEnv(cfg, 1, auto_reset=False)
    env.reset()
    for k in range(1000):
        o = env.step(np.zeros((1, 2), dtype=np.int32))
        assert bool(o["trunc"][0]) == (k == 999)
    # tagging: a dead agent's tag still counts (quirk C-6)
    cfg = oracle.default_config("tagging", n_crew=3, n_jobs=1, is_action_order_random=False)
    env = oracle.OracleEnv(cfg, 1, auto_reset=False)
    flat = np.array([[0, 0, 1, 1, 2, 2, 3, 3, 1, 1, 1, 0, 5, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 50]])
    env.import_flat(flat, np.array([[1, 0, 0, 0]]))
    o = env.step(np.array([[0, 0, 0, 6]], dtype=np.int32))  # agent 3 (dead crew) tags agent 0
    assert o["next_flat"][0, 19:23].tolist() == [1, 0, 0, 0]
