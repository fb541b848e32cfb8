"""TEST INFRASTRUCTURE (oracle/) -- never imported by the product path.

Python/numpy restatement of the susnet-b200 random-draw SPEC.  The reference
(Sus-Net) uses numpy's global Mersenne Twister (SURVEY.md A.6 lists the draw
sites R1-R6); the B200 build replaces it with counter-based Philox4x32-10
keyed by (seed, env, tick).  This module turns raw Philox words into the
*semantic* draws (imposter ids, spawn cells, action order, victim pick, random
actions) exactly the way the CUDA kernels and the C oracle do, so the unmodified
Python reference can be replayed on the same draws (`oracle/ref_harness.py`).

Spec
----
word(env, tick, purpose, slot) = philox4x32_10(
        counter = (env_id, tick & 0xffffffff, tick >> 32, purpose | (slot >> 2) << 8),
        key     = (seed & 0xffffffff, seed >> 32))[slot & 3]
bounded(u, k) = (u * k) >> 32

purpose 0  STEP       slots [0, A-1)      Fisher-Yates words for the action order (R4)
                      slots A-1+e         victim word of the e-th KILL event of the step
                                          whose candidate list is non-empty (R5)
purpose 1  AUTORESET  reset draws of an env that finished at this step tick
purpose 3  RESET      reset draws of an explicit reset (tick = reset epoch)
                      slots [0, n_imp)            imposter ids (R1, ascending-sorted subset)
                      slots n_imp + i             spawn cell of agent i (R2, with replacement)
                      slots n_imp + A + j         cell of job j (R3, without replacement)
purpose 2  ACT        slots i             random action of agent i (R6), tick = act epoch
purpose 4  ACT_FUSED  same, drawn inside a step launch (tick = step tick)
purpose 5  POLICY     epsilon-greedy acting (train.py:349-381), tick = act epoch
                      slots 2i            explore word of agent i: explore iff word * 2**-32 <= eps
                      slots 2i + 1        its random action (bounded by the role list's length)
"""
import numpy as np

P_STEP, P_AUTORESET, P_ACT, P_RESET, P_ACT_FUSED, P_POLICY = 0, 1, 2, 3, 4, 5

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k = np.asarray(key, dtype=np.uint64).copy()
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = k[..., 0], k[..., 1]
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK
        n1 = p1 & _MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def words(seed, env_ids, tick, purpose, n_slots):
    """Raw words for slots [0, n_slots) of every env in env_ids -> (len(env_ids), n_slots) uint32."""
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    n_blocks = max(1, (n_slots + 3) // 4)
    out = np.zeros((env_ids.size, n_blocks * 4), dtype=np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    for b in range(n_blocks):
        ctr = np.zeros((env_ids.size, 4), dtype=np.uint64)
        ctr[:, 0] = env_ids & _MASK
        ctr[:, 1] = tick & 0xFFFFFFFF
        ctr[:, 2] = (tick >> 32) & 0xFFFFFFFF
        ctr[:, 3] = purpose | (b << 8)
        out[:, 4 * b : 4 * b + 4] = philox4x32_10(ctr, np.broadcast_to(key, (env_ids.size, 2)))
    return out[:, :n_slots]


def bounded(u, k):
    return (int(u) * int(k)) >> 32


def n_step_slots(n_agents):
    return 2 * n_agents - 1


def n_reset_slots(n_imposters, n_agents, n_jobs):
    return n_imposters + n_agents + n_jobs


def action_order(w, n_agents):
    """Fisher-Yates from the back: word s pairs with k = A-1-s."""
    order = list(range(n_agents))
    for k in range(n_agents - 1, 0, -1):
        j = bounded(w[n_agents - 1 - k], k + 1)
        order[k], order[j] = order[j], order[k]
    return order


def pick_distinct(ws, n_total):
    """len(ws) distinct ids in [0, n_total): word m picks the r-th smallest unchosen id, r = bounded(w, n_total - m)."""
    chosen_sorted, picked = [], []
    for m, w in enumerate(ws):
        r = bounded(w, n_total - m)
        for c in chosen_sorted:
            if r >= c:
                r += 1
        picked.append(r)
        chosen_sorted.append(r)
        chosen_sorted.sort()
    return picked


def reset_draws(w, n_imposters, n_agents, n_jobs, n_valid, shuffle_imposter_index):
    """-> (imposter_idxs ascending, agent valid-cell indices, job valid-cell indices)."""
    if shuffle_imposter_index:
        imps = sorted(pick_distinct(w[:n_imposters], n_agents))
    else:
        imps = list(range(n_imposters))
    agent_cells = [bounded(w[n_imposters + i], n_valid) for i in range(n_agents)]
    base = n_imposters + n_agents
    job_cells = pick_distinct(w[base : base + n_jobs], n_valid)
    return imps, agent_cells, job_cells
