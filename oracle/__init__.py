"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of `oracle/susnet_oracle.c`, the scalar CPU restatement of Sus-Net's env and
featurizers (see that file's header for the reference file:line anchors and how parity is
pinned against the real reference).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package; nothing under
`sus_net_b200/` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsusnet_oracle.so")

VARIANT_IDS = {"base": 0, "tagging": 1, "training_ground": 2}
N_METRICS = 8
N_STATS = 10
METRIC_KEYS = ("total_time_steps", "imp_killed_crew", "completed_jobs", "sabotaged_jobs",
               "imp_voted_out", "crew_voted_out", "crew_won", "imposter_won")
STAT_KEYS = ("episodes", "crew_won", "imposter_won", "imp_killed_crew", "completed_jobs", "sabotaged_jobs",
             "imp_voted_out", "crew_voted_out", "total_time_steps", "truncated_episodes")
FLAT_COMPONENTS = {
    "onehot_pos": 0, "coords": 1, "alive_crew": 2, "closest_crew": 3, "l1_crew": 4, "dist_to_imposter": 5,
    "walls": 6, "rooms": 7, "scent": 8, "state_alive": 9, "state_job_status": 10, "state_used_tags": 11,
    "state_tag_counts": 12,
}


class _Cfg(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("n_imposters", C.c_int32), ("n_crew", C.c_int32), ("n_jobs", C.c_int32),
        ("include_walls", C.c_int32), ("is_action_order_random", C.c_int32), ("shuffle_imposter_index", C.c_int32),
        ("max_time_steps", C.c_int32), ("tag_reset_interval", C.c_int32),
        ("kill_reward", C.c_double), ("complete_job_reward", C.c_double), ("sabotage_reward", C.c_double),
        ("time_step_reward", C.c_double), ("game_end_reward", C.c_double), ("dead_penalty", C.c_double),
        ("vote_reward", C.c_double),
    ]


def default_config(variant="base", **kw):
    """Config dict in the reference's constructor vocabulary with the reference's defaults
    (base.py:103-120, tagging.py:10-12, pred_prey.py:26-66)."""
    cfg = dict(
        variant=variant, n_imposters=1, n_crew=4, n_jobs=5, include_walls=True,
        is_action_order_random=True, shuffle_imposter_index=True, max_time_steps=1000, tag_reset_interval=50,
        kill_reward=-5.0, complete_job_reward=3.0, sabotage_reward=3.0, time_step_reward=0.0,
        game_end_reward=10.0, dead_penalty=-2.0, vote_reward=3.0,
    )
    if variant == "training_ground":  # pred_prey.py:52-66
        cfg.update(n_imposters=1, dead_penalty=0.0, is_action_order_random=False, shuffle_imposter_index=False)
    cfg.update(kw)
    return cfg


def build(force=False):
    """Compile the C restatement with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "susnet_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libsusnet_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.argtypes = [C.POINTER(_Cfg), C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_destroy.restype = None
        for name in ("orc_reset", "orc_sample_actions", "orc_export_flat", "orc_export_metrics", "orc_imposter_mask",
                     "orc_stats"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_inject_words.argtypes = [C.c_void_p] * 4
        L.orc_step.argtypes = [C.c_void_p] * 8
        L.orc_import_flat.argtypes = [C.c_void_p] * 4
        L.orc_flat_size.argtypes = [C.POINTER(_Cfg)]
        L.orc_n_role_actions.argtypes = [C.POINTER(_Cfg), C.c_int]
        L.orc_global_nonspatial_size.argtypes = [C.POINTER(_Cfg)]
        L.orc_perspective_nonspatial_size.argtypes = [C.POINTER(_Cfg)]
        L.orc_encode_global.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_encode_perspective.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_flat_feature_size.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int]
        L.orc_encode_flat.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_track_returns.argtypes = [C.c_void_p, C.c_double]
        L.orc_return_sums.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_export_returns.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _cfg_struct(cfg):
    s = _Cfg()
    for name, _t in _Cfg._fields_:
        v = cfg[name]
        setattr(s, name, VARIANT_IDS[v] if name == "variant" else v)
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    return lib().orc_set_threads(int(n))


class OracleEnv:
    """Batched CPU env with the same launch/tick contract as the CUDA env (auto-reset, Philox spec)."""

    def __init__(self, cfg, num_envs, seed=0, env_id_base=0, auto_reset=True):
        self.cfg = dict(cfg)
        self._c = _cfg_struct(cfg)
        self.N = int(num_envs)
        self.A = cfg["n_imposters"] + cfg["n_crew"]
        self.J = cfg["n_jobs"]
        self.S = lib().orc_flat_size(C.byref(self._c))
        self._h = C.c_void_p()
        rc = lib().orc_create(C.byref(self._c), self.N, seed, env_id_base, int(auto_reset), C.byref(self._h))
        if rc != 0:
            raise ValueError(f"oracle rejected config {cfg}")
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def inject_words(self, step_words=None, reset_words=None, act_words=None):
        arrs = [None if w is None else np.ascontiguousarray(w, dtype=np.uint32) for w in (step_words, reset_words, act_words)]
        self._keep = arrs
        lib().orc_inject_words(self._h, *[_p(a) for a in arrs])

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(self._h, _p(m))
        return self.flat_states()

    def flat_states(self, out=None):
        if out is None:
            out = np.zeros((self.N, self.S), dtype=np.int64)
        lib().orc_export_flat(self._h, _p(out))
        return out

    def metrics(self):
        out = np.zeros((self.N, N_METRICS), dtype=np.int64)
        lib().orc_export_metrics(self._h, _p(out))
        return out

    def imposter_mask(self):
        out = np.zeros((self.N, self.A), dtype=np.uint8)
        lib().orc_imposter_mask(self._h, _p(out))
        return out

    def stats(self):
        out = np.zeros(N_STATS, dtype=np.int64)
        lib().orc_stats(self._h, _p(out))
        return out

    def track_returns(self, gamma):
        lib().orc_track_returns(self._h, float(gamma))

    def return_sums(self):
        out = np.zeros(2, dtype=np.float64)
        lib().orc_return_sums(self._h, _p(out))
        return out

    def returns(self):
        out = np.zeros((self.N, self.A), dtype=np.float64)
        lib().orc_export_returns(self._h, _p(out))
        return out

    def sample_actions(self, out=None):
        if out is None:
            out = np.zeros((self.N, self.A), dtype=np.int32)
        lib().orc_sample_actions(self._h, _p(out))
        return out

    def import_flat(self, flat, imposter_mask, t=None):
        f = np.ascontiguousarray(flat, dtype=np.int64)
        m = np.ascontiguousarray(imposter_mask, dtype=np.uint8)
        tt = None if t is None else np.ascontiguousarray(t, dtype=np.int32)
        lib().orc_import_flat(self._h, _p(f), _p(m), _p(tt))

    def step(self, actions=None, want_flat=True, want_metrics=True, out=None):
        """-> dict(rewards f64 (N,A), done, trunc, next_flat (pre-reset), metrics (pre-reset), actions).
        `out`: a dict returned by an earlier call, whose buffers are reused (timing loops)."""
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        if out is not None:
            a_out, rewards, done, trunc, nf, met = (out["actions"], out["rewards"], out["done"], out["trunc"],
                                                    out["next_flat"], out["metrics"])
        else:
            a_out = np.zeros((self.N, self.A), dtype=np.int32)
            rewards = np.zeros((self.N, self.A), dtype=np.float64)
            done = np.zeros(self.N, dtype=np.uint8)
            trunc = np.zeros(self.N, dtype=np.uint8)
            nf = np.zeros((self.N, self.S), dtype=np.int64) if want_flat else None
            met = np.zeros((self.N, N_METRICS), dtype=np.int64) if want_metrics else None
        rc = lib().orc_step(self._h, _p(a), _p(a_out), _p(rewards), _p(done), _p(trunc), _p(nf), _p(met))
        if rc != 0:
            raise IndexError("invalid action index for an agent's role list")
        return dict(rewards=rewards, done=done, trunc=trunc, next_flat=nf, metrics=met, actions=a_out)


def flat_size(cfg):
    return lib().orc_flat_size(C.byref(_cfg_struct(cfg)))


def n_role_actions(cfg, is_imposter):
    return lib().orc_n_role_actions(C.byref(_cfg_struct(cfg)), int(is_imposter))


def encode_global(cfg, flat, out=None):
    """flat (n, S) ints -> spatial (n, A+2, 9, 9) f32, non_spatial (A, n, F) f32.  `out` = (sp, ns) to reuse."""
    c = _cfg_struct(cfg)
    f = np.ascontiguousarray(flat, dtype=np.int64)
    n, A = f.shape[0], cfg["n_imposters"] + cfg["n_crew"]
    F = lib().orc_global_nonspatial_size(C.byref(c))
    if out is not None:
        sp, ns = out
    else:
        sp = np.zeros((n, A + 2, 9, 9), dtype=np.float32)
        ns = np.zeros((A, n, F), dtype=np.float32)
    if lib().orc_encode_global(C.byref(c), _p(f), n, _p(sp), _p(ns)) != 0:
        raise IndexError("GlobalFeaturizer needs n_jobs > 0")
    return sp, ns


def encode_perspective(cfg, flat):
    """flat (n, S) ints -> spatial (A, n, A+2, 9, 9) f32, non_spatial (A, n, F) f32."""
    c = _cfg_struct(cfg)
    f = np.ascontiguousarray(flat, dtype=np.int64)
    n, A = f.shape[0], cfg["n_imposters"] + cfg["n_crew"]
    F = lib().orc_perspective_nonspatial_size(C.byref(c))
    sp = np.zeros((A, n, A + 2, 9, 9), dtype=np.float32)
    ns = np.zeros((A, n, F), dtype=np.float32)
    if lib().orc_encode_perspective(C.byref(c), _p(f), n, _p(sp), _p(ns)) != 0:
        raise IndexError("PerspectiveFeaturizer needs n_jobs > 0")
    return sp, ns


def encode_flat(cfg, components, flat):
    """components: names from FLAT_COMPONENTS -> (n, F) f32."""
    c = _cfg_struct(cfg)
    comps = np.array([FLAT_COMPONENTS[x] for x in components], dtype=np.int32)
    f = np.ascontiguousarray(flat, dtype=np.int64)
    F = lib().orc_flat_feature_size(C.byref(c), _p(comps), len(comps))
    if F < 0:
        raise ValueError(f"component list {components} not valid for this env")
    out = np.zeros((f.shape[0], F), dtype=np.float32)
    lib().orc_encode_flat(C.byref(c), _p(comps), len(comps), _p(f), f.shape[0], _p(out))
    return out


def select_actions(cfg, seed, env_ids, act_epoch, alive, imposter_mask, q_imposter, q_crew, eps, imposter_per_view=False):
    """numpy restatement of the acting part of train() (train.py:349-381) on the susnet draw spec (rng_spec, purpose 5):
    alive / imposter_mask (N, A); q_imposter (N, nia) [row e = env e's imposter on its own view] or (A, N, nia) or None;
    q_crew (A, N, nca) or None; -> (N, A) int32 role-list indices."""
    from . import rng_spec as R

    alive = np.asarray(alive) != 0
    imp = np.asarray(imposter_mask) != 0
    N, A = alive.shape
    w = R.words(seed, env_ids, act_epoch, R.P_POLICY, 2 * A)
    nia, nca = n_role_actions(cfg, True), n_role_actions(cfg, False)
    out = np.zeros((N, A), dtype=np.int32)
    for e in range(N):
        for i in range(A):
            if not alive[e, i]:
                continue  # train.py:352: agent_actions starts as zeros
            if imp[e, i]:
                q = None if q_imposter is None else (q_imposter[i, e] if imposter_per_view else q_imposter[e])
                n = nia
            else:
                q = None if q_crew is None else q_crew[i, e]
                n = nca
            explore = q is None or float(w[e, 2 * i]) * 2.0 ** -32 <= float(np.float32(eps))
            out[e, i] = R.bounded(w[e, 2 * i + 1], n) if explore else int(np.argmax(q[:n]))
    return out
