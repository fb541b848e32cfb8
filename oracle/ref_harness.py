"""TEST INFRASTRUCTURE (oracle/) -- never imported by the product path.

Runs the UNMODIFIED Sus-Net reference (`/root/reference/src`, imported through the
gymnasium stand-in in `tests/_shims`) on *injected* random draws, so the C oracle
and the CUDA kernels can be compared with the real thing bit for bit.

The reference draws from numpy's global RNG at six call sites (SURVEY.md A.6,
`src/environment/base.py:274,288,295,329,374,497`, `tagging.py:167`).
`np.random.choice` / `np.random.shuffle` are looked up at call time, so assigning
wrappers on the module replaces them without touching the reference.  The wrappers
discriminate the call sites by signature and answer with the semantic draw derived
from the susnet Philox spec (`oracle/rng_spec.py`).

Usable where the reference's sources exist: `/root/reference` (the build container) or the staged
copy `baseline/_ref` made by `tools/stage_reference.py` (git-ignored; it travels to the GPU box with the
working-directory snapshot).  Golden fixtures made with this module live in `tests/golden/`.
"""
import os
import sys

import numpy as np

from . import rng_spec as R

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SHIMS = os.path.join(_REPO, "tests", "_shims")


def _find_reference_root():
    for cand in (os.environ.get("SUSNET_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "environment")):
            return cand
    return os.environ.get("SUSNET_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "environment"))


def import_reference():
    """Put the shims + reference on sys.path and return (env module, features module)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (_SHIMS, os.path.join(_SHIMS, "stubs"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import src.environment as env_mod  # noqa: E402
    import src.features as feat_mod  # noqa: E402

    return env_mod, feat_mod


def import_reference_training():
    """The reference's callers of the hot path (needs the pygame / matplotlib / ipywidgets / IPython import stubs):
    returns (src.train, src.replay_memory, src.models.dqn, src.metrics, src.scheduler)."""
    import_reference()
    import src.metrics as metrics_mod  # noqa: E402
    import src.models.dqn as dqn_mod  # noqa: E402
    import src.replay_memory as replay_mod  # noqa: E402
    import src.scheduler as sched_mod  # noqa: E402
    import src.train as train_mod  # noqa: E402

    return train_mod, replay_mod, dqn_mod, metrics_mod, sched_mod


class DrawContext:
    """What the patched numpy functions answer with for the env currently being driven."""

    def __init__(self):
        self.active = False
        self.cfg = None
        self.step_words = None
        self.reset_words = None
        self.act_words = None
        self.kill_events = 0
        self.act_calls = 0


_CTX = DrawContext()
_ORIG = {}


def _choice(a, size=None, replace=True, p=None):
    c = _CTX
    if not c.active:
        return _ORIG["choice"](a, size=size, replace=replace, p=p)
    cfg = c.cfg
    A, n_imp, J, V = cfg["n_agents"], cfg["n_imposters"], cfg["n_jobs"], cfg["n_valid"]
    if isinstance(a, range):  # R1: imposter ids
        assert size == n_imp and replace is False
        imps = sorted(R.pick_distinct(c.reset_words[:n_imp], A))
        return np.array(imps, dtype=np.int64)
    if isinstance(a, (int, np.integer)) and size is not None and replace is True:  # R2: spawn cells
        assert int(a) == V and size == A
        return np.array([R.bounded(c.reset_words[n_imp + i], V) for i in range(A)], dtype=np.int64)
    if isinstance(a, (int, np.integer)) and size is not None and replace is False:  # R3: job cells
        assert int(a) == V and size == J
        base = n_imp + A
        return np.array(R.pick_distinct(c.reset_words[base : base + J], V), dtype=np.int64)
    if isinstance(a, list):  # R5: victim
        assert size is None and len(a) >= 1
        w = c.step_words[A - 1 + c.kill_events]
        c.kill_events += 1
        return a[R.bounded(w, len(a))]
    if isinstance(a, (int, np.integer)) and size is None:  # R6: random action
        w = c.act_words[c.act_calls]
        c.act_calls += 1
        return R.bounded(w, int(a))
    raise AssertionError(f"unexpected np.random.choice call: {a!r}, size={size}, replace={replace}")


def _shuffle(x):
    c = _CTX
    if not c.active:
        return _ORIG["shuffle"](x)
    A = c.cfg["n_agents"]
    assert isinstance(x, list) and len(x) == A  # R4
    x[:] = R.action_order(c.step_words, A)


def install():
    if "choice" not in _ORIG:
        _ORIG["choice"] = np.random.choice
        _ORIG["shuffle"] = np.random.shuffle
    np.random.choice = _choice
    np.random.shuffle = _shuffle


def uninstall():
    if "choice" in _ORIG:
        np.random.choice = _ORIG["choice"]
        np.random.shuffle = _ORIG["shuffle"]


VARIANTS = ("base", "tagging", "training_ground")


def make_reference_env(cfg):
    """cfg: dict in the susnet-b200 config vocabulary (see sus_net_b200.config.EnvConfig)."""
    env_mod, _ = import_reference()
    v = cfg["variant"]
    if v == "training_ground":
        return env_mod.ImposterTrainingGround(
            n_crew=cfg["n_crew"],
            n_jobs=cfg["n_jobs"],
            time_step_reward=cfg["time_step_reward"],
            kill_reward=cfg["kill_reward"],
            sabotage_reward=cfg["sabotage_reward"],
            end_of_game_reward=cfg["game_end_reward"],
            shuffle_imposter_index=cfg["shuffle_imposter_index"],
            include_walls=cfg["include_walls"],
        )
    kw = dict(
        n_imposters=cfg["n_imposters"],
        n_crew=cfg["n_crew"],
        n_jobs=cfg["n_jobs"],
        is_action_order_random=cfg["is_action_order_random"],
        kill_reward=cfg["kill_reward"],
        complete_job_reward=cfg["complete_job_reward"],
        sabotage_reward=cfg["sabotage_reward"],
        time_step_reward=cfg["time_step_reward"],
        game_end_reward=cfg["game_end_reward"],
        dead_penalty=cfg["dead_penalty"],
        shuffle_imposter_index=cfg["shuffle_imposter_index"],
        max_time_steps=cfg["max_time_steps"],
        include_walls=cfg["include_walls"],
    )
    if v == "tagging":
        return env_mod.FourRoomEnvWithTagging(
            **kw, tag_reset_interval=cfg["tag_reset_interval"], vote_reward=cfg["vote_reward"]
        )
    assert v == "base"
    return env_mod.FourRoomEnv(**kw)


METRIC_KEYS = (
    "total_time_steps",
    "imp_killed_crew",
    "completed_jobs",
    "sabotaged_jobs",
    "imp_voted_out",
    "crew_voted_out",
    "crew_won",
    "imposter_won",
)


class ReferenceBatch:
    """N independent reference envs driven in lock step with the batched auto-reset contract
    (SURVEY.md A.7): a step that ends an episode reports the terminal state, then the env is
    reset with the AUTORESET draws of the same tick."""

    def __init__(self, cfg, num_envs, seed, env_id_base=0, auto_reset=True):
        self.cfg = dict(cfg)
        self.N = num_envs
        self.seed = seed
        self.env_ids = np.arange(env_id_base, env_id_base + num_envs, dtype=np.uint64)
        self.auto_reset = auto_reset
        install()
        self.envs = [make_reference_env(self.cfg) for _ in range(num_envs)]
        e0 = self.envs[0]
        self.cfg["n_agents"] = e0.n_agents
        self.cfg["n_imposters"] = e0.n_imposters
        self.cfg["n_valid"] = len(e0.valid_positions)
        self.A, self.J = e0.n_agents, e0.n_jobs
        self.step_tick = 0
        self.reset_epoch = 0
        self.act_epoch = 0
        self.states = [None] * num_envs
        self.S = None

    # -- helpers ---------------------------------------------------------------------------
    def _flat(self, i):
        return np.asarray(self.envs[i].flatten_state(self.states[i]), dtype=np.int64)

    def _metrics(self, i):
        m = self.envs[i].metrics.metrics
        return np.array([int(m[k]) for k in METRIC_KEYS], dtype=np.int64)

    def _reset_one(self, i, words):
        _CTX.active, _CTX.cfg, _CTX.reset_words = True, self.cfg, words
        try:
            s, _ = self.envs[i].reset()
        finally:
            _CTX.active = False
        self.states[i] = s

    # -- API -------------------------------------------------------------------------------
    def reset(self, reset_words=None):
        """reset_words: optional (N, n_imp+A+J) raw words replacing the Philox stream (parity mode)."""
        w = R.words(self.seed, self.env_ids, self.reset_epoch, R.P_RESET,
                    R.n_reset_slots(self.cfg["n_imposters"], self.A, self.J))
        if reset_words is not None:
            w = np.asarray(reset_words, dtype=np.uint32)
        self.reset_epoch += 1
        for i in range(self.N):
            self._reset_one(i, w[i])
        return self.flat_states()

    def flat_states(self):
        return np.stack([self._flat(i) for i in range(self.N)])

    def imposter_idxs(self):
        return np.stack([np.asarray(e.imposter_idxs, dtype=np.int64) for e in self.envs])

    def sample_actions(self, act_words=None):
        w = R.words(self.seed, self.env_ids, self.act_epoch, R.P_ACT, self.A)
        if act_words is not None:
            w = np.asarray(act_words, dtype=np.uint32)
        self.act_epoch += 1
        out = np.zeros((self.N, self.A), dtype=np.int64)
        for i in range(self.N):
            _CTX.active, _CTX.cfg, _CTX.act_words, _CTX.act_calls = True, self.cfg, w[i], 0
            try:
                out[i] = self.envs[i].sample_actions()
            finally:
                _CTX.active = False
        return out

    def step(self, actions, step_words=None, reset_words=None):
        """actions (N, A) ints -> dict of next_flat (pre-reset), rewards f64, done, trunc, metrics (pre-reset)."""
        A = self.A
        ws = R.words(self.seed, self.env_ids, self.step_tick, R.P_STEP, R.n_step_slots(A))
        wr = R.words(self.seed, self.env_ids, self.step_tick, R.P_AUTORESET,
                     R.n_reset_slots(self.cfg["n_imposters"], A, self.J))
        if step_words is not None:
            ws = np.asarray(step_words, dtype=np.uint32)
        if reset_words is not None:
            wr = np.asarray(reset_words, dtype=np.uint32)
        self.step_tick += 1
        S = self.envs[0].flattened_state_size
        out = dict(
            next_flat=np.zeros((self.N, S), dtype=np.int64),
            rewards=np.zeros((self.N, A), dtype=np.float64),
            done=np.zeros(self.N, dtype=np.uint8),
            trunc=np.zeros(self.N, dtype=np.uint8),
            metrics=np.zeros((self.N, len(METRIC_KEYS)), dtype=np.int64),
        )
        for i in range(self.N):
            _CTX.active, _CTX.cfg, _CTX.step_words, _CTX.kill_events = True, self.cfg, ws[i], 0
            try:
                s, r, d, t, _info = self.envs[i].step(np.asarray(actions[i]))
            finally:
                _CTX.active = False
            self.states[i] = s
            out["next_flat"][i] = self._flat(i)
            out["rewards"][i] = r
            out["done"][i] = d
            out["trunc"][i] = t
            out["metrics"][i] = self._metrics(i)
            if self.auto_reset and (d or t):
                self._reset_one(i, wr[i])
        return out


class DrawDrivenReferenceEnv:
    """ONE unmodified reference env whose `reset` / `sample_actions` / `step` consume the draws a susnet env with the same
    (seed, env id) makes, with the same tick contract as the CUDA env's reference mode (no auto-reset: a `reset()` call
    uses the next RESET epoch).  Everything else is the reference object's own attribute, so the reference's callers
    (`ReplayBuffer.populate`, `train()`) can drive it unchanged and be compared with the same callers driving the GPU env."""

    def __init__(self, cfg, seed, env_id=0):
        install()
        self._cfg = dict(cfg)
        self._env = make_reference_env(cfg)
        e = self._env
        self._cfg.update(n_agents=e.n_agents, n_imposters=e.n_imposters, n_valid=len(e.valid_positions))
        self._seed, self._ids = seed, np.array([env_id], dtype=np.uint64)
        self._step_tick = self._reset_epoch = self._act_epoch = 0

    def __getattr__(self, name):
        return getattr(self._env, name)

    def reset(self, *a, **kw):
        A, J, nI = self._env.n_agents, self._env.n_jobs, self._cfg["n_imposters"]
        w = R.words(self._seed, self._ids, self._reset_epoch, R.P_RESET, R.n_reset_slots(nI, A, J))
        self._reset_epoch += 1
        _CTX.active, _CTX.cfg, _CTX.reset_words = True, self._cfg, w[0]
        try:
            return self._env.reset(*a, **kw)
        finally:
            _CTX.active = False

    def sample_actions(self):
        w = R.words(self._seed, self._ids, self._act_epoch, R.P_ACT, self._env.n_agents)
        self._act_epoch += 1
        _CTX.active, _CTX.cfg, _CTX.act_words, _CTX.act_calls = True, self._cfg, w[0], 0
        try:
            return self._env.sample_actions()
        finally:
            _CTX.active = False

    def step(self, agent_actions):
        w = R.words(self._seed, self._ids, self._step_tick, R.P_STEP, R.n_step_slots(self._env.n_agents))
        self._step_tick += 1
        _CTX.active, _CTX.cfg, _CTX.step_words, _CTX.kill_events = True, self._cfg, w[0], 0
        try:
            return self._env.step(agent_actions)
        finally:
            _CTX.active = False
