"""Per-episode counters and finished-episode statistics (src/metrics.py:7-64)."""
from enum import StrEnum, auto

try:  # inside a Sus-Net checkout reuse its enum so `info` dict keys are the caller's own
    from src.metrics import SusMetrics  # type: ignore
except Exception:  # noqa: BLE001

    class SusMetrics(StrEnum):  # metrics.py:7-20
        IMP_KILLED_CREW = auto()
        IMP_VOTED_OUT = auto()
        CREW_VOTED_OUT = auto()
        SABOTAGED_JOBS = auto()
        COMPLETED_JOBS = auto()
        TOTAL_STALEMATES = auto()
        TOTAL_TIME_STEPS = auto()
        IMPOSTER_WON = auto()
        CREW_WON = auto()
        AVG_CREW_RETURNS = auto()
        AVG_IMPOSTER_RETURNS = auto()
        CREW_LOSS = auto()
        IMPOSTER_LOSS = auto()


# order of SusMetricIndex in include/susnet_b200.h
METRIC_ORDER = (SusMetrics.TOTAL_TIME_STEPS, SusMetrics.IMP_KILLED_CREW, SusMetrics.COMPLETED_JOBS,
                SusMetrics.SABOTAGED_JOBS, SusMetrics.IMP_VOTED_OUT, SusMetrics.CREW_VOTED_OUT, SusMetrics.CREW_WON,
                SusMetrics.IMPOSTER_WON)
# order of SusStatIndex
STAT_KEYS = ("episodes", "crew_won", "imposter_won", "imp_killed_crew", "completed_jobs", "sabotaged_jobs",
             "imp_voted_out", "crew_voted_out", "total_time_steps", "truncated_episodes")


class EnvMetricView:
    """Read-only stand-in for `EnvMetricHandler` (metrics.py:35-64): the counters live in the env's device state;
    `metrics` / `get_metrics()` return the reference's 13-key dict for the single env of reference mode."""

    def __init__(self, env):
        self._env = env

    def get_metrics(self):
        vals = getattr(self._env, "_host_metrics", None)
        out = {m: 0 for m in SusMetrics}
        if vals is not None:
            for k, v in zip(METRIC_ORDER, vals):
                out[k] = int(v)
        return out

    @property
    def metrics(self):
        return self.get_metrics()

    def __repr__(self):
        import json

        return json.dumps(self.get_metrics(), indent=4)


class EpisodicMetricHandler:
    """`EpisodicMetricHandler` (metrics.py:67-95) for batched runs.  The reference appends one `info` dict per finished episode
    and averages the lists; a batched run finishes millions of episodes between two looks at the host, so the lists hold one
    entry per LOGGING INTERVAL instead: the mean over the episodes that finished in it (`log_interval`), computed from the
    device-side finished-episode accumulators (already reduced over ranks).  `save_metrics` writes exactly the reference's
    schema -- `{metric name: list}` for all 13 `SusMetrics`, loadable by the reference's own `load_metrics` / `compute`
    (metrics.py:88-95) -- and the exact totals go to a `*_totals.json` sidecar.  `compute()` returns the exact per-episode
    averages from the totals (the reference's `sum(values) / len(values)` over per-episode values)."""

    _COUNTERS = {"crew_won": SusMetrics.CREW_WON, "imposter_won": SusMetrics.IMPOSTER_WON,
                 "imp_killed_crew": SusMetrics.IMP_KILLED_CREW, "completed_jobs": SusMetrics.COMPLETED_JOBS,
                 "sabotaged_jobs": SusMetrics.SABOTAGED_JOBS, "imp_voted_out": SusMetrics.IMP_VOTED_OUT,
                 "crew_voted_out": SusMetrics.CREW_VOTED_OUT, "total_time_steps": SusMetrics.TOTAL_TIME_STEPS}

    def __init__(self):
        self.metrics = {m: [] for m in SusMetrics}  # metrics.py:72-73
        self.totals = {k: 0 for k in STAT_KEYS}
        self.return_sums = [0.0, 0.0]  # summed imposter / crew returns over finished episodes (train.py:421-424)
        self._last = ({k: 0 for k in STAT_KEYS}, [0.0, 0.0])
        self.extra = {}

    def update_from_stats(self, stats, return_sums=None):
        """stats: (10,) int64 tensor / sequence in STAT_KEYS order (already reduced over ranks); return_sums: optional (2,)."""
        vals = stats.tolist() if hasattr(stats, "tolist") else list(stats)
        self.totals = dict(zip(STAT_KEYS, (int(v) for v in vals)))
        if return_sums is not None:
            self.return_sums = [float(x) for x in (return_sums.tolist() if hasattr(return_sums, "tolist") else return_sums)]

    def log_interval(self, stats, return_sums=None):
        """Append the per-episode means of the episodes finished since the last call (nothing if none finished)."""
        self.update_from_stats(stats, return_sums)
        last_t, last_r = self._last
        n = self.totals["episodes"] - last_t["episodes"]
        if n > 0:
            for k, m in self._COUNTERS.items():
                self.metrics[m].append((self.totals[k] - last_t[k]) / n)
            self.metrics[SusMetrics.TOTAL_STALEMATES].append(0.0)  # never incremented by the reference's envs either
            if return_sums is not None:
                self.metrics[SusMetrics.AVG_IMPOSTER_RETURNS].append((self.return_sums[0] - last_r[0]) / n)
                self.metrics[SusMetrics.AVG_CREW_RETURNS].append((self.return_sums[1] - last_r[1]) / n)
        self._last = (dict(self.totals), list(self.return_sums))
        return n

    def step(self, metrics):  # metrics.py:75-77 (one finished episode of a single env)
        for metric, value in metrics.items():
            self.metrics[metric].append(value)

    def set(self, metrics):  # metrics.py:79-82
        for k, v in metrics.items():
            assert any(m.value == k for m in SusMetrics), f"Invalid metric: {k}"
            self.metrics[k] = v

    def compute(self):
        """Per-episode averages: exact from the device totals where the env maintains the counter, else (returns, losses,
        single-env `step()` use) the reference's mean over the list."""
        out = {}
        n = self.totals["episodes"]
        by_metric = {m: k for k, m in self._COUNTERS.items()}
        for m in SusMetrics:
            vals = self.metrics[m]
            if n > 0 and m in by_metric:
                out[m] = self.totals[by_metric[m]] / n
            elif n > 0 and m == SusMetrics.AVG_IMPOSTER_RETURNS and any(self.return_sums):
                out[m] = self.return_sums[0] / n
            elif n > 0 and m == SusMetrics.AVG_CREW_RETURNS and any(self.return_sums):
                out[m] = self.return_sums[1] / n
            else:
                out[m] = sum(vals) / len(vals) if len(vals) else 0.0
        return out

    def save_metrics(self, save_file_path):
        """metrics.py:88-90: `json.dump(self.metrics)` -- {metric name: list}.  Lists that would be empty get the overall mean
        as their single entry so that the reference's `compute()` can average every key of a loaded file."""
        import json

        avg = self.compute()
        out = {str(m.value): (list(v) if len(v) else [avg[m]]) for m, v in self.metrics.items()}
        with open(save_file_path, "w") as f:
            json.dump(out, f)
        side = str(save_file_path)
        side = (side[:-5] if side.endswith(".json") else side) + "_totals.json"
        with open(side, "w") as f:
            json.dump({"episodes": self.totals["episodes"], "truncated_episodes": self.totals["truncated_episodes"],
                       "totals": self.totals, "return_sums": self.return_sums,
                       "averages": {str(k.value): v for k, v in avg.items()}, **self.extra}, f)

    def load_metrics(self, metrics_file_path):  # metrics.py:92-95
        import json

        with open(metrics_file_path, "r") as f:
            self.metrics = json.load(f)
