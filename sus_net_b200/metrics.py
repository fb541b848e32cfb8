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
    """`EpisodicMetricHandler` (metrics.py:67-95) fed from the device-side finished-episode accumulators instead of
    one `info` dict per episode: `compute()` returns the same per-episode averages the reference's
    `sum(values) / len(values)` gives for the counters the env maintains; `save_metrics` writes them as JSON."""

    def __init__(self):
        self.totals = {k: 0 for k in STAT_KEYS}
        self.extra = {}

    def update_from_stats(self, stats):
        """stats: (10,) int64 tensor / sequence in STAT_KEYS order (already reduced over ranks)."""
        vals = stats.tolist() if hasattr(stats, "tolist") else list(stats)
        self.totals = dict(zip(STAT_KEYS, (int(v) for v in vals)))

    def set(self, metrics):
        for k, v in metrics.items():
            self.extra[str(k)] = v

    def compute(self):
        n = max(self.totals["episodes"], 1)
        name_of = {"crew_won": SusMetrics.CREW_WON, "imposter_won": SusMetrics.IMPOSTER_WON,
                   "imp_killed_crew": SusMetrics.IMP_KILLED_CREW, "completed_jobs": SusMetrics.COMPLETED_JOBS,
                   "sabotaged_jobs": SusMetrics.SABOTAGED_JOBS, "imp_voted_out": SusMetrics.IMP_VOTED_OUT,
                   "crew_voted_out": SusMetrics.CREW_VOTED_OUT, "total_time_steps": SusMetrics.TOTAL_TIME_STEPS}
        out = {m: 0.0 for m in SusMetrics}
        for k, m in name_of.items():
            out[m] = self.totals[k] / n
        return out

    def save_metrics(self, save_file_path):
        import json

        with open(save_file_path, "w") as f:
            json.dump({"episodes": self.totals["episodes"], "truncated_episodes": self.totals["truncated_episodes"],
                       "totals": self.totals, "averages": {str(k): v for k, v in self.compute().items()},
                       **self.extra}, f)
