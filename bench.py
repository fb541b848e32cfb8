#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused step + observation-encode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU):
`FourRoomEnv(n_imposters=1, n_crew=4, n_jobs=5)` with the reference's default rewards on the walled 9x9 map,
random policy (`env.sample_actions()`), `GlobalFeaturizer` observation encoding, T = 1, auto-reset.
One "step" = every env of the shard advances one time step: `a = env.sample_actions()` (kernel K3, actions
materialised in HBM) then the fused step + encode launch (K1+K2: move/kill/fix/sabotage, win/reward/done,
auto-reset, Global feature tensors of the state the next action is taken from).

Numbers on the JSON line
  value     whole-job env-steps/s with state and actions resident in HBM (CUDA events, max over ranks)
  e2e       the same work driven through the public Python API with HOST buffers (sus_net_b200.HostStepper): per step
            the actions come from pinned host memory (H2D) and rewards/dones/truncations go back to pinned host
            memory (D2H) in the compact host protocol (bit-packed role-list indices in, reward codes + done/trunc
            bits out; `e2e.dense_protocol` = the same with uint8 actions / float32 rewards); the feature tensors
            stay on the device, where the Q-network consumes them
  roofline  achieved = 2650 algorithmic bytes per env-step x envs per launch / mean duration of the
            fused step+encode kernel, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.  The feature
            tensors live in L2-compressible memory, so the DRAM bytes per launch (`traffic`, from the ncu capture in
            profiles/) are about a third of the algorithmic bytes and `frac` can exceed 1; `store_stream` gives the
            fraction of the bare SM -> L2 bulk-store stream, which is what binds the kernel then
  cpu_baseline   the UNMODIFIED Python reference's loop (sample_actions + step + GlobalFeaturizer) on every host core of this
            box, one env per worker process (kind "reference"; its sources travel as the git-ignored baseline/_ref staged by
            tools/stage_reference.py), with the oracle's C port of the same loop beside it (`port`)
Environments are independent: ranks own disjoint env-id ranges (weak scaling, no collective on the step path);
the only collective is the final NCCL all-reduce of the episode statistics.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_STEP_ENCODE = 2650  # 82 B step + 2268 B spatial + 300 B non-spatial
ALGO_BYTES_STEP_ONLY = 82
WORKLOAD = "cfg4: FourRoomEnv 1 imposter vs 4 crew, 5 jobs, walled 9x9 map, random policy, fused step + GlobalFeaturizer encode (T=1), auto-reset"
METRIC = "env-steps/sec (step+obs encode)"
UNIT = "env-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the BASELINE configs[1], [2], cfg4-alt and small-batch extras")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the cpu_baseline sample")
    ap.add_argument("--reference-sample", type=float, default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm
class OracleLoop:
    """The reference loop on the oracle port: a = sample_actions(); step(a); Global encode of the current state.
    All buffers are allocated once (the loop body does no allocation)."""

    def __init__(self, n_envs, threads, seed=1234):
        import oracle

        self.oracle = oracle
        oracle.set_threads(threads)
        self.cfg = oracle.default_config("base", n_crew=4, n_jobs=5)
        self.env = oracle.OracleEnv(self.cfg, n_envs, seed=seed)
        self.n_envs = n_envs
        self.env.reset()
        self.actions = self.env.sample_actions()
        self.out = self.env.step(self.actions, want_flat=False, want_metrics=False)
        self.flat = self.env.flat_states()
        self.feat = oracle.encode_global(self.cfg, self.flat)

    def step(self):
        self.env.sample_actions(out=self.actions)
        self.env.step(self.actions, out=self.out)
        self.env.flat_states(out=self.flat)
        self.oracle.encode_global(self.cfg, self.flat, out=self.feat)

    def rate(self, n_steps):
        t0 = time.perf_counter()
        for _ in range(n_steps):
            self.step()
        dt = time.perf_counter() - t0
        return self.n_envs * n_steps / dt, dt


# ---- the UNMODIFIED reference (pure Python; /root/reference here, the staged baseline/_ref on the GPU box) on all host cores
_REF = {}


def _ref_worker_init(seed_base, counter):
    """One reference env + GlobalFeaturizer per worker process (one env = one 9x9 game, like the reference runs it)."""
    import numpy as np
    import torch

    from oracle import ref_harness as H

    env_mod, feat_mod = H.import_reference()
    torch.set_num_threads(1)
    with counter.get_lock():
        ident = counter.value
        counter.value += 1
    np.random.seed(seed_base + ident)
    env = env_mod.FourRoomEnv(n_imposters=1, n_crew=4, n_jobs=5)  # cfg4, the reference's defaults
    feat = feat_mod.GlobalFeaturizer(env)
    state, _ = env.reset()
    _REF.update(env=env, feat=feat, state=state, torch=torch, seq=np.zeros((1, env.flattened_state_size)))


def _ref_worker_run(m):
    """m env-steps of the reference loop of SURVEY.md 8(d): a = env.sample_actions(); env.step(a); featurizer.fit(seq[None]);
    generate_featurized_states(); reset on done / truncation."""
    env, feat, torch, seq = _REF["env"], _REF["feat"], _REF["torch"], _REF["seq"]
    t0 = time.perf_counter()
    for _ in range(m):
        a = env.sample_actions()
        state, _r, done, trunc, _info = env.step(a)
        if done or trunc:
            state, _ = env.reset()
        seq[0] = env.flatten_state(state)
        feat.fit(torch.tensor(seq).unsqueeze(0))
        feat.generate_featurized_states()
    return time.perf_counter() - t0


class ReferencePool:
    def __init__(self, procs):
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        self.procs = procs
        self.pool = ctx.Pool(procs, initializer=_ref_worker_init, initargs=(1234, ctx.Value("i", 0)))

    def step(self, m):
        """Every worker advances its env m steps; returns (env-steps done, wall seconds)."""
        t0 = time.perf_counter()
        self.pool.map(_ref_worker_run, [m] * self.procs, chunksize=1)
        return self.procs * m, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    """Cores this process may run on (torchrun / cgroup affinity respected), not just the machine's count."""
    try:
        return len(os.sched_getaffinity(0)) or (os.cpu_count() or 1)
    except AttributeError:
        return os.cpu_count() or 1


def reference_available():
    from oracle import ref_harness as H

    return H.reference_available()


def port_baseline(target_seconds):
    threads = host_cores()
    n_envs = 65536
    loop = OracleLoop(n_envs, threads)
    rate, _ = loop.rate(3)  # calibration
    n_steps = max(3, int(rate * target_seconds / n_envs))
    rate, dt = loop.rate(n_steps)
    return {
        "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{n_envs} envs x {n_steps} steps of sample_actions+step+Global encode (oracle C port, OpenMP, {dt:.1f} s)",
    }


def cpu_baseline(target_seconds):
    """The reference's own CPU implementation of the path timed on this box's host cores: the unmodified Python reference
    (one env per worker process, every core) when its sources are staged, with the oracle's C port of the same loop
    beside it; the port alone otherwise."""
    port = port_baseline(min(target_seconds, 6.0))
    if not reference_available():
        port["note"] = "reference sources not staged (tools/stage_reference.py): C port only"
        return port
    # in a fresh interpreter: this process holds a CUDA context, which must not be forked into the worker processes
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--reference-sample", str(target_seconds)],
                         capture_output=True, text=True, timeout=600)
    try:
        ref = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        port["note"] = "timing the staged reference failed: " + (res.stderr.strip().splitlines() or ["no output"])[-1][:200]
        return port
    ref["port"] = port
    return ref


def reference_sample(target_seconds):
    """`bench.py --reference-sample S`: ~S seconds of the unmodified reference loop on every host core -> one JSON line."""
    procs = host_cores()
    pool = ReferencePool(procs)
    try:
        pool.step(20)            # imports, first-call costs
        m = 200
        n, dt = pool.step(m)     # calibration, then grow the sample until it fills the target time
        for _ in range(4):
            if dt >= 0.7 * target_seconds:
                break
            m = max(m + 1, int(m * target_seconds / max(dt, 1e-3)))
            n, dt = pool.step(m)
    finally:
        pool.close()
    print(json.dumps({
        "value": n / dt, "unit": UNIT, "cores": procs, "kind": "reference",
        "sample": f"{procs} worker processes x {m} env-steps of the unmodified reference loop (FourRoomEnv(1, 4, 5).sample_actions "
                  f"+ step + GlobalFeaturizer.fit + generate_featurized_states, one env per process, {dt:.1f} s)",
        "per_core": n / dt / procs}))


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, all of them.  With the
    reference's sources present (/root/reference, or baseline/_ref staged by tools/stage_reference.py) this is the UNMODIFIED
    Python reference, one env per worker process; each bench step is a bounded sample (every worker advances its env m
    steps).  Without them: the oracle's C port of the same loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_cores()
    total_steps = args.steps + args.warmup
    budget = 75.0  # seconds for the whole --steps K --warmup W run
    if reference_available():
        pool = ReferencePool(threads)
        try:
            pool.step(20)  # imports, first-call costs
            pool.step(200)
            n, dt = pool.step(600)
            per_worker = n / dt / threads
            m = max(20, int(per_worker * budget / total_steps))
            for _ in range(args.warmup):
                pool.step(m)
            t0 = time.perf_counter()
            done = 0
            for _ in range(args.steps):
                done += pool.step(m)[0]
            dt = time.perf_counter() - t0
        finally:
            pool.close()
        value = done / dt
        kind = "reference"
        envs_per_step = threads * m
        sample = (f"{threads} worker processes (one unmodified reference env each) x {m} env-steps per bench step: "
                  "sample_actions + step + GlobalFeaturizer.fit + generate_featurized_states")
    else:
        n_envs = 65536
        loop = OracleLoop(n_envs, threads)
        rate, _ = loop.rate(3)
        if n_envs * total_steps / rate > budget:
            n_envs = max(1024, int(rate * budget / total_steps))
            loop = OracleLoop(n_envs, threads)
        loop.rate(args.warmup)
        value, dt = loop.rate(args.steps)
        kind = "port"
        envs_per_step = n_envs
        sample = f"{n_envs} envs per step on {threads} host threads (oracle C port of the reference loop; reference sources not staged)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "env_steps_per_bench_step": envs_per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~2 ms from a thread (the timed region
    is tens of milliseconds, too short for `nvidia-smi -lms`); falls back to nvidia-smi if NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, torch_device):
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.h = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None
        self.index = torch_device.index

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(get_reasons(self.h))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self, t_begin, t_end):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        inside = [(mhz, r) for (t, mhz, r) in self.samples if t_begin <= t <= t_end]
        if not inside:
            return self._smi_fallback()
        mask = 0
        for _, r in inside:
            mask |= r
        return {"sm_mhz": statistics.median(m for m, _ in inside), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit), "samples": len(inside),
                "source": "NVML polled every 2 ms inside the timed region"}

    def _smi_fallback(self):
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                    "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi right after the timed region (NVML polling unavailable)"}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def bind_to_gpu_numa_node(torch_device):
    """Run this rank on the CPUs NVML reports as local to its GPU (first-touch then places the pinned staging buffers
    on that NUMA node); matters for the e2e leg when 8 ranks share the host.  Silent no-op if unavailable."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(torch_device).uuid)).encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpus + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001
        pass


# ------------------------------------------------------------------------------------------------ extras
def other_configs(S, dev, rank, K, peak):
    """Not bench lines (BASELINE configs[1], [2] and the Flat recipe are parity-test cases): the fused kernels of the other
    configurations timed the same way -- CUDA events around the step launch, actions from `sample_actions()` in HBM."""
    import torch

    def timed(env, feat, n_steps, graph_steps=0):
        env.emit_next_states = False
        env.reset()
        for _ in range(5):
            env.step(env.sample_actions(), featurizer=feat)
        torch.cuda.synchronize(dev)
        evs = []
        for _ in range(n_steps):
            a = env.sample_actions()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.step(a, featurizer=feat); e.record()
            evs.append((s, e))
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_steps):
            env.step(env.sample_actions(), featurizer=feat)
        t1.record()
        torch.cuda.synchronize(dev)
        ms = sorted(s.elapsed_time(e) for s, e in evs)
        return ms[len(ms) // 2], t0.elapsed_time(t1) / n_steps

    out = {}
    N = 1 << 20
    # cfg4-alt: ImposterTrainingGround 1v4 + FlatFeaturizer(OneHot + AliveCrew + ClosestAliveCrew) = the reference's training recipe
    env = S.BatchedImposterTrainingGround(n_crew=4, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                          end_of_game_reward=0, num_envs=N, seed=1234, env_id_base=rank * N, device=dev)
    feat = S.FlatFeaturizer(env, S.CompositeFeaturizer([S.OneHotAgentPositionFeaturizer(env), S.AliveCrewFeaturizer(env),
                                                        S.ClosestAliveCrewFeaturizer(env)]))
    k_ms, step_ms = timed(env, feat, K)
    gbs = 453 * N / (k_ms * 1e-3) / 1e9
    out["cfg4alt_itg_1v4_flat98"] = {"envs": N, "kernel": "k_step_flat (fused step + Flat-98 encode, byte-staged rows)", "kernel_ms": k_ms,
                                     "algorithmic_bytes_per_env_step": 453, "achieved_gbs": gbs, "frac_of_hbm_copy_peak": gbs / peak,
                                     "env_steps_per_s_with_sampler": N / (step_ms * 1e-3)}
    del env, feat
    # BASELINE configs[3] at its smallest sweep point: Global features at 65 536 envs
    n = 65536
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=n, seed=1234, device=dev)
    feat = S.GlobalFeaturizer(env)
    k_ms, step_ms = timed(env, feat, K)
    gbs = ALGO_BYTES_STEP_ENCODE * n / (k_ms * 1e-3) / 1e9
    out["cfg4_global_65536_envs"] = {"envs": n, "kernel_ms": k_ms, "achieved_gbs": gbs, "frac_of_hbm_copy_peak": gbs / peak,
                                     "env_steps_per_s_with_sampler": n / (step_ms * 1e-3)}
    # opt-in byte planes (SUS_ENCODE_PLANES_U8): same step + Global encode with one byte per plane cell; its OWN algorithmic
    # bytes (82 step + 567 planes + 300 non-spatial = 949 B per env-step) -- never folded into roofline.frac
    N8 = 1 << 20
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N8, seed=1234, env_id_base=rank * N8, device=dev)
    feat = S.GlobalFeaturizer(env, plane_dtype=torch.uint8)
    k_ms, step_ms = timed(env, feat, K)
    out["cfg4_global_byte_planes_opt_in"] = {"envs": N8, "kernel_ms": k_ms, "algorithmic_bytes_per_env_step": 949,
                                             "achieved_gbs": 949 * N8 / (k_ms * 1e-3) / 1e9,
                                             "env_steps_per_s_with_sampler": N8 / (step_ms * 1e-3),
                                             "what": "GlobalFeaturizer(env, plane_dtype=torch.uint8): uint8 planes for a consumer that casts in its first layer"}
    del env, feat
    # BASELINE configs[2]: tagging 1v2, 5 jobs, 65 536 envs per GPU, step only (issue / latency bound: < 100 B per env-step)
    env = S.BatchedFourRoomEnvWithTagging(1, 2, 5, num_envs=n, seed=1234, device=dev)
    k_ms, step_ms = timed(env, None, K)
    out["cfg3_tagging_1v2_65536_envs_step_only"] = {"envs": n, "kernel_ms": k_ms, "env_steps_per_s_with_sampler": n / (step_ms * 1e-3),
                                                   "algorithmic_gbs": 74 * n / (k_ms * 1e-3) / 1e9}
    env.reset()
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); env.rollout(200); t1.record()
    torch.cuda.synchronize(dev)
    out["cfg3_tagging_1v2_65536_envs_step_only"]["rollout_env_steps_per_s"] = n * 200 / (t0.elapsed_time(t1) * 1e-3)
    del env
    # BASELINE configs[1]: ImposterTrainingGround 1v1 walled, 4096 envs on one GPU (launch-latency bound per step; one-launch rollout)
    n = 4096
    env = S.BatchedImposterTrainingGround(n_crew=1, n_jobs=0, time_step_reward=0, kill_reward=-3, sabotage_reward=0,
                                          end_of_game_reward=0, num_envs=n, seed=1234, device=dev)
    k_ms, step_ms = timed(env, None, K)
    env.reset()
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); env.rollout(1000); t1.record()
    torch.cuda.synchronize(dev)
    out["cfg2_itg_1v1_wall_4096_envs_step_only"] = {"envs": n, "kernel_ms": k_ms, "env_steps_per_s_with_sampler": n / (step_ms * 1e-3),
                                                   "rollout_env_steps_per_s": n * 1000 / (t0.elapsed_time(t1) * 1e-3)}
    del env
    out["small_kernels"] = small_kernels(S, dev, K, peak)
    return out


def small_kernels(S, dev, K, peak):
    """The path's two small kernels with their own byte counts (cfg4, 1 Mi envs): K3 `k_sample_actions` (reads the 16-byte aux
    record, writes A int32 actions: 36 B per env) and row f1's `k_replay_push` (T = 1: reads the sequence row, next_flat,
    cur_flat, actions, rewards, flags, imposters; writes states, next_states, the next sequence row, int64 actions, rewards,
    done, imposters: 24 S + 20 A + 7 = 827 B per transition at S = 30, A = 5)."""
    import ctypes as C

    import torch

    from sus_net_b200 import _lib as L

    N = 1 << 20
    env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=1234, device=dev)
    env.reset()
    buf = S.ReplayBuffer(2 * N, env.flattened_state_size, 1, env.n_agents, 1, device=dev)
    buf.attach(env)
    for _ in range(3):
        buf.collect_step(env.sample_actions())
    torch.cuda.synchronize(dev)
    ev_s, ev_p = [], []
    for _ in range(K):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); acts = env.sample_actions(); a1.record()
        ev_s.append((a0, a1))
        next_flat, rewards, dones, truncated, _ = env.step(acts)
        env.flat_states(out=buf._cur_flat)
        p = L.SusReplayPush(N=N, M=buf.max_size, idx=buf.idx, T=1, S=buf.state_size, A=buf.n_agents, n_imposters=1,
                            seq_in=buf._seq[0].data_ptr(), seq_out=buf._seq[1].data_ptr(), next_flat=next_flat.data_ptr(),
                            cur_flat=buf._cur_flat.data_ptr(), actions=acts.data_ptr(), actions_dtype=L.I32,
                            rewards=rewards.data_ptr(), done=dones.data_ptr(), truncated=truncated.data_ptr(),
                            imposters=env._imposters_buf.data_ptr(), states=buf.states.data_ptr(), r_actions=buf.actions.data_ptr(),
                            r_rewards=buf.rewards.data_ptr(), next_states=buf.next_states.data_ptr(), r_dones=buf.dones.data_ptr(),
                            r_imposters=buf.imposters.data_ptr())
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(); L.check(env.lib.sus_replay_push(C.byref(p), dev.index, env._stream())); p1.record()
        ev_p.append((p0, p1))
        buf._seq.reverse()
        buf.idx = (buf.idx + N) % buf.max_size
    torch.cuda.synchronize(dev)
    med = lambda evs: sorted(s.elapsed_time(e) for s, e in evs)[len(evs) // 2]  # noqa: E731
    ms_s, ms_p = med(ev_s), med(ev_p)
    S_, A_ = env.flattened_state_size, env.n_agents
    b_push = (3 * S_ * 4) + (A_ * 4) + (A_ * 4) + 2 + 2 + (3 * S_ * 4) + (A_ * 8) + (A_ * 4) + 1 + 2
    out = {"k_sample_actions": {"envs": N, "kernel_ms": ms_s, "algorithmic_bytes_per_env": 16 + 4 * A_,
                                "achieved_gbs": (16 + 4 * A_) * N / (ms_s * 1e-3) / 1e9},
           "k_replay_push": {"transitions": N, "kernel_ms": ms_p, "algorithmic_bytes_per_transition": b_push,
                             "achieved_gbs": b_push * N / (ms_p * 1e-3) / 1e9}}
    for v in out.values():
        v["frac_of_hbm_copy_peak"] = v["achieved_gbs"] / peak
    out["k_replay_push"]["kernel"] = "k_replay_push_v (128-bit streams + per-agent rows; round 1's one-thread-per-float kernel: 0.344 ms)"
    del env, buf
    try:
        out["k_mlp_forward"] = mlp_inference(S, dev, K)
    except Exception as exc:  # noqa: BLE001 -- an extra never costs the bench line
        out["k_mlp_forward"] = {"error": repr(exc)}
    return out


def mlp_inference(S, dev, K):
    """Row f2's dominant kernel at BASELINE configs[4]: the Q-network of the cfg5 recipe, MLP [98, 256, 128, 64, 16, 6] with PReLU
    (notebooks/experiment_1v1.ipynb cell 1), evaluated for 131 072 envs in one launch of k_mlp_forward (fp32 FFMA, no tensor cores)
    plus the weight-repacking launch in front of it.  Compute-bound: reported against the fp32 FFMA peak of the part
    (148 SMs x 128 lanes x 2 x the SM clock) and beside the torch module (cuBLAS SGEMMs + elementwise passes)."""
    import torch
    from torch import nn

    dims, rows = [98, 256, 128, 64, 16, 6], 131072
    layers = []
    for i, d in enumerate(dims[:-1]):
        layers += [nn.Linear(d, dims[i + 1]), nn.PReLU()]
    net = nn.Sequential(*layers[:-1]).to(dev)

    class Q(nn.Module):  # the reference's MLP keeps its stack in `.model` and ignores the spatial input (dqn.py:72-93)
        def __init__(self):
            super().__init__()
            self.model = net

        def forward(self, spatial, non_spatial):
            return self.model(non_spatial.view(non_spatial.size(0), -1))

    q = Q()
    fused = S.FusedMLP(q)
    x = (torch.rand(rows, 1, dims[0], device=dev) < 0.15).float()
    sp = torch.zeros(rows, 1, 1, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        evs = []
        for _ in range(max(K, 10)):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            evs.append((s, e))
        torch.cuda.synchronize(dev)
        return sorted(s.elapsed_time(e) for s, e in evs)[len(evs) // 2]

    flop = 2.0 * rows * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    with torch.no_grad():
        ms_torch = timed(lambda: q(sp, x))
        ms = timed(lambda: fused(sp, x))
        err = float((fused(sp, x) - q(sp, x)).abs().max())
    sm_mhz = torch.cuda.get_device_properties(dev).clock_rate / 1e3 if hasattr(torch.cuda.get_device_properties(dev), "clock_rate") else 1965.0
    peak_tf = torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    return {"rows": rows, "dims": dims, "ms": ms, "tflops_fp32": flop / ms / 1e9, "fp32_ffma_peak_tflops": peak_tf,
            "frac_of_fp32_ffma_peak": flop / ms / 1e9 / peak_tf, "torch_module_ms": ms_torch, "max_abs_diff_vs_torch": err,
            "launches": 2}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from sus_net_b200 import build as B

    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        B.build()  # no-op when the in-tree library matches the sources; compiles on a fresh checkout (needs nvcc)
    else:
        t_wait = time.time()
        while B.needs_build() and time.time() - t_wait < 900:  # local rank 0 is compiling
            time.sleep(1.0)
    import sus_net_b200 as S
    from sus_net_b200.distributed import max_over_ranks, reduce_episode_stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(dev)  # pinned host buffers are then allocated next to this GPU's PCIe root
    N, K, W, A = args.envs_per_gpu, args.steps, args.warmup, 5
    seed = 1234
    lib = S.lib()

    def make_env():
        env = S.BatchedFourRoomEnv(1, 4, 5, num_envs=N, seed=seed, env_id_base=rank * N, device=dev)
        env.emit_next_states = False  # the observation IS the feature tensors; raw replay rows are a separate option
        return env

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- arm 1: device-resident throughput (value) + kernel roofline
    env = make_env()
    feat = S.GlobalFeaturizer(env)
    env.reset()
    for _ in range(W):
        env.step(env.sample_actions(), featurizer=feat)
    sampler = ClockSampler(dev)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    barrier()
    launches0 = int(lib.sus_launch_count())
    t_begin = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(K):
        ev[k][0].record()
        a = env.sample_actions()          # K3: actions into HBM
        ev[k][1].record()
        env.step(a, featurizer=feat)      # K1+K2: fused step + encode
        ev[k][2].record()
    stop.record()
    barrier()
    t_end = time.perf_counter()
    launches = int(lib.sus_launch_count()) - launches0
    clocks = sampler.stop(t_begin, t_end)
    elapsed_ms = max_over_ranks(start.elapsed_time(stop), device=dev)
    value = world * N * K / (elapsed_ms * 1e-3)
    kern_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    sample_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    env.check_actions()
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_STEP_ENCODE * N / (kern_ms * 1e-3) / 1e9
    traffic, stream_peak = None, None
    from sus_net_b200.memory import is_compressible

    compressible = is_compressible(feat._sp_buf)
    # DRAM bytes per launch come from an ncu capture (profiles/traffic.json, written by tools/update_traffic.py); they are only
    # reported while the library is the build the capture was made with (content hash of the CUDA sources + flags)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_note = None
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("lib_hash") != B._source_hash():
                traffic_note = "profiles/traffic.json was captured with another build of the library: traffic not reported"
            elif int(tj.get("envs_per_launch", -1)) == N:
                traffic = (tj if compressible else tj.get("uncompressed", {})).get("dram_bytes_per_launch")
            stream_peak = tj.get("store_ceiling_gbs", {}).get("compressible" if compressible else "cudaMalloc")
        except Exception:  # noqa: BLE001
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_step_ws<BASE> (fused step + Global encode, warp-specialised TMA path)", "kernel_ms": kern_ms,
                "algorithmic_bytes_per_env_step": ALGO_BYTES_STEP_ENCODE, "peak_source": peak_src,
                "feature_memory": "L2-compressible (cuMemCreate + COMP_GENERIC)" if compressible else "cudaMalloc"}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    if traffic:
        roofline["dram_gbs"] = traffic / (kern_ms * 1e-3) / 1e9
    if compressible:
        roofline["note"] = ("the feature tensors (97 % of the bytes, almost all zeros) live in L2-compressible memory: the L2 "
                            "writes them to HBM compressed, so the DRAM traffic per launch (`traffic`, ncu) is about a third of "
                            "the algorithmic bytes and `frac` (algorithmic bytes over the measured HBM copy peak) can exceed 1; "
                            "what binds the kernel then is the SM -> L2 store stream, see `store_stream`")
    if stream_peak:
        roofline["store_stream"] = {"peak": stream_peak, "frac": achieved / stream_peak, "unit": "GB/s",
                                    "what": "best SM -> L2 store stream into the same kind of memory over every pattern of "
                                            "tools/micro/store_ceiling_bench.cu (STG.128 at 8-64 warps/SM, bulk tiles of 9-36 KB with "
                                            "1-4 in flight from 1-12 issuing warps, per-lane bulk issue)"}

    # extra (not the headline): the same step with the random policy fused into the step kernel (one launch)
    for _ in range(2):
        env.step(None, featurizer=feat)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(K):
        env.step(None, featurizer=feat)
    e2.record()
    barrier()
    fused_policy_value = world * N * K / (max_over_ranks(s2.elapsed_time(e2), device=dev) * 1e-3)
    # extra: step only (no encode), actions from HBM
    s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a = env.sample_actions().clone()
    env.step(None)
    barrier()
    s3.record()
    for _ in range(K):
        env.step(None)
    e3.record()
    barrier()
    step_only_value = world * N * K / (max_over_ranks(s3.elapsed_time(e3), device=dev) * 1e-3)
    del a
    stats_local = env.episode_stats()
    del env
    other = None
    if world == 1 and not args.no_extra:
        try:  # extras must never cost the headline line
            other = other_configs(S, dev, rank, K, peak)
        except Exception as exc:  # noqa: BLE001
            other = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- arm 2: end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        # record a valid action stream (roles change at every auto-reset, so actions are recorded from a shadow run of
        # the same deterministic trajectory: same seed and env ids => identical states), in both wire formats
        rec = make_env()
        rec.reset()
        cp = rec.compact
        host_packed = torch.empty((W + K, N, cp.action_bytes), dtype=torch.uint8).pin_memory()
        host_dense = torch.empty((W + K, N, A), dtype=torch.uint8).pin_memory()
        for k in range(W + K):
            a = rec.sample_actions()
            host_packed[k].copy_(cp.pack_actions(a), non_blocking=True)
            host_dense[k].copy_(a.to(torch.uint8), non_blocking=True)
            rec.step(a, featurizer=feat)
        torch.cuda.synchronize(dev)
        stats_shadow = rec.episode_stats().clone()
        del rec

        def run_stepper(protocol, host_actions):
            env = make_env()
            env.reset()
            torch.cuda.synchronize(dev)
            stepper = S.HostStepper(env, featurizer=feat, protocol=protocol)  # public API: 3 streams x 2 slots
            for k in range(W):
                stepper.step(host_actions[k])
            stepper.drain()
            barrier()
            s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s4.record()
            for k in range(W, W + K):
                slot = stepper.step(host_actions[k])
            stepper.drain()
            e4.record()
            barrier()
            res = stepper.wait(slot)
            env.check_actions()  # every replayed action was valid for its agent's role
            ok = bool(torch.equal(env.episode_stats(), stats_shadow))  # the host-driven run walked the recorded trajectory
            ms = max_over_ranks(s4.elapsed_time(e4), device=dev)
            return ms, stepper.h2d_bytes_per_step, stepper.d2h_bytes_per_step, res, ok

        e2e_ms, h2d, d2h, res, same = run_stepper("compact", host_packed)
        t0 = time.perf_counter()
        rewards_last = res.rewards  # lazy host decode of the last step's records (numpy, outside the timed region)
        decode_s = time.perf_counter() - t0
        dense_ms, dense_h2d, dense_d2h, _, same_dense = run_stepper("dense", host_dense)
        # the compact chain on ONE stream (no overlap between the copies and the kernel), for reference
        seq = make_env()
        seq.reset()
        d_actions = torch.empty((N, cp.action_bytes), dtype=torch.uint8, device=dev)
        d_res = torch.empty((N, cp.result_bytes), dtype=torch.uint8, device=dev)
        h_res = torch.empty((N, cp.result_bytes), dtype=torch.uint8).pin_memory()

        def seq_step(k):
            d_actions.copy_(host_packed[k], non_blocking=True)
            seq.step(d_actions, featurizer=feat, packed_actions=True, packed_out=d_res)
            h_res.copy_(d_res, non_blocking=True)

        for k in range(W):
            seq_step(k)
        barrier()
        s5, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s5.record()
        for k in range(W, W + K):
            seq_step(k)
        e5.record()
        barrier()
        seq_ms = max_over_ranks(s5.elapsed_time(e5), device=dev)
        e2e = {"value": world * N * K / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h,
               "protocol": f"compact: {cp.action_bytes} B of bit-packed actions in, {cp.result_bytes} B of reward codes + done/trunc "
                           f"bits out per env-step (sus_net_b200/compact.py; decoded rewards are bit-equal to the float64 rewards)",
               "single_stream_value": world * N * K / (seq_ms * 1e-3),
               "dense_protocol": {"value": world * N * K / (dense_ms * 1e-3), "h2d_bytes_per_step": world * dense_h2d,
                                  "d2h_bytes_per_step": world * dense_d2h,
                                  "what": "uint8 actions in, float32 rewards + done + truncated bytes out (round 1's wire format)"},
               "trajectory_check": bool(same and same_dense),
               "host_decode_ms_last_step": 1e3 * decode_s, "host_decoded_reward_sum_last_step": float(rewards_last.sum()),
               "note": "sus_net_b200.HostStepper: actions H2D from pinned host memory, fused step+encode, results D2H to pinned "
                       "host memory every step (3 streams, 2 slots, per-slot feature tensors); feature tensors stay in HBM for "
                       "the Q-network; no sampler kernel in this loop (the actions arrive from the host), which is why it can "
                       "exceed `value`; the host decodes the result records lazily (not in the timed region)"}
        del seq

    # ---- the one collective of the path: final episode-statistics reduce (NCCL)
    stats = reduce_episode_stats(stats_local)
    torch.cuda.synchronize(dev)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": N, "global_envs": world * N,
                       "parallelism": f"env-sharded x{world}, no step-path collective",
                       "l2": f"per-step working set {N * (ALGO_BYTES_STEP_ENCODE + 48) / 1e6:.0f} MB per GPU rewritten every step "
                             + (f"({roofline['traffic'] / 1e6:.0f} MB of DRAM traffic per step after L2 compression) "
                                if roofline.get("traffic") and roofline["traffic"] < 0.9 * N * ALGO_BYTES_STEP_ENCODE else "")
                             + "(> 126 MB L2); no L2 flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "extra": {"sample_actions_kernel_ms": sample_ms, "fused_random_policy_env_steps_per_s": fused_policy_value,
                      "step_only_env_steps_per_s": step_only_value,
                      "step_only_hbm_gbs_algorithmic": ALGO_BYTES_STEP_ONLY * step_only_value / world / 1e9,
                      "other_configs": other,
                      "episode_stats": dict(zip(S.STAT_KEYS, [int(x) for x in stats.tolist()]))},
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
            try:
                line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "",
                                        "error": f"{type(exc).__name__}: {exc}"[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.reference_sample is not None:
        reference_sample(args.reference_sample)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
