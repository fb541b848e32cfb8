#!/bin/bash
# One GPU-box session for the kernels changed in this commit range: parity tests of the paths that changed, then A/B timings.
# Everything lands under gpurun_out/callA/.  Usage (from the repo root): gpurun --timeout 540 -- 'bash tools/gpu_call_a.sh'
set -u
O=gpurun_out/callA
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > $O/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_replay.py tests/test_gpu_mlp.py tests/test_gpu_train.py tests/test_gpu_policy.py -x -q > $O/pytest_changed.log 2>&1
echo "pytest exit $?" >> $O/pytest_changed.log
tail -3 $O/pytest_changed.log
timeout 120 python tools/bench_replay_push.py > $O/replay_push.json 2> $O/replay_push.err; tail -c 1500 $O/replay_push.json
timeout 120 python tools/bench_mlp.py > $O/mlp.json 2> $O/mlp.err; tail -c 1200 $O/mlp.json; tail -2 $O/mlp.err
timeout 120 python tools/profile_train_loop.py > $O/train_phases_rows128.json 2> $O/train_phases_rows128.err
SUSNET_MLP_ROWS=64 timeout 120 python tools/profile_train_loop.py > $O/train_phases_rows64.json 2> $O/train_phases_rows64.err
grep -h "env_steps_per_s" -A3 $O/train_phases_rows128.json $O/train_phases_rows64.json
timeout 200 python tools/small_batch_sweep.py > $O/small_batch_sweep.jsonl 2> $O/small_batch_sweep.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/callA/small_batch_sweep.jsonl') if l.startswith('{')]
for n in sorted({r['envs'] for r in rows}):
    rs=sorted((r for r in rows if r['envs']==n), key=lambda r: r['median_ms'])
    d=[r for r in rs if r['geometry']=='default'][0]
    print(n, 'default', round(d['median_ms']*1e3,1), 'us; best', rs[0]['geometry'], round(rs[0]['median_ms']*1e3,1), 'us')
PY
