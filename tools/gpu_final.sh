#!/bin/bash
# The round's closing GPU session with the final library: full GPU test suite, DRAM traffic of the headline kernel (keyed by the
# library hash, so that the bench line carries `roofline.traffic`), the bench line, ncu summaries of the headline kernel and of
# the MLP inference kernel, the launch list, the training-loop phases.  Everything lands under gpurun_out/final/.
#   gpurun --timeout 900 -- 'bash tools/gpu_final.sh'
set -u
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -2 $O/pytest_gpu.log
# traffic of k_step_ws per launch (compressible / cudaMalloc feature memory) + launch list, then the traffic file for THIS build
bash tools/ncu_limiter.sh --traffic-only > $O/ncu_limiter.log 2>&1
python tools/update_traffic.py gpurun_out/r02_limiter_metrics.csv gpurun_out/r02_limiter_metrics_plainmem.csv profiles/r02_store_ceiling.json gpurun_out/r02_lib_hash.txt > $O/update_traffic.log 2>&1
cp profiles/traffic.json $O/traffic.json
timeout 400 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench exit $?"; head -c 900 $O/bench_1gpu.json; echo
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 3 -c 1 -f -o $O/k_step_ws_full $BENCH > $O/k_step_ws_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_mlp_forward -s 6 -c 1 -f -o $O/k_mlp_forward_full python tools/bench_mlp.py --reps 2 > $O/k_mlp_forward_full.log 2>&1
timeout 120 python tools/profile_train_loop.py > $O/train_loop_phases_1gpu.json 2> $O/train_loop_phases_1gpu.err; grep -o '"cuda_graphs": [0-9.]*' $O/train_loop_phases_1gpu.json
timeout 100 python tools/train_demo.py --seconds 5 > $O/train_demo_1gpu.json 2> $O/train_demo_1gpu.err; tail -c 400 $O/train_demo_1gpu.json
ls -la $O
