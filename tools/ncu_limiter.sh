#!/bin/bash
# Counters that name the limiter of the headline kernel (k_step_ws, fused step + Global encode, 1 Mi envs), run on the GPU box:
#   bash tools/ncu_limiter.sh   -> gpurun_out/r02_limiter_*.{ncu-rep,csv,txt}
# 1. one `--set full` capture with source import (per-instruction stall samples -> compute warps vs emitter warp by source line)
# 2. a metrics pass with the SM->L2 write path, L2 write sectors and (if the part exposes them) the L2 compression counters
# 3. the metric names this ncu knows about compression, for the record
set -x
OUT=gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
python -c "from sus_net_b200 import build as B; print(B._source_hash())" > $OUT/r02_lib_hash.txt
$BENCH > $OUT/r02_limiter_plain.json 2> $OUT/r02_limiter_plain.err || exit 1
if [ "$1" == "--traffic-only" ]; then
  M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
  ncu --metrics $M --clock-control none -k regex:k_step_ws -s 3 -c 2 --csv --log-file $OUT/r02_limiter_metrics.csv $BENCH > $OUT/r02_limiter_metrics.log 2>&1
  SUSNET_COMPRESSIBLE=0 ncu --metrics $M --clock-control none -k regex:k_step_ws -s 3 -c 1 --csv --log-file $OUT/r02_limiter_metrics_plainmem.csv $BENCH > $OUT/r02_limiter_metrics_plainmem.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $OUT/r02_launches.log 2>&1
  exit 0
fi
ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 3 -c 1 -f -o $OUT/r02_limiter_full $BENCH > $OUT/r02_limiter_full.log 2>&1
ncu --query-metrics 2>/dev/null | grep -i -E "compress|ltcfabric|l1tex2xbar|lts__t_sectors_op_write|lts__t_sectors_srcunit_tex_op_write|lts__d_sectors" > $OUT/r02_limiter_metric_names.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum"
M="$M,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_bytes.sum,lts__t_sectors.sum"
M="$M,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed"
M="$M,sm__inst_executed_pipe_uniform.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
M="$M,lts__t_sectors_op_write_lookup_hit.sum,lts__t_sectors_op_write_lookup_miss.sum"
M="$M,lts__average_t_sectors_per_request_op_write.ratio,l1tex__m_l1tex2xbar_throughput.avg.pct_of_peak_sustained_elapsed"
M="$M,lts__t_sectors_srcunit_ltcfabric.sum,lts__ltcfabric2lts_cycles_active.avg.pct_of_peak_sustained_elapsed"
M="$M,lts__d_sectors_fill_device.sum,lts__t_sectors_srcnode_gpc_op_write.sum"
ncu --metrics $M --clock-control none -k regex:k_step_ws -s 3 -c 2 --csv --log-file $OUT/r02_limiter_metrics.csv $BENCH > $OUT/r02_limiter_metrics.log 2>&1
COMP=$(grep -i compress $OUT/r02_limiter_metric_names.txt | awk '{print $1}' | head -12 | tr '\n' ',' | sed 's/,$//')
if [ -n "$COMP" ]; then
  ncu --metrics $COMP --clock-control none -k regex:k_step_ws -s 3 -c 1 --csv --log-file $OUT/r02_limiter_compression.csv $BENCH > $OUT/r02_limiter_compression.log 2>&1
fi
# the same two passes with the features in ordinary memory (no L2 compression), to separate the fabric from the compressor
SUSNET_COMPRESSIBLE=0 ncu --metrics $M --clock-control none -k regex:k_step_ws -s 3 -c 1 --csv --log-file $OUT/r02_limiter_metrics_plainmem.csv $BENCH > $OUT/r02_limiter_metrics_plainmem.log 2>&1
# launch list of the bench command (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02_launches.log 2>&1
./tools/micro/store_ceiling_bench > $OUT/r02_store_ceiling.json 2> $OUT/r02_store_ceiling.err
ls -la $OUT | grep r02_limiter
