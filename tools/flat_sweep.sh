#!/bin/bash
# A/B of the Flat-98 fused step (cfg4-alt, 1 Mi envs): staged bytes kernel vs the warp-specialised record kernel and its geometry
OUT=gpurun_out/r02_flat_sweep.jsonl
: > $OUT
SUSNET_PATH=staged python tools/ab_flat.py --steps 100 | sed 's/^/{"path":"staged","r":/; s/$/}/' >> $OUT
for e in 2 3 4 6; do for w in 16 20 24 28; do
  SUSNET_PATH=ws SUSNET_FLATWS_EMITTERS=$e SUSNET_FLATWS_WARPS=$w python tools/ab_flat.py --steps 100 | sed "s/^/{\"path\":\"ws\",\"emitters\":$e,\"warps\":$w,\"r\":/; s/\$/}/" >> $OUT
done; done
SUSNET_COMPRESSIBLE=0 SUSNET_PATH=staged python tools/ab_flat.py --steps 100 | sed 's/^/{"path":"staged","mem":"cudaMalloc","r":/; s/$/}/' >> $OUT
SUSNET_COMPRESSIBLE=0 SUSNET_PATH=ws python tools/ab_flat.py --steps 100 | sed 's/^/{"path":"ws","mem":"cudaMalloc","r":/; s/$/}/' >> $OUT
cat $OUT
