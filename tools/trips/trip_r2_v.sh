#!/usr/bin/env bash
# round-2 trip V (1 GPU): ncu --set full of the rank kernel with and without the candidate store (debug bit 128), to diff the metrics
mkdir -p gpurun_out
export TUNE_ONLY=c5
for st in 2 4; do
for m in 0 128; do
  B200_RANK_CTA=2 B200_RANK_STRIPS=$st TUNE_DEBUG_AFTER_WARMUP=$m timeout -s KILL 600 ncu --set full --clock-control none -k regex:rank_tc_kernel -s 1 -c 1 -o gpurun_out/rank_store_st${st}_m$m -f python tools/tune_rank.py > gpurun_out/ncu_v_${st}_$m.log 2>&1
done
done
ls -la gpurun_out/rank_store_*.ncu-rep
