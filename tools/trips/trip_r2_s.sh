#!/usr/bin/env bash
# round-2 trip S (1 GPU): rank candidate store through the TMA (timing experiment, debug bit 1024) vs LSU store vs no store
mkdir -p gpurun_out
export TUNE_ONLY=c5
rm -f gpurun_out/rank_s.log
for cfg in "2 4 0" "2 4 128" "2 4 1024" "2 2 0" "2 2 128" "2 2 1024"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG_AFTER_WARMUP=$3" >> gpurun_out/rank_s.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_DEBUG_AFTER_WARMUP=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_s.log 2>&1
done
grep -E "^==|^rank|rror" gpurun_out/rank_s.log
