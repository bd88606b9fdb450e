#!/usr/bin/env bash
# round-2 trip Z (1 GPU): rank, 128-column strips with the chunk loop not unrolled (2 screening instances instead of 4)
mkdir -p gpurun_out
export TUNE_ONLY=c5
rm -f gpurun_out/rank_z.log
for cfg in "2 2" "1 2"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2" >> gpurun_out/rank_z.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_z.log 2>&1
done
grep -E "^==|^rank|rror" gpurun_out/rank_z.log
