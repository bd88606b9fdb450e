#!/usr/bin/env bash
# round-2 trip Y (1 GPU): rank -- strips per shape (full sweep with ST forced to 2 and to 4), default picks
mkdir -p gpurun_out
rm -f gpurun_out/rank_y.log
for st in 2 4; do
  echo "== full sweep, B200_RANK_STRIPS=$st" >> gpurun_out/rank_y.log
  B200_RANK_STRIPS=$st timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_y.log 2>&1
done
echo "== full sweep, default" >> gpurun_out/rank_y.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_y.log 2>&1
grep -E "^==|^rank|rror" gpurun_out/rank_y.log
