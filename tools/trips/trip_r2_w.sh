#!/usr/bin/env bash
# round-2 trip W (1 GPU): rank with threshold warm-up stages (B200_RANK_PRE = number of warm-up stages, 0 = off) -- parity tests, timings
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py tests/test_full_size_gpu.py -q -x --timeout 300 ) > gpurun_out/pytest_w.log 2>&1
echo "exit $?" >> gpurun_out/pytest_w.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_w.log
for cfg in "2 4 0" "2 4 -" "2 4 25" "2 4 100" "2 4 200" "2 2 0" "2 2 -" "2 2 100" "1 2 -"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 PRE=$3" >> gpurun_out/rank_w.log
  if [ "$3" = "-" ]; then
    B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_w.log 2>&1
  else
    B200_RANK_PRE=$3 B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_w.log 2>&1
  fi
done
echo "== top10, default / off" >> gpurun_out/rank_w.log
TUNE_TOPK=10 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_w.log 2>&1
B200_RANK_PRE=0 TUNE_TOPK=10 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_w.log 2>&1
unset TUNE_ONLY
echo "== full sweep, default" >> gpurun_out/rank_w.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_w.log 2>&1
tail -6 gpurun_out/pytest_w.log; grep -E "^==|^rank|rror" gpurun_out/rank_w.log
