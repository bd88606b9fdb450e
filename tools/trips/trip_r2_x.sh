#!/usr/bin/env bash
# round-2 trip X (1 GPU): rank production kernel without the timing-experiment branches (separate DBG instantiation) -- tests, timings
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 300 ) > gpurun_out/pytest_x.log 2>&1
echo "exit $?" >> gpurun_out/pytest_x.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_x.log
for cfg in "2 4" "2 2" "1 2" "1 4"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2" >> gpurun_out/rank_x.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_x.log 2>&1
done
echo "== CTA=2 STRIPS=4 DEBUG=8 (DBG instantiation)" >> gpurun_out/rank_x.log
B200_RANK_DEBUG=8 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_x.log 2>&1
unset TUNE_ONLY
echo "== full sweep, default" >> gpurun_out/rank_x.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_x.log 2>&1
tail -6 gpurun_out/pytest_x.log; grep -E "^==|^rank|rror" gpurun_out/rank_x.log
