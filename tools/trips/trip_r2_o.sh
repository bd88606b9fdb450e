#!/usr/bin/env bash
# round-2 trip O (1 GPU): rank with shared-memory staged appends -- parity tests, timings; VEBPR / SBPR tests
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py tests/test_bprx_gpu.py -q --timeout 300 ) > gpurun_out/pytest_o.log 2>&1
echo "exit $?" >> gpurun_out/pytest_o.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_o.log
for cfg in "1 2 100" "2 2 100" "2 4 100" "2 4 10" "2 2 10"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 topk=$3" >> gpurun_out/rank_o.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_TOPK=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_o.log 2>&1
done
for cfg in "2 2 32" "2 4 32" "2 2 8"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG_AFTER_WARMUP=$3" >> gpurun_out/rank_o.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_DEBUG_AFTER_WARMUP=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_o.log 2>&1
done
unset TUNE_ONLY
echo "== full sweep, default" >> gpurun_out/rank_o.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_o.log 2>&1
tail -8 gpurun_out/pytest_o.log; grep -E "^==|^rank" gpurun_out/rank_o.log
