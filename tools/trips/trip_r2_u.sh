#!/usr/bin/env bash
# round-2 trip U (1 GPU): rank with L2 prefetch of the candidate-list rows (debug bit 2048 switches it off) -- parity tests, timings
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 300 ) > gpurun_out/pytest_u.log 2>&1
echo "exit $?" >> gpurun_out/pytest_u.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_u.log
for cfg in "2 4 0" "2 4 2048" "2 2 0" "2 2 2048" "1 2 0" "1 2 2048"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG=$3" >> gpurun_out/rank_u.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 B200_RANK_DEBUG=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_u.log 2>&1
done
tail -6 gpurun_out/pytest_u.log; grep -E "^==|^rank|rror" gpurun_out/rank_u.log
