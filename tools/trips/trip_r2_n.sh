#!/usr/bin/env bash
# round-2 trip N (1 GPU): VEBPR / SBPR tests; rank append-store experiments (no store / store to shared memory)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_bprx_gpu.py -q --timeout 300 ) > gpurun_out/pytest_bprx.log 2>&1
echo "bprx exit $?" >> gpurun_out/pytest_bprx.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_n.log
for cfg in "2 4 0" "2 4 128" "2 4 256" "2 2 0" "2 2 128" "2 2 256" "1 2 128"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG_AFTER_WARMUP=$3" >> gpurun_out/rank_n.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_DEBUG_AFTER_WARMUP=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_n.log 2>&1
done
tail -25 gpurun_out/pytest_bprx.log; grep -E "^==|^rank" gpurun_out/rank_n.log
