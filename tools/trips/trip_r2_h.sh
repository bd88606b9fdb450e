#!/usr/bin/env bash
# round-2 trip H (1 GPU): rank kernel after the relaxed hand-back arrive of the pair kernel -- parity tests, timings, ncu --set full per variant
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 420 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 200 ) > gpurun_out/pytest_rank.log 2>&1
echo "rank exit $?" >> gpurun_out/pytest_rank.log
export TUNE_ONLY=c5
for cta in 1 2; do
  for st in 2 4; do
    for dbg in 0 8; do
      echo "== B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg" >> gpurun_out/rank_h.log
      B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_h.log 2>&1
    done
  done
done
for v in "1 2" "2 2" "2 4"; do
  set -- $v
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 1 -c 1 -o gpurun_out/rank_tc_full_cta$1_st$2 -f python tools/tune_rank.py > gpurun_out/ncu_full_$1_$2.log 2>&1
done
tail -4 gpurun_out/pytest_rank.log; cat gpurun_out/rank_h.log
