#!/usr/bin/env bash
# round-2 trip D (1 GPU): rank epilogue v3 (64-column strips, early accumulator hand-back) + cache-blocked BPR order
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 420 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 200 ) > gpurun_out/pytest_rank.log 2>&1
rc=$?
echo "rank exit $rc" >> gpurun_out/pytest_rank.log
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
export TUNE_ONLY=c5
echo "== round-1 rank kernels" >> gpurun_out/rank_v3.log
B200_ALT_LIB=cornac_b200/lib/libb200rank_r1.so timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_v3.log 2>&1
for cta in 1 2; do
  for st in 4 2; do
    for dbg in 0 4; do
      echo "== B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg" >> gpurun_out/rank_v3.log
      B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_v3.log 2>&1
    done
  done
done
unset TUNE_ONLY
echo "== default, full sweep" >> gpurun_out/rank_v3.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/rank_v3.log 2>&1
for blocked in 0 1; do
  echo "== c3 block, blocked=$blocked" >> gpurun_out/tune_bpr_d.log
  B200_TUNE_EXPERIMENT=1 timeout -s KILL 200 python tools/tune_bpr.py --c3 1 --epochs 4 --blocked $blocked >> gpurun_out/tune_bpr_d.log 2>&1
  echo "== c2, blocked=$blocked" >> gpurun_out/tune_bpr_d.log
  B200_TUNE_EXPERIMENT=1 timeout -s KILL 200 python tools/tune_bpr.py --k 64 --epochs 4 --blocked $blocked >> gpurun_out/tune_bpr_d.log 2>&1
done
echo "== c2, streamed D=2" >> gpurun_out/tune_bpr_d.log
B200_TUNE_EXPERIMENT=1 B200_BPR_TUNE=2,256,0 timeout -s KILL 200 python tools/tune_bpr.py --k 64 --epochs 4 >> gpurun_out/tune_bpr_d.log 2>&1
echo "== c2, streamed D=2, blocked" >> gpurun_out/tune_bpr_d.log
B200_TUNE_EXPERIMENT=1 B200_BPR_TUNE=2,256,0 timeout -s KILL 200 python tools/tune_bpr.py --k 64 --epochs 4 --blocked 1 >> gpurun_out/tune_bpr_d.log 2>&1
tail -4 gpurun_out/pytest_rank.log; tail -12 gpurun_out/pytest.log; cat gpurun_out/rank_v3.log; cat gpurun_out/tune_bpr_d.log
