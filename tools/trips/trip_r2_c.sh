#!/usr/bin/env bash
# round-2 trip C (1 GPU): streamed BPR kernel (tests + A/B timings), rank kernel breakdown (round-1 build vs new, CTA 1 vs 2)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_bpr_gpu.py tests/test_full_size_gpu.py tests/test_models_gpu.py -q --timeout 300 ) > gpurun_out/pytest_bpr.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_bpr.log
for tune in 0,256,0 2,256,0 64,256,0 0,256,3; do
  echo "== c3 block, B200_BPR_TUNE=$tune" >> gpurun_out/tune_bpr_r2.log
  B200_TUNE_EXPERIMENT=1 B200_BPR_TUNE=$tune timeout -s KILL 200 python tools/tune_bpr.py --c3 1 --epochs 4 >> gpurun_out/tune_bpr_r2.log 2>&1
done
for tune in 0,256,0 64,256,0; do
  echo "== c2, B200_BPR_TUNE=$tune" >> gpurun_out/tune_bpr_r2.log
  B200_TUNE_EXPERIMENT=1 B200_BPR_TUNE=$tune timeout -s KILL 200 python tools/tune_bpr.py --k 64 --epochs 4 >> gpurun_out/tune_bpr_r2.log 2>&1
done
# rank: kernel-level breakdown
export TUNE_ONLY=c5
echo "== round-1 rank kernels" >> gpurun_out/rank_ab.log
B200_ALT_LIB=cornac_b200/lib/libb200rank_r1.so timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_ab.log 2>&1
B200_ALT_LIB=cornac_b200/lib/libb200rank_r1.so timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches_r1.csv python tools/tune_rank.py > /dev/null 2>&1
for cta in 1 2; do
  for dbg in 0 2 8; do
    echo "== B200_RANK_CTA=$cta B200_RANK_DEBUG=$dbg" >> gpurun_out/rank_ab.log
    B200_RANK_CTA=$cta B200_RANK_DEBUG=$dbg timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_ab.log 2>&1
  done
  B200_RANK_CTA=$cta timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches_cta$cta.csv python tools/tune_rank.py > /dev/null 2>&1
done
tail -8 gpurun_out/pytest_bpr.log; cat gpurun_out/tune_bpr_r2.log; cat gpurun_out/rank_ab.log
