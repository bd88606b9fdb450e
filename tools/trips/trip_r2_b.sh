#!/usr/bin/env bash
# round-2 trip B (1 GPU): the CTA-pair rank kernel first (short leash), then the whole GPU suite and the rank timings
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 420 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 200 ) > gpurun_out/pytest_rank_cta2.log 2>&1
rc=$?
echo "rank cta2 exit $rc" >> gpurun_out/pytest_rank_cta2.log
nvidia-smi --query-gpu=index,name,utilization.gpu --format=csv >> gpurun_out/pytest_rank_cta2.log 2>&1
if [ $rc -ne 0 ]; then
  export B200_RANK_CTA=1
  echo "FALLING BACK TO B200_RANK_CTA=1 for the rest of the trip" >> gpurun_out/pytest_rank_cta2.log
  ( time timeout -s KILL 420 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q --timeout 200 ) > gpurun_out/pytest_rank_cta1.log 2>&1
fi
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
for cta in 2 1; do
  for dbg in 0 1; do
    echo "== B200_RANK_CTA=$cta B200_RANK_DEBUG=$dbg" >> gpurun_out/tune_rank_r2.log
    B200_RANK_CTA=$cta B200_RANK_DEBUG=$dbg TUNE_ONLY=c5 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/tune_rank_r2.log 2>&1
  done
done
echo "== full sweep (default)" >> gpurun_out/tune_rank_r2.log
timeout -s KILL 300 python tools/tune_rank.py >> gpurun_out/tune_rank_r2.log 2>&1
tail -5 gpurun_out/pytest_rank_cta2.log; tail -15 gpurun_out/pytest.log; cat gpurun_out/tune_rank_r2.log
