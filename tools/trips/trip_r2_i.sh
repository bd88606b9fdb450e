#!/usr/bin/env bash
# round-2 trip I (1 GPU): rank raise-path experiments (no exclusions / top-10 / no compaction / denser raise schedule), replay pipelining check
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
export TUNE_ONLY=c5
for cfg in "2 4 0 100 100" "2 4 0 100 0" "2 4 0 10 100" "2 4 16 100 100" "2 4 4 100 100" "2 4 20 100 100" "1 2 0 100 0" "1 2 16 100 100" "2 2 16 100 100"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG=$3 topk=$4 excl=$5" >> gpurun_out/rank_i.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 B200_RANK_DEBUG=$3 TUNE_TOPK=$4 TUNE_EXCL=$5 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_i.log 2>&1
done
unset TUNE_ONLY
( time timeout -s KILL 600 python examples/bpr_experiment.py ) > gpurun_out/example.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_bpr_gpu.py tests/test_models_gpu.py -q -x --timeout 300 ) > gpurun_out/pytest_bpr.log 2>&1
cat gpurun_out/rank_i.log; tail -14 gpurun_out/example.log; tail -5 gpurun_out/pytest_bpr.log
