#!/usr/bin/env bash
# round-2 trip K (1 GPU): rank epilogue ceiling experiments -- perfect thresholds from the warm-up call (debug 32), without the scheduled raises (96)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
export TUNE_ONLY=c5
rm -f gpurun_out/rank_k.log
for cfg in "2 4 0" "2 4 32" "2 4 96" "2 2 32" "2 2 96" "1 2 32" "1 2 96" "2 4 8"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG_AFTER_WARMUP=$3" >> gpurun_out/rank_k.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_DEBUG_AFTER_WARMUP=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_k.log 2>&1
done
grep -E "^==|^rank" gpurun_out/rank_k.log
