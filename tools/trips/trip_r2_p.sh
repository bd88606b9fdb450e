#!/usr/bin/env bash
# round-2 trip P (1 GPU): rank append-store cache-policy experiments on the committed kernel (plain / none / .cg / .cs)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
export TUNE_ONLY=c5
rm -f gpurun_out/rank_p.log
for cfg in "2 4 0" "2 4 128" "2 4 256" "2 4 512" "2 2 0" "2 2 128" "2 2 256" "2 2 512" "1 2 0" "1 2 256"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG_AFTER_WARMUP=$3" >> gpurun_out/rank_p.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 TUNE_DEBUG_AFTER_WARMUP=$3 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_p.log 2>&1
done
grep -E "^==|^rank" gpurun_out/rank_p.log
