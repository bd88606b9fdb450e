#!/usr/bin/env bash
# round-2 trip G (1 GPU): rank kernel breakdown at the configs[4] shape -- CTA group x strips x debug bits, launch list, ncu --set full
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
export TUNE_ONLY=c5
for cta in 1 2; do
  for st in 2 4; do
    for dbg in 0 1 2 8; do
      echo "== B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg" >> gpurun_out/rank_g.log
      B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_g.log 2>&1
    done
  done
done
for cta in 1 2; do
  for st in 2 4; do
  B200_RANK_CTA=$cta B200_RANK_STRIPS=$st timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches_cta${cta}_st${st}.csv python tools/tune_rank.py > /dev/null 2>&1
  done
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 1 -c 1 -o gpurun_out/rank_tc_full -f python tools/tune_rank.py > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/rank_g.log
