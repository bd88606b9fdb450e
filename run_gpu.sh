#!/usr/bin/env bash
# GPU trip 13: BPR scatter experiments + C5-shape ncu capture of the rank kernel
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 900 bash tools/bpr_experiments.sh > gpurun_out/bpr_experiments.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 21 -c 1 -f -o gpurun_out/prof_rank_tc_c5 python tools/tune_rank.py > gpurun_out/ncu_rank_c5.log 2>&1
cat gpurun_out/bpr_experiments.log; tail -3 gpurun_out/ncu_rank_c5.log
