#!/usr/bin/env bash
# GPU trip 24: prefetch depth experiment at the C3-shard shape and at C2
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
export B200_TUNE_DEPTH=1
timeout -s KILL 600 python tools/tune_bpr.py --k 128 --users 1250000 --items 1000000 --nnz 125000000 > gpurun_out/tune_depth_c3.log 2>&1
timeout -s KILL 600 python tools/tune_bpr.py --k 64 > gpurun_out/tune_depth_c2.log 2>&1
cat gpurun_out/tune_depth_c3.log gpurun_out/tune_depth_c2.log
