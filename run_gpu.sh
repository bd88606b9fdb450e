#!/usr/bin/env bash
# GPU trip 35: BPR k=128 shard shape, 4 blocks x 256 threads (60 registers)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
F="--workload c3shard --no-e2e --no-rank --no-cpu-baseline --steps 6 --warmup 3"
for t in "33,256,0" "33,128,0" "0,128,0" "33,64,0"; do
  B200_BPR_TUNE=$t timeout -s KILL 300 python bench.py $F 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
done 2>&1 | tee gpurun_out/c3shard_variants2.log
