#!/usr/bin/env bash
# GPU trip 23: one GPU's share of BASELINE configs[2] (1.25M users x 1M items x 125M interactions, k=128)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1200 python bench.py --workload c3shard --steps 10 --warmup 3 ) > gpurun_out/bench_c3shard.json 2> gpurun_out/bench_c3shard.err
cat gpurun_out/bench_c3shard.json; tail -5 gpurun_out/bench_c3shard.err
