#!/usr/bin/env bash
# GPU trip 38: BPR tests + configs[2]-shard bench with the 4-block default for 32-lane groups
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 900 python -m pytest tests/test_bpr_gpu.py tests/test_full_size_gpu.py tests/test_models_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --workload c3shard --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3shard.json 2> gpurun_out/bench_c3shard.err
tail -4 gpurun_out/pytest.log; cut -c1-200 gpurun_out/bench_c3shard.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c3shard.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['rank'], d['mf'])"
