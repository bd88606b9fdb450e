#!/usr/bin/env bash
# GPU trip 17: e2e overlap, norm/finish tweaks -> tests, rank tuning numbers, bench
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log; cat gpurun_out/bench.json
