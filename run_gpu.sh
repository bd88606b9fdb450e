#!/usr/bin/env bash
# GPU trip 42: sanity of the strip heuristic (rank + model tests, smoke)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py tests/test_models_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
tail -4 gpurun_out/pytest.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/tune_rank.log
