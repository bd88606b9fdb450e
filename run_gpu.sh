#!/usr/bin/env bash
# GPU trip 39: 16 epilogue warps (4 column strips) vs 8 (2 strips)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
B200_RANK_STRIPS=4 timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py tests/test_models_gpu.py -q -m gpu > gpurun_out/pytest_s4.log 2>&1
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
B200_RANK_STRIPS=4 timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank_s4.log 2>&1
tail -3 gpurun_out/pytest.log; tail -5 gpurun_out/pytest_s4.log; cat gpurun_out/tune_rank.log gpurun_out/tune_rank_s4.log
