#!/usr/bin/env bash
# GPU trip 5: warm the box, tensor-core rank validation (hard timeout), chunk-kernel tuning, bench
mkdir -p gpurun_out
( time python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" ) > gpurun_out/warm.log 2>&1
timeout -s KILL 500 python -m pytest tests/test_rank_tc_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
nvidia-smi --query-gpu=name,memory.used --format=csv >> gpurun_out/pytest_tc.log 2>&1
timeout -s KILL 600 python tools/tune_bpr.py --k 64 > gpurun_out/tune_k64.log 2>&1
timeout -s KILL 600 python tools/tune_bpr.py --k 128 --scale 0.5 > gpurun_out/tune_k128.log 2>&1
timeout -s KILL 900 python -m pytest tests/test_bpr_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
cat gpurun_out/warm.log; tail -30 gpurun_out/pytest_tc.log; cat gpurun_out/tune_k64.log gpurun_out/tune_k128.log; tail -6 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
