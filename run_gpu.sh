#!/usr/bin/env bash
# GPU trip 22: windowed replay with batched metadata: tests + throughput + example
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 600 python -m pytest tests/test_bpr_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 python tools/tune_replay.py > gpurun_out/tune_replay.log 2>&1
( time timeout -s KILL 600 python examples/bpr_experiment.py ) > gpurun_out/example.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/tune_replay.log; tail -12 gpurun_out/example.log
