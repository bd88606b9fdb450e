#!/usr/bin/env bash
# GPU trip 26: device metrics kernel, BaselineOnly, staged rank finish + batched votes
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 900 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py tests/test_models_gpu.py tests/test_mf_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
B200_RANK_FINISH_DIRECT=1 timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank_direct.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_' -c 24 --csv --log-file gpurun_out/launches_rank.csv python tools/tune_rank.py > /dev/null 2>&1
tail -15 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log; cat gpurun_out/tune_rank_direct.log
