#!/usr/bin/env bash
# GPU trip 31: TMEM-drain experiment, float4 norms, ncu full of the final rank kernel, bench with rank_c5
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 900 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
export TUNE_ONLY=c5
B200_RANK_DEBUG=1 timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/c5_dbg1.log 2>&1
B200_RANK_DEBUG=2 timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/c5_dbg2.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_|scale_' -c 16 --csv --log-file gpurun_out/launches_rank_c5.csv python tools/tune_rank.py > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 1 -c 1 -f -o gpurun_out/rank_tc_c5 python tools/tune_rank.py > gpurun_out/ncu1.log 2>&1
unset TUNE_ONLY
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log gpurun_out/c5_dbg1.log gpurun_out/c5_dbg2.log; cat gpurun_out/bench.json; tail -4 gpurun_out/bench.err
