#!/usr/bin/env bash
# GPU trip 12: rank kernel v5 (vote-screened nomination), full tests, bench with MF metric
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 400 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
timeout -s KILL 600 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_rank_tc5 python tools/tune_rank.py > gpurun_out/ncu_rank_full.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_rank_tc_gpu.py --deselect tests/test_rank_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_tc.log; cat gpurun_out/tune_rank.log; tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-200; tail -3 gpurun_out/bench.err
