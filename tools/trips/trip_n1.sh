#!/usr/bin/env bash
# One-GPU trip used during development:  gpurun --timeout 2400 -- bash tools/trips/trip_n1.sh
# full GPU test suite, smoke, default bench, rank sweep; everything lands in gpurun_out/.
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1500 python -m pytest tests -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/tune_rank.log
