#!/usr/bin/env bash
# GPU trip 19: full validation of the round-1 state: tests (incl. full-size properties), smoke, example, bench
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 400 --durations=8 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
( time timeout -s KILL 600 python examples/bpr_experiment.py ) > gpurun_out/example.log 2>&1
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -22 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; tail -22 gpurun_out/example.log; cat gpurun_out/bench.json | cut -c1-400
