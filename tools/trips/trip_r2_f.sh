#!/usr/bin/env bash
# round-2 trip F (1 GPU): re-entry baseline -- whole GPU suite, default bench (configs[2] on one GPU), reference arm, launch list
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
( time timeout -s KILL 600 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
TUNE_ONLY=c5 timeout -s KILL 200 python tools/tune_rank.py > gpurun_out/tune_rank_c5.log 2>&1
timeout -s KILL 300 python tools/tune_replay.py > gpurun_out/tune_replay.log 2>&1
( time timeout -s KILL 600 python examples/bpr_experiment.py ) > gpurun_out/example.log 2>&1
tail -15 gpurun_out/pytest.log; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; cat gpurun_out/bench_ref.json; cat gpurun_out/tune_rank_c5.log; cat gpurun_out/tune_replay.log; tail -14 gpurun_out/example.log
