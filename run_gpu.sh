#!/usr/bin/env bash
# GPU trip 37: full validation of HEAD + records for profiles/
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
( time timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout -s KILL 600 python bench.py --workload c3shard --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3shard.json 2> gpurun_out/bench_c3shard.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bpr_|mf_|score_|topk_|delta_|rank_tc|pack_|norm_|scale_' -c 160 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --no-cpu-baseline > /dev/null 2>&1
tail -6 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -4 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.json; cut -c1-900 gpurun_out/bench_c3shard.json
