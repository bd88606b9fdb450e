#!/usr/bin/env bash
# round-2 trip Q (1 GPU): evidence pass -- whole GPU suite, smoke, default bench + reference arm, launch list, ncu --set full of the two timed kernels
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
( time timeout -s KILL 600 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
TUNE_ONLY=c5 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 1 -c 1 -o gpurun_out/r02_rank_tc_full -f python tools/tune_rank.py > gpurun_out/ncu_rank.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:bpr_hogwild_stream -s 3 -c 1 -o gpurun_out/r02_bpr_stream_c3_full -f python bench.py --steps 1 --warmup 3 --no-rank --no-mf --no-e2e --no-cpu-baseline > gpurun_out/ncu_bpr.log 2>&1
tail -8 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_n1.json | cut -c1-5000; cat gpurun_out/bench_ref.json | cut -c1-800; tail -3 gpurun_out/ncu_rank.log; tail -3 gpurun_out/ncu_bpr.log
