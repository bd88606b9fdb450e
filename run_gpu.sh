#!/usr/bin/env bash
# GPU trip 3: tests, bench (atomic default), tuning, ncu on the atomic kernel
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout -s KILL 600 python tools/tune_bpr.py --k 64 > gpurun_out/tune_k64.log 2>&1
timeout -s KILL 600 python tools/tune_bpr.py --k 128 --scale 0.5 > gpurun_out/tune_k128.log 2>&1
timeout -s KILL 600 python tools/tune_mf.py > gpurun_out/tune_mf.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bpr_|mf_|score_|topk_|delta_' -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:bpr_hogwild -s 3 -c 1 -f -o gpurun_out/prof_bpr_atomic python bench.py --steps 1 --warmup 3 --no-e2e --no-rank --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -8 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/tune_k64.log gpurun_out/tune_k128.log gpurun_out/tune_mf.log; tail -3 gpurun_out/ncu_full.log
