#!/usr/bin/env bash
# round-2 trip A (1 GPU):  gpurun --timeout 1500 -- bash tools/trips/trip_r2_a.sh
# GPU test suite, smoke, the new default bench (configs[2] on one GPU), reference arm, ncu baselines of the timed kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python - > gpurun_out/host.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, '.')
import bench
print(bench.host_cores())
try: print(open('/sys/fs/cgroup/cpu.max').read().strip())
except Exception as e: print('cpu.max', e)
try: print(open('/sys/fs/cgroup/memory.max').read().strip())
except Exception as e: print('memory.max', e)
os.system("nproc; free -g | head -2")
PY
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
( time timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# ncu: the BPR kernel at the configs[2] block shape, the rank kernels at the configs[4] shape
B200_TUNE_EXPERIMENT=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:bpr_hogwild -s 2 -c 1 \
  -o gpurun_out/r02_bpr_c3 -f python tools/tune_bpr.py --c3 1 --epochs 1 > gpurun_out/ncu_bpr.log 2>&1
TUNE_ONLY=c5 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc -s 3 -c 2 \
  -o gpurun_out/r02_rank_c5 -f python tools/tune_rank.py > gpurun_out/ncu_rank.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-6000; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_ref.json | cut -c1-1500; cat gpurun_out/host.txt; tail -3 gpurun_out/ncu_bpr.log; tail -3 gpurun_out/ncu_rank.log
