#!/usr/bin/env bash
# first GPU trip: tests, smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --maxfail=30 -x --deselect tests/test_models_gpu.py > gpurun_out/pytest_kernels.log 2>&1
echo "pytest kernels exit $?" >> gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q --timeout 300 --maxfail=30 > gpurun_out/pytest_models.log 2>&1
echo "pytest models exit $?" >> gpurun_out/pytest_models.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest_kernels.log; tail -5 gpurun_out/pytest_models.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
