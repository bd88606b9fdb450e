#!/usr/bin/env bash
# GPU trip 4: tensor-core rank validation first (under a hard timeout), then everything else
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py -m gpu -q --timeout 200 -x > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
nvidia-smi --query-gpu=name,memory.used --format=csv >> gpurun_out/pytest_tc.log 2>&1
B200_RANK_TC=0 timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_rank_tc_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout -s KILL 600 python tools/tune_mf.py > gpurun_out/tune_mf.log 2>&1
tail -25 gpurun_out/pytest_tc.log; tail -6 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/tune_mf.log
