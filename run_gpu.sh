#!/usr/bin/env bash
# GPU trip 43: sanity of the strip heuristic
mkdir -p gpurun_out
( time timeout -s KILL 400 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
TUNE_ONLY=c5 timeout -s KILL 200 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
tail -4 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log
