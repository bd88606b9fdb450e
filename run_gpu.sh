#!/usr/bin/env bash
mkdir -p gpurun_out
timeout -s KILL 100 python -m pytest tests/test_mf_gpu.py -q -m gpu -k "sharded" > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
