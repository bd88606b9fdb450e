#!/usr/bin/env bash
# 2-GPU trip used during development:  gpurun --gpus 2 --timeout 1500 -- bash tools/trips/trip_n2.sh
# NCCL path of the bench + sharded-fit functional check (BPR and MF) + the GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
python -c "
import torch, sys
sys.path.insert(0, '.')
for i in range(torch.cuda.device_count()): torch.zeros(1, device='cuda:%d' % i)
torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok', torch.cuda.device_count())
" > gpurun_out/warm.log 2>&1
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/mgpu_check.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?" >> gpurun_out/bench_n2.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
cat gpurun_out/warm.log; tail -4 gpurun_out/mgpu_check.log; cat gpurun_out/bench_n2.json | cut -c1-2500; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2_ref.json | cut -c1-600; tail -6 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log
