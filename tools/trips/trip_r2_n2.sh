#!/usr/bin/env bash
# round-2 2-GPU trip:  gpurun --gpus 2 --timeout 1500 -- bash tools/trips/trip_r2_n2.sh
# sharded-fit functional check (BPR + MF, NVLink peer exchange), the bench at N = 2 (default = p2p exchange; then NCCL for A/B), reference arm
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
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?" >> gpurun_out/bench_n2.err
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --exchange nccl --no-rank --no-mf --no-e2e --no-cpu-baseline ) > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
echo "bench n2 nccl exit $?" >> gpurun_out/bench_n2_nccl.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
cat gpurun_out/warm.log; tail -6 gpurun_out/mgpu_check.log; cat gpurun_out/bench_n2.json | cut -c1-6000; tail -5 gpurun_out/bench_n2.err; cut -c1-1500 gpurun_out/bench_n2_nccl.json; tail -3 gpurun_out/bench_n2_nccl.err; cat gpurun_out/bench_n2_ref.json | cut -c1-600
