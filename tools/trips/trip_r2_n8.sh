#!/usr/bin/env bash
# round-2 8-GPU trip:  gpurun --gpus 8 --timeout 900 -- bash tools/trips/trip_r2_n8.sh      (also used with --gpus 4: N follows the visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi --query-gpu=index --format=csv,noheader | wc -l)
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt 2>&1
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/mgpu_check_n$N.log
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?" >> gpurun_out/bench_n$N.err
grep -v "^\[\|^W09\|^E09\|Traceback\|File \|^    \|^  \|torch\.\|^=\|^-\|^\*\|Setting OMP\|^$" gpurun_out/mgpu_check_n$N.log | tail -30; cut -c1-3000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
