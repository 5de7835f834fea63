#!/bin/bash
# GPU call 48 (--gpus N): bench at N GPUs with passes in flight
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/c48_bench_wgs_n$N.json 2> gpurun_out/c48_bench_wgs_n$N.err
echo "bench rc=$?" >> gpurun_out/c48_bench_wgs_n$N.err
python tools/bench_line.py gpurun_out/c48_bench_wgs_n$N.json
tail -2 gpurun_out/c48_bench_wgs_n$N.err
