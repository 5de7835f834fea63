#!/bin/bash
# GPU call 45 (--gpus 2): bench at N=2 with passes in flight; one-GPU run of an eighth of the genome (the M=6 path of N=8)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c45_bench_wgs_n2.json 2> gpurun_out/c45_bench_wgs_n2.err
echo "bench rc=$?" >> gpurun_out/c45_bench_wgs_n2.err
python tools/bench_line.py gpurun_out/c45_bench_wgs_n2.json
tail -3 gpurun_out/c45_bench_wgs_n2.err
timeout 300 python bench.py --scale 0.125 --no-cpu-baseline > gpurun_out/c45_bench_eighth.json 2> gpurun_out/c45_bench_eighth.err
python tools/bench_line.py gpurun_out/c45_bench_eighth.json
