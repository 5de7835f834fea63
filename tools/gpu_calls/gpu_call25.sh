#!/bin/bash
# GPU call 25: small per-rank batches: thread stage vs warp kernels at 1/8 and 1/4 of the genome
mkdir -p gpurun_out
for sc in 0.125 0.25; do
timeout 600 python bench.py --scale $sc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_bench_s${sc}_thread.json 2> gpurun_out/c25_bench_s${sc}_thread.err
AVK_THREAD_MIN_REGIONS=100000000 timeout 600 python bench.py --scale $sc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_bench_s${sc}_warp.json 2> gpurun_out/c25_bench_s${sc}_warp.err
AVK_THREAD_MIN_REGIONS=100000000 AVK_DENSE_N=8 timeout 600 python bench.py --scale $sc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_bench_s${sc}_warp_d8.json 2> gpurun_out/c25_bench_s${sc}_warp_d8.err
AVK_THREAD_POP_BUDGET=24 AVK_DENSE_N=8 timeout 600 python bench.py --scale $sc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_bench_s${sc}_thread_b24d8.json 2> gpurun_out/c25_bench_s${sc}_thread_b24d8.err
done
python tools/bench_line.py gpurun_out/c25_bench_s*.json
