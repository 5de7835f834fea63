#!/bin/bash
# GPU call 28: thread kernel batch threshold sweep at three batch sizes (thread stage forced on)
mkdir -p gpurun_out
for bm in 1 4 8 16 24; do
for sc in 0.125 0.25 1.0; do
AVK_THREAD_MIN_REGIONS=0 AVK_THREAD_BATCH_MIN=$bm timeout 600 python bench.py --scale $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c28_bench_s${sc}_bm${bm}.json 2> gpurun_out/c28_bench_s${sc}_bm${bm}.err
done; done
python tools/bench_line.py gpurun_out/c28_bench_s*.json
