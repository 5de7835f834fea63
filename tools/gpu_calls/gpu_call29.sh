#!/bin/bash
# GPU call 29: batch threshold fine sweep at full scale
mkdir -p gpurun_out
for bm in 12 14 16 18 20; do
AVK_THREAD_BATCH_MIN=$bm timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c29_bench_bm${bm}.json 2> gpurun_out/c29_bench_bm${bm}.err
done
for bm in 12 16 20; do
AVK_THREAD_BATCH_MIN=$bm timeout 600 python bench.py --scale 0.5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c29_bench_s05_bm${bm}.json 2> gpurun_out/c29_bench_s05_bm${bm}.err
done
python tools/bench_line.py gpurun_out/c29_bench_*.json
