#!/bin/bash
# GPU call 44: passes in flight sweep
mkdir -p gpurun_out
for m in 2 3 6 8; do
timeout 600 python bench.py --in-flight $m --no-cpu-baseline > gpurun_out/c44_bench_m$m.json 2> gpurun_out/c44_bench_m$m.err
python tools/bench_line.py gpurun_out/c44_bench_m$m.json | cut -c1-200
done
