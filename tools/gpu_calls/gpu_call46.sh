#!/bin/bash
# GPU call 46: eighth / quarter of the genome with passes in flight, thread stage forced on
mkdir -p gpurun_out
for t in 100000; do
AVK_THREAD_MIN_REGIONS=$t timeout 300 python bench.py --scale 0.125 --no-cpu-baseline > gpurun_out/c46_eighth_t$t.json 2> gpurun_out/c46_eighth_t$t.err
python tools/bench_line.py gpurun_out/c46_eighth_t$t.json | cut -c1-330
AVK_THREAD_MIN_REGIONS=$t timeout 300 python bench.py --scale 0.125 --in-flight 1 --no-cpu-baseline > gpurun_out/c46_eighth_t${t}_m1.json 2> gpurun_out/c46_eighth_t$t.err
python tools/bench_line.py gpurun_out/c46_eighth_t${t}_m1.json | cut -c1-330
done
timeout 300 python bench.py --scale 0.25 --no-cpu-baseline > gpurun_out/c46_quarter.json 2> gpurun_out/c46_quarter.err
python tools/bench_line.py gpurun_out/c46_quarter.json | cut -c1-330
