#!/bin/bash
# GPU call 19: where does a quarter-size WGS pass spend its time?  (launch lists at scale 0.25 and 0.5)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/c19_launches_wgs025.csv python bench.py --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c19_under_ncu_wgs025.log 2>&1
timeout 600 python bench.py --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c19_bench_wgs025.json 2> gpurun_out/c19_bench_wgs025.err
timeout 600 python bench.py --scale 0.5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c19_bench_wgs05.json 2> gpurun_out/c19_bench_wgs05.err
python tools/bench_line.py gpurun_out/c19_bench_wgs025.json gpurun_out/c19_bench_wgs05.json
python tools/launch_summary.py gpurun_out/c19_launches_wgs025.csv 6 2>/dev/null | head -12
