#!/bin/bash
# GPU call 12: launch lists (ncu, serialised per-kernel durations) of a chr20 pass and a WGS pass; heaviest clusters alone
mkdir -p gpurun_out
AVK_PIPELINE_BINS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c12_launches_chr20.csv python bench.py --config chr20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c12_under_ncu_chr20.log 2>&1
AVK_PIPELINE_BINS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/c12_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c12_under_ncu_wgs.log 2>&1
timeout 300 python tools/heavy_clusters.py > gpurun_out/c12_heavy.txt 2>&1
cat gpurun_out/c12_heavy.txt
