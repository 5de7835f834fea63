#!/bin/bash
# GPU call 15: ncu --set full of k_search_spec on a chr20 pass
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_search_spec -c 1 -s 6 -o gpurun_out/c15_spec_full python bench.py --config chr20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c15_ncu_full.log 2>&1
tail -3 gpurun_out/c15_ncu_full.log | cut -c1-300
