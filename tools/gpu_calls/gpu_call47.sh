#!/bin/bash
# GPU call 47: eighth of the genome with passes in flight: lanes pick the thread stage themselves; pop budget 32 (default) / 64
mkdir -p gpurun_out
timeout 300 python bench.py --scale 0.125 --no-cpu-baseline > gpurun_out/c47_eighth.json 2> gpurun_out/c47_eighth.err
python tools/bench_line.py gpurun_out/c47_eighth.json | cut -c1-330
AVK_LANE_POP_BUDGET=64 timeout 300 python bench.py --scale 0.125 --no-cpu-baseline > gpurun_out/c47_eighth_p64.json 2> gpurun_out/c47_eighth_p64.err
python tools/bench_line.py gpurun_out/c47_eighth_p64.json | cut -c1-330
AVK_LANE_POP_BUDGET=16 timeout 300 python bench.py --scale 0.125 --no-cpu-baseline > gpurun_out/c47_eighth_p16.json 2> gpurun_out/c47_eighth_p16.err
python tools/bench_line.py gpurun_out/c47_eighth_p16.json | cut -c1-330
AVK_LANE_POP_BUDGET=64 timeout 300 python bench.py --scale 0.25 --no-cpu-baseline > gpurun_out/c47_quarter_p64.json 2> gpurun_out/c47_quarter_p64.err
python tools/bench_line.py gpurun_out/c47_quarter_p64.json | cut -c1-330
AVK_LANE_POP_BUDGET=32 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c47_full_p32.json 2> gpurun_out/c47_full_p32.err
python tools/bench_line.py gpurun_out/c47_full_p32.json | cut -c1-330
