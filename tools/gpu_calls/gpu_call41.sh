#!/bin/bash
# GPU call 41: several whole-genome compares in flight on one GPU (one context per lane)
mkdir -p gpurun_out
timeout 900 python tools/overlap_bench.py 1.0 6 5 > gpurun_out/c41_overlap.json 2> gpurun_out/c41_overlap.err
tail -3 gpurun_out/c41_overlap.err
cat gpurun_out/c41_overlap.json
