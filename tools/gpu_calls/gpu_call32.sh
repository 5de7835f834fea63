#!/bin/bash
# GPU call 32: the FULL SV / long-indel config (BASELINE configs[3], 250 Mbp, 20 k events) on one GPU, one timed pass after one warm-up
mkdir -p gpurun_out
AVK_DEBUG=1 timeout 1500 python tools/sv_timing.py 1.0 > gpurun_out/c32_sv_full.log 2>&1
tail -6 gpurun_out/c32_sv_full.log
