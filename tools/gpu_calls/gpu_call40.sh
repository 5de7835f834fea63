#!/bin/bash
# GPU call 40: ncu source-level capture of the byte-staged CTA DWFA after the sentinel/word-LCP rewrite
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wfa_ed_cta -c 1 -s 1 -o gpurun_out/c40_wfa_bytes python tools/wfa_bench.py 148 > gpurun_out/c40_ncu_bytes.log 2>&1
tail -3 gpurun_out/c40_ncu_bytes.log
