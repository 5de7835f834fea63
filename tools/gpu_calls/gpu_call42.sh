#!/bin/bash
# GPU call 42: concurrent lanes of one genome on one GPU: parity test + wall time per genome by lane count
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "lanes or pipelined or multi" > gpurun_out/c42_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c42_pytest.log
tail -5 gpurun_out/c42_pytest.log
AVK_TIMING=1 timeout 900 python tools/overlap_bench.py 1.0 6 1,2,3,4,6,8 24,33,42 > gpurun_out/c42_overlap.json 2> gpurun_out/c42_overlap.err
tail -30 gpurun_out/c42_overlap.err | cut -c1-250
cat gpurun_out/c42_overlap.json
