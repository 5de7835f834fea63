#!/bin/bash
# GPU call 39: sentinel-wavefront / uniform-word-LCP CTA DWFA: parity of the WFA + SV paths, microbench, SV sample timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "wfa or sv or last_resort or coop or golden" > gpurun_out/c39_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c39_pytest.log
timeout 300 python tools/wfa_bench.py > gpurun_out/c39_wfa_bench.json 2> gpurun_out/c39_wfa_bench.err
AVK_DEBUG=1 timeout 600 python tools/sv_timing.py 0.05 > gpurun_out/c39_sv_timing.log 2>&1
tail -3 gpurun_out/c39_pytest.log
cat gpurun_out/c39_wfa_bench.json | cut -c1-1500
tail -5 gpurun_out/c39_sv_timing.log
