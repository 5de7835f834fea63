#!/bin/bash
# GPU call 18: 2-bit packed, eight-diagonals-per-thread CTA-wide DWFA: parity, microbenchmark, SV sample; streamed bins check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "wfa or sv or last_resort" > gpurun_out/c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
timeout 300 python tools/wfa_bench.py > gpurun_out/c18_wfa_bench.json 2> gpurun_out/c18_wfa_bench.err
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 0.05 > gpurun_out/c18_sv_timing.log 2>&1
AVK_PIPELINE_BINS=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c18_bench_wgs_bins2.json 2> gpurun_out/c18_bench_wgs_bins2.err
AVK_PIPELINE_BINS=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c18_bench_wgs_bins3.json 2> gpurun_out/c18_bench_wgs_bins3.err
tail -3 gpurun_out/c18_pytest.log
cat gpurun_out/c18_wfa_bench.json; tail -2 gpurun_out/c18_wfa_bench.err
tail -4 gpurun_out/c18_sv_timing.log
python tools/bench_line.py gpurun_out/c18_bench_wgs_bins2.json gpurun_out/c18_bench_wgs_bins3.json
