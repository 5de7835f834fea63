#!/bin/bash
# GPU call 22: full GPU suite + benches (WGS N=1 with CPU leg, chr20, quarter scale), launch list and ncu of the top kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c22_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c22_bench_wgs.json 2> gpurun_out/c22_bench_wgs.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c22_bench_chr20.json 2> gpurun_out/c22_bench_chr20.err
timeout 600 python bench.py --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c22_bench_wgs025.json 2> gpurun_out/c22_bench_wgs025.err
timeout 300 python tools/wfa_bench.py > gpurun_out/c22_wfa_bench.json 2> gpurun_out/c22_wfa_bench.err
tail -3 gpurun_out/c22_pytest.log
python tools/bench_line.py gpurun_out/c22_bench_wgs.json gpurun_out/c22_bench_chr20.json gpurun_out/c22_bench_wgs025.json
cut -c1-400 gpurun_out/c22_wfa_bench.json
