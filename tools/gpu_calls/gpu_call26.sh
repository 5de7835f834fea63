#!/bin/bash
# GPU call 26: full suite (VCF ingest + BED builder tests included), small-batch thresholds
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c26_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c26_pytest.log
tail -25 gpurun_out/c26_pytest.log
timeout 600 python bench.py --scale 0.125 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c26_bench_s0125.json 2> gpurun_out/c26_bench_s0125.err
python tools/bench_line.py gpurun_out/c26_bench_s0125.json
