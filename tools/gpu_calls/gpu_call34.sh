#!/bin/bash
# GPU call 34: zero-flip shortcut fix (skipped no-op ALTs), adversarial random clusters through every pipeline variant; full suite; bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c34_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c34_pytest.log
tail -12 gpurun_out/c34_pytest.log
timeout 900 python bench.py > gpurun_out/c34_bench_wgs.json 2> gpurun_out/c34_bench_wgs.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c34_bench_chr20.json 2> gpurun_out/c34_bench_chr20.err
python tools/bench_line.py gpurun_out/c34_bench_wgs.json gpurun_out/c34_bench_chr20.json
