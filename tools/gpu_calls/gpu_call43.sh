#!/bin/bash
# GPU call 43: lanes (avk_create_lane) parity tests + bench with passes in flight
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "lanes or pool or pipelined or multi" > gpurun_out/c43_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c43_pytest.log
tail -5 gpurun_out/c43_pytest.log
timeout 900 python bench.py > gpurun_out/c43_bench_wgs.json 2> gpurun_out/c43_bench_wgs.err
tail -5 gpurun_out/c43_bench_wgs.err
python tools/bench_line.py gpurun_out/c43_bench_wgs.json
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c43_bench_chr20.json 2> gpurun_out/c43_bench_chr20.err
python tools/bench_line.py gpurun_out/c43_bench_chr20.json
