#!/bin/bash
# GPU call 17: lane groups with converged job loops
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "dense or speculative or synthetic or golden" > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c17_bench_chr20.json 2> gpurun_out/c17_bench_chr20.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c17_bench_wgs.json 2> gpurun_out/c17_bench_wgs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c17_launches_chr20.csv python bench.py --config chr20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c17_under_ncu_chr20.log 2>&1
tail -3 gpurun_out/c17_pytest.log
python tools/bench_line.py gpurun_out/c17_bench_chr20.json gpurun_out/c17_bench_wgs.json
python tools/launch_summary.py gpurun_out/c17_launches_chr20.csv 8 2>/dev/null | head -6
