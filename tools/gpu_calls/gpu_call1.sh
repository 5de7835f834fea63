#!/bin/bash
# GPU call 1 of round 2: full GPU test-suite, WGS bench line, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_wgs.json 2> gpurun_out/c1_bench_wgs.err
echo "bench rc=$?" >> gpurun_out/c1_bench_wgs.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c1_bench_chr20.json 2> gpurun_out/c1_bench_chr20.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c1_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_under_ncu.log 2>&1
tail -5 gpurun_out/c1_pytest.log; cut -c1-600 gpurun_out/c1_bench_wgs.json; tail -3 gpurun_out/c1_bench_wgs.err
