#!/bin/bash
# GPU call 4: thread-per-cluster stage: parity suite, WGS bench, ncu of the new kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench_wgs.json 2> gpurun_out/c4_bench_wgs.err
echo "bench rc=$?" >> gpurun_out/c4_bench_wgs.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_chr20.json 2> gpurun_out/c4_bench_chr20.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c4_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_compare_thread -c 1 -s 3 -o gpurun_out/c4_thread_full python bench.py --config wgs --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c4_ncu_full.log 2>&1
tail -5 gpurun_out/c4_pytest.log; cut -c1-400 gpurun_out/c4_bench_wgs.json; tail -3 gpurun_out/c4_bench_wgs.err
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 0.05 > gpurun_out/c4_sv_timing.log 2>&1
tail -4 gpurun_out/c4_sv_timing.log
