#!/bin/bash
# GPU call 5: micro-stepped thread kernel, staged CTA-wide DWFA with dynamic layout, pipelined e2e, wfa microbench, racecheck
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c5_bench_wgs.json 2> gpurun_out/c5_bench_wgs.err
echo "bench rc=$?" >> gpurun_out/c5_bench_wgs.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_chr20.json 2> gpurun_out/c5_bench_chr20.err
timeout 300 python tools/wfa_bench.py > gpurun_out/c5_wfa_bench.json 2> gpurun_out/c5_wfa_bench.err
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 0.05 > gpurun_out/c5_sv_timing.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c5_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_compare_thread -c 1 -s 3 -o gpurun_out/c5_thread_full python bench.py --config wgs --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_ncu_full.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --kernel-name kns=k_compare_team python tools/sanitize_driver.py > gpurun_out/c5_racecheck_team.log 2>&1
tail -5 gpurun_out/c5_pytest.log; cut -c1-400 gpurun_out/c5_bench_wgs.json; tail -3 gpurun_out/c5_bench_wgs.err; cat gpurun_out/c5_wfa_bench.json; tail -3 gpurun_out/c5_sv_timing.log; tail -3 gpurun_out/c5_racecheck_team.log
