#!/bin/bash
# GPU call 24: round-2 evidence on one GPU: suite, bench (WGS default + reference arm + chr20), launch list, DRAM traffic, ncu of the two top kernels, DWFA microbench, SV sample
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c24_pytest.log
timeout 900 python bench.py > gpurun_out/c24_bench_wgs.json 2> gpurun_out/c24_bench_wgs.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c24_bench_ref.json 2> gpurun_out/c24_bench_ref.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c24_bench_chr20.json 2> gpurun_out/c24_bench_chr20.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c24_launches_wgs.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c24_under_ncu_wgs.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/c24_traffic_wgs.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c24_under_ncu_traffic.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_compare_thread -c 1 -s 3 -o gpurun_out/c24_thread_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c24_ncu_thread.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_search_spec -c 1 -s 6 -o gpurun_out/c24_spec_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c24_ncu_spec.log 2>&1
timeout 300 python tools/wfa_bench.py > gpurun_out/c24_wfa_bench.json 2> gpurun_out/c24_wfa_bench.err
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 0.05 > gpurun_out/c24_sv_timing.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=k_search_spec python tools/sanitize_driver.py > gpurun_out/c24_racecheck_spec.log 2>&1
tail -3 gpurun_out/c24_pytest.log
python tools/bench_line.py gpurun_out/c24_bench_wgs.json gpurun_out/c24_bench_chr20.json
cut -c1-300 gpurun_out/c24_bench_ref.json
tail -3 gpurun_out/c24_racecheck_spec.log
