#!/bin/bash
# GPU call 14: clusters finished inside k_search_spec (lane-parallel metrics), host scan beside the uploads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_chr20.json 2> gpurun_out/c14_bench_chr20.err
AVK_DENSE_N=8 timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_chr20_d8.json 2> gpurun_out/c14_bench_chr20_d8.err
AVK_DENSE_N=4 timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_chr20_d4.json 2> gpurun_out/c14_bench_chr20_d4.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c14_bench_wgs.json 2> gpurun_out/c14_bench_wgs.err
timeout 300 python tools/seed_timings.py > gpurun_out/c14_seed_timings.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c14_launches_chr20.csv python bench.py --config chr20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c14_under_ncu_chr20.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/c14_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c14_under_ncu_wgs.log 2>&1
tail -4 gpurun_out/c14_pytest.log
python tools/bench_line.py gpurun_out/c14_bench_chr20.json gpurun_out/c14_bench_chr20_d8.json gpurun_out/c14_bench_chr20_d4.json gpurun_out/c14_bench_wgs.json
cat gpurun_out/c14_seed_timings.txt
python tools/launch_summary.py gpurun_out/c14_launches_chr20.csv 8 | head -8
python tools/launch_summary.py gpurun_out/c14_launches_wgs.csv 6 | head -8
