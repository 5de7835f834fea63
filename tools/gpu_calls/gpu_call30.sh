#!/bin/bash
# GPU call 30: final single-GPU evidence (batch threshold 16): suite, WGS bench (default flags) + reference arm, chr20, launch list, ncu of the thread kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c30_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c30_pytest.log
timeout 900 python bench.py > gpurun_out/c30_bench_wgs.json 2> gpurun_out/c30_bench_wgs.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c30_bench_ref.json 2> gpurun_out/c30_bench_ref.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c30_bench_chr20.json 2> gpurun_out/c30_bench_chr20.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c30_launches_wgs.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c30_under_ncu_wgs.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_compare_thread -c 1 -s 3 -o gpurun_out/c30_thread_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c30_ncu_thread.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c30_smoke.log 2>&1
tail -3 gpurun_out/c30_pytest.log; tail -1 gpurun_out/c30_smoke.log
python tools/bench_line.py gpurun_out/c30_bench_wgs.json gpurun_out/c30_bench_chr20.json
cut -c1-200 gpurun_out/c30_bench_ref.json
