#!/bin/bash
# GPU call 31: thread kernel advances state by state (lanes in the same coroutine state run together)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "synthetic or wgs or golden or edge or closed_form or expansion" > gpurun_out/c31_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c31_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench_wgs.json 2> gpurun_out/c31_bench_wgs.err
AVK_THREAD_MIN_REGIONS=0 timeout 600 python bench.py --scale 0.125 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench_s0125_thread.json 2> gpurun_out/c31_bench_s0125_thread.err
timeout 600 python bench.py --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench_s025.json 2> gpurun_out/c31_bench_s025.err
for bm in 8 24; do AVK_THREAD_BATCH_MIN=$bm timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench_wgs_bm$bm.json 2> gpurun_out/c31_bench_wgs_bm$bm.err; done
tail -3 gpurun_out/c31_pytest.log
python tools/bench_line.py gpurun_out/c31_bench_*.json
