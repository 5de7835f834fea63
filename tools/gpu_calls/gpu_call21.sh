#!/bin/bash
# GPU call 21: balanced packed DWFA; thread-stage pop budget sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "wfa or sv or last_resort" > gpurun_out/c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c21_pytest.log
timeout 300 python tools/wfa_bench.py > gpurun_out/c21_wfa_bench.json 2> gpurun_out/c21_wfa_bench.err
AVK_PACKED_DWFA=1 timeout 300 python tools/wfa_bench.py > gpurun_out/c21_wfa_bench_bytes.json 2> gpurun_out/c21_wfa_bench_bytes.err
for b in 8 16 32; do
AVK_THREAD_POP_BUDGET=$b timeout 600 python bench.py --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c21_bench_wgs025_b$b.json 2> gpurun_out/c21_bench_wgs025_b$b.err
AVK_THREAD_POP_BUDGET=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c21_bench_wgs_b$b.json 2> gpurun_out/c21_bench_wgs_b$b.err
done
tail -3 gpurun_out/c21_pytest.log
cut -c1-420 gpurun_out/c21_wfa_bench.json; cut -c1-420 gpurun_out/c21_wfa_bench_bytes.json
python tools/bench_line.py gpurun_out/c21_bench_wgs025_b*.json gpurun_out/c21_bench_wgs_b*.json
