#!/bin/bash
# GPU call 20: ncu of k_wfa_ed_cta, packed vs byte-staged; thread-solver pop budget
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "synthetic or dense or speculative or wgs or wfa_ed_cta" > gpurun_out/c20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wfa_ed_cta -c 1 -s 1 -o gpurun_out/c20_wfa_packed python tools/wfa_bench.py 148 > gpurun_out/c20_ncu_packed.log 2>&1
AVK_NO_PACKED_DWFA=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wfa_ed_cta -c 1 -s 1 -o gpurun_out/c20_wfa_bytes python tools/wfa_bench.py 148 > gpurun_out/c20_ncu_bytes.log 2>&1
timeout 600 python bench.py --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c20_bench_wgs025.json 2> gpurun_out/c20_bench_wgs025.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c20_bench_wgs.json 2> gpurun_out/c20_bench_wgs.err
AVK_PIPELINE_BINS=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c20_bench_wgs_bins3.json 2> gpurun_out/c20_bench_wgs_bins3.err
tail -3 gpurun_out/c20_pytest.log
python tools/bench_line.py gpurun_out/c20_bench_wgs025.json gpurun_out/c20_bench_wgs.json gpurun_out/c20_bench_wgs_bins3.json
