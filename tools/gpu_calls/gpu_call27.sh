#!/bin/bash
# GPU call 27: reference window + alleles staged in shared memory in k_search_spec
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "dense or speculative or synthetic or golden or sv_vs or edge" > gpurun_out/c27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c27_pytest.log
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c27_bench_chr20.json 2> gpurun_out/c27_bench_chr20.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c27_bench_wgs.json 2> gpurun_out/c27_bench_wgs.err
timeout 600 python bench.py --scale 0.125 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c27_bench_s0125.json 2> gpurun_out/c27_bench_s0125.err
tail -3 gpurun_out/c27_pytest.log
python tools/bench_line.py gpurun_out/c27_bench_chr20.json gpurun_out/c27_bench_wgs.json gpurun_out/c27_bench_s0125.json
