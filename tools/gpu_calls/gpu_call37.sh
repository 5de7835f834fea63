#!/bin/bash
# GPU call 37: warp-cooperative metric alignments beyond ED 6 in k_search_spec
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "dense or speculative or synthetic or golden or adversarial or sv_vs" > gpurun_out/c37_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c37_pytest.log
tail -3 gpurun_out/c37_pytest.log
timeout 300 python tools/spec_profile.py > gpurun_out/c37_spec_profile.txt 2>&1
cat gpurun_out/c37_spec_profile.txt
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c37_bench_chr20.json 2> gpurun_out/c37_bench_chr20.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c37_bench_wgs.json 2> gpurun_out/c37_bench_wgs.err
python tools/bench_line.py gpurun_out/c37_bench_chr20.json gpurun_out/c37_bench_wgs.json
