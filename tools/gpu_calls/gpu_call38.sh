#!/bin/bash
# GPU call 38: final state: full suite, smoke, default bench + reference arm, chr20
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c38_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c38_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c38_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/c38_bench_wgs.json 2> gpurun_out/c38_bench_wgs.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c38_bench_ref.json 2> gpurun_out/c38_bench_ref.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c38_bench_chr20.json 2> gpurun_out/c38_bench_chr20.err
tail -3 gpurun_out/c38_pytest.log; tail -1 gpurun_out/c38_smoke.log | cut -c1-120
python tools/bench_line.py gpurun_out/c38_bench_wgs.json gpurun_out/c38_bench_chr20.json
cut -c1-200 gpurun_out/c38_bench_ref.json
