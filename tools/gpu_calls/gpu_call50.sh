#!/bin/bash
# GPU call 50: BGZF inflate on the device: parity tests + timing of a whole-genome-sized VCF
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "bgzf" > gpurun_out/c50_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c50_pytest.log
tail -15 gpurun_out/c50_pytest.log
timeout 600 python tools/bgzf_bench.py > gpurun_out/c50_bgzf_bench.json 2> gpurun_out/c50_bgzf_bench.err
cat gpurun_out/c50_bgzf_bench.json; tail -3 gpurun_out/c50_bgzf_bench.err
