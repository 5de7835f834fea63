#!/bin/bash
# GPU call 51: full GPU suite + smoke on the final library, BGZF timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c51_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c51_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c51_smoke.log 2>&1
timeout 600 python tools/bgzf_bench.py > gpurun_out/c51_bgzf_bench.json 2> gpurun_out/c51_bgzf_bench.err
tail -4 gpurun_out/c51_pytest.log
tail -1 gpurun_out/c51_smoke.log
cat gpurun_out/c51_bgzf_bench.json; tail -3 gpurun_out/c51_bgzf_bench.err
