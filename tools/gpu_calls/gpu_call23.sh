#!/bin/bash
# GPU call 23: full GPU suite (BED region builder included)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c23_pytest.log
tail -15 gpurun_out/c23_pytest.log
