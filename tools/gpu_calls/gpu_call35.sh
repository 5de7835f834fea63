#!/bin/bash
# GPU call 35: adversarial compare (two shapes) and merge tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "adversarial" > gpurun_out/c35_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c35_pytest.log
tail -25 gpurun_out/c35_pytest.log
