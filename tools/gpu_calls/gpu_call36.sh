#!/bin/bash
# GPU call 36: phase profile of k_search_spec (clock64 per phase)
mkdir -p gpurun_out
timeout 300 python tools/spec_profile.py > gpurun_out/c36_spec_profile.txt 2>&1
cat gpurun_out/c36_spec_profile.txt
