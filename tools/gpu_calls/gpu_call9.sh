#!/bin/bash
# GPU call 9: streamed single-GPU compare (bins through two lanes, copy streams), bit-iterating replay
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c9_bench_wgs.json 2> gpurun_out/c9_bench_wgs.err
echo "bench rc=$?" >> gpurun_out/c9_bench_wgs.err
AVK_TIMING=1 AVK_PIPELINE_BINS=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_wgs_bins0.json 2> gpurun_out/c9_bench_wgs_bins0.err
AVK_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_wgs_timing.json 2> gpurun_out/c9_bench_wgs_timing.err
AVK_PIPELINE_BINS=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_wgs_bins3.json 2> gpurun_out/c9_bench_wgs_bins3.err
AVK_PIPELINE_BINS=6 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_wgs_bins6.json 2> gpurun_out/c9_bench_wgs_bins6.err
AVK_PIPELINE_BINS=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_wgs_bins8.json 2> gpurun_out/c9_bench_wgs_bins8.err
tail -4 gpurun_out/c9_pytest.log
for f in wgs wgs_bins0 wgs_timing wgs_bins3 wgs_bins6 wgs_bins8; do python - "$f" <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/c9_bench_{f}.json") if l.startswith("{")][-1])
    print(f, "ms", round(d["ms_per_step"],2), "e2e_ms", round(d["e2e"]["ms_per_step"],2), d.get("cpu_baseline",{}).get("seconds_per_genome"), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("matches_gpu_bit_exact"))
except Exception as e: print(f, "failed", e)
P
done
tail -12 gpurun_out/c9_bench_wgs_timing.err
