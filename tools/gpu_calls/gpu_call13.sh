#!/bin/bash
# GPU call 13: hard clusters (thread-stage rejects / search rejects) through speculative search + team; dense_n 6 for small batches; word-wise LCP
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_chr20.json 2> gpurun_out/c13_bench_chr20.err
AVK_DENSE_N=8 timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_chr20_d8.json 2> gpurun_out/c13_bench_chr20_d8.err
AVK_DENSE_N=4 timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_chr20_d4.json 2> gpurun_out/c13_bench_chr20_d4.err
AVK_PIPELINE_BINS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c13_bench_wgs_bins0.json 2> gpurun_out/c13_bench_wgs_bins0.err
AVK_PIPELINE_BINS=0 AVK_DENSE_N=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_wgs_d8.json 2> gpurun_out/c13_bench_wgs_d8.err
AVK_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_wgs_bins4.json 2> gpurun_out/c13_bench_wgs_bins4.err
timeout 300 python tools/seed_timings.py > gpurun_out/c13_seed_timings.txt 2>&1
AVK_PIPELINE_BINS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c13_launches_chr20.csv python bench.py --config chr20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c13_under_ncu_chr20.log 2>&1
AVK_PIPELINE_BINS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/c13_launches_wgs.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c13_under_ncu_wgs.log 2>&1
tail -4 gpurun_out/c13_pytest.log
for f in chr20 chr20_d8 chr20_d4 wgs_bins0 wgs_d8 wgs_bins4; do python - "$f" <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/c13_bench_{f}.json") if l.startswith("{")][-1])
    print(f, "ms", round(d["ms_per_step"],2), "e2e_ms", round(d["e2e"]["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d.get("cpu_baseline",{}).get("seconds_per_genome"), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("matches_gpu_bit_exact"))
except Exception as e: print(f, "failed", e)
P
done
grep "avk\] streamed" gpurun_out/c13_bench_wgs_bins4.err | tail -6
cat gpurun_out/c13_seed_timings.txt
