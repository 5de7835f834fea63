#!/bin/bash
# GPU call 11: speculative dense search (k_search_spec) + streamed compare
mkdir -p gpurun_out
strings aardvark_b200/csrc/libaardvark_b200.so | grep -c "streamed step" > gpurun_out/c11_sanity.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_chr20.json 2> gpurun_out/c11_bench_chr20.err
AVK_NO_SPEC_SEARCH=1 timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_chr20_nospec.json 2> gpurun_out/c11_bench_chr20_nospec.err
AVK_PIPELINE_BINS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c11_bench_wgs_bins0.json 2> gpurun_out/c11_bench_wgs_bins0.err
AVK_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_wgs_bins4.json 2> gpurun_out/c11_bench_wgs_bins4.err
AVK_PIPELINE_BINS=8 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_wgs_bins8.json 2> gpurun_out/c11_bench_wgs_bins8.err
timeout 300 python tools/seed_timings.py > gpurun_out/c11_seed_timings.txt 2>&1
tail -4 gpurun_out/c11_pytest.log
for f in chr20 chr20_nospec wgs_bins0 wgs_bins4 wgs_bins8; do python - "$f" <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/c11_bench_{f}.json") if l.startswith("{")][-1])
    print(f, "ms", round(d["ms_per_step"],2), "e2e_ms", round(d["e2e"]["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d.get("cpu_baseline",{}).get("seconds_per_genome"), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("matches_gpu_bit_exact"))
except Exception as e: print(f, "failed", e)
P
done
grep "avk\]" gpurun_out/c11_bench_wgs_bins4.err | tail -8
cat gpurun_out/c11_seed_timings.txt
