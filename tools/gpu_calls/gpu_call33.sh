#!/bin/bash
# GPU call 33: where does the end-to-end call spend its host time?  + suite after the blob-table change
mkdir -p gpurun_out
AVK_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c33_bench_timing.json 2> gpurun_out/c33_bench_timing.err
grep "avk\]" gpurun_out/c33_bench_timing.err | tail -8
python - <<'P'
import torch, time
n = 353_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw pinned H2D of 353 MB: %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw pinned D2H of 353 MB: %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
P
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c33_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c33_pytest.log
tail -3 gpurun_out/c33_pytest.log
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 0.25 > gpurun_out/c33_sv_quarter.log 2>&1
tail -3 gpurun_out/c33_sv_quarter.log
