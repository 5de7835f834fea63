#!/bin/bash
# GPU call 49: final single-GPU evidence: suite, smoke, bench (WGS default + reference arm + chr20), ncu of the CTA DWFA, full SV config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c49_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c49_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c49_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/c49_bench_wgs.json 2> gpurun_out/c49_bench_wgs.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c49_bench_ref.json 2> gpurun_out/c49_bench_ref.err
timeout 300 python bench.py --config chr20 --steps 10 --warmup 3 > gpurun_out/c49_bench_chr20.json 2> gpurun_out/c49_bench_chr20.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wfa_ed_cta -c 1 -s 1 -o gpurun_out/c49_wfa_bytes python tools/wfa_bench.py 148 > gpurun_out/c49_ncu_bytes.log 2>&1
timeout 300 python tools/wfa_bench.py > gpurun_out/c49_wfa_bench.json 2> gpurun_out/c49_wfa_bench.err
AVK_DEBUG=1 timeout 900 python tools/sv_timing.py 1.0 > gpurun_out/c49_sv_full.log 2>&1
tail -3 gpurun_out/c49_pytest.log
tail -1 gpurun_out/c49_smoke.log
python tools/bench_line.py gpurun_out/c49_bench_wgs.json gpurun_out/c49_bench_chr20.json
cut -c1-300 gpurun_out/c49_bench_ref.json
tail -4 gpurun_out/c49_sv_full.log | cut -c1-300
