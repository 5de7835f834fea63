#!/bin/bash
# GPU call 8 (2 GPUs): parity suite, N=1 and N=2 bench lines (torchrun, NCCL gather), reference arm at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c8_bench_wgs_n1.json 2> gpurun_out/c8_bench_wgs_n1.err
echo "bench rc=$?" >> gpurun_out/c8_bench_wgs_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c8_bench_wgs_n2.json 2> gpurun_out/c8_bench_wgs_n2.err
echo "bench2 rc=$?" >> gpurun_out/c8_bench_wgs_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c8_bench_ref_n2.json 2> gpurun_out/c8_bench_ref_n2.err
tail -4 gpurun_out/c8_pytest.log; cut -c1-260 gpurun_out/c8_bench_wgs_n1.json; cut -c1-260 gpurun_out/c8_bench_wgs_n2.json; tail -5 gpurun_out/c8_bench_wgs_n2.err; cut -c1-400 gpurun_out/c8_bench_ref_n2.json
