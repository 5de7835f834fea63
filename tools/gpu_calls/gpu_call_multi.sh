#!/bin/bash
# multi-GPU evidence: bash tools/gpu_calls/gpu_call_multi.sh N   (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_gpus.txt
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "multi_context or wgs_24" > gpurun_out/m${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m${N}_pytest.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench_wgs.json 2> gpurun_out/m${N}_bench_wgs.err
echo "bench rc=$?" >> gpurun_out/m${N}_bench_wgs.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/m${N}_bench_ref.json 2> gpurun_out/m${N}_bench_ref.err
tail -2 gpurun_out/m${N}_pytest.log 2>/dev/null
python tools/bench_line.py gpurun_out/m${N}_bench_wgs.json
cut -c1-260 gpurun_out/m${N}_bench_ref.json
tail -3 gpurun_out/m${N}_bench_wgs.err
