#!/bin/bash
# 2-GPU round (gpurun --gpus 2): per-shard parity / broadcast integrity tests, and the driver's multi-GPU bench invocation
TAG=${1:-r2z}
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/${TAG}_smi_2gpu.txt
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q --no-header -rf -p no:cacheprovider > $O/${TAG}_pytest_2gpu.log 2>&1; echo "2-GPU tests rc=$?"; tail -5 $O/${TAG}_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
echo "bench n=2 rc=$?"; tail -c 1200 $O/${TAG}_bench_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload vqgan --steps 3 --warmup 3 > $O/${TAG}_vqgan_n2.json 2> $O/${TAG}_vqgan_n2.err
echo "vqgan n=2 rc=$?"; tail -c 600 $O/${TAG}_vqgan_n2.json; echo
