#!/bin/bash
# 2-GPU data-parallel validation: DDP over NCCL, one process per GPU
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
cat gpurun_out/bench_2gpu.json | cut -c1-1200; tail -5 gpurun_out/bench_2gpu.err | cut -c1-300; cat gpurun_out/bench_2gpu_ref.json | cut -c1-600; tail -3 gpurun_out/bench_2gpu_ref.err | cut -c1-200
