#!/bin/bash
# run 15 (2 GPUs): flat-bucket data parallel with CUDA graphs vs torch DDP eager; reference arm under torchrun
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
grep -v Warning gpurun_out/bench_2gpu.json | cut -c1-400; grep -iE "error|Traceback" gpurun_out/bench_2gpu.err | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 3 --dp ddp --no-e2e > gpurun_out/bench_2gpu_ddp.json 2> gpurun_out/bench_2gpu_ddp.err
cut -c1-300 gpurun_out/bench_2gpu_ddp.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
cut -c1-700 gpurun_out/bench_2gpu_ref.json
