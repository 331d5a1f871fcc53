#!/bin/bash
# 2-GPU re-validation of the flat-bucket + CUDA-graph data-parallel step with the final kernels
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
cut -c1-300 gpurun_out/bench_2gpu.json; grep -iE "error|Traceback" gpurun_out/bench_2gpu.err | head -5
