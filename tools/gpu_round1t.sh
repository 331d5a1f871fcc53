#!/bin/bash
# run 20: ReLU mask recomputed in the BatchNorm backward (y not read): tests + bench + per-shape numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-200
timeout 300 python tools/bench_norm.py --shapes 56x64,56x256,28x128 --json gpurun_out/bench_norm.json > gpurun_out/bench_norm.log 2>&1; cut -c1-700 gpurun_out/bench_norm.log
