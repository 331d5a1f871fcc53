#!/bin/bash
# run 14: bn_apply_batch, fused SE batch_norm, stem pad off; per-op conv backend bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python tools/bench_conv.py --json gpurun_out/bench_conv.json > gpurun_out/bench_conv.log 2>&1; cat gpurun_out/bench_conv.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
