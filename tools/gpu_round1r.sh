#!/bin/bash
# run 18: in-graph A/B of the training convolution backends of the CoT layer (whole-model bench, CUDA graph)
mkdir -p gpurun_out
for be in cudnn tc_e0 tc_1x1 tc; do
  COTB200_TRAIN_CONV=$be timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_conv_$be.json 2> gpurun_out/bench_conv_$be.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_conv_$be.json")); print("$be", round(d["value"],1), round(d["ms_per_step"],2), d["launch_mode"])
except Exception as e:
    print("$be", "FAILED", e); print(open("gpurun_out/bench_conv_$be.err").read()[-1500:])
PY
done
