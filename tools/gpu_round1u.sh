#!/bin/bash
# run 21: cuDNN autotune limit A/B
mkdir -p gpurun_out
COTB200_CUDNN_BENCH_LIMIT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cudnn_all.json 2> gpurun_out/bench_cudnn_all.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cudnn_all.json")); print("cudnn limit 0:", round(d["value"],1), round(d["ms_per_step"],2))
PY
