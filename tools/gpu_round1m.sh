#!/bin/bash
# run 13: persistent gn72_stats, atomic-free finish, unrolled tails; stem micro-bench; clean launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python tools/bench_stem.py > gpurun_out/bench_stem.json 2> gpurun_out/bench_stem.err; cat gpurun_out/bench_stem.json; tail -2 gpurun_out/bench_stem.err | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
COTB200_STEM_PAD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_nopad.json 2> gpurun_out/bench_nopad.err; cut -c1-300 gpurun_out/bench_nopad.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 -c 4700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --graph off --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
