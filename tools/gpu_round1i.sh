#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/debug_pool.py > gpurun_out/debug_pool.log 2>&1
timeout 300 python tools/bench_ops.py --json gpurun_out/bench_ops.json --only tap > gpurun_out/bench_ops.log 2>&1
cat gpurun_out/debug_pool.log; cut -c1-330 gpurun_out/bench_ops.log
