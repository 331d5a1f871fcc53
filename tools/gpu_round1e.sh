#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py > gpurun_out/debug_tc.log 2>&1
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 900 python -m pytest tests/test_agg_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_ops.py --json gpurun_out/bench_ops.json --only tap > gpurun_out/bench_ops.log 2>&1
timeout 600 python tools/bench_block.py --train --json gpurun_out/bench_block.json > gpurun_out/bench_block.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg3_.*tma|tc_gemm" -c 16 -o gpurun_out/prof_tma python tools/bench_block.py --iters 1 --train > gpurun_out/ncu_tma.log 2>&1
cat gpurun_out/debug_tc.log; grep -E "passed|failed" gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu2.log | tail -4; grep -hE "^FAILED|^ERROR" gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu2.log | head -30; cat gpurun_out/bench.json | cut -c1-1500; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ops.log | cut -c1-330; cat gpurun_out/bench_block.log
