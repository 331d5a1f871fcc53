#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --graph off --no-cpu-baseline --no-e2e > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
timeout 600 python tools/bench_block.py --train --json gpurun_out/bench_block.json > gpurun_out/bench_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg3_d" -c 6 -o gpurun_out/prof_aggbwd python tools/bench_ops.py --once --stages 0 --only tap > gpurun_out/ncu_aggbwd.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu2.log | tail -2; grep -hE "^FAILED|^ERROR" gpurun_out/pytest_gpu2.log | head -20; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-2500; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_nograph.json | cut -c1-400; cat gpurun_out/bench_block.log
