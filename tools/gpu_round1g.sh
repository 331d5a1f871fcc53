#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py tests/test_agg_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python tools/bench_ops.py --json gpurun_out/bench_ops.json --only tap > gpurun_out/bench_ops.log 2>&1
timeout 600 python tools/bench_block.py --train --json gpurun_out/bench_block.json > gpurun_out/bench_block.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ops.csv python tools/bench_ops.py --once --only tap > gpurun_out/ncu_ops.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 4500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|agg3_dw_tma|agg3_fwd_tma" -c 14 -o gpurun_out/prof_tc python tools/bench_block.py --iters 1 --train > gpurun_out/ncu_tc.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu2.log | tail -2; grep -hE "^FAILED|^ERROR" gpurun_out/pytest_gpu2.log | head -20; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ops.log | cut -c1-330; cat gpurun_out/bench_block.log
