#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -x --deselect tests/test_agg_gpu.py::test_full_size_properties > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_agg_gpu.py::test_full_size_properties tests/test_fused_gpu.py tests/test_cot_layer_gpu.py tests/test_tc_gemm_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python tools/bench_ops.py --json gpurun_out/bench_ops.json --only tap > gpurun_out/bench_ops.log 2>&1
timeout 600 python tools/bench_block.py --train --json gpurun_out/bench_block.json > gpurun_out/bench_block.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg3_fwd_tma|tc_gemm" -c 12 -o gpurun_out/prof_tma python tools/bench_block.py --iters 1 > gpurun_out/ncu_tma.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu2.log | tail -4; grep -hE "^FAILED|^ERROR" gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu2.log | head -30; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ops.log; cat gpurun_out/bench_block.log
