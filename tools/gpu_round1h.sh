#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/bench_block.py --train --json gpurun_out/bench_block.json > gpurun_out/bench_block.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -hE "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -20; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench_block.log | cut -c1-600
