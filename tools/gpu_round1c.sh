#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python tools/bench_ops.py --json gpurun_out/bench_ops.json > gpurun_out/bench_ops.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg3|gn_|tail_|bn_|col_stats" -c 40 -o gpurun_out/prof_v2 python tools/bench_ops.py --once --stages 0 > gpurun_out/ncu_v2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3; grep FAILED gpurun_out/pytest_gpu.log | head -30; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ops.log
