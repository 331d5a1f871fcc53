#!/bin/bash
# final validation of round 1: full GPU suite, smoke, both bench arms, launch list + DRAM traffic of the BatchNorm kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-200
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-500 gpurun_out/bench_ref.json
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"bn_bwd_apply|bn_bwd_sums|bn_apply|col_stats" -c 712 --csv --log-file gpurun_out/traffic_bench.csv python bench.py --graph off --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 13500 -c 4300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --graph off --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
