#!/bin/bash
# run 11: full GPU suite + smoke + per-shape norm bench + reference-kernel bench + model bench + traffic capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os;print('cpus',os.cpu_count(),len(os.sched_getaffinity(0)));print(open('/sys/fs/cgroup/cpu.max').read())" >> gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/bench_ref_kernels.py --json gpurun_out/bench_ref_kernels.json > gpurun_out/bench_ref_kernels.log 2>&1; tail -4 gpurun_out/bench_ref_kernels.log | cut -c1-700
timeout 400 python tools/bench_norm.py --json gpurun_out/bench_norm.json > gpurun_out/bench_norm.log 2>&1; tail -16 gpurun_out/bench_norm.log | cut -c1-420
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"bn_bwd_apply|bn_bwd_sums|bn_apply|col_stats" -c 712 --csv --log-file gpurun_out/traffic_bench.csv python bench.py --graph off --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; tail -2 gpurun_out/ncu_traffic.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply -s 6 -c 1 -o gpurun_out/prof_bn -f python tools/bench_norm.py --iters 2 --shapes 56x256 > gpurun_out/ncu_bn.log 2>&1; tail -2 gpurun_out/ncu_bn.log | cut -c1-300
