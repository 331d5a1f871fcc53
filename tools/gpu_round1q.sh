#!/bin/bash
# run 17: full GPU suite after the 16-bit NCHW TMA alignment fix + op bench for bf16 NCHW
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/bench_ops.py --only nchw --json gpurun_out/bench_ops_nchw.json > gpurun_out/bench_ops_nchw.log 2>&1; cut -c1-330 gpurun_out/bench_ops_nchw.log
