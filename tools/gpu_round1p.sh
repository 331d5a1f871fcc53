#!/bin/bash
# run 16: NCHW TMA kernels (+ split/fused register kernels for comparison), zero arena, new model-level tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/bench_ref_kernels.py --json gpurun_out/bench_ref_kernels.json > gpurun_out/bench_ref_kernels.log 2>&1; cut -c1-600 gpurun_out/bench_ref_kernels.log
COTB200_AGG_NCHW_TMA=0 timeout 300 python tools/bench_ref_kernels.py --json gpurun_out/bench_ref_kernels_notma_split.json > gpurun_out/bench_ref_kernels_notma_split.log 2>&1; cut -c1-330 gpurun_out/bench_ref_kernels_notma_split.log
COTB200_AGG_NCHW_TMA=0 COTB200_NCHW_SPLIT=0 timeout 300 python tools/bench_ref_kernels.py --json gpurun_out/bench_ref_kernels_notma_fused.json > gpurun_out/bench_ref_kernels_notma_fused.log 2>&1; cut -c1-330 gpurun_out/bench_ref_kernels_notma_fused.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
