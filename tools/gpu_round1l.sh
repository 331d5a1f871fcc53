#!/bin/bash
# run 12: gn72 kernels, tanh sigmoid, stem padding, col_stats FHFMA: tests + bench + launch list + ncu of gn72/tail kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
timeout 400 python tools/bench_norm.py --shapes 112x64,56x64,56x256,28x512 --json gpurun_out/bench_norm.json > gpurun_out/bench_norm.log 2>&1; cut -c1-200 gpurun_out/bench_norm.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --graph off --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gn72|tail_|col_stats" -s 40 -c 14 -o gpurun_out/prof_gn_tail -f python tools/bench_block.py --train --iters 1 --batch 256 > gpurun_out/ncu_gn_tail.log 2>&1; tail -2 gpurun_out/ncu_gn_tail.log | cut -c1-200
