#!/bin/bash
# round 2, GPU call D: GEMM epilogue rework + eval-path kernels: tests, per-shape conv bench, backend A/B, CoTNeXt
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_trainer_gpu.py -m gpu -q --maxfail=60 2>&1 | tail -200 ) > gpurun_out/d_tests.log 2>&1
( timeout 900 python -m pytest tests/test_cot_layer_gpu.py tests/test_variants_gpu.py -m gpu -q --maxfail=60 2>&1 | tail -100 ) > gpurun_out/d_tests2.log 2>&1
( timeout 600 python tools/bench_conv.py --iters 10 --json gpurun_out/d_bench_conv.json ) > gpurun_out/d_bench_conv.log 2>&1
for be in tc_e0 tc_all1x1; do
  ( COTB200_TRAIN_CONV=$be timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_bench_$be.json ) 2> gpurun_out/d_bench_$be.err
done
( COTB200_TRAIN_CONV=tc_all1x1 COTB200_TC_TRUNK_MAX_WEIGHT=4194304 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg > gpurun_out/d_bench_tc_all1x1_nolimit.json ) 2> gpurun_out/d_bench_tc_all1x1_nolimit.err
( timeout 400 python bench.py --steps 10 --warmup 3 --model cotnext50_2x48d --no-cpu-baseline > gpurun_out/d_bench_cotnext50.json ) 2> gpurun_out/d_bench_cotnext50.err
( timeout 300 python tools/profile_step.py --model cotnext50_2x48d --batch 256 --out gpurun_out/d_prof_cotnext50_train.md ) > gpurun_out/d_prof_cotnext.log 2>&1
tail -12 gpurun_out/d_tests.log; tail -8 gpurun_out/d_tests2.log; tail -17 gpurun_out/d_bench_conv.log | cut -c1-330
for be in tc_e0 tc_all1x1 tc_all1x1_nolimit cotnext50; do head -c 250 gpurun_out/d_bench_$be.json; echo; tail -2 gpurun_out/d_bench_$be.err; done
