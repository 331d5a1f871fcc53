#!/bin/bash
# round 2, GPU call C: tests after calibration, CoTNeXt on the dense-masked path, per-shape conv kernels vs cuDNN, ncu of the GEMM / wgrad kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_trainer_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=60 2>&1 | tail -300 ) > gpurun_out/c_tests.log 2>&1
( timeout 400 python bench.py --steps 10 --warmup 3 --model cotnext50_2x48d --no-cpu-baseline > gpurun_out/c_bench_cotnext50.json ) 2> gpurun_out/c_bench_cotnext50.err
( timeout 600 python tools/bench_conv.py --iters 10 --json gpurun_out/c_bench_conv.json ) > gpurun_out/c_bench_conv.log 2>&1
( COTB200_TRAIN_CONV=tc_all1x1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|tc_wgrad_kernel" -s 60 -c 24 -o gpurun_out/c_prof_tc python bench.py --steps 1 --warmup 3 --graph off --no-e2e --no-cpu-baseline --no-cot-leg --batch 256 > gpurun_out/c_ncu.log 2>&1 )
tail -30 gpurun_out/c_tests.log; head -c 400 gpurun_out/c_bench_cotnext50.json; echo; tail -3 gpurun_out/c_bench_cotnext50.err; tail -20 gpurun_out/c_bench_conv.log
