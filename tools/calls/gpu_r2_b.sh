#!/bin/bash
# round 2, GPU call B: full test suite (no -x, full logs), wgrad bring-up, conv-backend A/B inside the bench step, CoTNeXt profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -s --maxfail=50 2>&1 | grep -v "^$" | tail -150 ) > gpurun_out/b_tests_tc.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 --deselect tests/test_tc_gemm_gpu.py 2>&1 | tail -400 ) > gpurun_out/b_tests_all.log 2>&1
for be in tc_e0 tc_all1x1 cudnn; do
  ( COTB200_TRAIN_CONV=$be timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg > gpurun_out/b_bench_$be.json ) 2> gpurun_out/b_bench_$be.err
done
( timeout 300 python tools/profile_step.py --model cotnext50_2x48d --batch 64 --out gpurun_out/b_prof_cotnext50_train.md ) > gpurun_out/b_prof_cotnext.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnext50_2x48d --batch 64 --eval --out gpurun_out/b_prof_cotnext50_eval.md ) >> gpurun_out/b_prof_cotnext.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --out gpurun_out/b_prof_cotnet50_train.md ) > gpurun_out/b_prof_cotnet.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/b_prof_cotnet50_eval.md ) >> gpurun_out/b_prof_cotnet.log 2>&1
tail -30 gpurun_out/b_tests_tc.log; tail -25 gpurun_out/b_tests_all.log
for be in tc_e0 tc_all1x1 cudnn; do head -c 300 gpurun_out/b_bench_$be.json; echo; tail -2 gpurun_out/b_bench_$be.err; done
