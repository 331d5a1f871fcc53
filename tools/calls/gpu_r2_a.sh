#!/bin/bash
# round 2, GPU call A: full GPU test suite, headline bench, configs 3-5, GPU reference arm, launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -x --deselect tests/test_trainer_gpu.py 2>&1 | tail -40 ) > gpurun_out/a_tests_old.log 2>&1
( timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_variants_gpu.py -m gpu -q --maxfail=40 2>&1 | tail -80 ) > gpurun_out/a_tests_new.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench_cotnet50.json ) 2> gpurun_out/a_bench_cotnet50.err
( timeout 400 python bench.py --steps 10 --warmup 3 --model cotnext50_2x48d --no-cpu-baseline > gpurun_out/a_bench_cotnext50.json ) 2> gpurun_out/a_bench_cotnext50.err
( timeout 400 python bench.py --steps 10 --warmup 3 --model se_cotnetd_101 --batch 128 --no-cpu-baseline > gpurun_out/a_bench_secotnetd101.json ) 2> gpurun_out/a_bench_secotnetd101.err
( timeout 400 python bench.py --steps 10 --warmup 3 --model se_cotnetd_152 --batch 64 --res 320 --no-cpu-baseline > gpurun_out/a_bench_secotnetd152.json ) 2> gpurun_out/a_bench_secotnetd152.err
( timeout 400 python tools/bench_reference_gpu.py --model cotnet50 --batch 256 --amp bf16 --steps 5 --json gpurun_out/a_refgpu_bf16.json ) > gpurun_out/a_refgpu_bf16.log 2>&1
( timeout 400 python tools/bench_reference_gpu.py --model cotnet50 --batch 256 --amp fp32 --steps 5 --json gpurun_out/a_refgpu_fp32.json ) > gpurun_out/a_refgpu_fp32.log 2>&1
( timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 10500 -c 4200 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-e2e --no-cpu-baseline --no-cot-leg > gpurun_out/a_ncu_bench.log 2>&1 )
tail -5 gpurun_out/a_tests_old.log; tail -15 gpurun_out/a_tests_new.log; cat gpurun_out/a_smoke.log | tail -3
head -c 600 gpurun_out/a_bench_cotnet50.json; echo; tail -3 gpurun_out/a_bench_cotnet50.err
