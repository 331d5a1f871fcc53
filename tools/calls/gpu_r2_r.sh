#!/bin/bash
# round 2, GPU call R (last GPU minutes): the tests that changed since call N + smoke + one headline line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 150 python -m pytest tests/test_trainer_gpu.py tests/test_fused_gpu.py -m gpu -q -x --maxfail=5 2>&1 | tail -15 ) > gpurun_out/r_tests.log 2>&1
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r_smoke.log 2>&1
( timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cot-leg > gpurun_out/r_bench.json ) 2> gpurun_out/r_bench.err
tail -6 gpurun_out/r_tests.log | cut -c1-300; tail -2 gpurun_out/r_smoke.log; head -c 400 gpurun_out/r_bench.json; echo; tail -2 gpurun_out/r_bench.err | cut -c1-300
