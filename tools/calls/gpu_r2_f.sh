#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=60 2>&1 | tail -60 ) > gpurun_out/f_tests.log 2>&1
( timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_bench_default.json ) 2> gpurun_out/f_bench_default.err
( COTB200_EVAL_FUSED_AGG=0 timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_bench_nofuse.json ) 2> gpurun_out/f_bench_nofuse.err
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/f_prof_cotnet50_eval.md ) > gpurun_out/f_prof.log 2>&1
( timeout 300 python tools/bench_block.py --json gpurun_out/f_bench_block.json ) > gpurun_out/f_bench_block.log 2>&1
( timeout 400 python tools/bench_ref_kernels.py --iters 10 --json gpurun_out/f_bench_ref_kernels.json ) > gpurun_out/f_bench_ref_kernels.log 2>&1
tail -6 gpurun_out/f_tests.log | cut -c1-250
python - <<'PY'
import json
for n in ("default","nofuse"):
    try:
        d=json.loads(open("gpurun_out/f_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), "cot_forward", {k:v for k,v in d.get("cot_forward",{}).items() if k!='mode'})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/f_bench_%s.err"%n).read()[-600:])
PY
head -14 gpurun_out/f_prof_cotnet50_eval.md | cut -c1-140
tail -8 gpurun_out/f_bench_ref_kernels.log | cut -c1-400
