#!/bin/bash
# round 2, GPU call H (re-entry): state of the whole GPU suite, smoke, headline bench, CoT-block bench, eval/train kernel profiles
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 2>&1 | tail -120 ) > gpurun_out/h_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/h_smoke.log 2>&1
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/h_bench_default.json ) 2> gpurun_out/h_bench_default.err
( timeout 300 python tools/bench_block.py --json gpurun_out/h_bench_block.json ) > gpurun_out/h_bench_block.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/h_prof_cotnet50_eval.md ) > gpurun_out/h_prof.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --out gpurun_out/h_prof_cotnet50_train.md ) > gpurun_out/h_prof_train.log 2>&1
( timeout 400 python tools/bench_ref_kernels.py --iters 10 --json gpurun_out/h_bench_ref_kernels.json ) > gpurun_out/h_bench_ref_kernels.log 2>&1
tail -25 gpurun_out/h_tests.log | cut -c1-250
tail -3 gpurun_out/h_smoke.log
python - <<'PY'
import json
for n in ("default",):
    try:
        d=json.loads(open("gpurun_out/h_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f e2e %s"%(d["value"], d["ms_per_step"], d.get("e2e")), "cot_forward", {k:v for k,v in (d.get("cot_forward") or {}).items() if k!='mode'})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/h_bench_%s.err"%n).read()[-600:])
PY
tail -8 gpurun_out/h_bench_block.log | cut -c1-300
head -16 gpurun_out/h_prof_cotnet50_eval.md | cut -c1-140
