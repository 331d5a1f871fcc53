#!/bin/bash
# round 2, GPU call L: new default training backend (tc_all1x1+k): full GPU suite, smoke, headline bench (full line), configs 3-5,
# trunk weight-cap A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | tail -60 ) > gpurun_out/l_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/l_smoke.log 2>&1
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/l_bench_cotnet50.json ) 2> gpurun_out/l_bench_cotnet50.err
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/l_bench_$name.json ) 2> gpurun_out/l_bench_$name.err; }
COTB200_TC_TRUNK_MAX_WEIGHT=262144 b trunk256k --no-e2e --no-cot-leg
COTB200_TC_TRUNK_MAX_WEIGHT=2097152 b trunk2m --no-e2e --no-cot-leg
COTB200_TC_KEY_MAX_DIM=256 b key256 --no-e2e --no-cot-leg
b cotnext50 --model cotnext50_2x48d
COTB200_TRAIN_CONV=tc_e0 b cotnext50_tc_e0 --model cotnext50_2x48d --no-e2e --no-cot-leg
b secotnetd101 --model se_cotnetd_101 --batch 128
COTB200_TRAIN_CONV=tc_e0 b secotnetd101_tc_e0 --model se_cotnetd_101 --batch 128 --no-e2e --no-cot-leg
b secotnetd152 --model se_cotnetd_152 --batch 64 --res 320
tail -25 gpurun_out/l_tests.log | cut -c1-250
tail -3 gpurun_out/l_smoke.log
python - <<'PY'
import json
for n in ("cotnet50","trunk256k","trunk2m","key256","cotnext50","cotnext50_tc_e0","secotnetd101","secotnetd101_tc_e0","secotnetd152"):
    try:
        d=json.loads(open("gpurun_out/l_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), "e2e", (d.get("e2e") or {}).get("value"), "cot_forward", {k:v for k,v in (d.get("cot_forward") or {}).items() if k in ("cot_layers_ms","frac","model_forward_ms")})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/l_bench_%s.err"%n).read()[-600:])
PY
