#!/bin/bash
# round 2, GPU call Q: step bookkeeping (BatchNorm gradient sums from the step arena, one multi-tensor counter bump, cached conv index
# tensors): tests + A/B on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -30 ) > gpurun_out/q_tests.log 2>&1
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg "$@" > gpurun_out/q_bench_$name.json ) 2> gpurun_out/q_bench_$name.err; }
b c50_new
COTB200_BOOKKEEPING=0 b c50_old
b d101_new --model se_cotnetd_101 --batch 128
COTB200_BOOKKEEPING=0 b d101_old --model se_cotnetd_101 --batch 128
tail -8 gpurun_out/q_tests.log | cut -c1-250
python - <<'PY'
import json
for n in ("c50_new","c50_old","d101_new","d101_old"):
    try:
        d=json.loads(open("gpurun_out/q_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["roofline"]["all_kernels"]; lib=sum(v["ms_per_step"] for v in k.values())
        print(n, "img/s %.0f ms %.2f lib %.2f"%(d["value"], d["ms_per_step"], lib), d["launch_mode"])
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/q_bench_%s.err"%n).read()[-500:])
PY
