#!/bin/bash
# round 2, GPU call E: fused inference LocalConv + SE kernel + GEMM epilogue round 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=60 2>&1 | tail -150 ) > gpurun_out/e_tests.log 2>&1
( timeout 600 python tools/bench_conv.py --iters 10 --json gpurun_out/e_bench_conv.json ) > gpurun_out/e_bench_conv.log 2>&1
( timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_default.json ) 2> gpurun_out/e_bench_default.err
( COTB200_EVAL_FUSED_AGG=0 timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/e_bench_nofuse.json ) 2> gpurun_out/e_bench_nofuse.err
( COTB200_TRAIN_CONV=tc_all1x1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg > gpurun_out/e_bench_tc_all1x1.json ) 2> gpurun_out/e_bench_tc_all1x1.err
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/e_prof_cotnet50_eval.md ) > gpurun_out/e_prof.log 2>&1
( timeout 300 python tools/bench_block.py --json gpurun_out/e_bench_block.json ) > gpurun_out/e_bench_block.log 2>&1
tail -15 gpurun_out/e_tests.log | cut -c1-250
python - <<'PY'
import json
for n in ("default","nofuse","tc_all1x1"):
    try:
        d=json.loads(open("gpurun_out/e_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), "cot_forward", d.get("cot_forward"))
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/e_bench_%s.err"%n).read()[-600:])
PY
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/e_bench_conv.json")):
        w=r["raw"]; print(r["name"], "fwd_stats %.0f dgrad %.0f wgrad %.0f | cudnn fprop %.0f wgrad %.0f | roof %.0f"%(w["tc_fwd_stats_us"],w["tc_dgrad_us"],w["tc_wgrad_us"],w["cudnn_fprop_us"],w["cudnn_wgrad_us"],w["roof_us_at_6485GBps"]))
except Exception as e: print("conv ERR", e)
PY
tail -12 gpurun_out/e_bench_block.log | cut -c1-300
