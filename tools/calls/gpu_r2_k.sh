#!/bin/bash
# round 2, GPU call K: warp-uniform MMA issue (elect.sync) in the tcgen05 kernels, stem weight gradient on the MN-major wgrad kernel,
# key convolution on the haloed kernel for dim <= 128 (+k)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_cot_layer_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -40 ) > gpurun_out/k_tests.log 2>&1
( timeout 600 python -m pytest tests/test_trainer_gpu.py -m gpu -q --maxfail=10 2>&1 | tail -30 ) > gpurun_out/k_tests_trainer.log 2>&1
( timeout 600 python tools/bench_halo.py gpurun_out/k_bench_halo.json ) > gpurun_out/k_bench_halo.log 2>&1
( timeout 600 python tools/bench_conv.py --iters 10 --json gpurun_out/k_bench_conv.json ) > gpurun_out/k_bench_conv.log 2>&1
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/k_bench_$name.json ) 2> gpurun_out/k_bench_$name.err; }
b default
COTB200_STEM_WGRAD_TC=0 b stem_wgrad_cudnn --no-cot-leg
COTB200_TRAIN_CONV=tc_e0+k b tc_e0_k --no-cot-leg
COTB200_TRAIN_CONV=tc_all1x1 b tc_all1x1 --no-cot-leg
COTB200_TRAIN_CONV=tc_all1x1+k b tc_all1x1_k --no-cot-leg
( timeout 300 python tools/bench_block.py --json gpurun_out/k_bench_block.json ) > gpurun_out/k_bench_block.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/k_prof_cotnet50_eval.md ) > gpurun_out/k_prof.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --out gpurun_out/k_prof_cotnet50_train.md ) > gpurun_out/k_prof_train.log 2>&1
tail -12 gpurun_out/k_tests.log | cut -c1-250
tail -8 gpurun_out/k_tests_trainer.log | cut -c1-250
cat gpurun_out/k_bench_halo.log | cut -c1-700
python - <<'PY'
import json
for n in ("default","stem_wgrad_cudnn","tc_e0_k","tc_all1x1","tc_all1x1_k"):
    try:
        d=json.loads(open("gpurun_out/k_bench_%s.json"%n).read().strip().splitlines()[-1])
        lib=sum(v["ms_per_step"] for v in d["roofline"]["all_kernels"].values())
        print(n, "img/s %.0f ms %.2f lib-kernels %.2f ms"%(d["value"], d["ms_per_step"], lib), "cot_forward", {k:v for k,v in (d.get("cot_forward") or {}).items() if k!='mode'})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/k_bench_%s.err"%n).read()[-600:])
try:
    for r in json.load(open("gpurun_out/k_bench_conv.json")):
        w=r["raw"]; print(r["name"], "fwd_stats %.0f dgrad %.0f wgrad %.0f | cudnn fprop %.0f wgrad %.0f | roof %.0f | fwdbwd tc %.0f cudnn %.0f"%(w["tc_fwd_stats_us"],w["tc_dgrad_us"],w["tc_wgrad_us"],w["cudnn_fprop_us"],w["cudnn_wgrad_us"],w["roof_us_at_6485GBps"], r["tc_fwdbwd_us"], r["cudnn_fwdbwd_us"]))
except Exception as e: print("conv ERR", e)
PY
tail -5 gpurun_out/k_bench_block.log | cut -c1-330
head -14 gpurun_out/k_prof_cotnet50_eval.md | cut -c1-140
