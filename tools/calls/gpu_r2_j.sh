#!/bin/bash
# round 2, GPU call J: unfused inference default, haloed conv variants, plane-kernel staging, two TMEM loads in flight in the GEMM
# epilogue, key convolution on the haloed tcgen05 kernel in training (+k)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_cot_layer_gpu.py tests/test_fused_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -40 ) > gpurun_out/j_tests.log 2>&1
( COTB200_AGG_PLANE=2 timeout 600 python -m pytest tests/test_agg_gpu.py tests/test_ref_kernels_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -20 ) > gpurun_out/j_tests_plane2.log 2>&1
( COTB200_TRAIN_CONV=tc_e0+k timeout 400 python -m pytest tests/test_trainer_gpu.py -m gpu -q -k "bench_path and cotnet50" 2>&1 | tail -8 ) > gpurun_out/j_tests_k.log 2>&1
( timeout 600 python tools/bench_halo.py gpurun_out/j_bench_halo.json ) > gpurun_out/j_bench_halo.log 2>&1
( COTB200_AGG_PLANE=2 timeout 400 python tools/bench_ref_kernels.py --iters 10 --json gpurun_out/j_bench_ref_kernels_plane2.json ) > gpurun_out/j_bench_ref_kernels_plane2.log 2>&1
( timeout 400 python tools/bench_ref_kernels.py --iters 10 --json gpurun_out/j_bench_ref_kernels.json ) > gpurun_out/j_bench_ref_kernels.log 2>&1
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/j_bench_$name.json ) 2> gpurun_out/j_bench_$name.err; }
b default
COTB200_TRAIN_CONV=tc_e0+k b tc_e0_k --no-cot-leg
COTB200_TRAIN_CONV=tc_all1x1 b tc_all1x1 --no-cot-leg
COTB200_TRAIN_CONV=tc_all1x1+k b tc_all1x1_k --no-cot-leg
( timeout 300 python tools/bench_block.py --json gpurun_out/j_bench_block.json ) > gpurun_out/j_bench_block.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/j_prof_cotnet50_eval.md ) > gpurun_out/j_prof.log 2>&1
( COTB200_TRAIN_CONV=tc_e0+k timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --out gpurun_out/j_prof_cotnet50_train_k.md ) > gpurun_out/j_prof_train.log 2>&1
tail -12 gpurun_out/j_tests.log | cut -c1-250
tail -5 gpurun_out/j_tests_plane2.log | cut -c1-250
tail -4 gpurun_out/j_tests_k.log | cut -c1-250
cat gpurun_out/j_bench_halo.log | cut -c1-700
python - <<'PY'
import json
for n in ("default","tc_e0_k","tc_all1x1","tc_all1x1_k"):
    try:
        d=json.loads(open("gpurun_out/j_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), "cot_forward", {k:v for k,v in (d.get("cot_forward") or {}).items() if k!='mode'})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/j_bench_%s.err"%n).read()[-600:])
for n in ("j_bench_ref_kernels_plane2","j_bench_ref_kernels"):
    try:
        for r in json.load(open("gpurun_out/%s.json"%n)):
            if r["tag"] in ("s3_b256","s4_b256"): print(n, r["tag"], {k:v for k,v in r.items() if k.startswith(("ref_fp32","ours_fp32"))})
    except Exception as e: print(n,"ERR",e)
PY
tail -5 gpurun_out/j_bench_block.log | cut -c1-330
head -16 gpurun_out/j_prof_cotnet50_eval.md | cut -c1-140
