#!/bin/bash
# round 2, GPU call I: stem on the tcgen05 implicit GEMM, tiled SE kernel; ncu --set full of the eval CoT layer and the training GEMMs;
# A/B of the training convolution backends and of the fused / unfused inference aggregation; CoTNeXt gradient-parity A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_tc_gemm_gpu.py tests/test_fused_gpu.py tests/test_cot_layer_gpu.py tests/test_agg_gpu.py tests/test_ref_kernels_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -60 ) > gpurun_out/i_tests.log 2>&1
( COTB200_CONV_HALO=2 timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -k conv3x3 2>&1 | tail -15 ) > gpurun_out/i_tests_halo_baseoff.log 2>&1
( COTB200_CONV_HALO=0 timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -k conv3x3 2>&1 | tail -5 ) > gpurun_out/i_tests_halo_off.log 2>&1
( timeout 400 python tools/bench_ref_kernels.py --iters 10 --json gpurun_out/i_bench_ref_kernels.json ) > gpurun_out/i_bench_ref_kernels.log 2>&1
( timeout 400 python -m pytest tests/test_trainer_gpu.py -m gpu -q -k "bench_path or plain_pytorch_loop" 2>&1 | tail -30 ) > gpurun_out/i_tests_trainer.log 2>&1
cp gpurun_out/parity_measured.json gpurun_out/i_parity_default.json 2>/dev/null
( COTB200_TRAIN_CONV=cudnn timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -k "bench_path and cotnext" 2>&1 | tail -8 ) > gpurun_out/i_tests_cotnext_cudnn.log 2>&1
cp gpurun_out/parity_measured.json gpurun_out/i_parity_cudnn.json 2>/dev/null
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/i_bench_$name.json ) 2> gpurun_out/i_bench_$name.err; }
b default
COTB200_STEM_TC=0 b stem_cudnn --no-cot-leg
COTB200_TRAIN_CONV=tc_all1x1 b tc_all1x1 --no-cot-leg
COTB200_EVAL_FUSED_AGG=0 b eval_nofuse
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/i_prof_cotnet50_eval.md ) > gpurun_out/i_prof.log 2>&1
( COTB200_EVAL_FUSED_AGG=0 timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --eval --out gpurun_out/i_prof_cotnet50_eval_nofuse.md ) > gpurun_out/i_prof2.log 2>&1
( timeout 300 python tools/profile_step.py --model cotnet50 --batch 256 --out gpurun_out/i_prof_cotnet50_train.md ) > gpurun_out/i_prof_train.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/i_targets python tools/ncu_targets.py --what eval,gemm ) > gpurun_out/i_ncu.log 2>&1
ls -la gpurun_out/i_targets.ncu-rep
tail -25 gpurun_out/i_tests.log | cut -c1-250
echo '--- halo base offset variant'; tail -6 gpurun_out/i_tests_halo_baseoff.log | cut -c1-250; tail -2 gpurun_out/i_tests_halo_off.log
tail -8 gpurun_out/i_tests_trainer.log | cut -c1-250
tail -4 gpurun_out/i_tests_cotnext_cudnn.log | cut -c1-250
python - <<'PY'
import json
for n in ("default","stem_cudnn","tc_all1x1","eval_nofuse"):
    try:
        d=json.loads(open("gpurun_out/i_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), "cot_forward", {k:v for k,v in (d.get("cot_forward") or {}).items() if k!='mode'})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/i_bench_%s.err"%n).read()[-600:])
for n in ("default","cudnn"):
    try:
        d=json.load(open("gpurun_out/i_parity_%s.json"%n))["bench_path_cotnext50_2x48d"]
        print(n, {k:{a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items()} for k in ("bench_graph","bench_eager","plain_amp")})
    except Exception as e:
        print(n, "ERR", e)
PY
head -14 gpurun_out/i_prof_cotnet50_eval.md | cut -c1-140
tail -6 gpurun_out/i_bench_ref_kernels.log | cut -c1-500
head -14 gpurun_out/i_prof_cotnet50_eval_nofuse.md | cut -c1-140
