#!/bin/bash
# round 2, GPU call M: wave-quantised geometry + vector reductions in the BatchNorm / tail kernels; backend policy A/B (one box):
# tc_all1x1+k vs tc_e0 vs pixel thresholds on CoTNet-50 bs256 and SE-CoTNetD-101 bs128 / -152 320^2 bs64
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_cot_layer_gpu.py tests/test_trainer_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -40 ) > gpurun_out/m_tests.log 2>&1
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg "$@" > gpurun_out/m_bench_$name.json ) 2> gpurun_out/m_bench_$name.err; }
b c50_default
COTB200_TRAIN_CONV=tc_e0 b c50_tc_e0
COTB200_TC_MIN_PIXELS=30000 b c50_px30k
COTB200_TC_MIN_PIXELS=100000 b c50_px100k
b d101_default --model se_cotnetd_101 --batch 128
COTB200_TRAIN_CONV=tc_e0 b d101_tc_e0 --model se_cotnetd_101 --batch 128
COTB200_TC_MIN_PIXELS=30000 b d101_px30k --model se_cotnetd_101 --batch 128
COTB200_TC_MIN_PIXELS=100000 b d101_px100k --model se_cotnetd_101 --batch 128
b d152_default --model se_cotnetd_152 --batch 64 --res 320
COTB200_TC_MIN_PIXELS=100000 b d152_px100k --model se_cotnetd_152 --batch 64 --res 320
COTB200_TRAIN_CONV=tc_e0 b d152_tc_e0 --model se_cotnetd_152 --batch 64 --res 320
tail -8 gpurun_out/m_tests.log | cut -c1-250
python - <<'PY'
import json
for n in ("c50_default","c50_tc_e0","c50_px30k","c50_px100k","d101_default","d101_tc_e0","d101_px30k","d101_px100k","d152_default","d152_px100k","d152_tc_e0"):
    try:
        d=json.loads(open("gpurun_out/m_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["roofline"]["all_kernels"]; lib=sum(v["ms_per_step"] for v in k.values())
        print(n, "img/s %.0f ms %.2f lib %.2f"%(d["value"], d["ms_per_step"], lib), {x:k[x]["ms_per_step"] for x in ("bn_bwd_sums","bn_bwd_apply","bn_apply_batch","col_stats","tail_pool") if x in k})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/m_bench_%s.err"%n).read()[-400:])
PY
