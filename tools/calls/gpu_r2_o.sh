#!/bin/bash
# round 2, GPU call O: forked block outputs (the two gradients of a block output summed inside bn3's backward kernels): tests, A/B on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_cot_layer_gpu.py tests/test_trainer_gpu.py tests/test_tc_gemm_gpu.py -m gpu -q --maxfail=30 2>&1 | tail -40 ) > gpurun_out/o_tests.log 2>&1
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-cot-leg "$@" > gpurun_out/o_bench_$name.json ) 2> gpurun_out/o_bench_$name.err; }
b c50_fork
COTB200_FORK=0 b c50_nofork
b d101_fork --model se_cotnetd_101 --batch 128
COTB200_FORK=0 b d101_nofork --model se_cotnetd_101 --batch 128
b next50_fork --model cotnext50_2x48d
COTB200_FORK=0 b next50_nofork --model cotnext50_2x48d
( timeout 400 python tools/bench_block.py --train --json gpurun_out/o_bench_block_train.json ) > gpurun_out/o_bench_block_train.log 2>&1
tail -10 gpurun_out/o_tests.log | cut -c1-250
python - <<'PY'
import json
for n in ("c50_fork","c50_nofork","d101_fork","d101_nofork","next50_fork","next50_nofork"):
    try:
        d=json.loads(open("gpurun_out/o_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["roofline"]["all_kernels"]; lib=sum(v["ms_per_step"] for v in k.values())
        print(n, "img/s %.0f ms %.2f lib %.2f"%(d["value"], d["ms_per_step"], lib), {x:k[x]["ms_per_step"] for x in ("bn_bwd_sums","bn_bwd_apply") if x in k})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/o_bench_%s.err"%n).read()[-500:])
PY
tail -5 gpurun_out/o_bench_block_train.log | cut -c1-500
