#!/bin/bash
# run 19: tests of the hybrid backends + in-graph A/B (tc_e0 default vs tc_e0e3 vs old GroupNorm kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cot_layer_gpu.py tests/test_fused_gpu.py tests/test_tc_gemm_gpu.py -m gpu -q --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for cfg in "tc_e0 1" "tc_e0e3 1" "tc_e0 0"; do
  set -- $cfg
  COTB200_TRAIN_CONV=$1 COTB200_GN72=$2 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab_$1_$2.json 2> gpurun_out/bench_ab_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab_$1_$2.json")); print("$1 gn72=$2", round(d["value"],1), round(d["ms_per_step"],2), d["launch_mode"])
except Exception as e:
    print("$1 $2", "FAILED", e); print(open("gpurun_out/bench_ab_$1_$2.err").read()[-1500:])
PY
done
