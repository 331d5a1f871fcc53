#!/bin/bash
# round 2, GPU call N: validation of the final defaults -- whole GPU suite, smoke, final bench lines (configs 2-5 + reference arm),
# ncu launch list and DRAM-traffic pass of the same bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | tail -40 ) > gpurun_out/n_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/n_smoke.log 2>&1
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench_cotnet50.json ) 2> gpurun_out/n_bench_cotnet50.err
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/n_bench_reference_arm.json ) 2> gpurun_out/n_bench_reference_arm.err
b() { name=$1; shift; ( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/n_bench_$name.json ) 2> gpurun_out/n_bench_$name.err; }
b cotnext50 --model cotnext50_2x48d
b secotnetd101 --model se_cotnetd_101 --batch 128
b secotnetd152 --model se_cotnetd_152 --batch 64 --res 320
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10500 -c 4000 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-e2e --no-cpu-baseline --no-cot-leg > gpurun_out/n_ncu_bench.log 2>&1 )
( timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"bn_bwd_apply|bn_bwd_sums|bn_apply|col_stats" -s 1020 -c 340 --csv --log-file gpurun_out/n_traffic.csv python bench.py --steps 1 --warmup 3 --graph off --no-e2e --no-cpu-baseline --no-cot-leg > gpurun_out/n_ncu_traffic.log 2>&1 )
tail -12 gpurun_out/n_tests.log | cut -c1-250
tail -3 gpurun_out/n_smoke.log
python - <<'PY'
import json
for n in ("cotnet50","reference_arm","cotnext50","secotnetd101","secotnetd152"):
    try:
        d=json.loads(open("gpurun_out/n_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.1f ms %s"%(d["value"], d.get("ms_per_step")), "e2e", (d.get("e2e") or {}).get("value"), "roof", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), "step_frac", (d.get("step_roofline") or {}).get("frac"), "cot", {k:v for k,v in (d.get("cot_forward") or {}).items() if k in ("cot_layers_ms","frac")}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/n_bench_%s.err"%n).read()[-600:])
PY
wc -l gpurun_out/n_launches.csv gpurun_out/n_traffic.csv
