#!/bin/bash
# round 2, GPU call P (2 GPUs): the data-parallel step with the final defaults -- NCCL captured in the step graph with chunked overlap,
# exposed communication, and the no-overlap / eager-collective variants for comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cot-leg --no-cpu-baseline "$@" > gpurun_out/p_$name.json ) 2> gpurun_out/p_$name.err; tail -2 gpurun_out/p_$name.err | cut -c1-300; }
run default
run nooverlap --no-overlap
( timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cot-leg --no-cpu-baseline > gpurun_out/p_single.json ) 2> gpurun_out/p_single.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/p_reference.json ) 2> gpurun_out/p_reference.err
python - <<'PY'
import json
for n in ("single","default","nooverlap","reference"):
    try:
        d=json.loads(open("gpurun_out/p_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "n_gpus", d.get("n_gpus"), "img/s %.0f ms %s"%(d["value"], d.get("ms_per_step")), d.get("launch_mode"), d.get("comm"), d.get("comm_error"), (d.get("config") or {}).get("step"))
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/p_%s.err"%n).read()[-500:])
PY
