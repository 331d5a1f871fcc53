#!/bin/bash
# round 2, GPU call E (2 GPUs): the data-parallel step -- NCCL captured in the step graph with chunked overlap vs eager collectives
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; ( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e "$@" > gpurun_out/e2_$name.json ) 2> gpurun_out/e2_$name.err; head -c 300 gpurun_out/e2_$name.json; echo; tail -3 gpurun_out/e2_$name.err; }
run default
run nocapture --no-nccl-capture
run nooverlap --no-overlap
python - <<'PY'
import json
for n in ("default","nocapture","nooverlap"):
    try:
        d=json.loads(open("gpurun_out/e2_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f ms %.2f"%(d["value"], d["ms_per_step"]), d.get("launch_mode"), d.get("comm"), d.get("comm_error"))
    except Exception as e:
        print(n, "ERR", e)
PY
