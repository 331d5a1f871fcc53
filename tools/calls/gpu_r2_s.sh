#!/bin/bash
# round 2, GPU call S (2 GPUs, last minutes): does the N > 1 bench leave by itself (exit code 0, no teardown hang)?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cot-leg --no-cpu-baseline > gpurun_out/s_2gpu.json 2> gpurun_out/s_2gpu.err
rc=$?
t1=$(date +%s)
echo "exit code $rc after $((t1 - t0)) s" | tee gpurun_out/s_exit.txt
head -c 300 gpurun_out/s_2gpu.json; echo; tail -3 gpurun_out/s_2gpu.err | cut -c1-200
