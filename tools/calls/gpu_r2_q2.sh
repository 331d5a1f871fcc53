#!/bin/bash
# round 2, GPU call Q2: the fp32 TrainStep-vs-PyTorch-loop test three times with and without the step bookkeeping (is its worst-parameter
# spread run-to-run noise or the bookkeeping?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for i in 1 2 3; do
  ( timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -k "plain_pytorch_loop or forked_block" 2>&1 | grep -E "^E  |passed|failed" | head -8 ) > gpurun_out/q2_new_$i.log 2>&1
  ( COTB200_BOOKKEEPING=0 timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -k "plain_pytorch_loop" 2>&1 | grep -E "^E  |passed|failed" | head -8 ) > gpurun_out/q2_old_$i.log 2>&1
done
for i in 1 2 3; do echo "new $i"; cat gpurun_out/q2_new_$i.log | cut -c1-300; echo "old $i"; cat gpurun_out/q2_old_$i.log | cut -c1-300; done
