#!/usr/bin/env python
"""Kernel-time table of ONE eager training step (torch.profiler / CUPTI): which kernels a model's step spends its time in.

    python tools/profile_step.py --model cotnext50_2x48d --batch 256 [--res 224] [--eval] [--out gpurun_out/prof.md]

Complements the ncu launch lists (which serialise launches): same step as bench.py (cotnet_b200.trainer.TrainStep, eager)."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cotnet_b200 import trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cotnet50")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--eval", action="store_true", help="profile the eval-mode bf16 forward instead of the training step")
    ap.add_argument("--out", default=None)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    m = bench.build_model(a.model, zero_init_last_bn=False).to(dev).to(memory_format=torch.channels_last).train()
    x = torch.randn(a.batch, 3, a.res, a.res, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device=dev)
    ts = trainer.TrainStep(m, lr=0.05, ema_decay=0.9999)
    if a.eval:
        m.eval()

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                m(x)
    else:
        def run():
            ts.step_eager(x, y)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        run()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += 1
            agg[ev.name][1] += ev.device_time
            total += ev.device_time
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    lines = ["# kernel time of one eager %s of %s bs%d %d^2 (torch.profiler): %.2f ms of kernels, %d launches" % (
        "eval forward" if a.eval else "training step", a.model, a.batch, a.res, total / 1e3, sum(v[0] for v in agg.values())), "",
        "| ms | share % | launches | kernel |", "|---|---|---|---|"]
    for name, (n, us) in rows[:a.top]:
        lines.append("| %.3f | %.1f | %d | `%s` |" % (us / 1e3, 100.0 * us / total, n, name[:110]))
    text = "\n".join(lines)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
