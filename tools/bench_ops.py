#!/usr/bin/env python
"""Op-level timing of the LocalConv kernels at the BASELINE stage shapes.

    python tools/bench_ops.py [--batch 256] [--iters 20] [--json out.json] [--only nhwc|nchw] [--once]

CUDA-event timing on the launch stream, >= 3 warm-ups, inputs rotated through more than 126 MB (L2) so every
timed launch streams from HBM.  Prints GB/s of ALGORITHMIC bytes and the fraction of the measured HBM peak.
``--once`` runs each kernel exactly once per shape (for ncu captures)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cotnet_b200 import _lib, aggregation_zeropad, aggregation_zeropad_mix, fused  # noqa: E402

STAGES = [(64, 8, 56), (128, 16, 28), (256, 32, 14), (512, 64, 7)]


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_fn(fn, nsets, iters, once):
    if once:
        fn(0)
        torch.cuda.synchronize()
        return 0.0
    for i in range(3):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--stages", default="0,1,2,3")
    a = ap.parse_args()
    _lib.load()
    dev = torch.device("cuda")
    pk = peak()
    rows = []
    B = a.batch
    for si in [int(s) for s in a.stages.split(",")]:
        C, wc, HW = STAGES[si]
        for dtype, es in ((torch.bfloat16, 2), (torch.float32, 4)):
            for layout in ("tap", "nhwc", "nchw"):
                if a.only and a.only != layout:
                    continue
                elems = (2 * C + 9 * wc) * HW * HW * B
                nsets = 1 if a.once else max(2, int(300e6 // (elems * es)) + 1)
                xs, ws, gs = [], [], []
                for _ in range(nsets):
                    x = torch.randn(B, C, HW, HW, device=dev, dtype=dtype)
                    w = torch.randn(B, 1, wc, 9, HW, HW, device=dev, dtype=dtype)
                    g = torch.randn(B, C, HW, HW, device=dev, dtype=dtype)
                    if layout == "tap":      # block-internal order: w [B, 9*wc, H, W] channels_last, tap-major chunks of 8
                        x = x.contiguous(memory_format=torch.channels_last)
                        g = g.contiguous(memory_format=torch.channels_last)
                        w = torch.randn(B, 9 * wc, HW, HW, device=dev, dtype=dtype).contiguous(memory_format=torch.channels_last)
                    if layout == "nhwc":
                        x = x.contiguous(memory_format=torch.channels_last)
                        g = g.contiguous(memory_format=torch.channels_last)
                        w = w.permute(0, 4, 5, 1, 2, 3).contiguous().permute(0, 3, 4, 5, 1, 2)
                    xs.append(x.requires_grad_(True)); ws.append(w.requires_grad_(True)); gs.append(g)
                ys = [None] * nsets

                def fwd(i):
                    if layout == "tap":
                        ys[i] = fused.AggTapFn.apply(xs[i], ws[i], 1, 8)
                    else:
                        ys[i] = aggregation_zeropad(xs[i], ws[i], 3, 1, 1, 1)

                def bwd(i):
                    torch.autograd.grad(ys[i], (xs[i], ws[i]), gs[i], retain_graph=True)

                t_f = time_fn(fwd, nsets, a.iters, a.once)
                for i in range(nsets):
                    fwd(i)
                t_b = time_fn(bwd, nsets, a.iters, a.once)
                bf = elems * es
                bb = (3 * C + 2 * 9 * wc) * HW * HW * B * es
                row = {"C": C, "HW": HW, "B": B, "dtype": str(dtype).split(".")[-1], "layout": layout,
                       "fwd_us": 1e3 * t_f, "fwd_GBs": bf / (t_f * 1e6) if t_f else None,
                       "fwd_frac": bf / (t_f * 1e6) / pk if t_f else None,
                       "bwd_us": 1e3 * t_b, "bwd_GBs": bb / (t_b * 1e6) if t_b else None,
                       "bwd_frac": bb / (t_b * 1e6) / pk if t_b else None}
                rows.append(row)
                print(json.dumps(row), flush=True)
                del xs, ws, gs, ys
                torch.cuda.empty_cache()
    # mix op at its SURVEY shape (fp32 NCHW)
    if not a.only or a.only == "nchw":
        C, wc, HW = 64, 8, 56
        x = torch.randn(B, C, HW, HW, device=dev)
        w1 = torch.randn(B, 1, wc, 9, HW, HW, device=dev)
        w2 = torch.randn(B, 1, wc, 25, HW, HW, device=dev)
        t = time_fn(lambda i: aggregation_zeropad_mix(x, w1, w2, 3, 5, 1, 1, 2, 1), 1, a.iters, a.once)
        bm = (3 * C + 34 * wc) * HW * HW * B * 4
        row = {"op": "mix_fwd", "C": C, "HW": HW, "B": B, "dtype": "float32", "us": 1e3 * t,
               "GBs": bm / (t * 1e6) if t else None, "frac": bm / (t * 1e6) / pk if t else None,
               "note": "single buffer set (1.6 GB > L2)"}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.json:
        json.dump({"peak_gbs": pk, "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
