#!/usr/bin/env python
"""The reference's own LocalConv kernels (oracle/_ref cubins) timed beside libcotb200 at the CoTNet-50 bs256 stage shapes.

    python tools/bench_ref_kernels.py [--iters 10] [--json out.json]

Reference arm: fp32 NCHW (the only layout/dtypes the CuPy path takes, cupy_layers/utils.py:8-12), launched exactly as
aggregation_zeropad.py:140-143.  Ours: fp32 NCHW through the same operator API (drop-in case) and bf16 channels_last with
tap-major weights (the layout the fused block uses).  CUDA-event timing, inputs rotated through > 126 MB."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cotnet_b200  # noqa: E402
from cotnet_b200 import fused  # noqa: E402
from oracle import ref_kernels  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rk = ref_kernels.RefKernels()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    rows = []
    flush = torch.empty(200 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for tag in ("s1_b256", "s2_b256", "s3_b256", "s4_b256"):
        e, x, w, dy = rk.make_inputs(tag, "float")
        N, C, H, W, wc = e["N"], e["C"], e["H"], e["W"], e["wc"]
        alg = (2 * C + 9 * wc) * H * W * N                       # elements per pass (x, w, y)
        rec = {"tag": tag, "C": C, "HW": H, "B": N}

        def t(fn):
            def g():
                flush.zero_()                                     # evict L2 between timed launches
                fn()
            base = timeit(lambda: flush.zero_(), a.iters)
            return timeit(g, a.iters) - base

        y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.empty_like(w)
        ef = rk.entry("aggregation_zeropad_forward_kernel", tag, "float")
        ei = rk.entry("aggregation_zeropad_input_backward_kernel", tag, "float")
        ew = rk.entry("aggregation_zeropad_weight_backward_kernel", tag, "float")
        rec["ref_fp32_fwd_us"] = t(lambda: rk.launch(ef, x, w, y))
        rec["ref_fp32_dx_us"] = t(lambda: rk.launch(ei, dy, w, dx))
        rec["ref_fp32_dw_us"] = t(lambda: rk.launch(ew, dy, x, dw))
        # ours, drop-in contract: fp32 NCHW through the operator mirror
        xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        rec["ours_fp32_nchw_fwd_us"] = t(lambda: cotnet_b200.aggregation_zeropad(x, w, 3, 1, 1, 1))
        yo = cotnet_b200.aggregation_zeropad(xo, wo, 3, 1, 1, 1)
        rec["ours_fp32_nchw_bwd_us"] = t(lambda: torch.autograd.grad(yo, (xo, wo), dy, retain_graph=True))
        # ours, block-internal layout: bf16 channels_last, tap-major weights
        gc = fused.tap_chunk(wc)
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        wb = torch.randn(N, 9 * wc, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        db = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            rec["ours_bf16_tap_fwd_us"] = t(lambda: fused.AggTapFn.apply(xb, wb, 1, gc))
        yb = fused.AggTapFn.apply(xb, wb, 1, gc)
        rec["ours_bf16_tap_bwd_us"] = t(lambda: torch.autograd.grad(yb, (xb, wb), db, retain_graph=True))
        rec["ref_fp32_fwd_frac"] = alg * 4 / rec["ref_fp32_fwd_us"] / 1e3 / peak
        rec["ours_fp32_nchw_fwd_frac"] = alg * 4 / rec["ours_fp32_nchw_fwd_us"] / 1e3 / peak
        rec["ours_bf16_tap_fwd_frac"] = alg * 2 / rec["ours_bf16_tap_fwd_us"] / 1e3 / peak
        rec["speedup_fwd_fp32"] = rec["ref_fp32_fwd_us"] / rec["ours_fp32_nchw_fwd_us"]
        rec["speedup_bwd_fp32"] = (rec["ref_fp32_dx_us"] + rec["ref_fp32_dw_us"]) / rec["ours_fp32_nchw_bwd_us"]
        rec["speedup_fwd_bf16_tap"] = rec["ref_fp32_fwd_us"] / rec["ours_bf16_tap_fwd_us"]
        rec["speedup_bwd_bf16_tap"] = (rec["ref_fp32_dx_us"] + rec["ref_fp32_dw_us"]) / rec["ours_bf16_tap_bwd_us"]
        rec = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in rec.items()}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
        del x, w, dy, y, dx, dw, xo, wo, yo, xb, wb, db, yb
        torch.cuda.empty_cache()
    # ---- the mix op and the three cupy_layers variants at the shapes their cubins were built for
    def t2(fn):
        def g():
            flush.zero_()
            fn()
        return timeit(g, a.iters) - timeit(lambda: flush.zero_(), a.iters)

    e, x, w1, w2, dy = rk.make_inputs("mix_s1_b32", "float", op="mix")
    rec = {"tag": "mix_s1_b32", "op": "aggregation_zeropad_mix", "shape": [e["N"], e["C"], e["H"], e["W"]]}
    y = torch.empty(e["N"], 2 * e["C"], e["H"], e["W"], device="cuda")
    ef = rk.entry("aggregation_zeropad_mix_forward_kernel", "mix_s1_b32", "float")
    rec["ref_fp32_fwd_us"] = t2(lambda: rk.launch(ef, x, w1, w2, y))
    rec["ref_fp32_bwd_us"] = t2(lambda: rk.mix_backward("mix_s1_b32", dy, x, w1, w2))
    xo, w1o, w2o = [t_.clone().requires_grad_(True) for t_ in (x, w1, w2)]
    rec["ours_fp32_fwd_us"] = t2(lambda: cotnet_b200.aggregation_zeropad_mix(x, w1, w2, 3, 5, 1, 1, 2, 1))
    yo = cotnet_b200.aggregation_zeropad_mix(xo, w1o, w2o, 3, 5, 1, 1, 2, 1)
    rec["ours_fp32_bwd_us"] = t2(lambda: torch.autograd.grad(yo, (xo, w1o, w2o), dy, retain_graph=True))
    rec["speedup_fwd"] = rec["ref_fp32_fwd_us"] / rec["ours_fp32_fwd_us"]
    rec["speedup_bwd"] = rec["ref_fp32_bwd_us"] / rec["ours_fp32_bwd_us"]
    rows.append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in rec.items()})
    print(json.dumps(rows[-1]), flush=True)
    for op, tag, fwd, bwd, ours in (
            ("refpad", "refpad_s2_b8", rk.refpad_forward, rk.refpad_backward,
             lambda x_, w_, *r: cotnet_b200.aggregation_refpad(x_, w_, 3, 1, 1, 1)),
            ("dilate", "dilate_s2_b8", rk.dilate_forward, rk.dilate_backward,
             lambda x_, w_, dil: cotnet_b200.aggregation_zeropad_dilate(x_, w_, dil, 3, 1)),
            ("merge", "merge_s1_b8", rk.merge_forward, rk.merge_backward,
             lambda x_, w_, e_: cotnet_b200.aggregation_zeropad_mix_merge(x_, w_, e_["heads"], e_["wc"], 3, 5, 1, 1, 2, 1))):
        inp = rk.make_variant_inputs(tag, "float", op)
        e, x, w = inp[0], inp[1], inp[2]
        dy = inp[-1]
        extra = (inp[3],) if op == "dilate" else ((e,) if op == "merge" else ())
        rec = {"tag": tag, "op": op}
        rec["ref_fp32_fwd_us"] = t2(lambda: fwd(tag, x, w, *(extra if op == "dilate" else ())))
        rec["ref_fp32_bwd_us"] = t2(lambda: bwd(tag, dy, x, w, *(extra if op == "dilate" else ())))
        xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        rec["ours_fp32_fwd_us"] = t2(lambda: ours(x, w, *extra))
        yo = ours(xo, wo, *extra)
        rec["ours_fp32_bwd_us"] = t2(lambda: torch.autograd.grad(yo, (xo, wo), dy, retain_graph=True))
        rec["speedup_fwd"] = rec["ref_fp32_fwd_us"] / rec["ours_fp32_fwd_us"]
        rec["speedup_bwd"] = rec["ref_fp32_bwd_us"] / rec["ours_fp32_bwd_us"]
        rows.append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in rec.items()})
        print(json.dumps(rows[-1]), flush=True)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
