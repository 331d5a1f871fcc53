"""Arithmetic-intensity roofline of a CoT network step: the number bench.py divides by.

Per layer  T = max(FLOPs / P_tensor, bytes / BW_hbm)  with perfect fusion assumed: BatchNorm / activations / pooling /
residual adds ride along with the convolution that produces or consumes their tensor (zero cost), every convolution reads
its input and weights once and writes its output once, and a CoT block moves only its input and output
(2*C*H*W*B*s bytes, section 8d "CoT block forward").  Backward of a layer = data gradient + weight gradient:
2x the FLOPs, reads dY twice and x / w once, writes dX and dW.

    layer_table(model, x)         one real forward with shape-recording hooks; one row per unit
    step_roofline(rows, peaks)    sums the per-layer bounds for forward and forward+backward

This is measurement support, not product math.
"""
import torch
import torch.nn as nn

from .cot_layer import CotLayer, CoXtLayer


def cot_block_counts(C, H, W, B, s=2, groups_x=False):
    """FLOPs and algorithmic bytes of one CoT block forward: key 3x3 grouped conv (groups 4: 4.5 C^2/px;
    CoXt groups 8: 2.25 C^2), embed 2C->C/2 (2 C^2; CoXt groups 2: C^2), C/2 -> 9C/8 (1.125 C^2; CoXt 0.5625 C^2),
    value C->C (2 C^2; CoXt C^2), aggregation 18 C.  Bytes: input + output."""
    px = B * H * W
    c2 = float(C) * C
    if groups_x:
        flops = (2.25 + 1.0 + 0.5625 + 1.0) * c2 * px + 18.0 * C * px
        params = (9.0 / 8 + 0.5 + 9.0 / 32 + 0.5) * c2
    else:
        flops = (4.5 + 2.0 + 1.125 + 2.0) * c2 * px + 18.0 * C * px
        params = (9.0 / 4 + 1.0 + 9.0 / 16 + 1.0) * c2
    A = max(C // 2, 32)
    flops += 2.0 * B * (C * A + 2 * C * A)
    params += 3.0 * C * A
    return flops, 2.0 * C * px * s + params * s


def layer_table(model, x, s=2, batch=None):
    """Rows {kind, name, flops_fwd, bytes_fwd, flops_bwd, bytes_bwd} for every convolution / linear layer outside CoT
    blocks and one row per CoT block, recorded by forward hooks during ONE real forward of `model` on `x` (any small
    batch); `batch` rescales the activation terms to the batch actually benchmarked (weights are counted once)."""
    rows = []
    inside = set()
    for name, m in model.named_modules():
        if isinstance(m, (CotLayer, CoXtLayer)):
            for sub_name, _ in m.named_modules():
                if sub_name:
                    inside.add(name + "." + sub_name)
    hooks = []
    B0 = x.shape[0]
    k = 1.0 if batch is None else float(batch) / B0

    def conv_hook(name):
        def fn(mod, inp, out):
            xi = inp[0]
            if isinstance(mod, nn.Conv2d):
                Ho, Wo = out.shape[2], out.shape[3]
                macs = xi.shape[0] * Ho * Wo * mod.out_channels * (mod.in_channels // mod.groups) * mod.kernel_size[0] * mod.kernel_size[1]
            else:
                macs = xi.numel() // xi.shape[-1] * mod.in_features * mod.out_features
            fl = 2.0 * macs * k
            bi, bo, bw = xi.numel() * s * k, out.numel() * s * k, mod.weight.numel() * s
            rows.append(dict(kind="conv" if isinstance(mod, nn.Conv2d) else "linear", name=name, flops_fwd=fl,
                             bytes_fwd=bi + bo + bw, flops_bwd=2.0 * fl, bytes_bwd=2.0 * bo + 2.0 * bi + 2.0 * bw))
        return fn

    def cot_hook(name):
        def fn(mod, inp, out):
            B, C, H, W = inp[0].shape
            fl, by = cot_block_counts(C, H, W, int(round(B * k)), s, isinstance(mod, CoXtLayer))
            rows.append(dict(kind="cot", name=name, C=C, H=H, W=W, flops_fwd=fl, bytes_fwd=by, flops_bwd=2.0 * fl,
                             bytes_bwd=2.0 * by))
        return fn

    for name, m in model.named_modules():
        if isinstance(m, (CotLayer, CoXtLayer)):
            hooks.append(m.register_forward_hook(cot_hook(name)))
        elif isinstance(m, (nn.Conv2d, nn.Linear)) and name not in inside:
            hooks.append(m.register_forward_hook(conv_hook(name)))
    try:
        with torch.no_grad():
            model(x)
    finally:
        for h in hooks:
            h.remove()
    return rows


def step_roofline(rows, hbm_gbs, tflops):
    """Sum of per-layer bounds (ms) for the forward and for forward + backward, split CoT blocks / rest."""
    def t(fl, by):
        return max(fl / (tflops * 1e12), by / (hbm_gbs * 1e9)) * 1e3
    out = {"fwd_ms": 0.0, "fwd_bwd_ms": 0.0, "cot_fwd_ms": 0.0, "cot_fwd_bwd_ms": 0.0, "flops_fwd": 0.0, "bytes_fwd": 0.0}
    for r in rows:
        f, b = t(r["flops_fwd"], r["bytes_fwd"]), t(r["flops_bwd"], r["bytes_bwd"])
        out["fwd_ms"] += f
        out["fwd_bwd_ms"] += f + b
        out["flops_fwd"] += r["flops_fwd"]
        out["bytes_fwd"] += r["bytes_fwd"]
        if r["kind"] == "cot":
            out["cot_fwd_ms"] += f
            out["cot_fwd_bwd_ms"] += f + b
    return out
