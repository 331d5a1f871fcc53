#!/usr/bin/env python
"""Debug helper: fused avg-pool backward vs torch (CUDA NCHW, CUDA channels_last, CPU)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from cotnet_b200 import fused

torch.manual_seed(0)
B, C, H, W = 2, 16, 8, 8
x = torch.randn(B, C, H, W, device="cuda")
cot = torch.randn(B, C, 4, 4, device="cuda")
pool = nn.AvgPool2d(3, 2, padding=1)


def grad(xin, fn, c):
    xin = xin.clone().requires_grad_(True)
    y = fn(xin)
    (g,) = torch.autograd.grad(y, xin, c)
    return y.detach(), g


xcl = x.contiguous(memory_format=torch.channels_last)
ccl = cot.contiguous(memory_format=torch.channels_last)
y0, g0 = grad(x, pool, cot)
y1, g1 = grad(xcl, pool, ccl)
y2, g2 = grad(x.cpu(), pool, cot.cpu())
y3, g3 = grad(xcl, fused.avg_pool3x3s2, ccl)
print("fwd  ours-vs-nchw %.3e  cl-vs-nchw %.3e  cpu-vs-nchw %.3e" % ((y3 - y0).abs().max(), (y1 - y0).abs().max(), (y2.cuda() - y0).abs().max()))
print("bwd  ours-vs-nchw %.3e  cl-vs-nchw %.3e  cpu-vs-nchw %.3e" % ((g3 - g0).abs().max(), (g1 - g0).abs().max(), (g2.cuda() - g0).abs().max()))
print("ours[0,0]:\n", g3[0, 0])
print("ref [0,0]:\n", g0[0, 0])
print("cot [0,0]:\n", cot[0, 0])
