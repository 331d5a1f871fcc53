#!/usr/bin/env python
"""Stem convolution (7x7/s2, 3 -> 64, bs256, 224^2, bf16 autocast, channels_last) forward + weight gradient, with the input
zero-padded to 3 / 4 / 8 channels -- decides CoTResNet.stem_pad (cuDNN kernel selection depends on the channel alignment)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    x3 = torch.randn(256, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    w3 = torch.randn(64, 3, 7, 7, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = {}
    for pad_to in (3, 4, 8):
        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                if pad_to > 3:
                    x = F.pad(x3, (0, 0, 0, 0, 0, pad_to - 3)).contiguous(memory_format=torch.channels_last)
                    w = F.pad(w3, (0, 0, 0, 0, 0, pad_to - 3)).contiguous(memory_format=torch.channels_last)
                else:
                    x, w = x3, w3
                y = F.conv2d(x, w, None, 2, 3)
            (g,) = torch.autograd.grad(y, w3, torch.ones_like(y))
            return g
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        out["pad_to_%d_fwd_wgrad_us" % pad_to] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
