"""CPU restatement of the LocalConv ("aggregation_zeropad") operator.  TEST INFRASTRUCTURE ONLY.

Two independent restatements of the same operator:

* ``*_loops``  -- numpy, follows the reference CUDA kernels' index math tap by tap:
    forward          cupy_layers/aggregation_zeropad.py:20-46
    input backward   cupy_layers/aggregation_zeropad.py:48-79
    weight backward  cupy_layers/aggregation_zeropad.py:81-110
    mix forward      cupy_layers/aggregation_zeropad_mix.py:20-74
    mix backward     cupy_layers/aggregation_zeropad_mix.py:76-207
* ``*_unfold`` -- torch, the identity the reference's self-tests assert the kernels equal
    (cupy_layers/aggregation_zeropad.py:249-251, aggregation_zeropad_mix.py:360-366);
    differentiable, so autograd gives dX / dW.

``tests/test_oracle.py`` pins one against the other at the reference self-test shapes with
the reference's own gate (max-abs < 1e-9 in fp64).

Layouts are the reference's: x [N,C,H,W]; w [N,heads,wc,kh*kw,Ho,Wo]; out [N,heads*C,Ho,Wo].
"""
import numpy as np
import torch


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def out_size(H, W, k, s, p, d):
    k, s, p, d = _pair(k), _pair(s), _pair(p), _pair(d)
    # aggregation_zeropad.py:119-120
    Ho = int((H + 2 * p[0] - (d[0] * (k[0] - 1) + 1)) / s[0] + 1)
    Wo = int((W + 2 * p[1] - (d[1] * (k[1] - 1) + 1)) / s[1] + 1)
    return Ho, Wo


# --------------------------------------------------------------------------- numpy loops
def agg_zeropad_fwd_loops(x, w, k=3, s=1, p=0, d=1):
    """aggregation_zeropad.py:20-46.  out[n, head*C + c, h, w] = sum_tap w[n,head,c%wc,tap,h,w] * x[n,c,h_in,w_in]."""
    k, s, p, d = _pair(k), _pair(s), _pair(p), _pair(d)
    N, C, H, W = x.shape
    _, heads, wc, K2, Ho, Wo = w.shape
    assert K2 == k[0] * k[1] and C % wc == 0
    out = np.zeros((N, heads, C, Ho, Wo), dtype=np.result_type(x.dtype, w.dtype))
    cidx = np.arange(C) % wc
    for kh in range(k[0]):
        for kw in range(k[1]):
            tap = kh * k[1] + kw
            # valid output range for this tap
            hs = [h for h in range(Ho) if 0 <= -p[0] + h * s[0] + kh * d[0] < H]
            ws = [q for q in range(Wo) if 0 <= -p[1] + q * s[1] + kw * d[1] < W]
            if not hs or not ws:
                continue
            hs = np.array(hs)
            ws = np.array(ws)
            hin = -p[0] + hs * s[0] + kh * d[0]
            win = -p[1] + ws * s[1] + kw * d[1]
            xs = x[:, :, hin[:, None], win[None, :]]                      # [N,C,h,w]
            wt = w[:, :, :, tap][:, :, :, hs[:, None], ws[None, :]]       # [N,heads,wc,h,w]
            wt = wt[:, :, cidx]                                           # [N,heads,C,h,w]
            out[:, :, :, hs[:, None], ws[None, :]] += wt * xs[:, None]
    return out.reshape(N, heads * C, Ho, Wo)


def agg_zeropad_bwd_loops(dy, x, w, k=3, s=1, p=0, d=1):
    """aggregation_zeropad.py:48-110 -> (dX [N,C,H,W], dW like w)."""
    k, s, p, d = _pair(k), _pair(s), _pair(p), _pair(d)
    N, C, H, W = x.shape
    _, heads, wc, K2, Ho, Wo = w.shape
    dy5 = dy.reshape(N, heads, C, Ho, Wo)
    dx = np.zeros_like(x, dtype=np.result_type(x.dtype, w.dtype))
    dw = np.zeros_like(w, dtype=np.result_type(x.dtype, w.dtype))
    cidx = np.arange(C) % wc
    for kh in range(k[0]):
        for kw in range(k[1]):
            tap = kh * k[1] + kw
            hs = [h for h in range(Ho) if 0 <= -p[0] + h * s[0] + kh * d[0] < H]
            ws = [q for q in range(Wo) if 0 <= -p[1] + q * s[1] + kw * d[1] < W]
            if not hs or not ws:
                continue
            hs = np.array(hs)
            ws = np.array(ws)
            hin = -p[0] + hs * s[0] + kh * d[0]
            win = -p[1] + ws * s[1] + kw * d[1]
            g = dy5[:, :, :, hs[:, None], ws[None, :]]                    # [N,heads,C,h,w]
            wt = w[:, :, :, tap][:, :, :, hs[:, None], ws[None, :]][:, :, cidx]
            # dX (kernel :48-79): sum over heads of w * dY at the output positions that read this input
            dx[:, :, hin[:, None], win[None, :]] += (wt * g).sum(1)   # (hin, win) unique within one tap
            # dW (kernel :81-110): sum over the C/wc channels sharing weight channel cw
            xs = x[:, :, hin[:, None], win[None, :]]                      # [N,C,h,w]
            prod = g * xs[:, None]                                        # [N,heads,C,h,w]
            prod = prod.reshape(N, heads, C // wc, wc, len(hs), len(ws)).sum(2)
            dw[:, :, :, tap][:, :, :, hs[:, None], ws[None, :]] = prod
    return dx, dw


def agg_zeropad_mix_fwd_loops(x, w1, w2, k1=3, k2=5, s=1, p1=0, p2=0, d=1):
    """aggregation_zeropad_mix.py:20-74: channel order [n, (kernel_idx*heads + head)*C + c]."""
    N, C, H, W = x.shape
    heads = w1.shape[1]
    y1 = agg_zeropad_fwd_loops(x, w1, k1, s, p1, d)
    y2 = agg_zeropad_fwd_loops(x, w2, k2, s, p2, d)
    return np.concatenate([y1, y2], axis=1)


def agg_zeropad_mix_bwd_loops(dy, x, w1, w2, k1=3, k2=5, s=1, p1=0, p2=0, d=1):
    """aggregation_zeropad_mix.py:76-207.  NOTE: the reference's dX kernel hard-codes head=0
    (:88); for heads==1 (the only configuration its self-test and callers use) that is the
    full gradient, which is what is restated here for any heads."""
    N, C, H, W = x.shape
    heads = w1.shape[1]
    half = heads * C
    dx1, dw1 = agg_zeropad_bwd_loops(dy[:, :half], x, w1, k1, s, p1, d)
    dx2, dw2 = agg_zeropad_bwd_loops(dy[:, half:], x, w2, k2, s, p2, d)
    return dx1 + dx2, dw1, dw2


# --------------------------------------------------------------------------- torch Unfold identity
def agg_zeropad_unfold(x, w, k=3, s=1, p=0, d=1):
    """aggregation_zeropad.py:247-251 (the self-test's right-hand side), any dtype, differentiable."""
    N, C, H, W = x.shape
    _, heads, wc, K2, Ho, Wo = w.shape
    unfold = torch.nn.Unfold(kernel_size=k, dilation=d, padding=p, stride=s)
    x2 = unfold(x).view(N, C // wc, wc, K2, Ho, Wo)
    y = (w.unsqueeze(2) * x2.unsqueeze(1)).sum(-3)
    return y.view(N, heads * C, Ho, Wo)


def agg_zeropad_mix_unfold(x, w1, w2, k1=3, k2=5, s=1, p1=0, p2=0, d=1):
    """aggregation_zeropad_mix.py:360-366."""
    return torch.cat([agg_zeropad_unfold(x, w1, k1, s, p1, d),
                      agg_zeropad_unfold(x, w2, k2, s, p2, d)], dim=1)
