"""CPU restatement of the CoT block.  TEST INFRASTRUCTURE ONLY.

Functional (state-dict in, tensor out) restatement of
    CotLayer   models/cotnet.py:36-104
    CoXtLayer  models/cotnet.py:106-178
    CoTLayer   models/cotnet_hybrid.py:48-116   (same arithmetic as CotLayer)
following SURVEY.md Appendix A step by step.  Pinned by ``tests/golden/cot_*.npz`` -- outputs
and gradients of the reference's own unmodified module code (``oracle/make_golden.py``).

Everything is plain torch on whatever dtype/device the inputs have (fp64 CPU in tests), so
autograd through this function is the backward oracle.  BatchNorm follows ``nn.BatchNorm2d``
defaults (eps 1e-5, momentum 0.1, biased variance for normalisation / unbiased for the running
estimate); in training mode the running buffers in ``sd`` are updated in place like the module
would, so buffer parity can be checked too.
"""
import torch
import torch.nn.functional as F

from . import agg_ref

EPS = 1e-5


def _bn(t, sd, prefix, training, momentum=0.1):
    """nn.BatchNorm2d (models/cotnet.py:45,53,61,65,74) on [B,C,H,W] or [B,C,1,1]."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        dims = (0, 2, 3)
        mean = t.mean(dims)
        var = t.var(dims, unbiased=False)
        n = t.numel() / t.shape[1]
        with torch.no_grad():
            rm.mul_(1 - momentum).add_(momentum * mean.detach().to(rm.dtype))
            rv.mul_(1 - momentum).add_(momentum * (var.detach() * n / max(n - 1, 1)).to(rv.dtype))
            key = prefix + ".num_batches_tracked"
            if key in sd:
                sd[key] += 1
    else:
        mean, var = rm.to(t.dtype), rv.to(t.dtype)
    scale = w / torch.sqrt(var + EPS)
    return (t - mean[None, :, None, None]) * scale[None, :, None, None] + b[None, :, None, None]


def _group_norm9(l, gamma, beta, groups):
    """nn.GroupNorm(C/8, 9C/8) (models/cotnet.py:56): group g = the 9 tap-channels 9g..9g+8,
    statistics over 9*H*W values of one sample, biased variance, per-channel affine."""
    B, J, H, W = l.shape
    t = l.reshape(B, groups, (J // groups) * H * W)
    m = t.mean(-1, keepdim=True)
    v = t.var(-1, unbiased=False, keepdim=True)
    t = ((t - m) / torch.sqrt(v + EPS)).reshape(B, J, H, W)
    return t * gamma[None, :, None, None] + beta[None, :, None, None]


def _radix_tail(y, k, sd, training):
    """models/cotnet.py:92-104: split attention over (y, k) with radix 2."""
    B, C, H, W = y.shape
    gap = (y + k).mean((2, 3), keepdim=True)                                    # :95-98
    q = F.conv2d(gap, sd["se.0.weight"], sd["se.0.bias"])                       # :73
    q = torch.relu(_bn(q, sd, "se.1", training))                               # :74-75
    z = F.conv2d(q, sd["se.3.weight"], sd["se.3.bias"])                         # :76
    a = torch.softmax(z.view(B, C, 2), dim=2)                                   # :100-101 (channel 2c+r)
    out = y * a[:, :, 0].reshape(B, C, 1, 1) + k * a[:, :, 1].reshape(B, C, 1, 1)  # :102
    return out.contiguous()


def cot_layer(x, sd, training=False, kernel_size=3):
    """CotLayer.forward, models/cotnet.py:79-104."""
    B, C, H, W = x.shape
    ks = kernel_size
    k = F.conv2d(x, sd["key_embed.0.weight"], None, 1, ks // 2, 1, 4)           # :44 groups=4
    k = torch.relu(_bn(k, sd, "key_embed.1", training))                        # :45-46
    qk = torch.cat([x, k], dim=1)                                               # :81  x-block then k-block
    e = F.conv2d(qk, sd["embed.0.weight"])                                      # :52
    e = torch.relu(_bn(e, sd, "embed.1", training))                            # :53-54
    l = F.conv2d(e, sd["embed.3.weight"], sd["embed.3.bias"])                   # :55
    g = _group_norm9(l, sd["embed.4.weight"], sd["embed.4.bias"], C // 8)       # :56
    w = g.view(B, 1, C // 8, ks * ks, H, W)                                     # :85
    v = _bn(F.conv2d(x, sd["conv1x1.0.weight"]), sd, "conv1x1.1", training)     # :59-62, :87
    u = agg_ref.agg_zeropad_unfold(v, w, ks, 1, (ks - 1) // 2, 1)               # :88
    y = F.silu(_bn(u, sd, "bn", training))                                      # :89-90
    return _radix_tail(y, k, sd, training)


def coxt_layer(x, sd, training=False, kernel_size=3):
    """CoXtLayer.forward, models/cotnet.py:150-178."""
    B, C, H, W = x.shape
    ks = kernel_size
    k = F.conv2d(x, sd["key_embed.0.weight"], None, 1, ks // 2, 1, 8)           # :114 groups=8
    k = torch.relu(_bn(k, sd, "key_embed.1", training))
    qk = torch.stack([x, k], dim=2).reshape(B, 2 * C, H, W)                     # :153-154 interleaved x0,k0,x1,k1,..
    e = F.conv2d(qk, sd["embed.0.weight"], None, 1, 0, 1, 2)                    # :123 groups=2
    e = torch.relu(_bn(e, sd, "embed.1", training))
    l = F.conv2d(e, sd["embed.3.weight"], sd["embed.3.bias"], 1, 0, 1, 2)       # :126 groups=2
    g = _group_norm9(l, sd["embed.4.weight"], sd["embed.4.bias"], C // 8)       # :127
    w = g.view(B * 2, 1, -1, ks * ks, H, W)                                     # :157
    v = _bn(F.conv2d(x, sd["conv1x1.0.weight"], None, 1, 0, 1, 2), sd, "conv1x1.1", training)  # :130-133
    v = v.view(B * 2, -1, H, W)                                                 # :160
    u = agg_ref.agg_zeropad_unfold(v, w, ks, 1, (ks - 1) // 2, 1).view(B, C, H, W)  # :161-162
    y = F.silu(_bn(u, sd, "bn", training))
    return _radix_tail(y, k, sd, training)


def init_state_dict(kind, dim, gen, dtype=torch.float64, perturb=True, kernel_size=3):
    """Random parameters with the reference's names/shapes.

    Conv weights: kaiming-normal fan_out like models/resnet.py:575-580; when ``perturb`` the
    BN/GN affines, running stats and biases are randomised so that
    BN-folding / affine bugs cannot hide behind the fresh-module identity values."""
    C, ks = dim, kernel_size
    A = max(C * 2 // 4, 32)
    if kind == "cot":
        gk, ge = 4, 1
    elif kind == "coxt":
        gk, ge = 8, 2
    else:
        raise ValueError(kind)

    def conv(co, ci, k):
        std = (2.0 / (co * k * k)) ** 0.5
        return torch.randn(co, ci, k, k, generator=gen, dtype=torch.float64).mul_(std).to(dtype)

    def vec(n, kind_):
        if not perturb:
            return {"w": torch.ones(n), "b": torch.zeros(n), "m": torch.zeros(n), "v": torch.ones(n)}[kind_].to(dtype)
        if kind_ == "w":
            return (torch.rand(n, generator=gen, dtype=torch.float64) + 0.5).to(dtype)
        if kind_ == "b":
            return (torch.randn(n, generator=gen, dtype=torch.float64) * 0.2).to(dtype)
        if kind_ == "m":
            return (torch.randn(n, generator=gen, dtype=torch.float64) * 0.5).to(dtype)
        if kind_ == "v":
            return (torch.rand(n, generator=gen, dtype=torch.float64) * 1.5 + 0.5).to(dtype)

    sd = {}

    def bn(prefix, n):
        sd[prefix + ".weight"] = vec(n, "w")
        sd[prefix + ".bias"] = vec(n, "b")
        sd[prefix + ".running_mean"] = vec(n, "m")
        sd[prefix + ".running_var"] = vec(n, "v")
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    sd["key_embed.0.weight"] = conv(C, C // gk, ks)
    bn("key_embed.1", C)
    sd["embed.0.weight"] = conv(C // 2, 2 * C // ge, 1)
    bn("embed.1", C // 2)
    sd["embed.3.weight"] = conv(ks * ks * C // 8, (C // 2) // ge, 1)
    sd["embed.3.bias"] = (torch.randn(ks * ks * C // 8, generator=gen, dtype=torch.float64) * 0.1).to(dtype)
    sd["embed.4.weight"] = vec(ks * ks * C // 8, "w")
    sd["embed.4.bias"] = vec(ks * ks * C // 8, "b")
    sd["conv1x1.0.weight"] = conv(C, C // ge, 1)
    bn("conv1x1.1", C)
    bn("bn", C)
    sd["se.0.weight"] = conv(A, C, 1)
    sd["se.0.bias"] = (torch.randn(A, generator=gen, dtype=torch.float64) * 0.1).to(dtype)
    bn("se.1", A)
    sd["se.3.weight"] = conv(2 * C, A, 1)
    sd["se.3.bias"] = (torch.randn(2 * C, generator=gen, dtype=torch.float64) * 0.1).to(dtype)
    return sd
