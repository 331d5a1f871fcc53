"""Autograd functions over the fused normalisation / split-attention kernels (csrc/norm_tail.cu).

They turn the reference block's long eager chains into a few HBM passes while keeping autograd semantics:

  GroupNorm9Fn   nn.GroupNorm(C/8, 9C/8) on the attention logits          /root/reference/models/cotnet.py:56
                 (under autocast the eager op runs in fp32 and writes 4-byte logits; here: 2 reads + 1 write)
  CotTailFn      bn -> SiLU -> (y + k) pooled -> se -> softmax(radix 2) -> a0*y + a1*k        :89-104
                 fwd: [stats] + pool + combine = 3 passes; bwd: sums + dz-sums + apply = 3 passes;
                 the tiny SE MLP on [B, C] stays PyTorch and is differentiated by autograd inside the function.

All tensors are channels_last (NHWC memory); math is fp32; outputs keep the input dtype.
"""
import torch
from torch.autograd import Function

from . import _lib


def _is_cl(t):
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def supported(t):
    return t.is_cuda and t.dtype in (torch.float32, torch.bfloat16, torch.float16) and _is_cl(t) and not t.is_contiguous()


def _f32(t):
    return t.detach().float().contiguous()


class GroupNorm9Fn(Function):
    """l [B, 9*wc, H, W] channels_last -> GroupNorm with wc groups of 9 consecutive channels (gc=0) or tap-major
    chunks (gc>0); gamma/beta [9*wc]."""

    @staticmethod
    def forward(ctx, l, gamma, beta, eps, gc=0):
        assert _is_cl(l)
        B, J, H, W = l.shape
        wc, HW = J // 9, H * W
        lib, st, dt = _lib.load(), _lib.stream_ptr(l), _lib.dtype_code(l)
        l = l.detach()
        stats = torch.zeros(2, B, wc, dtype=torch.float32, device=l.device)
        _lib.check(lib.cotb200_gn9_stats(dt, B, HW, wc, gc, l.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), st), "gn9_stats")
        n = 9.0 * HW
        mean = stats[0] / n
        var = (stats[1] / n - mean * mean).clamp_min_(0.0)
        rstd = torch.rsqrt(var + eps)
        g32, b32 = _f32(gamma), _f32(beta)
        out = torch.empty_like(l, memory_format=torch.channels_last)
        _lib.check(lib.cotb200_gn9_apply(dt, B, HW, wc, gc, l.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g32.data_ptr(),
                                         b32.data_ptr(), out.data_ptr(), st), "gn9_apply")
        ctx.save_for_backward(l, mean, rstd, g32)
        ctx.gc = gc
        ctx.param_dtype = (gamma.dtype, beta.dtype)
        return out

    @staticmethod
    def backward(ctx, dg):
        l, mean, rstd, g32 = ctx.saved_tensors
        B, J, H, W = l.shape
        wc, HW = J // 9, H * W
        dg = dg.contiguous(memory_format=torch.channels_last)
        lib, st, dt = _lib.load(), _lib.stream_ptr(l), _lib.dtype_code(l)
        sums = torch.zeros(2, B, wc, dtype=torch.float32, device=l.device)
        dgb = torch.zeros(2, J, dtype=torch.float32, device=l.device)
        _lib.check(lib.cotb200_gn9_bwd_sums(dt, B, HW, wc, ctx.gc, dg.data_ptr(), l.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                            g32.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), dgb[0].data_ptr(),
                                            dgb[1].data_ptr(), st), "gn9_bwd_sums")
        dl = None
        if ctx.needs_input_grad[0]:
            dl = torch.empty_like(l, memory_format=torch.channels_last)
            _lib.check(lib.cotb200_gn9_bwd_apply(dt, B, HW, wc, ctx.gc, dg.data_ptr(), l.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                 g32.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), dl.data_ptr(), st),
                       "gn9_bwd_apply")
        return dl, dgb[0].to(ctx.param_dtype[0]), dgb[1].to(ctx.param_dtype[1]), None, None


class CotTailFn(Function):
    """(u, k) -> out of models/cotnet.py:89-104.  `bn` is the nn.BatchNorm2d(dim) module (its buffers are updated in
    training mode exactly like the module would), `se` the nn.Sequential producing the radix-2 logits."""

    @staticmethod
    def forward(ctx, u, k, bn_weight, bn_bias, bn, se, *se_params):
        assert _is_cl(u) and _is_cl(k) and u.dtype == k.dtype and u.shape == k.shape
        B, C, H, W = u.shape
        HW, n = H * W, float(B * H * W)
        lib, st, dt = _lib.load(), _lib.stream_ptr(u), _lib.dtype_code(u)
        u, k = u.detach(), k.detach()
        training = bn.training
        if training or bn.running_mean is None:
            stats = torch.zeros(2, C, dtype=torch.float32, device=u.device)
            _lib.check(lib.cotb200_col_stats(dt, B, HW, C, u.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), st), "col_stats")
            mean = stats[0] / n
            var = (stats[1] / n - mean * mean).clamp_min_(0.0)
            if bn.running_mean is not None and bn.track_running_stats:
                with torch.no_grad():
                    bn.num_batches_tracked += 1
                    mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                    bn.running_mean.mul_(1 - mom).add_(mean.to(bn.running_mean.dtype), alpha=mom)
                    bn.running_var.mul_(1 - mom).add_((var * (n / max(n - 1.0, 1.0))).to(bn.running_var.dtype), alpha=mom)
        else:
            mean, var = bn.running_mean.float(), bn.running_var.float()
        rstd = torch.rsqrt(var + bn.eps)
        scale = (bn_weight.detach().float() * rstd).contiguous()
        shift = (bn_bias.detach().float() - mean * scale).contiguous()
        mean, rstd = mean.contiguous(), rstd.contiguous()
        psum = torch.zeros(B, C, dtype=torch.float32, device=u.device)
        _lib.check(lib.cotb200_tail_pool(dt, B, HW, C, u.data_ptr(), k.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                         psum.data_ptr(), st), "tail_pool")
        # the SE MLP on [B, C] (3 tiny GEMV-sized ops) stays PyTorch; its graph is kept for backward
        with torch.enable_grad():
            p_leaf = (psum / HW).requires_grad_(True)
            z = se(p_leaf.view(B, C, 1, 1).to(u.dtype if not torch.is_autocast_enabled() else torch.float32))
            a = torch.softmax(z.float().view(B, C, 2), dim=2)
        a_c = a.detach().contiguous()
        out = torch.empty_like(u, memory_format=torch.channels_last)
        _lib.check(lib.cotb200_tail_combine(dt, B, HW, C, u.data_ptr(), k.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                            a_c.data_ptr(), out.data_ptr(), st), "tail_combine")
        ctx.save_for_backward(u, k, scale, shift, mean, rstd, a_c)
        ctx.graph = (p_leaf, a, [p for p in se_params])
        ctx.training = training
        ctx.bn_dtypes = (bn_weight.dtype, bn_bias.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        u, k, scale, shift, mean, rstd, a_c = ctx.saved_tensors
        p_leaf, a, se_params = ctx.graph
        B, C, H, W = u.shape
        HW, n = H * W, float(B * H * W)
        lib, st, dt = _lib.load(), _lib.stream_ptr(u), _lib.dtype_code(u)
        dout = dout.contiguous(memory_format=torch.channels_last)
        S = torch.zeros(B, C, 2, dtype=torch.float32, device=u.device)
        _lib.check(lib.cotb200_tail_bwd_sums(dt, B, HW, C, dout.data_ptr(), u.data_ptr(), k.data_ptr(), scale.data_ptr(),
                                             shift.data_ptr(), S.data_ptr(), st), "tail_bwd_sums")
        grads = torch.autograd.grad(a, [p_leaf] + se_params, grad_outputs=S, allow_unused=True)
        dpn = (grads[0] / HW).contiguous()
        se_grads = [None if g is None else g for g in grads[1:]]
        sums = torch.zeros(2, C, dtype=torch.float32, device=u.device)
        need_param = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        if ctx.training or need_param:
            _lib.check(lib.cotb200_tail_bwd_dz_sums(dt, B, HW, C, dout.data_ptr(), u.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                    mean.data_ptr(), rstd.data_ptr(), a_c.data_ptr(), dpn.data_ptr(),
                                                    sums[0].data_ptr(), sums[1].data_ptr(), st), "tail_bwd_dz_sums")
        c1 = c2 = None
        if ctx.training:
            c1, c2 = (sums[0] / n).contiguous(), (sums[1] / n).contiguous()
        du = torch.empty_like(u, memory_format=torch.channels_last)
        dk = torch.empty_like(u, memory_format=torch.channels_last)
        _lib.check(lib.cotb200_tail_bwd_apply(dt, B, HW, C, dout.data_ptr(), u.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                              mean.data_ptr(), rstd.data_ptr(), a_c.data_ptr(), dpn.data_ptr(), _lib.ptr(c1),
                                              _lib.ptr(c2), du.data_ptr(), dk.data_ptr(), st), "tail_bwd_apply")
        dgamma = sums[1].to(ctx.bn_dtypes[0]) if ctx.needs_input_grad[2] else None
        dbeta = sums[0].to(ctx.bn_dtypes[1]) if ctx.needs_input_grad[3] else None
        ctx.graph = None
        return (du, dk, dgamma, dbeta, None, None) + tuple(se_grads)


def group_norm9(l, gn: torch.nn.GroupNorm, gc=0):
    return GroupNorm9Fn.apply(l, gn.weight, gn.bias, gn.eps, gc)


def cot_tail(u, k, bn: torch.nn.BatchNorm2d, se: torch.nn.Module):
    params = [p for p in se.parameters()]
    return CotTailFn.apply(u, k, bn.weight, bn.bias, bn, se, *params)
