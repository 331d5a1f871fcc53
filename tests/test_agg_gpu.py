"""GPU parity tests of the LocalConv op against the CPU oracle (call through the C ABI via the op mirror).

Tolerances (BASELINE.json north_star / SURVEY.md section 8d): fp64 1e-9 (the reference self-tests' gate),
fp32 atol=rtol=1e-3 (observed ~1e-6), bf16/fp16 atol=rtol=1e-2 on bf16/fp16-representable inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import agg_ref

pytestmark = pytest.mark.gpu

TOL = {torch.float64: (1e-9, 1e-9), torch.float32: (1e-3, 1e-3), torch.bfloat16: (1e-2, 1e-2), torch.float16: (1e-2, 1e-2)}


def _ops():
    import cotnet_b200
    return cotnet_b200


def _close(got, want, dtype, what=""):
    atol, rtol = TOL[dtype]
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    err = (got - want).abs()
    ok = bool((err <= atol + rtol * want.abs()).all())
    assert ok, "%s: max abs err %.3e (max |ref| %.3e)" % (what, err.max().item(), want.abs().max().item())


def _rand(shape, dtype, gen):
    t = torch.randn(*shape, generator=gen, dtype=torch.float64)
    return t.to(dtype).double()       # representable in `dtype`, kept in fp64 for the oracle


def _run_case(N, C, wc, H, W, k, s, p, d, heads, dtype, channels_last=False, seed=0, check_bwd=True):
    gen = torch.Generator().manual_seed(seed)
    Ho, Wo = agg_ref.out_size(H, W, k, s, p, d)
    x64 = _rand((N, C, H, W), dtype, gen)
    w64 = _rand((N, heads, wc, k * k, Ho, Wo), dtype, gen)
    cot64 = _rand((N, heads * C, Ho, Wo), dtype, gen)
    x = x64.to(dtype).cuda()
    w = w64.to(dtype).cuda()
    cot = cot64.to(dtype).cuda()
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
        w = w.permute(0, 4, 5, 1, 2, 3).contiguous().permute(0, 3, 4, 5, 1, 2)
        cot = cot.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y = _ops().aggregation_zeropad(x, w, k, s, p, d)
    assert y.shape == (N, heads * C, Ho, Wo) and y.dtype == dtype
    if channels_last and C > 1:
        assert y.is_contiguous(memory_format=torch.channels_last)
    xr = x64.clone().requires_grad_(True)
    wr = w64.clone().requires_grad_(True)
    yr = agg_ref.agg_zeropad_unfold(xr, wr, k, s, p, d)
    _close(y, yr, dtype, "forward")
    if check_bwd:
        gx, gw = torch.autograd.grad(y, (x, w), cot)
        gxr, gwr = torch.autograd.grad(yr, (xr, wr), cot64)
        assert gx.shape == x.shape and gw.shape == w.shape
        _close(gx, gxr, dtype, "dX")
        _close(gw, gwr, dtype, "dW")


# config 1 of BASELINE.json: B=2 C=64 H=W=32 k=3 (the numerics gate)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cl", [False, True])
def test_config1_gate(dtype, cl):
    _run_case(2, 64, 8, 32, 32, 3, 1, 1, 1, 1, dtype, channels_last=cl)


# the reference self-test shapes (aggregation_zeropad.py:238-246, :266-274): k=5/k=1, heads=2, 9x9, fp64
@pytest.mark.parametrize("k", [5, 1])
@pytest.mark.parametrize("cl", [False, True])
def test_reference_selftest_shapes_fp64(k, cl):
    _run_case(2, 8, 4, 9, 9, k, 1, (k - 1 + 1) // 2 if k > 1 else 0, 1, 2, torch.float64, channels_last=cl)


@pytest.mark.parametrize("name", ["agg_selftest_k5.npz", "agg_selftest_k1.npz", "agg_cot_k3.npz"])
def test_golden_fixtures(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    k, s, p, d, heads = [int(v) for v in g["meta"]]
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    w = torch.from_numpy(g["w"]).cuda().requires_grad_(True)
    y = _ops().aggregation_zeropad(x, w, k, s, p, d)
    gx, gw = torch.autograd.grad(y, (x, w), torch.from_numpy(g["cot"]).cuda())
    assert (y.cpu() - torch.from_numpy(g["y"])).abs().max() < 1e-9
    assert (gx.cpu() - torch.from_numpy(g["gx"])).abs().max() < 1e-9
    assert (gw.cpu() - torch.from_numpy(g["gw"])).abs().max() < 1e-9


# CoTNet-50 stage shapes at a small batch, CoXt fold shapes (wc = 6/12), odd batch
@pytest.mark.parametrize("C,wc,HW", [(64, 8, 56), (128, 16, 28), (256, 32, 14), (512, 64, 7), (48, 6, 56), (96, 12, 28),
                                     (192, 24, 14), (384, 48, 7)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cl", [False, True])
def test_stage_shapes(C, wc, HW, dtype, cl):
    _run_case(3, C, wc, HW, HW, 3, 1, 1, 1, 1, dtype, channels_last=cl, seed=C + HW)


# generic path: stride / dilation / heads / rectangular maps / k=7 (san_lowrank uses k=7)
@pytest.mark.parametrize("k,s,p,d,heads", [(3, 2, 1, 1, 1), (3, 1, 2, 2, 2), (5, 2, 2, 1, 1), (7, 1, 3, 1, 1), (3, 1, 0, 1, 1),
                                           (3, 1, 1, 1, 2), (5, 1, 2, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("cl", [False, True])
def test_generic_configs(k, s, p, d, heads, dtype, cl):
    _run_case(2, 12, 4, 11, 9, k, s, p, d, heads, dtype, channels_last=cl, seed=k * 10 + s)


def test_needs_input_grad_variants_and_noncontiguous():
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 10, 10, generator=gen).cuda()
    w = torch.randn(2, 1, 4, 9, 10, 10, generator=gen).cuda()
    ops = _ops()
    xr, wr = x.double().cpu().requires_grad_(True), w.double().cpu().requires_grad_(True)
    yr = agg_ref.agg_zeropad_unfold(xr, wr, 3, 1, 1, 1)
    gxr, gwr = torch.autograd.grad(yr.sum(), (xr, wr))
    xg = x.clone().requires_grad_(True)
    y = ops.aggregation_zeropad(xg, w, 3, 1, 1, 1)
    (gx,) = torch.autograd.grad(y.sum(), (xg,))
    _close(gx, gxr, torch.float32, "dX only")
    wg = w.clone().requires_grad_(True)
    y = ops.aggregation_zeropad(x, wg, 3, 1, 1, 1)
    (gw,) = torch.autograd.grad(y.sum(), (wg,))
    _close(gw, gwr, torch.float32, "dW only")
    # non-contiguous input (sliced): must behave like .contiguous()
    xbig = torch.randn(2, 32, 10, 10, generator=gen).cuda()
    xs = xbig[:, ::2]
    y1 = ops.aggregation_zeropad(xs, w, 3, 1, 1, 1)
    y2 = ops.aggregation_zeropad(xs.contiguous(), w, 3, 1, 1, 1)
    assert torch.equal(y1, y2)


def test_gradcheck_fp64():
    from functools import partial
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 5, 5, generator=gen, dtype=torch.float64).cuda().requires_grad_(True)
    w = torch.randn(2, 1, 2, 9, 5, 5, generator=gen, dtype=torch.float64).cuda().requires_grad_(True)
    assert torch.autograd.gradcheck(partial(_ops().aggregation_zeropad, kernel_size=3, stride=1, padding=1, dilation=1), (x, w))


def test_empty_batch_and_module_api():
    ops = _ops()
    m = ops.LocalConvolution(8, 8, kernel_size=3, stride=1, padding=1, dilation=1)
    assert m.kernel_size == 3 and m.in_channels == 8
    x = torch.zeros(0, 8, 6, 6).cuda()
    w = torch.zeros(0, 1, 4, 9, 6, 6).cuda()
    assert m(x, w).shape == (0, 8, 6, 6)
    with pytest.raises(AssertionError):
        ops.aggregation_zeropad(torch.zeros(2, 8, 6, 6).cuda(), torch.zeros(2, 1, 3, 9, 6, 6).cuda(), 3, 1, 1, 1)
    with pytest.raises(AssertionError):   # Ho*Wo != weight H*W (aggregation_zeropad.py:122)
        ops.aggregation_zeropad(torch.zeros(2, 8, 6, 6).cuda(), torch.zeros(2, 1, 4, 9, 5, 6).cuda(), 3, 1, 1, 1)


def test_cpu_tensor_bounce():
    """aggregation_zeropad.py:192-196: CPU tensors are bounced through the GPU and come back on the CPU."""
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(1, 8, 6, 6, generator=gen)
    w = torch.randn(1, 1, 4, 9, 6, 6, generator=gen)
    y = _ops().aggregation_zeropad(x, w, 3, 1, 1, 1)
    assert not y.is_cuda
    _close(y, agg_ref.agg_zeropad_unfold(x.double(), w.double(), 3, 1, 1, 1), torch.float32)


# ---- mix op (aggregation_zeropad_mix.py:344-383 self-test shapes + a larger one)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 8, 4, 6, 6), (2, 32, 4, 14, 13)])
def test_mix(dtype, shape):
    N, C, wc, H, W = shape
    gen = torch.Generator().manual_seed(21)
    x64, w164, w264 = _rand((N, C, H, W), dtype, gen), _rand((N, 1, wc, 9, H, W), dtype, gen), _rand((N, 1, wc, 25, H, W), dtype, gen)
    cot64 = _rand((N, 2 * C, H, W), dtype, gen)
    x, w1, w2 = [t.to(dtype).cuda().requires_grad_(True) for t in (x64, w164, w264)]
    y = _ops().aggregation_zeropad_mix(x, w1, w2, 3, 5, 1, 1, 2, 1)
    xr, w1r, w2r = [t.clone().requires_grad_(True) for t in (x64, w164, w264)]
    yr = agg_ref.agg_zeropad_mix_unfold(xr, w1r, w2r, 3, 5, 1, 1, 2, 1)
    _close(y, yr, dtype, "mix fwd")
    g = torch.autograd.grad(y, (x, w1, w2), cot64.to(dtype).cuda())
    gr = torch.autograd.grad(yr, (xr, w1r, w2r), cot64)
    for a, b, n in zip(g, gr, ("dX", "dW1", "dW2")):
        _close(a, b, dtype, "mix " + n)


def test_mix_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "agg_mix_selftest.npz"))
    x, w1, w2 = [torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("x", "w1", "w2")]
    y = _ops().aggregation_zeropad_mix(x, w1, w2, 3, 5, 1, 1, 2, 1)
    gx, g1, g2 = torch.autograd.grad(y, (x, w1, w2), torch.from_numpy(g["cot"]).cuda())
    for a, k in ((y, "y"), (gx, "gx"), (g1, "gw1"), (g2, "gw2")):
        assert (a.cpu() - torch.from_numpy(g[k])).abs().max() < 1e-9, k


# ---- BASELINE full size (bs256 stage-1 shape): size-independent properties + exact check on a slice
def test_full_size_properties():
    ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    N, C, wc, H, W = 256, 64, 8, 56, 56
    x = torch.randn(N, C, H, W, generator=gen, device="cuda")
    w1 = torch.randn(N, 1, wc, 9, H, W, generator=gen, device="cuda")
    w2 = torch.randn(N, 1, wc, 9, H, W, generator=gen, device="cuda")
    y1 = ops.aggregation_zeropad(x, w1, 3, 1, 1, 1)
    y2 = ops.aggregation_zeropad(x, w2, 3, 1, 1, 1)
    y12 = ops.aggregation_zeropad(x, w1 + w2, 3, 1, 1, 1)
    assert torch.allclose(y12, y1 + y2, atol=1e-4, rtol=1e-4)                     # linearity in w
    ycl = ops.aggregation_zeropad(x.contiguous(memory_format=torch.channels_last),
                                  w1.permute(0, 4, 5, 1, 2, 3).contiguous().permute(0, 3, 4, 5, 1, 2), 3, 1, 1, 1)
    assert torch.allclose(ycl, y1, atol=1e-5, rtol=1e-5)                          # NHWC kernel == NCHW kernel
    # identity weights (centre tap = 1) reproduce x exactly
    wi = torch.zeros_like(w1)
    wi[:, :, :, 4] = 1
    assert torch.equal(ops.aggregation_zeropad(x, wi, 3, 1, 1, 1), x)
    # exact check of the first and last samples against the oracle
    for n in (0, N - 1):
        yr = agg_ref.agg_zeropad_unfold(x[n:n + 1].double().cpu(), w1[n:n + 1].double().cpu(), 3, 1, 1, 1)
        _close(y1[n:n + 1], yr, torch.float32, "sample %d" % n)
    # adjoint identity  <agg(x,w), g> == <x, dX> == <w, dW>
    g = torch.randn(N, C, H, W, generator=gen, device="cuda")
    xg, wg = x.clone().requires_grad_(True), w1.clone().requires_grad_(True)
    y = ops.aggregation_zeropad(xg, wg, 3, 1, 1, 1)
    gx, gw = torch.autograd.grad(y, (xg, wg), g)
    lhs = (y.double() * g.double()).sum()
    assert abs((x.double() * gx.double()).sum() - lhs) <= 1e-6 * abs(lhs) + 1e-3
    assert abs((w1.double() * gw.double()).sum() - lhs) <= 1e-6 * abs(lhs) + 1e-3
