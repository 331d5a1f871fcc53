"""GPU tests of the fused normalisation / tail autograd functions against eager PyTorch on the same inputs."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("wc,H", [(8, 14), (16, 9), (12, 7), (64, 7)])
def test_groupnorm9(dtype, tol, wc, H):
    from cotnet_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(wc + H)
    B, J = 5, 9 * wc
    gn = nn.GroupNorm(wc, J).cuda()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5, generator=g)
        gn.bias.normal_(0, 0.3, generator=g)
    l = _cl((torch.randn(B, J, H, H, generator=g, device="cuda") * 2 + 0.5).to(dtype)).requires_grad_(True)
    cot = _cl(torch.randn(B, J, H, H, generator=g, device="cuda").to(dtype))
    out = fused.group_norm9(l, gn)
    gl, gw, gb = torch.autograd.grad(out, (l, gn.weight, gn.bias), cot)
    lr = l.detach().double().requires_grad_(True)
    gnr = nn.GroupNorm(wc, J).cuda().double()
    gnr.load_state_dict(gn.state_dict())
    ref = gnr(lr)
    rl, rw, rb = torch.autograd.grad(ref, (lr, gnr.weight, gnr.bias), cot.double())
    for a, b, name in ((out, ref, "out"), (gl, rl, "dl"), (gw, rw, "dgamma"), (gb, rb, "dbeta")):
        err = (a.double() - b).abs().max().item()
        scale = max(1.0, b.abs().max().item())
        assert err <= tol * scale, "%s err %.3e scale %.3e" % (name, err, scale)
    assert out.is_contiguous(memory_format=torch.channels_last) and out.dtype == dtype


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-4), (torch.bfloat16, 4e-2)])
@pytest.mark.parametrize("C,H", [(64, 14), (128, 7), (48, 9), (512, 7)])
@pytest.mark.parametrize("training", [False, True])
def test_cot_tail(dtype, tol, C, H, training):
    from cotnet_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(C + H)
    B, A = 16, max(C // 2, 32)
    bn = nn.BatchNorm2d(C).cuda()
    se = nn.Sequential(nn.Conv2d(C, A, 1), nn.BatchNorm2d(A), nn.ReLU(inplace=True), nn.Conv2d(A, 2 * C, 1)).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(0, 0.3, generator=g)
        bn.running_mean.normal_(0, 0.3, generator=g); bn.running_var.uniform_(0.5, 2, generator=g)
        se[1].running_mean.normal_(0, 0.1, generator=g); se[1].running_var.uniform_(0.5, 2, generator=g)
    import copy
    bn_r, se_r = copy.deepcopy(bn).double(), copy.deepcopy(se).double()
    for m in (bn, se, bn_r, se_r):
        m.train(training)
    u = _cl(torch.randn(B, C, H, H, generator=g, device="cuda").to(dtype)).requires_grad_(True)
    k = _cl(torch.relu(torch.randn(B, C, H, H, generator=g, device="cuda")).to(dtype)).requires_grad_(True)
    cot = _cl(torch.randn(B, C, H, H, generator=g, device="cuda").to(dtype))
    out = fused.cot_tail(u, k, bn, se)
    params = [bn.weight, bn.bias] + list(se.parameters())
    grads = torch.autograd.grad(out, [u, k] + params, cot)
    # eager reference in fp64 (models/cotnet.py:89-104)
    ur, kr = u.detach().double().requires_grad_(True), k.detach().double().requires_grad_(True)
    y = F.silu(bn_r(ur))
    gap = (y + kr).mean((2, 3), keepdim=True)
    a = torch.softmax(se_r(gap).view(B, C, 2), 2)
    ref = y * a[:, :, 0].reshape(B, C, 1, 1) + kr * a[:, :, 1].reshape(B, C, 1, 1)
    params_r = [bn_r.weight, bn_r.bias] + list(se_r.parameters())
    grads_r = torch.autograd.grad(ref, [ur, kr] + params_r, cot.double())
    err = (out.double() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), "out err %.3e" % err
    for i, (a_, b_) in enumerate(zip(grads, grads_r)):
        e = (a_.double() - b_).abs().max().item()
        s = max(1.0, b_.abs().max().item())
        assert e <= 4 * tol * s, "grad %d err %.3e scale %.3e" % (i, e, s)
    if training:
        assert torch.allclose(bn.running_mean.double(), bn_r.running_mean, atol=1e-2 if dtype != torch.float32 else 1e-5)
        assert torch.allclose(bn.running_var.double(), bn_r.running_var, atol=1e-2 if dtype != torch.float32 else 1e-4)
        assert int(bn.num_batches_tracked) == 1
