"""CPU check of the closed-form backward formulas implemented by csrc/norm_tail.cu (GroupNorm-of-9-taps and the
radix-2 tail), restated here in torch fp64 and compared with autograd.  Guards the kernel maths before GPU time is
spent; the kernels themselves are checked against the oracle in the -m gpu tests."""
import torch


def test_groupnorm9_backward_formula():
    torch.manual_seed(0)
    B, wc, HW = 3, 4, 10
    J = 9 * wc
    l = torch.randn(B, HW, J, dtype=torch.float64, requires_grad=True)
    gamma = torch.rand(J, dtype=torch.float64) + 0.5
    beta = torch.randn(J, dtype=torch.float64)
    gamma.requires_grad_(True); beta.requires_grad_(True)
    eps = 1e-5
    # reference: nn.GroupNorm semantics on [B, J, HW]
    ref = torch.nn.functional.group_norm(l.permute(0, 2, 1), wc, gamma, beta, eps).permute(0, 2, 1)
    dg = torch.randn_like(ref)
    gl, gg, gb = torch.autograd.grad(ref, (l, gamma, beta), dg)
    # kernel maths (gn_stats / gn_bwd_sums / gn_bwd_apply)
    with torch.no_grad():
        n = 9.0 * HW
        lg = l.view(B, HW, wc, 9)
        mean = lg.sum((1, 3)) / n
        var = (lg * lg).sum((1, 3)) / n - mean * mean
        rstd = torch.rsqrt(var + eps)
        lhat = (lg - mean[:, None, :, None]) * rstd[:, None, :, None]
        dgg = dg.view(B, HW, wc, 9)
        gam = gamma.view(wc, 9)
        s1 = (dgg * gam).sum((1, 3))
        s2 = (dgg * gam * lhat).sum((1, 3))
        dl = rstd[:, None, :, None] * (dgg * gam - (s1 / n)[:, None, :, None] - lhat * (s2 / n)[:, None, :, None])
        dgamma = (dgg * lhat).sum((0, 1)).view(J)
        dbeta = dgg.sum((0, 1)).view(J)
    assert (dl.reshape(B, HW, J) - gl).abs().max() < 1e-10
    assert (dgamma - gg).abs().max() < 1e-10 and (dbeta - gb).abs().max() < 1e-10


def test_tail_backward_formula():
    torch.manual_seed(1)
    B, C, HW = 4, 6, 7
    A = 5
    u = torch.randn(B, HW, C, dtype=torch.float64, requires_grad=True)
    k = torch.relu(torch.randn(B, HW, C, dtype=torch.float64)).requires_grad_(True)
    gw = (torch.rand(C, dtype=torch.float64) + 0.5).requires_grad_(True)
    gb = torch.randn(C, dtype=torch.float64).requires_grad_(True)
    W1 = torch.randn(A, C, dtype=torch.float64, requires_grad=True)
    W2 = torch.randn(2 * C, A, dtype=torch.float64, requires_grad=True)
    eps = 1e-5

    def se(p):
        return torch.relu(p @ W1.t()) @ W2.t()

    # reference graph (training-mode BN over B*HW rows)
    mean = u.mean((0, 1)); var = u.var((0, 1), unbiased=False)
    z = (u - mean) / torch.sqrt(var + eps) * gw + gb
    y = z * torch.sigmoid(z)
    p = (y + k).mean(1)
    a = torch.softmax(se(p).view(B, C, 2), 2)
    out = y * a[:, None, :, 0] + k * a[:, None, :, 1]
    dout = torch.randn_like(out)
    gu, gk, ggw, ggb = torch.autograd.grad(out, (u, k, gw, gb), dout, retain_graph=True)

    # kernel maths
    with torch.no_grad():
        n = float(B * HW)
        mu = u.sum((0, 1)) / n
        va = (u * u).sum((0, 1)) / n - mu * mu
        rstd = torch.rsqrt(va + eps)
        scale = gw * rstd; shift = gb - mu * scale
        zz = u * scale + shift
        sg = torch.sigmoid(zz)
        yy = zz * sg
        S0 = (dout * yy).sum(1); S1 = (dout * k).sum(1)          # tail_bwd_sums
    S = torch.stack([S0, S1], 2)
    p_leaf = ((yy + k.detach()).sum(1) / HW).requires_grad_(True)
    a2 = torch.softmax(se(p_leaf).view(B, C, 2), 2)
    (dp,) = torch.autograd.grad(a2, (p_leaf,), S)
    with torch.no_grad():
        dpn = dp / HW
        a0, a1 = a2[:, :, 0].detach(), a2[:, :, 1].detach()
        dz = (a0[:, None] * dout + dpn[:, None]) * (sg * (1 + zz * (1 - sg)))
        xhat = (u - mu) * rstd
        sum_dz = dz.sum((0, 1)); sum_dzx = (dz * xhat).sum((0, 1))
        du = scale * (dz - sum_dz / n - xhat * sum_dzx / n)
        dk = a1[:, None] * dout + dpn[:, None]
    assert (du - gu).abs().max() < 1e-10
    assert (dk - gk).abs().max() < 1e-10
    assert (sum_dzx - ggw).abs().max() < 1e-10 and (sum_dz - ggb).abs().max() < 1e-10


def test_zero_arena_hands_out_clean_disjoint_slices():
    """fused._ZeroArena (accumulator scratch): slices are zero, 16-byte aligned, disjoint within a step, recycled (and
    cleared) by step_begin(), and the allocator falls back to torch.zeros when the arena is off or exhausted."""
    import torch
    from cotnet_b200 import fused
    dev = torch.device("cpu")
    fused._ARENA.buf.pop(dev, None)
    fused.step_arena_off()
    a = fused._zeros((2, 5), dev)                      # arena off: plain zeros
    assert a.shape == (2, 5) and float(a.abs().sum()) == 0.0
    old = fused._ZeroArena.SIZE
    fused._ZeroArena.SIZE = 64
    try:
        fused.step_begin(dev)
        x = fused._zeros((3, 3), dev)
        y = fused._zeros((7,), dev)
        assert x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0
        assert y.data_ptr() >= x.data_ptr() + 9 * 4                      # disjoint
        x.add_(1.0); y.add_(2.0)
        big = fused._zeros((100,), dev)                                  # does not fit: falls back, still zero
        assert float(big.abs().sum()) == 0.0 and big.numel() == 100
        fused.step_begin(dev)                                            # recycle: the used prefix is cleared
        x2 = fused._zeros((3, 3), dev)
        assert x2.data_ptr() == x.data_ptr() and float(x2.abs().sum()) == 0.0
    finally:
        fused._ZeroArena.SIZE = old
        fused._ARENA.buf.pop(dev, None)
        fused.step_arena_off()


def test_row_view_detects_pitched_nhwc_rows():
    """fused._row_view: the pixel pitch of a [B,C,H,W] tensor whose memory is NHWC rows (channels_last tensors and channel
    slices of them -- the gradients of a concat), None for anything the row kernels cannot address."""
    import torch
    from cotnet_b200 import fused
    t = torch.zeros(2, 12, 5, 7).contiguous(memory_format=torch.channels_last)
    assert fused._row_view(t) == 12
    assert fused._row_view(t[:, :4]) == 12 and fused._row_view(t[:, 4:]) == 12        # channel slices keep the pitch
    assert fused._row_view(torch.zeros(2, 12, 5, 7)) is None                            # NCHW: channel stride != 1
    assert fused._row_view(t[:, :, ::2]) is None                                        # strided rows
    assert fused._row_view(t[:, ::2]) is None                                           # strided channels
    assert fused._row_view(torch.zeros(3, 4)) is None
    one = torch.zeros(1, 8, 1, 1).contiguous(memory_format=torch.channels_last)
    assert fused._row_view(one) == 8


def test_tap_chunk_and_positions():
    """Tap-major layout bookkeeping used by GroupNorm / LocalConv: chunk width and the (g, t) -> position map of
    include/cotb200.h (COTB200_NHWC_TAP)."""
    from cotnet_b200 import fused
    assert fused.tap_chunk(8) == 8 and fused.tap_chunk(64) == 8
    gc = 8
    wc = 16
    seen = set()
    for g in range(wc):
        for t in range(9):
            pos = ((g // gc) * 9 + t) * gc + g % gc
            assert 0 <= pos < 9 * wc
            seen.add(pos)
            # each chunk of 8 groups occupies the same 72 positions in both orders (what csrc/gn72.cu relies on)
            assert pos // 72 == (g * 9 + t) // 72
    assert len(seen) == 9 * wc


def test_stem_space_to_depth_weight_packing():
    """conv 7x7 / stride 2 / pad 3 on 3 channels == 4 row-taps x (4 cells x 16 channels) contraction over the space-to-depth image,
    with the weight packed by cotnet_b200.tc.prepare_stem_weight -- the arithmetic of cotb200_stem7x7s2_bf16 (csrc/tc_gemm.cu
    stem mode) restated in torch fp64 (models/resnet.py:552)."""
    from cotnet_b200 import tc
    torch.manual_seed(3)
    B, H, W, N = 2, 12, 20, 8
    x = torch.randn(B, 3, H, W, dtype=torch.float64)
    w = torch.randn(N, 3, 7, 7, dtype=torch.float64)
    want = torch.nn.functional.conv2d(x, w, None, 2, 3)
    Hh, Wh = H // 2, W // 2
    # P[b, i, jp, (di*2+dj)*3 + c] = x[b, c, 2i+di, 2(jp-2)+dj], two zero cells either side, channels 12..15 zero
    P = torch.zeros(B, Hh, Wh + 4, 16, dtype=torch.float64)
    xs = x.view(B, 3, Hh, 2, Wh, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Hh, Wh, 12)      # [b, i, j, (di, dj, c)]
    P[:, :, 2:Wh + 2, :12] = xs
    wm = tc.prepare_stem_weight(w.float()).double()        # bf16-rounded values
    wq = torch.zeros_like(w)
    # undo the packing to know which (rounded) weight the kernel multiplies with
    wm4 = wm.view(N, 4, 4, 16)[..., :12].reshape(N, 4, 4, 2, 2, 3)
    for a in range(4):
        for a2 in range(4):
            for di in range(2):
                for dj in range(2):
                    kh, kw = 2 * a + di - 1, 2 * a2 + dj - 1
                    if kh >= 0 and kw >= 0:
                        wq[:, :, kh, kw] = wm4[:, a, a2, di, dj, :]
    assert (wq - w).abs().max() <= 2 ** -8 * w.abs().max()          # only bf16 rounding
    want_q = torch.nn.functional.conv2d(x, wq, None, 2, 3)
    # implicit GEMM: out[b, oh, ow, n] = sum_a  window(P[b, oh-2+a, ow .. ow+3, :]) . wm[n, a*64 : (a+1)*64]
    Pp = torch.nn.functional.pad(P, (0, 0, 0, 0, 2, 2))              # rows oh-2+a in [-2, Hh+1] -> zero rows (TMA out-of-bounds fill)
    out = torch.zeros(B, Hh, Wh, N, dtype=torch.float64)
    for a in range(4):
        rows = Pp[:, a:a + Hh]                                      # [B, Hh, Wh+4, 16]
        win = torch.stack([rows[:, :, k:k + Wh] for k in range(4)], dim=3).reshape(B, Hh, Wh, 64)
        out += win @ wm[:, a * 64:(a + 1) * 64].t()
    assert (out.permute(0, 3, 1, 2) - want_q).abs().max() < 1e-9
    assert (out.permute(0, 3, 1, 2) - want).abs().max() < 0.2


def test_stem_weight_grad_unpack_is_inverse_of_packing():
    from cotnet_b200 import tc
    torch.manual_seed(4)
    w = torch.randn(8, 3, 7, 7).bfloat16().float()
    assert torch.equal(tc.unpack_stem_weight_grad(tc.prepare_stem_weight(w).float()), w)


def test_deferred_batchnorm_counters_bump_once_per_step():
    """TrainStep bookkeeping (cotnet_b200/fused.py): with defer_bn_counters() the `num_batches_tracked += 1` of every BatchNorm is
    recorded and applied by ONE multi-tensor add in flush_bn_counters(); without it (or for momentum=None modules, whose momentum
    depends on the counter) the counter moves immediately -- nn.BatchNorm2d semantics either way."""
    from cotnet_b200 import fused
    bns = [torch.nn.BatchNorm2d(4) for _ in range(3)] + [torch.nn.BatchNorm2d(4, momentum=None)]
    fused.defer_bn_counters(True)
    try:
        for bn in bns:
            fused._bump_counter(bn)
        assert [int(b.num_batches_tracked) for b in bns] == [0, 0, 0, 1]        # momentum=None is never deferred
        fused.flush_bn_counters()
        assert [int(b.num_batches_tracked) for b in bns] == [1, 1, 1, 1]
        fused.flush_bn_counters()                                                # idempotent: the list was cleared
        assert [int(b.num_batches_tracked) for b in bns] == [1, 1, 1, 1]
    finally:
        fused.defer_bn_counters(False)
    fused._bump_counter(bns[0])
    assert int(bns[0].num_batches_tracked) == 2


def test_coxt_grouped_convs_as_dense_block_diagonal():
    """CoXtLayer's fast path runs its grouped convolutions as dense convolutions with block-diagonal weights built from the grouped
    parameters (cot_layer._dense_from_grouped) and embed.0 -- which consumes the channel-INTERLEAVED qk = [x0,k0,x1,k1,...]
    (models/cotnet.py:153-154) with groups=2 -- as  x @ Wx^T + k @ Wk^T  (cot_layer._coxt_embed0_dense).  Same arithmetic, and the
    gradient of the grouped parameter is the blocks of the dense weight gradient: checked against F.conv2d(groups=...) in fp64."""
    from cotnet_b200.cot_layer import _coxt_embed0_dense, _dense_from_grouped
    torch.manual_seed(5)
    F = torch.nn.functional
    B, C, H = 2, 48, 5
    x = torch.randn(B, C, H, H, dtype=torch.float64)
    k = torch.randn(B, C, H, H, dtype=torch.float64)
    # 3x3, groups 8 (key_embed) and 1x1, groups 2 (embed.3 / conv1x1)
    for groups, ks, cout in ((8, 3, C), (2, 1, 54), (2, 1, C)):
        w = torch.randn(cout, C // groups, ks, ks, dtype=torch.float64, requires_grad=True)
        want = F.conv2d(x, w, None, 1, ks // 2, 1, groups)
        got = F.conv2d(x, _dense_from_grouped(w, groups), None, 1, ks // 2)
        assert (got - want).abs().max() < 1e-10
        g = torch.randn_like(want)
        gw_want, = torch.autograd.grad(want, w, g, retain_graph=True)
        gw_got, = torch.autograd.grad(got, w, g)
        assert (gw_got - gw_want).abs().max() < 1e-9
    # embed.0 on the interleaved concat
    G = 2
    w0 = torch.randn(C // 2, 2 * C // G, 1, 1, dtype=torch.float64)
    qk = torch.stack([x, k], dim=2).reshape(B, 2 * C, H, H)
    want = F.conv2d(qk, w0, None, 1, 0, 1, G)
    wx, wk = _coxt_embed0_dense(w0, G)
    got = torch.einsum("bchw,nc->bnhw", x, wx) + torch.einsum("bchw,nc->bnhw", k, wk)
    assert (got - want).abs().max() < 1e-10


def test_haloed_tile_conv_index_math():
    """Index math of tc_conv3x3_halo_kernel (csrc/tc_gemm.cu), restated with torch: ONE haloed tile {W+2, R+2} per work item, output
    pixels enumerated in PADDED coordinates q = r*(W+2) + c, tap (dh, dw) reads tile row q + (dh+1)*(W+2) + (dw+1); outputs with
    c >= W are garbage columns that are dropped, and rows past the loaded box (NaN here, stale shared memory on the GPU) never reach a
    valid output.  == conv2d(3x3, pad 1)."""
    torch.manual_seed(0)
    F = torch.nn.functional
    B, C, H, W, R = 2, 4, 6, 5, 3
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    w = torch.randn(C, C, 3, 3, dtype=torch.float64)
    want = F.conv2d(x, w, None, 1, 1)
    Wp = W + 2
    xp = F.pad(x, (1, 1, 1, 1))                                   # the TMA out-of-bounds fill
    out = torch.full_like(want, float("nan"))
    for b in range(B):
        for h0 in range(0, H, R):
            tile = xp[b, :, h0:h0 + R + 2, :].permute(1, 2, 0).reshape((R + 2) * Wp, C)
            tile = torch.cat([tile, torch.full((2 * Wp + 4, C), float("nan"), dtype=torch.float64)])
            for q in range(R * Wp):
                ro, co = divmod(q, Wp)
                acc = sum(w[:, :, t // 3, t % 3] @ tile[q + (t // 3) * Wp + (t % 3)] for t in range(9))
                if co < W:                                         # the epilogue compacts the garbage columns away
                    out[b, :, h0 + ro, co] = acc
    assert (out - want).abs().max() < 1e-12


def test_stem_weight_gradient_in_packed_layout():
    """cotb200_stem7x7s2_wgrad_bf16 accumulates  dWm[n, a*64 + j] = sum_px dY[px, n] * window_a(px)[j]  over the space-to-depth image;
    cotnet_b200.tc.unpack_stem_weight_grad maps it back to [N, 3, 7, 7].  Restated in torch fp64 == conv2d weight gradient."""
    from cotnet_b200 import tc
    torch.manual_seed(6)
    F = torch.nn.functional
    B, H, W, N = 2, 8, 12, 8
    x = torch.randn(B, 3, H, W, dtype=torch.float64)
    dy = torch.randn(B, N, H // 2, W // 2, dtype=torch.float64)
    want = torch.nn.grad.conv2d_weight(x, (N, 3, 7, 7), dy, stride=2, padding=3)
    Hh, Wh = H // 2, W // 2
    P = torch.zeros(B, Hh, Wh + 4, 16, dtype=torch.float64)
    P[:, :, 2:Wh + 2, :12] = x.view(B, 3, Hh, 2, Wh, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Hh, Wh, 12)
    Pp = F.pad(P, (0, 0, 0, 0, 2, 2))                              # s2d rows -2, -1, Hh, Hh+1 are zero (TMA out-of-bounds fill)
    dwm = torch.zeros(N, 256, dtype=torch.float64)
    dyr = dy.permute(0, 2, 3, 1).reshape(-1, N)                    # [px, N]
    for a in range(4):
        rows = Pp[:, a:a + Hh]
        win = torch.stack([rows[:, :, k:k + Wh] for k in range(4)], dim=3).reshape(-1, 64)      # [px, 64]
        dwm[:, a * 64:(a + 1) * 64] = dyr.t() @ win
    got = tc.unpack_stem_weight_grad(dwm)
    assert got.shape == want.shape and (got - want).abs().max() < 1e-10
