"""tcgen05 conv / GEMM / wgrad kernels vs a plain fp32 PyTorch reference of the same op on the
same bf16-rounded operands (tolerance: fp32 accumulation-order noise only, 2e-3 of the output scale)."""
import pytest
import torch
import torch.nn.functional as F

from scene_generation_b200 import _lib, convspec, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def bf(x):
    return x.to(torch.bfloat16)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def to_nhwc5(x_nchw, cp=None):
    """(N,C,H,W) f32 -> bf16 (N,1,H,W,Cp) cuda"""
    N, C, H, W = x_nchw.shape
    cp = cp or ops.round_up(C, 8)
    out = torch.zeros(N, 1, H, W, cp, dtype=torch.bfloat16)
    out[:, 0, :, :, :C] = bf(x_nchw.permute(0, 2, 3, 1))
    return out.to(DEV)


def to_planes(x_nchw, cp=None):
    N, C, H, W = x_nchw.shape
    cp = cp or ops.round_up(C, 8)
    Hp, Wp = (H + 1) // 2, (W + 1) // 2
    out = torch.zeros(N, 4, Hp, Wp, cp, dtype=torch.bfloat16)
    xl = bf(x_nchw.permute(0, 2, 3, 1))
    for ph in range(2):
        for pw in range(2):
            sub = xl[:, ph::2, pw::2]
            out[:, ph * 2 + pw, :sub.shape[1], :sub.shape[2], :C] = sub
    return out.to(DEV)


def pack_w(w_oihw, cp=None):
    """(Cout,Cin,kh,kw) -> bf16 (Cout, kh*kw, Cp)"""
    Co, Ci, kh, kw = w_oihw.shape
    cp = cp or ops.round_up(Ci, 8)
    out = torch.zeros(Co, kh * kw, cp, dtype=torch.bfloat16)
    out[:, :, :Ci] = bf(w_oihw.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci))
    return out.to(DEV)


def r32(x):
    return bf(x).float()


def check(y, ref, tol=2e-3):
    y = y.float().cpu()
    scale = ref.abs().max().item()
    err = (y - ref).abs().max().item()
    assert err <= tol * max(scale, 1e-3), 'max err %.3e vs scale %.3e' % (err, scale)


@pytest.mark.parametrize('M,K,N', [(300, 454, 512), (128, 64, 64), (77, 512, 1152), (208, 512, 4), (1000, 1024, 172)])
def test_gemm(M, K, N):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.1), rnd(N, seed=3)
    Kp = ops.round_up(K, 8)
    x5 = torch.zeros(1, 1, 1, M, Kp, dtype=torch.bfloat16)
    x5[0, 0, 0, :, :K] = bf(x)
    w3 = torch.zeros(N, 1, Kp, dtype=torch.bfloat16)
    w3[:, 0, :K] = bf(w)
    y = torch.full((M, N), float('nan'), device=DEV)
    ops.conv_tc(x5.to(DEV), w3.to(DEV), y, (0, 0, N), 1, M, [(0, 0, 0, 0)], bias=b.to(DEV), act=_lib.ACT_RELU)
    ref = torch.relu(r32(x) @ r32(w).t() + b)
    check(y, ref)


@pytest.mark.parametrize('N,C,H,Co,k,out_bf16', [(3, 72, 8, 40, 3, False), (4, 128, 8, 256, 3, True),
                                                 (2, 24, 16, 3, 7, False), (5, 64, 4, 64, 3, True),
                                                 (1, 204, 32, 64, 7, True)])
def test_conv_s1_reflect_prepadded(N, C, H, Co, k, out_bf16):
    p = k // 2
    x, w, b = rnd(N, C, H, H, seed=4), rnd(Co, C, k, k, seed=5, scale=0.05), rnd(Co, seed=6)
    xp = F.pad(r32(x), (p, p, p, p), mode='reflect')
    ref = F.conv2d(xp, r32(w), b)
    taps, off = convspec.conv_s1(k, 0)
    y = torch.full((N, H, H, Co), float('nan'), device=DEV, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _, stats = ops.conv_tc(to_nhwc5(xp), pack_w(w), y, (H * H * Co, H * Co, Co), H, H, taps, bias=b.to(DEV), stats=True)
    check(y.permute(0, 3, 1, 2), ref, 1e-2 if out_bf16 else 2e-3)
    assert stats.shape[0] == N and stats.shape[2:] == (Co, 2)      # (N, slots, Cout, 2) partial sums
    check(stats.sum(1)[..., 0], ref.sum(dim=(2, 3)), 2e-3)
    check(stats.sum(1)[..., 1], (ref ** 2).sum(dim=(2, 3)), 2e-3)
    # fixed-order reductions: a second launch reproduces the statistics bit for bit
    y2 = torch.empty_like(y)
    _, stats2 = ops.conv_tc(to_nhwc5(xp), pack_w(w), y2, (H * H * Co, H * Co, Co), H, H, taps, bias=b.to(DEV), stats=True)
    assert torch.equal(stats, stats2) and torch.equal(y, y2)


@pytest.mark.parametrize('N,C,H,W,Co,k,p', [(2, 45, 13, 13, 64, 4, 2), (3, 64, 16, 16, 128, 3, 1), (2, 8, 32, 32, 64, 4, 0),
                                            (2, 207, 64, 64, 64, 4, 2)])
def test_conv_s2_planes(N, C, H, W, Co, k, p):
    x, w, b = rnd(N, C, H, W, seed=7), rnd(Co, C, k, k, seed=8, scale=0.05), rnd(Co, seed=9)
    ref = F.leaky_relu(F.conv2d(r32(x), r32(w), b, stride=2, padding=p), 0.2)
    Ho, Wo = ref.shape[2:]
    y = torch.full((N, Ho, Wo, Co), float('nan'), device=DEV)
    ops.conv_tc(to_planes(x), pack_w(w), y, (Ho * Wo * Co, Wo * Co, Co), Ho, Wo, convspec.conv_s2(k, p),
                bias=b.to(DEV), act=_lib.ACT_LEAKY, slope=0.2)
    check(y.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize('N,C,H,Co,k,p', [(2, 45, 9, 64, 4, 2), (2, 256, 8, 1, 3, 1), (3, 512, 10, 1, 4, 2)])
def test_conv_s1_zero_pad_oob(N, C, H, Co, k, p):
    x, w, b = rnd(N, C, H, H, seed=10), rnd(Co, C, k, k, seed=11, scale=0.05), rnd(Co, seed=12)
    ref = F.conv2d(r32(x), r32(w), b, padding=p)
    Ho = ref.shape[2]
    taps, off = convspec.conv_s1(k, p)
    y = torch.full((N, Ho, Ho, Co), float('nan'), device=DEV)
    ops.conv_tc(to_nhwc5(x), pack_w(w), y, (Ho * Ho * Co, Ho * Co, Co), Ho, Ho, taps, in_h0=off, in_w0=off, bias=b.to(DEV))
    check(y.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize('N,C,H,Co', [(2, 128, 8, 64), (3, 64, 16, 32), (1, 1024, 8, 512)])
def test_convT_phases(N, C, H, Co):
    x, w, b = rnd(N, C, H, H, seed=13), rnd(C, Co, 3, 3, seed=14, scale=0.05), rnd(Co, seed=15)
    ref = F.conv_transpose2d(r32(x), r32(w), b, stride=2, padding=1, output_padding=1)
    Ho = 2 * H
    taps, phases = convspec.convT_s2(3, 1)
    w3 = pack_w(w.permute(1, 0, 2, 3))          # (Cout, Cin, kh, kw) view of the ConvT weight
    y = torch.full((N, Ho, Ho, Co), float('nan'), device=DEV)
    _, stats = ops.conv_tc(to_nhwc5(x), w3, y, (Ho * Ho * Co, Ho * Co, Co), H, H, taps, phases=phases, oh_mul=2, ow_mul=2,
                           bias=b.to(DEV), stats=True)
    check(y.permute(0, 3, 1, 2), ref)
    assert stats.shape[1] % 4 == 0                 # one slot per (sub-pixel phase, M tile of the image)
    check(stats.sum(1)[..., 0], ref.sum(dim=(2, 3)), 2e-3)
    check(stats.sum(1)[..., 1], (ref ** 2).sum(dim=(2, 3)), 2e-3)


def test_dgrad_s1_matches_autograd():
    N, C, H, Co, k, p = 2, 64, 8, 128, 3, 1
    x, w = rnd(N, C, H, H, seed=16), rnd(Co, C, k, k, seed=17, scale=0.05)
    dy = rnd(N, Co, H, H, seed=18)
    xr = r32(x).requires_grad_(True)
    F.conv2d(xr, r32(w), padding=p).backward(r32(dy))
    wT = pack_w(w.permute(1, 0, 2, 3))          # [Cin][taps][Cout]
    dx = torch.full((N, H, H, C), float('nan'), device=DEV)
    ops.conv_tc(to_nhwc5(dy), wT, dx, (H * H * C, H * C, C), H, H, convspec.dgrad_s1(k, p))
    check(dx.permute(0, 3, 1, 2), xr.grad)


@pytest.mark.parametrize('N,C,H,Co,k,p', [(2, 64, 8, 128, 3, 1), (4, 128, 8, 256, 3, 0), (2, 204, 16, 64, 7, 0),
                                          (3, 40, 12, 72, 3, 1), (2, 64, 20, 64, 7, 3), (3, 8, 17, 64, 4, 2),
                                          (2, 24, 16, 200, 5, 2)])
def test_wgrad_s1(N, C, H, Co, k, p):
    # C <= 64: the taps are stacked in the N tile (3 or 4 per CTA; 49 and 25 taps leave a last group of one tap)
    x, w = rnd(N, C, H + 2 * (k // 2 - p), H + 2 * (k // 2 - p), seed=19), rnd(Co, C, k, k, seed=20, scale=0.05)
    wr = r32(w).requires_grad_(True)
    out = F.conv2d(r32(x), wr, padding=p)
    dy = rnd(*out.shape, seed=21)
    out.backward(r32(dy))
    Ho = out.shape[2]
    dw = torch.full((Co, k * k, C), float('nan'), device=DEV)      # overwritten, split-K included (no zero-fill needed)
    ops.wgrad_tc(to_nhwc5(dy), to_nhwc5(x), dw, Ho, Ho, convspec.wgrad_s1(k, p), Co, C)
    ref = wr.grad.permute(0, 2, 3, 1).reshape(Co, k * k, C)
    check(dw, ref)
    # the k-splits of a tile are added in split order: bit-reproducible, also with a forced deep split
    for ksplit in (0, 3):
        a = torch.full((Co, k * k, C), float('nan'), device=DEV)
        b = torch.full((Co, k * k, C), float('nan'), device=DEV)
        ops.wgrad_tc(to_nhwc5(dy), to_nhwc5(x), a, Ho, Ho, convspec.wgrad_s1(k, p), Co, C, ksplit=ksplit)
        ops.wgrad_tc(to_nhwc5(dy), to_nhwc5(x), b, Ho, Ho, convspec.wgrad_s1(k, p), Co, C, ksplit=ksplit)
        assert torch.equal(a, b)
        check(a, ref)


def test_wgrad_s2_and_convT():
    N, C, H, Co, k, p = 2, 64, 16, 128, 3, 1
    x, w = rnd(N, C, H, H, seed=22), rnd(Co, C, k, k, seed=23, scale=0.05)
    wr = r32(w).requires_grad_(True)
    out = F.conv2d(r32(x), wr, stride=2, padding=p)
    dy = rnd(*out.shape, seed=24)
    out.backward(r32(dy))
    dw = torch.zeros(Co, k * k, C, device=DEV)
    ops.wgrad_tc(to_nhwc5(dy), to_planes(x), dw, out.shape[2], out.shape[3], convspec.wgrad_s2(k, p), Co, C)
    check(dw, wr.grad.permute(0, 2, 3, 1).reshape(Co, k * k, C))
    # transposed conv: weight (Cin_t, Cout_t, k, k)
    wt = rnd(Co, C, k, k, seed=25, scale=0.05)        # here Cin_t = Co, Cout_t = C
    wtr = r32(wt).requires_grad_(True)
    xt = rnd(N, Co, 8, 8, seed=26)
    out = F.conv_transpose2d(r32(xt), wtr, stride=2, padding=1, output_padding=1)
    dy = rnd(*out.shape, seed=27)
    out.backward(r32(dy))
    dwt = torch.zeros(C, k * k, Co, device=DEV)       # [Cout_t][taps][Cin_t]
    ops.wgrad_tc(to_planes(dy), to_nhwc5(xt), dwt, 8, 8, convspec.wgrad_convT(k, 1), C, Co)
    check(dwt, wtr.grad.permute(1, 2, 3, 0).reshape(C, k * k, Co))


@pytest.mark.parametrize('H,W', [(16, 16), (21, 35)])
def test_output_conv_64_to_3_tanh_forward_and_all_adjoints(H, W):
    """The generator's last layer (reflpad3 + conv7 64->3 + tanh, generators.py:87) through ConvFn: fused tanh epilogue
    with f32 NCHW output, direct small-Cout dgrad kernel, tensor-core wgrad, bias column sums."""
    from scene_generation_b200 import functional as Fn
    from scene_generation_b200.functional import ConvSpec
    N, C, Co, k = 2, 64, 3, 7
    x, w, b = rnd(N, C, H, W, seed=30), rnd(Co, C, k, k, seed=31, scale=0.05), rnd(Co, seed=32)
    xp = F.pad(r32(x), (3, 3, 3, 3), mode='reflect')
    xr = xp.clone().requires_grad_(True)
    wr, br = r32(w).requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.tanh(F.conv2d(xr, wr, br))
    g = rnd(*ref.shape, seed=33)
    ref.backward(g)
    wd = torch.nn.Parameter(w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).to(DEV))     # channels-last master
    bd = torch.nn.Parameter(b.to(DEV))
    op = to_nhwc5(xp).requires_grad_(True)
    y = Fn.conv(op, wd, bd, ConvSpec('s1', 7, 0, act=3, out='f32_nchw'))
    check(y, ref.detach(), 1e-2)
    y.backward(g.to(DEV))
    check(op.grad[:, 0].permute(0, 3, 1, 2), xr.grad, 2e-2)
    check(wd.grad, wr.grad, 2e-2)
    check(bd.grad, br.grad, 2e-2)


# ---- MN-major weight operand: the adjoint reads the fprop weight copy (sg_conv_desc_t.w_mn) -------------
@pytest.mark.parametrize('M,Kdim,Ndim,c0,c1', [
    (300, 512, 454, 0, 454),        # Linear dgrad: K = out features, N = in features (not a multiple of 8 columns)
    (208, 4, 512, 0, 512),          # K tail: 4 weight rows (box_net's last layer)
    (130, 1152, 512, 0, 512),       # long K
    (77, 172, 1024, 0, 1024),       # K = 172 (class logits), N = 1024 -> BN 256
    (260, 192, 192, 0, 192),        # N = 192 -> the 192-wide tile
    (90, 64, 208, 168, 204),        # restricted columns [168, 204) of a 208-wide tensor (partial dgrad)
    (64, 8, 8, 0, 3),               # tiny everything
])
def test_gemm_mn_major_weights(M, Kdim, Ndim, c0, c1):
    """y[m, n] = sum_k x[m, k] * w[k, c0 + n] with w stored (K rows, 1 tap, N cols): must equal the K-major kernel
    fed with the explicitly transposed copy, and the fp32 reference."""
    x, w = rnd(M, Kdim, seed=1), rnd(Kdim, Ndim, seed=2, scale=0.1)
    Kp, Np = ops.round_up(Kdim, 8), ops.round_up(Ndim, 8)
    x5 = torch.zeros(1, 1, 1, M, Kp, dtype=torch.bfloat16)
    x5[0, 0, 0, :, :Kdim] = bf(x)
    w3 = torch.zeros(Kdim, 1, Np, dtype=torch.bfloat16)
    w3[:, 0, :Ndim] = bf(w)
    n = c1 - c0
    y = torch.full((M, n), 7.0, device=DEV)
    ops.conv_tc(x5.to(DEV), w3.to(DEV), y, (0, 0, n, 1), 1, M, [(0, 0, 0, 0)], mn_cols=(c0, c1))
    ref = r32(x) @ r32(w)[:, c0:c1]
    check(y, ref)
    wt3 = torch.zeros(n, 1, Kp, dtype=torch.bfloat16)
    wt3[:, 0, :Kdim] = bf(w)[:, c0:c1].t()
    y2 = torch.empty((M, n), device=DEV)
    ops.conv_tc(x5.to(DEV), wt3.to(DEV), y2, (0, 0, n, 1), 1, M, [(0, 0, 0, 0)])
    check(y, y2.float().cpu(), 1e-5)


@pytest.mark.parametrize('Cin,Cout,k,H', [(64, 128, 3, 16), (192, 192, 3, 8), (24, 1024, 3, 8), (8, 64, 4, 9), (304, 256, 3, 8)])
def test_conv_dgrad_mn_major_weights(Cin, Cout, k, H):
    """dgrad of a stride-1 zero-padded conv through the fprop weight copy == F.conv_transpose2d reference."""
    N, pad = 3, k // 2
    dy, w = rnd(N, Cout, H, H, seed=4), rnd(Cout, Cin, k, k, seed=5, scale=0.05)
    Ho = H
    Hx = Ho + k - 1 - 2 * pad
    ref = F.conv_transpose2d(r32(dy), r32(w), padding=pad)
    wk = pack_w(w)                                              # (Cout, k*k, Cin_p): the fprop operand
    Cx = wk.shape[2]
    dx = torch.zeros((N, Hx, Hx, Cx), dtype=torch.bfloat16, device=DEV)
    ops.conv_tc(to_nhwc5(dy), wk, dx, (Hx * Hx * Cx, Hx * Cx, Cx, 1), Hx, Hx, convspec.dgrad_s1(k, pad), mn_cols=(0, Cin))
    check(dx[..., :Cin].permute(0, 3, 1, 2), ref, 1e-2)         # bf16 output rounding


@pytest.mark.parametrize('N,C,H,Co', [(4, 1024, 8, 1024), (5, 512, 8, 512), (32, 1024, 8, 1024), (3, 576, 16, 256)])
def test_cta_pair_kernel_long_k_wide_n(N, C, H, Co):
    """The shapes served by the CTA-pair kernel (tcgen05 cta_group::2, M = 256 x N = 256 per pair, K split over
    blockIdx.z with the partial accumulators added in split order): ResnetBlock convolution (fused statistics, odd
    numbers of M tiles) and its input gradient through the fprop weight copy; bit-reproducible."""
    k, p = 3, 1
    x, w, b = rnd(N, C, H, H, seed=4), rnd(Co, C, k, k, seed=5, scale=0.03), rnd(Co, seed=6)
    xp = F.pad(r32(x), (p, p, p, p), mode='reflect')
    ref = F.conv2d(xp, r32(w), b)
    taps, _ = convspec.conv_s1(k, 0)
    outs = []
    for rep in range(2):
        y = torch.full((N, H, H, Co), float('nan'), device=DEV, dtype=torch.bfloat16)
        _, st = ops.conv_tc(to_nhwc5(xp), pack_w(w), y, (H * H * Co, H * Co, Co), H, H, taps, bias=b.to(DEV), stats=True)
        outs.append((y, st))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    y, st = outs[0]
    check(y.permute(0, 3, 1, 2), ref, 1e-2)
    check(st.sum(1)[..., 0], ref.sum(dim=(2, 3)), 2e-3)
    check(st.sum(1)[..., 1], (ref ** 2).sum(dim=(2, 3)), 2e-3)
    # input gradient: dx = conv_transpose(dy, w) with the SAME bf16 weight copy read MN-major
    dy = rnd(N, Co, H, H, seed=7)
    refd = F.conv_transpose2d(r32(dy), r32(w), padding=p)
    dx = torch.full((N, H, H, C), float('nan'), dtype=torch.bfloat16, device=DEV)
    ops.conv_tc(to_nhwc5(dy), pack_w(w), dx, (H * H * C, H * C, C, 1), H, H, convspec.dgrad_s1(k, p), mn_cols=(0, C))
    check(dx.permute(0, 3, 1, 2), refd, 1e-2)

