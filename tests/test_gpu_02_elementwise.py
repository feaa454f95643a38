"""Operand-writer / pooling / concat kernels (forward and adjoint) vs plain fp32 PyTorch of the same op.
bf16 storage -> tolerance 2e-2 of the output scale."""
import pytest
import torch
import torch.nn.functional as F

from scene_generation_b200 import _lib
from scene_generation_b200 import functional as Fn
from scene_generation_b200.functional import NapSpec

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 2 - 1


def close(a, b, tol=2e-2, name=''):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(b.abs().max().item(), 1e-3), '%s err %.3e scale %.3e' % (name, err, b.abs().max().item())


def stats_of(y_nhwc):
    yf = y_nhwc.float()
    # (N, slots = 1, C, 2): the partial-sum layout the conv epilogue writes (sg_conv_desc_t.stats)
    return torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1).unsqueeze(1).contiguous()


def planes_ref(x_nchw):
    N, C, H, W = x_nchw.shape
    out = torch.zeros(N, 4, (H + 1) // 2, (W + 1) // 2, C)
    for ph in range(2):
        for pw in range(2):
            s = x_nchw[:, :, ph::2, pw::2]
            out[:, ph * 2 + pw, :s.shape[2], :s.shape[3]] = s.permute(0, 2, 3, 1)
    return out


@pytest.mark.parametrize('H,W,C,pad,mode,planes,act', [(8, 8, 64, 1, 1, False, 1), (9, 7, 24, 3, 1, False, 1),
                                                       (12, 12, 128, 0, 0, True, 2), (5, 5, 16, 0, 0, True, 0),
                                                       (6, 6, 32, 2, 0, False, 1)])
def test_nap_instance_norm(H, W, C, pad, mode, planes, act):
    N = 3
    y = (rnd(N, H, W, C, seed=1) * 2 + 0.3).to(torch.bfloat16)
    yd = y.to(DEV).requires_grad_(True)
    spec = NapSpec(norm='in', act=act, slope=0.2, pad=pad, pad_mode=mode, planes=planes)
    out = Fn.nap(yd, stats_of(yd.detach()), spec=spec)
    yr = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    z = F.instance_norm(yr, eps=1e-5)
    z = F.relu(z) if act == 1 else (F.leaky_relu(z, 0.2) if act == 2 else z)
    if pad:
        z = F.pad(z, (pad,) * 4, mode='reflect' if mode else 'constant')
    ref = planes_ref(z) if planes else z.permute(0, 2, 3, 1).unsqueeze(1)
    close(out, ref, name='fwd')
    g = rnd(*ref.shape, seed=2)
    out.backward(g.to(DEV).to(torch.bfloat16))
    (ref * g.to(torch.bfloat16).float()).sum().backward()
    close(yd.grad.permute(0, 3, 1, 2), yr.grad, 3e-2, name='bwd')


def test_nap_batchnorm_upsample_and_running_stats():
    N, H, W, C = 5, 4, 4, 192
    y = (rnd(N, H, W, C, seed=3) * 1.5 - 0.2).to(torch.bfloat16)
    gamma, beta = rnd(C, seed=4) * 0.2 + 1.0, rnd(C, seed=5) * 0.1
    bn = torch.nn.BatchNorm2d(C)
    bn.weight.data.copy_(gamma)
    bn.bias.data.copy_(beta)
    bn.train()
    yd = y.to(DEV).requires_grad_(True)
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    out = Fn.nap(yd, stats_of(yd.detach()), gd, bd, None, (rm, rv), NapSpec(norm='bn', act=_lib.ACT_RELU, up=2))
    yr = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(F.relu(bn(yr)), scale_factor=2, mode='nearest')
    close(out, ref.permute(0, 2, 3, 1).unsqueeze(1), name='bn fwd')
    close(rm, bn.running_mean, 1e-3, 'running_mean')
    close(rv, bn.running_var, 1e-3, 'running_var')
    g = rnd(*ref.shape, seed=6)
    out.backward(g.permute(0, 2, 3, 1).unsqueeze(1).to(DEV).to(torch.bfloat16))
    (ref * g.to(torch.bfloat16).float()).sum().backward()
    close(yd.grad.permute(0, 3, 1, 2), yr.grad, 3e-2, 'bn dx')
    close(gd.grad, bn.weight.grad, 3e-2, 'dgamma')
    close(bd.grad, bn.bias.grad, 3e-2, 'dbeta')


def test_nap_residual_block_tail():
    N, H, C = 2, 8, 64
    y = rnd(N, H, H, C, seed=7).to(torch.bfloat16)
    xr = rnd(N, C, H, H, seed=8).to(torch.bfloat16)
    res_op = F.pad(xr.float(), (1, 1, 1, 1), mode='reflect').permute(0, 2, 3, 1).unsqueeze(1).contiguous().to(torch.bfloat16)
    yd = y.to(DEV).requires_grad_(True)
    rd = res_op.to(DEV).requires_grad_(True)
    out = Fn.nap(yd, stats_of(yd.detach()), None, None, rd, None, NapSpec(norm='in', pad=1, pad_mode=1, res_pad=1))
    yt = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    xt = xr.float().requires_grad_(True)
    ref = F.pad(xt + F.instance_norm(yt), (1, 1, 1, 1), mode='reflect')
    close(out, ref.permute(0, 2, 3, 1).unsqueeze(1), name='res fwd')
    g = rnd(*ref.shape, seed=9)
    out.backward(g.permute(0, 2, 3, 1).unsqueeze(1).to(DEV).to(torch.bfloat16))
    (ref * g.to(torch.bfloat16).float()).sum().backward()
    close(yd.grad.permute(0, 3, 1, 2), yt.grad, 3e-2, 'res dy')
    close(rd.grad[:, 0, 1:-1, 1:-1].permute(0, 3, 1, 2), xt.grad, 3e-2, 'res dx')
    assert rd.grad[:, 0, 0].abs().max().item() == 0


@pytest.mark.parametrize('H,W', [(64, 64), (33, 17), (8, 8)])
def test_avgpool_gap_concat_slot(H, W):
    N, C = 2, 48
    x = rnd(N, H, W, C, seed=10).to(torch.bfloat16)
    xd = x.to(DEV).requires_grad_(True)
    out = Fn.AvgPoolFn.apply(xd)
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.avg_pool2d(xt, 3, stride=2, padding=1, count_include_pad=False)
    close(out.permute(0, 3, 1, 2), ref, name='avgpool')
    g = rnd(*ref.shape, seed=11)
    out.backward(g.permute(0, 2, 3, 1).to(DEV).to(torch.bfloat16))
    ref.backward(g.to(torch.bfloat16).float())
    close(xd.grad.permute(0, 3, 1, 2), xt.grad, name='avgpool bwd')
    # global average pool
    xd2 = x.to(DEV).requires_grad_(True)
    gp = Fn.GapFn.apply(xd2)
    close(gp, x.float().mean(dim=(1, 2)), 1e-2, 'gap')
    gg = rnd(N, C, seed=12)
    gp.backward(gg.to(DEV))
    close(xd2.grad, (gg / (H * W)).view(N, 1, 1, C).expand(N, H, W, C), name='gap bwd')
    # one-hot concat
    cls = torch.tensor([3, 0])
    cc = Fn.ConcatCondFn.apply(x.to(DEV), cls.to(DEV), 10)
    assert cc.shape == (N, H, W, 64)
    oh = torch.zeros(N, 10)
    oh[torch.arange(N), cls] = 1
    ref_cc = torch.cat([x.float(), oh.view(N, 1, 1, 10).expand(N, H, W, 10), torch.zeros(N, H, W, 6)], dim=-1)
    close(cc, ref_cc, 1e-6, 'concat')
    # image slot
    lay = torch.zeros(N, H, W, C, dtype=torch.bfloat16)
    lay[..., :42] = x[..., :42]
    img = rnd(N, 3, H, W, seed=13)
    imd = img.to(DEV).requires_grad_(True)
    slot = Fn.ImageSlotFn.apply(lay.to(DEV), imd, 42)
    close(slot[..., 42:45].permute(0, 3, 1, 2), img, 1e-2, 'slot')
    close(slot[..., :42], lay[..., :42], 1e-6, 'slot keeps layout')
    gs = rnd(N, H, W, C, seed=14).to(torch.bfloat16)
    slot.backward(gs.to(DEV))
    close(imd.grad, gs[..., 42:45].float().permute(0, 3, 1, 2), 1e-6, 'slot bwd')


def test_linear_fn_forward_backward():
    M, K, Nout = 37, 454, 512
    x, w, b = rnd(M, K, seed=15), rnd(Nout, K, seed=16) * 0.1, rnd(Nout, seed=17)
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = Fn.linear(xd, wd, bd, _lib.ACT_RELU)
    xr, wr, br = (t.to(torch.bfloat16).float().requires_grad_(True) for t in (x, w, b))
    br = b.clone().requires_grad_(True)
    ref = F.relu(xr @ wr.t() + br)
    close(y, ref, 1e-2, 'linear fwd')
    g = rnd(M, Nout, seed=18)
    y.backward(g.to(DEV))
    ref.backward(g)
    close(xd.grad, xr.grad, 2e-2, 'dx')
    close(wd.grad, wr.grad, 2e-2, 'dw')
    close(bd.grad, br.grad, 2e-2, 'db')


@pytest.mark.parametrize('N,H,W,C,norm', [(3, 8, 8, 1024, 'in'), (2, 64, 64, 64, 'in'), (5, 33, 17, 128, 'in'),
                                          (7, 16, 16, 192, 'bn'), (40, 4, 4, 64, 'bn')])
def test_norm_backward_and_bias_gradient_are_bit_reproducible(N, H, W, C, norm):
    """No floating-point atomics on the path: two runs of the norm backward (partial sums per CTA added in a fixed
    order) and of the bias-gradient column sum give identical bits, and several slots of conv-epilogue partial
    sums finalize to the statistics of their total."""
    y = (rnd(N, H, W, C, seed=11) * 2 + 0.3).to(torch.bfloat16).to(DEV)
    g = rnd(N, 1, H, W, C, seed=12).to(torch.bfloat16).to(DEV)
    gamma = (rnd(C, seed=13) * 0.2 + 1.0).to(DEV).requires_grad_(True) if norm == 'bn' else None
    beta = (rnd(C, seed=14) * 0.1).to(DEV).requires_grad_(True) if norm == 'bn' else None
    st = stats_of(y)
    # the same statistics split into 3 slots must finalize identically up to fp32 rounding of the split
    st3 = torch.cat([st * 0.25, st * 0.5, st * 0.25], dim=1).contiguous()
    outs = []
    for stats in (st, st, st3):
        yd = y.clone().requires_grad_(True)
        if gamma is not None:
            gamma.grad = beta.grad = None
        out = Fn.nap(yd, stats, gamma, beta, None, None, NapSpec(norm=norm, act=_lib.ACT_LEAKY, slope=0.2))
        out.backward(g)
        outs.append((out.detach().clone(), yd.grad.clone(), None if gamma is None else gamma.grad.clone(),
                     None if beta is None else beta.grad.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert (a is None and b is None) or torch.equal(a, b)
    close(outs[2][0], outs[0][0], 1e-2, 'slots fwd')
    close(outs[2][1], outs[0][1], 2e-2, 'slots bwd')
    from scene_generation_b200 import ops
    x2 = g.reshape(-1, C)
    s1, s2 = ops.colsum(x2, C), ops.colsum(x2, C)
    assert torch.equal(s1, s2)
    close(s1, x2.float().sum(0), 2e-3, 'colsum')
