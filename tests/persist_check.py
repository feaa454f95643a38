"""Timing + parity of short-K, many-tile convolutions with the current SG_CONV_PERSIST setting (run it twice, with
SG_CONV_PERSIST=0 and =1, and compare).  EXPERIMENTAL: the persistent kernel was written without hardware access at
the end of round 1.  Shapes are the ones whose per-CTA fixed cost dominates in the round-1 profile."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_generation_b200 import _lib, convspec, ops          # noqa: E402
from tests.test_gpu_conv_tc import pack_w, r32, rnd, to_nhwc5   # noqa: E402

DEV = 'cuda'
tag = 'persist=%s' % os.environ.get('SG_CONV_PERSIST', '0')
bad = 0
#      N   Cin  H   Cout k  act
for (N, C, H, Co, k, act) in [(32, 64, 64, 64, 3, _lib.ACT_NONE), (32, 64, 64, 128, 3, _lib.ACT_RELU), (16, 128, 32, 256, 3, _lib.ACT_NONE),
                              (24, 64, 65, 64, 4, _lib.ACT_LEAKY), (8, 64, 128, 3, 7, _lib.ACT_TANH), (40, 192, 32, 192, 3, _lib.ACT_NONE),
                              (8, 256, 16, 1024, 3, _lib.ACT_NONE)]:
    p = k // 2
    x, w, b = rnd(N, C, H, H, seed=4), rnd(Co, C, k, k, seed=5, scale=0.03), rnd(Co, seed=6)
    xr = r32(x)
    ref = F.conv2d(xr, r32(w), b, padding=p)
    Ho = ref.shape[2]
    taps, off = convspec.conv_s1(k, p)
    Cop = ops.round_up(Co, 8)
    fp32_out = Co % 8 != 0
    if fp32_out:
        y = torch.full((N, Co, Ho, Ho), float('nan'), device=DEV)
        strides = (Co * Ho * Ho, Ho, 1, Ho * Ho)
    else:
        y = torch.full((N, Ho, Ho, Co), float('nan'), device=DEV)
        strides = (Ho * Ho * Co, Ho * Co, Co, 1)
    stats = torch.zeros(N, Co, 2, device=DEV)
    x5, w3, bd = to_nhwc5(x), pack_w(w), b.to(DEV)
    kw = dict(in_h0=off, in_w0=off, bias=bd, act=act, slope=0.2, stats=stats)
    ops.conv_tc(x5, w3, y, strides, Ho, Ho, taps, **kw)
    torch.cuda.synchronize()
    refa = {_lib.ACT_NONE: ref, _lib.ACT_RELU: ref.relu(), _lib.ACT_LEAKY: F.leaky_relu(ref, 0.2), _lib.ACT_TANH: ref.tanh()}[act]
    got = y.cpu() if fp32_out else y.permute(0, 3, 1, 2).cpu()
    err = (got - refa).abs().max().item()
    serr = (stats[..., 0].cpu() - ref.sum(dim=(2, 3))).abs().max().item()
    ok = err <= 3e-3 * max(refa.abs().max().item(), 1e-3) and serr <= 3e-3 * ref.sum(dim=(2, 3)).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.conv_tc(x5, w3, y, strides, Ho, Ho, taps, **kw)
    e0.record()
    for _ in range(20):
        ops.conv_tc(x5, w3, y, strides, Ho, Ho, taps, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * Ho * Ho * Co * C * k * k
    print('[%s] conv %-28s max err %.3e stats err %.3e %s  %.1f us  %.0f TFLOP/s' % (
        tag, (N, C, H, Co, k), err, serr, 'OK' if ok else 'MISMATCH', us, fl / us / 1e6))
    bad += not ok
sys.exit(1 if bad else 0)
