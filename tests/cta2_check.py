"""Run with SG_CONV_2CTA=1: the CTA-pair (tcgen05 cta_group::2, M = 256) variant of sg_conv_tc vs fp32 PyTorch on the
same bf16-rounded operands — K-major weights (fprop) and MN-major weights (dgrad through the fprop copy), even and odd
numbers of M tiles, fused statistics.  Spawned by test_gpu_conv_2cta.py so that the environment switch is read fresh;
EXPERIMENTAL: written without access to hardware at the end of round 1, first thing to run in round 2
(`SG_TEST_2CTA=1 python -m pytest tests/test_gpu_conv_2cta.py`, or this file directly under `timeout 120`)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_generation_b200 import convspec, ops          # noqa: E402
from tests.test_gpu_conv_tc import pack_w, r32, rnd, to_nhwc5   # noqa: E402

assert os.environ.get('SG_CONV_2CTA') == '1'
DEV = 'cuda'
bad = 0
# fprop, 3x3 on reflection-padded input (the resblock shape family); (N=5, H=8) gives an odd number of M tiles
for (N, C, H, Co, k) in [(4, 1024, 8, 1024, 3), (6, 512, 8, 256, 3), (5, 256, 8, 128, 3), (2, 512, 10, 512, 3), (3, 128, 16, 384, 3)]:
    p = k // 2
    x, w, b = rnd(N, C, H, H, seed=4), rnd(Co, C, k, k, seed=5, scale=0.03), rnd(Co, seed=6)
    xp = F.pad(r32(x), (p, p, p, p), mode='reflect')
    ref = F.conv2d(xp, r32(w), b)
    taps, _ = convspec.conv_s1(k, 0)
    y = torch.full((N, H, H, Co), float('nan'), device=DEV)
    stats = torch.zeros(N, Co, 2, device=DEV)
    ops.conv_tc(to_nhwc5(xp), pack_w(w), y, (H * H * Co, H * Co, Co), H, H, taps, bias=b.to(DEV), stats=stats)
    torch.cuda.synchronize()
    err = (y.permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    serr = (stats[..., 0].cpu() - ref.sum(dim=(2, 3))).abs().max().item()
    ok = err <= 2e-3 * ref.abs().max().item() and serr <= 2e-3 * ref.sum(dim=(2, 3)).abs().max().item()
    print('2-CTA conv fprop', (N, C, H, Co, k), 'max err %.3e' % err, 'stats err %.3e' % serr, 'OK' if ok else 'MISMATCH')
    bad += not ok
# dgrad through the fprop weight copy (MN-major B)
for (N, Cin, Cout, k, H) in [(4, 1024, 1024, 3, 8), (3, 256, 512, 3, 8), (5, 128, 256, 3, 16)]:
    pad = k // 2
    dy, w = rnd(N, Cout, H, H, seed=7), rnd(Cout, Cin, k, k, seed=8, scale=0.03)
    ref = F.conv_transpose2d(r32(dy), r32(w), padding=pad)
    wk = pack_w(w)
    dx = torch.zeros((N, H, H, Cin), dtype=torch.bfloat16, device=DEV)
    ops.conv_tc(to_nhwc5(dy), wk, dx, (H * H * Cin, H * Cin, Cin, 1), H, H, convspec.dgrad_s1(k, pad), mn_cols=(0, Cin))
    torch.cuda.synchronize()
    err = (dx.float().permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    ok = err <= 1e-2 * ref.abs().max().item()
    print('2-CTA conv dgrad', (N, Cin, Cout, k, H), 'max err %.3e' % err, 'OK' if ok else 'MISMATCH')
    bad += not ok
# timing of the shape the variant exists for
x5, w3 = to_nhwc5(F.pad(rnd(32, 1024, 8, 8, seed=1), (1, 1, 1, 1), mode='reflect')), pack_w(rnd(1024, 1024, 3, 3, seed=2, scale=0.02))
y = torch.empty((32, 8, 8, 1024), device=DEV, dtype=torch.bfloat16)
taps, _ = convspec.conv_s1(3, 0)
for _ in range(3):
    ops.conv_tc(x5, w3, y, (64 * 1024, 8 * 1024, 1024), 8, 8, taps)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv_tc(x5, w3, y, (64 * 1024, 8 * 1024, 1024), 8, 8, taps)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print('resblock conv (32x8x8, 1024->1024, 3x3) CTA pairs: %.1f us  %.0f TFLOP/s' % (us, 2 * 2048 * 1024 * 9216 / us / 1e6))
sys.exit(1 if bad else 0)
