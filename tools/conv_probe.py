#!/usr/bin/env python
"""ResnetBlock-shaped launches of sg_conv_tc for ncu captures and A/B timing:
    python tools/conv_probe.py                      # CUDA-event timing (20 launches each, L2-warm like inside a step)
    ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 6 -o gpurun_out/prof_conv python tools/conv_probe.py --once
Shapes: fprop 32 x 8x8, 1024 -> 1024, 3x3 (GEMM 2048 x 1024 x 9216) and its input gradient over the padded 10x10 grid."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_generation_b200 import convspec, ops      # noqa: E402

DEV = 'cuda'
once = '--once' in sys.argv
g = torch.Generator().manual_seed(0)
x = (torch.rand((32, 1, 10, 10, 1024), generator=g) - 0.5).to(torch.bfloat16).to(DEV)
w = ((torch.rand((1024, 9, 1024), generator=g) - 0.5) * 0.05).to(torch.bfloat16).to(DEV)
bias = torch.zeros(1024, device=DEV)
y = torch.empty((32, 8, 8, 1024), device=DEV, dtype=torch.bfloat16)
taps, _ = convspec.conv_s1(3, 0)
dz = (torch.rand((32, 1, 8, 8, 1024), generator=g) - 0.5).to(torch.bfloat16).to(DEV)
dx = torch.empty((32, 10, 10, 1024), device=DEV, dtype=torch.bfloat16)
dtaps = convspec.dgrad_s1(3, 2)          # full correlation: gradient w.r.t. the 10x10 padded operand


def fwd():
    ops.conv_tc(x, w, y, (64 * 1024, 8 * 1024, 1024), 8, 8, taps, bias=bias, stats=True)


def dgrad():
    ops.conv_tc(dz, w, dx, (100 * 1024, 10 * 1024, 1024, 1), 10, 10, dtaps, mn_cols=(0, 1024))


for name, fn, flops in (('fprop 2048x1024x9216 (+stats)', fwd, 2 * 2048 * 1024 * 9216), ('dgrad 3200x1024x9216', dgrad, 2 * 3200 * 1024 * 9216)):
    n = 1 if once else 20
    for _ in range(0 if once else 3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print('%-32s %.1f us  %.0f TFLOP/s' % (name, us, flops / us / 1e6))
