#!/usr/bin/env python
"""clocks per tcgen05.mma (M=128, K=16) issued back to back by one thread, by N and issue pattern (sg_probe_mma_rate)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_generation_b200 import _lib              # noqa: E402
from scene_generation_b200.ops import _ptr, _stream  # noqa: E402

out = torch.zeros(1, device='cuda')
for mode, what in ((0, 'k-steps 0..3, one accumulator'), (1, 'same k-step'), (2, 'four accumulators'),
                   (3, 'A = shifted halo view (+3 rows)'), (4, 'A groups 14 rows apart')):
    for BN in (16, 64, 128, 256):
        vals = []
        for rep in range(3):
            _lib.call('sg_probe_mma_rate', BN, 2048, mode, _ptr(out), _stream())
            torch.cuda.synchronize()
            vals.append(float(out))
        ideal = 128 * BN * 16 * 2 / 8192.0
        print('N=%3d  %-32s %7.1f clk/MMA   (math at 8192 flop/clk/SM: %5.1f)' % (BN, what, min(vals), ideal))
