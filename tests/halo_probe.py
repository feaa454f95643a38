"""Round-2 groundwork: does a tap-shifted UMMA descriptor over ONE shared-memory halo tile reproduce a 3x3 convolution?
Run on the GPU box: `timeout 60 python tests/halo_probe.py`.  Prints the max error of both descriptor modes against the
production kernel (sg_conv_tc) on the same operands; a mode with error ~1e-6 answers the question with yes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene_generation_b200 import _lib, convspec, ops          # noqa: E402
from scene_generation_b200.ops import _ptr, _stream            # noqa: E402

DEV = 'cuda'
g = torch.Generator().manual_seed(0)
x5 = (torch.rand((1, 1, 18, 10, 64), generator=g) - 0.5).to(torch.bfloat16).to(DEV)
w3 = ((torch.rand((64, 9, 64), generator=g) - 0.5) * 0.2).to(torch.bfloat16).to(DEV)
ref = torch.empty((1, 16, 8, 64), dtype=torch.float32, device=DEV)
taps, off = convspec.conv_s1(3, 0)
ops.conv_tc(x5, w3, ref, (16 * 8 * 64, 8 * 64, 64, 1), 16, 8, taps)
torch.cuda.synchronize()
ok = False
for mode in (0, 1):
    y = torch.full((128, 64), float('nan'), device=DEV)
    _lib.call('sg_probe_shifted_desc', _ptr(x5), _ptr(w3), _ptr(y), mode, _stream())
    torch.cuda.synchronize()
    err = (y.view(1, 16, 8, 64) - ref).abs().max().item()
    print('halo probe, descriptor base_offset mode %d: max |err| vs sg_conv_tc = %.3e (scale %.3e)' % (mode, err, ref.abs().max().item()))
    ok |= err < 1e-4 * ref.abs().max().item()
print('tap-shifted descriptors over a halo tile:', 'WORK' if ok else 'do NOT work this way')
