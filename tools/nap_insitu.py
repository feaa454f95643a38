#!/usr/bin/env python
"""Times every sg_norm_act_pad_fwd / sg_norm_act_pad_bwd call of one eager cfg2 training step IN PLACE (CUDA events
around the entry point, the inputs as L2-warm or -cold as the step leaves them) and prints the total per operand
shape with its algorithmic bytes:

    python tools/nap_insitu.py
"""
import collections
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scene_generation_b200 import _lib, args as sgargs, synthetic        # noqa: E402
from scene_generation_b200.trainer import Trainer                        # noqa: E402

H, NUM_OBJS, BATCH = 128, 172, 32
dev = torch.device('cuda', 0)
targs = sgargs.default_args(image_size=(H, H), num_objs=NUM_OBJS)
targs.cuda_graphs = False
torch.manual_seed(1234)
tr = Trainer(targs, synthetic.make_vocab(NUM_OBJS), {})
tr.use_graphs = False
hb = synthetic.make_batch(BATCH, (H, H), NUM_OBJS, 3, 8, seed=5)
meta = synthetic.HostMeta(hb)
batch = meta.attach(tuple(t.to(dev) for t in hb))
for i in range(3):
    tr.train_step(batch, use_gt=(i % 2 == 0))
torch.cuda.synchronize()

records = []
orig_call = _lib.call


def timed_call(name, *args):
    if name not in ('sg_norm_act_pad_fwd', 'sg_norm_act_pad_bwd'):
        return orig_call(name, *args)
    d = args[0]._obj
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *args)
    e1.record()
    elems = d.N * d.H * d.W * d.C
    Hp, Wp = d.H * d.up + 2 * d.pad, d.W * d.up + 2 * d.pad
    op_elems = d.N * Hp * Wp * d.C
    if name.endswith('fwd'):
        kind = 'fwd'
        nbytes = 2 * (elems + op_elems + (elems if d.res else 0))
    else:
        normed = bool(args[2])
        kind = 'bwd-norm' if normed else 'bwd'
        # reduce pass (operand gradient + source) and apply pass (the same again + the source gradient)
        nbytes = 2 * ((2 if normed else 1) * (op_elems + elems) + elems + (elems if args[9] else 0))
    key = (kind, d.N, d.H, d.W, d.C, d.up, d.pad, d.planes, bool(d.res) or (kind != 'fwd' and bool(args[9])))
    records.append((key, nbytes, e0, e1))


_lib.call = timed_call
tr.train_step(batch, use_gt=True)
torch.cuda.synchronize()
_lib.call = orig_call

tot = collections.OrderedDict()
for key, nbytes, e0, e1 in records:
    t = tot.setdefault(key, [0, 0.0, 0])
    t[0] += 1
    t[1] += e0.elapsed_time(e1) * 1e3
    t[2] += nbytes
print('%-9s %-34s %5s %10s %10s %9s' % ('kind', 'N,H,W,C,up,pad,planes,res', 'calls', 'us total', 'MB total', 'GB/s'))
all_us = collections.Counter()
all_b = collections.Counter()
for key, (n, us, nb) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-9s %-34s %5d %10.1f %10.1f %9.0f' % (key[0], ','.join(str(int(v)) for v in key[1:]), n, us, nb / 1e6, nb / us / 1e3))
    all_us[key[0]] += us
    all_b[key[0]] += nb
for k in all_us:
    print('TOTAL %-9s %10.1f us  %10.1f MB  %8.0f GB/s' % (k, all_us[k], all_b[k] / 1e6, all_b[k] / all_us[k] / 1e3))
print('TOTAL all       %10.1f us (events around each entry point: includes ~2-4 us of launch gap per call)' % sum(all_us.values()))
