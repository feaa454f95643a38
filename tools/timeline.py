#!/usr/bin/env python
"""Kernel timeline of captured (CUDA-graph) training steps from CUPTI (torch.profiler): how much of a step the GPU is
busy, how much of it runs two branches at once, and where the idle gaps are.

    python tools/timeline.py [--steps 3] > gpurun_out/timeline.txt
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scene_generation_b200 import args as sgargs, synthetic        # noqa: E402
from scene_generation_b200.trainer import Trainer                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--eager', action='store_true')
a = ap.parse_args()

H, NUM_OBJS, BATCH = 128, 172, 32
dev = torch.device('cuda', 0)
targs = sgargs.default_args(image_size=(H, H), num_objs=NUM_OBJS)
if a.eager:
    targs.cuda_graphs = False
torch.manual_seed(1234)
tr = Trainer(targs, synthetic.make_vocab(NUM_OBJS), {})
hb = synthetic.make_batch(BATCH, (H, H), NUM_OBJS, 3, 8, seed=5)
batch = synthetic.HostMeta(hb).attach(tuple(t.to(dev) for t in hb))
for i in range(8):                   # eager sightings, then the captures of both coin values
    tr.train_step(batch, use_gt=(i % 2 == 0))
torch.cuda.synchronize()

from torch.profiler import ProfilerActivity, profile              # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(a.steps):
        tr.train_step(batch, use_gt=(i % 2 == 0))
    torch.cuda.synchronize()

evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
if not evs:
    print('no CUDA events recorded')
    sys.exit(1)
t0, t1 = evs[0][0], max(e[1] for e in evs)
span = t1 - t0
busy = 0.0
overlap2 = 0.0
# sweep: time with >= 1 and >= 2 kernels in flight
pts = []
for s, e, _ in evs:
    pts.append((s, 1))
    pts.append((e, -1))
pts.sort()
depth, last = 0, pts[0][0]
for t, d in pts:
    if depth >= 1:
        busy += t - last
    if depth >= 2:
        overlap2 += t - last
    depth += d
    last = t
print('%d device activities over %d steps: span %.3f ms/step, busy %.3f ms/step (%.1f %%), >= 2 in flight %.3f ms/step, summed durations %.3f ms/step'
      % (len(evs), a.steps, span / a.steps / 1e3, busy / a.steps / 1e3, 100 * busy / span, overlap2 / a.steps / 1e3,
         sum(e - s for s, e, _ in evs) / a.steps / 1e3))
# idle gaps
gaps = []
cur_end, cur_name = evs[0][1], evs[0][2]
for s, e, name in evs[1:]:
    if s > cur_end:
        gaps.append((s - cur_end, cur_name, name))
    if e > cur_end:
        cur_end, cur_name = e, name
hist = collections.Counter()
for g, _, _ in gaps:
    hist['<2us' if g < 2 else '2-5us' if g < 5 else '5-20us' if g < 20 else '>=20us'] += g
print('idle time by gap length (us/step):', {k: round(v / a.steps, 1) for k, v in hist.items()}, 'gaps/step: %d' % (len(gaps) // a.steps))
print('largest gaps:')
for g, a_, b_ in sorted(gaps, reverse=True)[:25]:
    print('  %8.1f us  after %-60s before %s' % (g, a_[:60], b_[:60]))
# per (kernel, grid) totals from the chrome trace (kineto records the launch geometry of every kernel)
import json                                                       # noqa: E402
import tempfile                                                   # noqa: E402

with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, 'trace.json')
    prof.export_chrome_trace(path)
    trace = json.load(open(path))
shape_us = collections.Counter()
shape_n = collections.Counter()
for ev in trace.get('traceEvents', []):
    if ev.get('cat') == 'kernel':
        g = ev.get('args', {}).get('grid')
        key = (ev['name'][:64], tuple(g) if g else None)
        shape_us[key] += ev.get('dur', 0)
        shape_n[key] += 1
print('top (kernel, grid) by time (us/step, launches/step, us/launch):')
for key, us in shape_us.most_common(70):
    n = shape_n[key]
    print('  %9.1f  %5.1f  %8.1f  %-28s %s' % (us / a.steps, n / a.steps, us / n, key[1], key[0]))
by = collections.Counter()
cnt = collections.Counter()
for s, e, name in evs:
    by[name[:70]] += e - s
    cnt[name[:70]] += 1
print('top activities (us/step):')
for name, us in by.most_common(30):
    print('  %9.1f  %5d  %s' % (us / a.steps, cnt[name] // a.steps, name))
