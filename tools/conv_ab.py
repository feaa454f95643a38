#!/usr/bin/env python
"""A/B microbenchmark of the tensor-core conv kernel variants over the conv shapes of one training iteration.

    python tools/conv_ab.py                       # baseline vs SG_CONV_PERSIST=1 vs SG_CONV_2CTA=1
    python tools/conv_ab.py --configs base SG_CONV_PERSIST=1,SG_CONV_2CTA=1 --top 30

Each configuration runs in its own process (the switches are read once per process).  Shapes come from
profiles/r01_conv_shapes.json (bench.py's per-shape probe: N,H,W,Cout,Cin,taps,phases + launches per iteration); every
shape is timed as a stride-1 k x k convolution with the same GEMM dimensions (M = N*H*W pixels, N = Cout,
K = Cin*taps), 20 launches after 3 warm-ups, CUDA events.  Output: per-shape microseconds per configuration, the
iteration-weighted total, and the max abs difference of the outputs against the first configuration."""
import argparse
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(shapes):
    import torch
    sys.path.insert(0, ROOT)
    from scene_generation_b200 import convspec, ops
    dev = 'cuda'
    out = []
    g = torch.Generator().manual_seed(0)
    for (N, H, W, Co, Ci, taps, ph, launches) in shapes:
        k = int(round(math.sqrt(taps)))
        if k * k != taps:
            k = 1
        Cip, Cop = ops.round_up(Ci, 8), ops.round_up(Co, 8)
        x5 = (torch.rand((N, 1, H + k - 1, W + k - 1, Cip), generator=g) - 0.5).to(torch.bfloat16).to(dev)
        w3 = ((torch.rand((Co, k * k, Cip), generator=g) - 0.5) * 0.1).to(torch.bfloat16).to(dev)
        if Co % 8 == 0:
            y = torch.empty((N, H, W, Co), dtype=torch.bfloat16, device=dev)
            strides = (H * W * Co, W * Co, Co, 1)
        else:
            y = torch.empty((N, Co, H, W), dtype=torch.float32, device=dev)
            strides = (Co * H * W, W, 1, H * W)
        tp, off = convspec.conv_s1(k, 0)
        for _ in range(3):
            ops.conv_tc(x5, w3, y, strides, H, W, tp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_tc(x5, w3, y, strides, H, W, tp)
        e1.record()
        torch.cuda.synchronize()
        out.append({'us': e0.elapsed_time(e1) / 20 * 1e3, 'sum': float(y.float().abs().sum()), 'probe': y.flatten()[:4096].float().cpu().tolist()})
    print('RESULT ' + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', nargs='*', default=['base', 'SG_CONV_PERSIST=1', 'SG_CONV_2CTA=1'])
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--shapes', default=os.path.join(ROOT, 'profiles', 'r01_conv_shapes.json'))
    ap.add_argument('--worker', default=None)
    a = ap.parse_args()
    if a.worker:
        return worker(json.loads(a.worker))
    rows = [r for r in json.load(open(a.shapes)) if r['kernel'] == 'conv_tc'][:a.top]
    shapes = [r['N,H,W,Cout,Cin,taps,phases'] + [r['launches']] for r in rows]
    results = {}
    for cfg in a.configs:
        env = dict(os.environ)
        if cfg != 'base':
            for kv in cfg.split(','):
                k, v = kv.split('=')
                env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', json.dumps(shapes)], env=env,
                           capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        if r.returncode != 0 or not line:
            print('config %s FAILED:\n%s\n%s' % (cfg, r.stdout[-1500:], r.stderr[-1500:]))
            continue
        results[cfg] = json.loads(line[-1][7:])
    cfgs = list(results)
    if not cfgs:
        sys.exit(1)
    print('%-36s %4s ' % ('N,H,W,Cout,Cin,taps,phases', 'n') + ' '.join('%22s' % c[-22:] for c in cfgs) + '   max|diff| vs first')
    tot = {c: 0.0 for c in cfgs}
    for i, s in enumerate(shapes):
        us = [results[c][i]['us'] for c in cfgs]
        base = results[cfgs[0]][i]['probe']
        diff = max((max(abs(x - y) for x, y in zip(results[c][i]['probe'], base)) for c in cfgs[1:]), default=0.0)
        for c, u in zip(cfgs, us):
            tot[c] += u * s[7]
        print('%-36s %4d ' % (str(s[:7]), s[7]) + ' '.join('%19.1f us' % u for u in us) + '   %.3e' % diff)
    print('%-41s ' % 'per iteration (launch-weighted), ms' + ' '.join('%19.3f ms' % (tot[c] / 1e3) for c in cfgs))


if __name__ == '__main__':
    main()
