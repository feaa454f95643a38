#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound kernels of the path (layout scatter forward / adjoint, box crop, graph gather /
pooled scatter) at the BASELINE configs' shapes: CUDA events on the launching stream, a 256 MB L2 flush between
launches, algorithmic bytes per launch over the median launch time against MEASURED_PEAKS.json.

    python tools/hbm_kernels.py [--config cfg2|cfg4|cfg5] [--batch 32] [--json out.json]
    ncu --set full --clock-control none -k regex:layout_fwd_tile -c 2 -o gpurun_out/prof python tools/hbm_kernels.py --reps 1
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scene_generation_b200 import ops, synthetic                        # noqa: E402

DEV = 'cuda'


def timed(fn, reps, flush):
    times = []
    for r in range(reps + 2):
        flush.zero_()
        torch.cuda._sleep(300000)          # the host runs ahead: the interval is the kernel, not the launch gap
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 2 or reps == 1:
            times.append(e0.elapsed_time(e1))
    return sorted(times)[len(times) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='cfg2', choices=['cfg2', 'cfg4', 'cfg5'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    H, kmin, kmax = {'cfg2': (128, 3, 8), 'cfg4': (256, 8, 15), 'cfg5': (128, 29, 29)}[a.config]
    peak = 6550.0
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = json.load(open(p)).get('hbm_gbs', peak)
    hb = synthetic.make_batch(a.batch, (H, H), 172, kmin, kmax, seed=3)
    meta = synthetic.HostMeta(hb)
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = meta.attach(tuple(t.to(DEV) for t in hb))
    O, T, N, D = objs.numel(), triples.shape[0], a.batch, 204
    ranges = o2i._sg_ranges
    seg_ptr, seg_src = triples._sg_csr
    vecs = torch.zeros((O, D), device=DEV)
    vecs.scatter_(1, objs.view(-1, 1), 1.0)
    vecs[:, 172:] = torch.randn(O, 32, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    rows = []

    def rec(name, nbytes, fn):
        ms = timed(fn, a.reps, flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({'kernel': name, 'config': a.config, 'ms': round(ms, 4), 'algorithmic_bytes': int(nbytes),
                     'achieved_gbs': round(gbs, 1), 'peak_gbs': peak, 'frac': round(gbs / peak, 3)})
        print('%-34s %9.4f ms  %8.1f MB  %8.1f GB/s  %.3f of the measured HBM peak' % (name, ms, nbytes / 1e6, gbs, gbs / peak))

    # layout scatter, dense reference-shaped vectors cat(one_hot, appearance), bf16 NHWC (Cp = 208)
    Cp = 208
    out_bytes = N * H * H * Cp * 2
    in_bytes = O * (D * 4 + 32 * 32 * masks.element_size() + 16)
    rec('layout_fwd (dense, bf16 NHWC)', out_bytes + in_bytes,
        lambda: ops.masks_to_layout_fwd(vecs, boxes, masks, ranges, H, H, False, ops.NHWC_BF16, raw=True))
    # the compacted variant of the train step: 64 channels
    cvecs = torch.zeros((O, 56), device=DEV)
    cvecs[:, 24:] = vecs[:, 172:]
    cvecs.scatter_(1, objs._sg_compact[0].view(-1, 1).clamp(max=23), 1.0)
    rec('layout_fwd (compact 64 ch)', N * H * H * 64 * 2 + O * (56 * 4 + 32 * 32 * masks.element_size() + 16),
        lambda: ops.masks_to_layout_fwd(cvecs, boxes, masks, ranges, H, H, False, ops.NHWC_BF16, raw=True, Cp=64))
    g = torch.randn((N, H, H, 64), device=DEV).to(torch.bfloat16)
    rec('layout_bwd d vecs (compact 64 ch)', N * H * H * 64 * 2 + O * 56 * 4,
        lambda: ops.masks_to_layout_bwd(cvecs, boxes, masks, ranges, H, H, g))
    gd = torch.randn((N, H, H, Cp), device=DEV).to(torch.bfloat16)
    rec('layout_bwd d vecs (dense)', N * H * H * Cp * 2 + O * D * 4,
        lambda: ops.masks_to_layout_bwd(vecs, boxes, masks, ranges, H, H, gd))
    # box crops: 64x64 appearance crops (forward) and the 32x32 object-discriminator crops (forward + adjoint)
    rec('crop_fwd 64x64 (bf16 NHWC)', O * 64 * 64 * 8 * 2 + N * 3 * H * H * 4,
        lambda: ops.crop_bbox_fwd(imgs, boxes, o2i, 64, 64, False, ops.NHWC_BF16))
    gc = torch.randn((O, 32, 32, 8), device=DEV).to(torch.bfloat16)
    rec('crop_bwd 32x32 -> d image', O * 32 * 32 * 8 * 2 + N * 3 * H * H * 4,
        lambda: ops.crop_bbox_bwd(gc, boxes, o2i, N, 3, H, H, False, ops.NHWC_BF16))
    # graph convolution gather / pooled scatter (first layer: Do = 163, later layers: 128)
    for Do in (163, 128):
        ov, pv = torch.randn((O, Do), device=DEV), torch.randn((T, 128), device=DEV)
        edges = torch.stack([triples[:, 0], triples[:, 2]], 1).contiguous()
        ld = (2 * Do + 128 + 7) // 8 * 8
        rec('gconv_gather Do=%d (bf16 operand)' % Do, T * (2 * Do + 128) * 4 + T * ld * 2,
            lambda: ops.gconv_gather(ov, pv, edges, torch.bfloat16, ld))
    new_t = torch.randn((T, 1152), device=DEV)
    rec('gconv_pool H=512', T * 1024 * 4 + O * 512 * 4, lambda: ops.gconv_pool(new_t, 640, seg_ptr, seg_src, O, 512, True))
    dp = torch.randn((O, 512), device=DEV)
    dnp = torch.randn((T, 128), device=DEV)
    rec('gconv_pool_bwd', O * 512 * 4 + T * 128 * 4 + T * 1152 * 4,
        lambda: ops.gconv_pool_bwd(dp, dnp, edges, seg_ptr, T, 512, 128, True))
    dcur = torch.randn((T, 384), device=DEV)
    rec('gconv_gather_bwd', T * 384 * 4 + O * 128 * 4 + T * 128 * 4,
        lambda: ops.gconv_gather_bwd(dcur, seg_ptr, seg_src, O, T, 128, 128))
    print('O = %d objects, T = %d triples, N = %d images, %dx%d' % (O, T, N, H, H))
    if a.json:
        json.dump({'config': a.config, 'O': O, 'T': T, 'N': N, 'H': H, 'rows': rows}, open(a.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
