"""Channel-compacted layouts (csrc/compact.cu, Model._compact_plan): the per-image-weight convolutions and the
whole train step must reproduce the dense path — same products, minus the terms that multiply a zero channel.
Tolerances: the two paths differ only by fp32 summation order inside a K loop, then by bf16 rounding of activations."""
import random

import pytest
import torch

from oracle import cases, restate as R
from scene_generation_b200 import _lib, functional as Fn, layout as L, synthetic
from scene_generation_b200.functional import ConvSpec

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF = torch.bfloat16


def _ptr(t):
    return t.data_ptr()


def make_cmap(N, Cc, Cin, n_used, gen):
    cmap = torch.full((N, Cc), -1, dtype=torch.int32)
    for n in range(N):
        perm = torch.randperm(Cin, generator=gen)[:n_used]
        cmap[n, :n_used] = perm.to(torch.int32)
    return cmap


def test_pack_weight_cmap_and_scatter():
    gen = torch.Generator().manual_seed(3)
    Cout, taps, Cin, N, Cc = 12, 5, 37, 4, 16
    w = torch.randn(Cout, taps, Cin, generator=gen)
    cmap = make_cmap(N, Cc, Cin, 11, gen)
    cmap[1, 3] = Cin + 2                       # beyond the weight's channels -> zero, like the image slot in the generator
    wd, cd = w.to(DEV), cmap.to(DEV)
    Coutp = 16
    wk = torch.empty((N, Cout, taps, Cc), dtype=BF, device=DEV)
    wt = torch.empty((N, Cc, taps, Coutp), dtype=BF, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call('sg_pack_weight_cmap', _ptr(wd), Cout, taps, Cin, _ptr(cd), N, Cc, Coutp, _ptr(wk), _ptr(wt), st)
    ok = (cmap >= 0) & (cmap < Cin)
    idx = torch.where(ok, cmap, torch.zeros_like(cmap)).long()
    ref = w[:, :, idx]                                             # (Cout, taps, N, Cc)
    ref = (ref * ok[None, None].float()).permute(2, 0, 1, 3)
    assert torch.equal(wk.float().cpu(), ref.to(BF).float())
    ref_t = torch.zeros(N, Cc, taps, Coutp)
    ref_t[..., :Cout] = ref.permute(0, 3, 2, 1)
    assert torch.equal(wt.float().cpu(), ref_t.to(BF).float())
    # adjoint
    dwc = torch.randn(N, Cout, taps, Cc, generator=gen)
    dw = torch.full((Cout, taps, Cin), 7.0, device=DEV)           # must be overwritten
    _lib.call('sg_wgrad_cmap_scatter', _ptr(dwc.to(DEV)), _ptr(cd), N, Cout, taps, Cc, Cin, _ptr(dw), st)
    ref_dw = torch.zeros(Cout, taps, Cin)
    for n in range(N):
        for j in range(Cc):
            c = int(cmap[n, j])
            if 0 <= c < Cin:
                ref_dw[:, :, c] += dwc[n, :, :, j]
    assert torch.allclose(dw.cpu(), ref_dw, atol=1e-5)


@pytest.mark.parametrize('kind,k,pad,H', [('s1', 7, 0, 22), ('s1', 3, 1, 16), ('s2', 4, 2, 32)])
def test_conv_per_image_weights_matches_dense(kind, k, pad, H):
    """Compacted operand + cmap vs the same data scattered into the dense channel layout."""
    gen = torch.Generator().manual_seed(11)
    N, Cc, Cin, Cout, used = 3, 64, 104, 64, 40
    cmap = make_cmap(N, Cc, Cin, used, gen)
    xc = torch.zeros(N, H, H, Cc)
    xc[..., :used] = torch.randn(N, H, H, used, generator=gen)
    xc = xc.to(BF)
    xd = torch.zeros(N, H, H, Cin, dtype=BF)
    for n in range(N):
        xd[n][..., cmap[n, :used].long()] = xc[n][..., :used]
    conv = torch.nn.Conv2d(Cin, Cout, k).to(DEV)
    from scene_generation_b200.layers import channels_last_
    channels_last_(conv)
    with torch.no_grad():
        conv.weight.normal_(0, 0.05)
    res = {}
    c0, c1 = 8, used
    for tag, x, cm, dxc in (('dense', xd, None, None), ('compact', xc, cmap.to(DEV), (c0, c1))):
        conv.zero_grad()
        x = x.to(DEV).requires_grad_(True)
        if kind == 's2':
            op = Fn.to_planes_fn(x)
            spec = ConvSpec('s2', k, pad, in_hw=(H, H), stats=True, dx_channels=dxc)
        else:
            op = Fn.plain_fn(x)
            spec = ConvSpec('s1', k, pad, stats=True, dx_channels=dxc)
        y, stats = Fn.conv(op, conv.weight, conv.bias, spec, cm)
        g = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(DEV).to(BF)
        y.backward(g)
        res[tag] = (y.detach().float(), stats.clone(), conv.weight.grad.clone(), conv.bias.grad.clone(), x.grad.detach().float())
    yd, sd, wd, bd, dxd = res['dense']
    yc, sc, wc, bc, dxc_ = res['compact']
    assert (yd - yc).abs().max() <= 2e-2 * yd.abs().max()
    assert torch.allclose(sd, sc, rtol=2e-2, atol=2e-2 * float(sd.abs().max()))
    assert torch.allclose(bd, bc, rtol=1e-3, atol=1e-3)
    assert (wd - wc).abs().max() <= 2e-3 * wd.abs().max(), (wd - wc).abs().max() / wd.abs().max()
    # input gradient: compact channels [c0, c1) of image n are dense channels cmap[n, c0:c1]
    for n in range(N):
        ref = dxd[n][..., cmap[n, c0:c1].long()]
        got = dxc_[n][..., c0:c1]
        assert (ref - got).abs().max() <= 2e-2 * ref.abs().max() + 1e-6
        assert float(dxc_[n][..., :c0].abs().max()) == 0.0


def test_expand_layout_roundtrip():
    N, S, A, D = 2, 24, 32, 12 + 32
    cmap = torch.full((N, 64), -1, dtype=torch.int32)
    cmap[0, :3] = torch.tensor([5, 0, 11], dtype=torch.int32)
    cmap[1, :2] = torch.tensor([7, 5], dtype=torch.int32)
    cmap[:, S:S + A] = torch.arange(12, 12 + A, dtype=torch.int32)
    raw = torch.zeros(N, 4, 4, 64, device=DEV, dtype=BF)
    raw[0, ..., :3] = 1.0
    raw[1, ..., :2] = 2.0
    raw[..., S:S + A] = 0.5
    lay = raw.permute(0, 3, 1, 2)[:, :S + A]
    lay._sg_cmap = cmap.to(DEV)
    dense = L.expand_layout(lay, D)
    assert dense.shape == (N, D, 4, 4)
    assert float(dense[0, 5].min()) == 1.0 and float(dense[0, 0].min()) == 1.0 and float(dense[0, 11].min()) == 1.0
    assert float(dense[0, 7].abs().max()) == 0.0 and float(dense[1, 7].min()) == 2.0
    assert float(dense[:, 12:].min()) == 0.5
    assert float(dense.sum()) == float(raw.float().sum())


def _trainer(cfg, sds, compact):
    from scene_generation_b200 import args as sgargs
    from scene_generation_b200.trainer import Trainer
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'])
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    tr.model.load_state_dict(sds['g'])
    tr.obj_discriminator.load_state_dict(sds['obj'])
    tr.mask_discriminator.load_state_dict(sds['mask'])
    tr.netD.load_state_dict(sds['img'])
    tr.model.compact_layout = compact
    return tr


def _gen_step(tr, batch_cpu, seed):
    meta = synthetic.HostMeta(batch_cpu)
    batch = meta.attach([t.to(DEV) if t is not None else None for t in batch_cpu])
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = batch
    noise = cases.noise_for(seed).to(DEV)
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    try:
        random.seed(seed)
        out = tr.model(imgs, objs, triples, o2i, boxes_gt=boxes, masks_gt=masks, attributes=attrs)
    finally:
        torch.randn = orig
    imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
    tr.optimizer.step = lambda *a, **k: None
    tr.train_generator(imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, o2i, True)
    tr.optimizer_d_img.step = lambda *a, **k: None
    tr.train_image_discriminator(imgs, imgs_pred.detach(), layout, layout_wrong)
    return out


def _snapshot(tr, out):
    dense_dim = tr.model.num_objs + tr.model.rep_size
    return dict(
        imgs=out[0].detach().float(), layouts=[L.expand_layout(t, dense_dim).float() for t in out[3:6]],
        g={n: p.grad.detach().float().clone() for n, p in tr.model.named_parameters() if p.grad is not None},
        d={n: p.grad.detach().float().clone() for n, p in tr.netD.named_parameters() if p.grad is not None},
        gl=dict(tr.generator_losses.all_losses), dl=dict(tr.d_img_losses.all_losses))


def test_generator_and_image_d_step_is_bit_reproducible_and_compact_matches_dense():
    """(1) Every reduction of the CUDA path has a fixed order (no floating-point atomics): two runs of forward +
    generator step + image-discriminator step from the same weights give IDENTICAL bits — images, losses and every
    gradient.  (2) The channel-compacted layouts reproduce the dense run: same products minus the terms that multiply
    a zero channel, in a different summation order inside the first convolutions' K loop, so the comparison has fixed
    tolerances (cfg-1 — batch 2, InstanceNorm over 4x4 maps — amplifies bf16 rounding flips the most)."""
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    batch_cpu = cases.cfg1_batch()
    runs = {}
    for key, compact in (('dense2', False), ('dense', False), ('compact', True), ('compact2', True)):
        tr = _trainer(cfg, sds, compact)
        out = _gen_step(tr, batch_cpu, 21)
        assert (getattr(out[3], '_sg_cmap', None) is not None) == compact
        runs[key] = _snapshot(tr, out)
    for x, y in (('dense', 'dense2'), ('compact', 'compact2')):
        a, b = runs[x], runs[y]
        assert torch.equal(a['imgs'], b['imgs']), 'imgs_pred of two %s runs differ' % x
        assert a['gl'] == b['gl'] and a['dl'] == b['dl'], (a['gl'], b['gl'])
        for grads in ('g', 'd'):
            assert a[grads].keys() == b[grads].keys()
            bad = [n for n, ga in a[grads].items() if not torch.equal(ga, b[grads][n])]
            assert not bad, 'gradients differ between two %s runs: %s' % (x, bad[:8])
    a, b = runs['dense'], runs['compact']
    ncls = cases.CFG1['num_objs']
    for i, (la, lb) in enumerate(zip(a['layouts'], b['layouts'])):
        assert la.shape == lb.shape
        # class channels: the same sums in the same order (gt / wrong layouts scatter the given masks; the predicted
        # masks of both runs are bit-identical too, mask_net does not see the layout)
        assert torch.equal(la[:, :ncls], lb[:, :ncls])
        assert torch.equal(la[:, ncls:], lb[:, ncls:])      # appearance channels: the crop encoder is the same program
    d_ab = float((a['imgs'] - b['imgs']).abs().mean())
    print('imgs_pred mean |dense - compact| %.3e (scale %.3e)' % (d_ab, float(a['imgs'].abs().mean())))
    assert d_ab <= 2e-2, d_ab
    for name, v in a['gl'].items():
        print('G %-28s dense %.5f compact %.5f' % (name, v, b['gl'][name]))
        assert abs(v - b['gl'][name]) <= 3e-2 * abs(v) + 2e-3, (name, v, b['gl'][name])
    for name, v in a['dl'].items():
        print('D %-28s dense %.5f compact %.5f' % (name, v, b['dl'][name]))
        assert abs(v - b['dl'][name]) <= 3e-2 * abs(v) + 2e-3, (name, v, b['dl'][name])

    def cosine(x, y):
        return float(torch.dot(x.reshape(-1), y.reshape(-1)) / (x.norm() * y.norm() + 1e-30))

    rows = []
    for grads in ('g', 'd'):
        assert a[grads].keys() == b[grads].keys()
        for name, ga in a[grads].items():
            if ga.abs().max() < 1e-7 or name.endswith('.bias'):
                continue
            rows.append((grads + '.' + name, cosine(ga, b[grads][name]), float(b[grads][name].norm() / ga.norm())))
    print('\n'.join('%-60s cos(dense,compact) %.4f ratio %.3f' % r for r in rows))
    # the layers next to the losses see (almost) the same activations in both runs; deep in the 4x4 InstanceNorm stack
    # rounding flips decorrelate the two runs (that is the bf16 path's sensitivity at this size, not the compaction:
    # the per-image-weight kernels are checked exactly in test_conv_per_image_weights_matches_dense)
    by = dict((r[0], r) for r in rows)
    for name in ('g.layout_to_image.model.38.weight', 'd.scale0_layer4.0.weight', 'd.scale1_layer4.0.weight'):
        assert by[name][1] >= 0.98 and abs(by[name][2] - 1) < 0.1, by[name]
    mean_c = sum(r[1] for r in rows) / len(rows)
    assert mean_c >= 0.75, mean_c
