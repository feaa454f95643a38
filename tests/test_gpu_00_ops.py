"""HBM-bound kernels (layout scatter, graph gather/pool, box crop) through the C ABI vs the oracle
and the committed reference goldens.  fp32 paths: 1e-5 relative; bit-exact for the index gather."""
import os

import pytest
import torch

from oracle import cases, restate as R
from scene_generation_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def ranges_of(o2i):
    return torch.from_numpy(synthetic.image_ranges(o2i)).to(DEV)


def close(a, b, tol=1e-5):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def test_layout_golden_and_oracle():
    g = torch.load(os.path.join(GOLD, 'ops.pt'))
    vecs, boxes, masks, o2i = cases.layout_literals()
    out = ops.masks_to_layout_fwd(vecs.to(DEV), boxes.to(DEV), masks.to(DEV), ranges_of(o2i), 24, 20)
    close(out, g['layout_lit'])
    out = ops.masks_to_layout_fwd(vecs.to(DEV), boxes.to(DEV), masks.to(DEV), ranges_of(o2i), 24, 20, test_mode=True)
    close(out, g['layout_lit_test'])
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = cases.ragged_batch()
    lv = cases.rand((objs.numel(), 42), 3)
    pm = cases.rand((objs.numel(), 32, 32), 4, 0.0, 1.0)
    r = ranges_of(o2i)
    close(ops.masks_to_layout_fwd(lv.to(DEV), boxes.to(DEV), masks.to(DEV), r, 32, 32), g['layout_ragged_int'])
    close(ops.masks_to_layout_fwd(lv.to(DEV), boxes.to(DEV), pm.to(DEV), r, 32, 32), g['layout_ragged_float'])
    close(ops.masks_to_layout_fwd(lv.to(DEV), boxes.to(DEV), masks.to(DEV), r, 32, 32, test_mode=True), g['layout_ragged_test'])
    # bf16 channels-last variant
    o16 = ops.masks_to_layout_fwd(lv.to(DEV), boxes.to(DEV), masks.to(DEV), r, 32, 32, out_format=ops.NHWC_BF16)
    assert o16.shape == g['layout_ragged_int'].shape
    close(o16, g['layout_ragged_int'], 1e-2)
    # align_corners=True (PyTorch 1.0 semantics of the published checkpoints) vs the oracle
    ac = ops.masks_to_layout_fwd(lv.to(DEV), boxes.to(DEV), pm.to(DEV), r, 32, 32, align_corners=True)
    close(ac, R.masks_to_layout(lv, boxes, pm, o2i, 32, align_corners=True))


@pytest.mark.parametrize('H,W,kmax', [(64, 64, 8), (128, 128, 8), (32, 48, 40)])
def test_layout_fwd_bwd_vs_oracle(H, W, kmax):
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = synthetic.make_batch(3, (H, W), num_objs=20, kmin=1, kmax=kmax, seed=3)
    O = objs.numel()
    vecs = cases.rand((O, 52), 5).requires_grad_(True)
    pm = cases.rand((O, 32, 32), 6, 0.0, 1.0).requires_grad_(True)
    ref = R.masks_to_layout(vecs, boxes, pm, o2i, H, W)
    gout = cases.rand(tuple(ref.shape), 7)
    ref.backward(gout)
    r = ranges_of(o2i)
    out = ops.masks_to_layout_fwd(vecs.detach().to(DEV), boxes.to(DEV), pm.detach().to(DEV), r, H, W)
    close(out, ref.detach())
    dv, dm = ops.masks_to_layout_bwd(vecs.detach().to(DEV), boxes.to(DEV), pm.detach().to(DEV), r, H, W, gout.to(DEV),
                                     need_dmasks=True)
    close(dv, vecs.grad, 1e-4)
    close(dm, pm.grad, 1e-4)
    # bf16 NHWC gradient input
    g16 = gout.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dv16, _ = ops.masks_to_layout_bwd(vecs.detach().to(DEV), boxes.to(DEV), pm.detach().to(DEV), r, H, W, g16)
    close(dv16, vecs.grad, 2e-2)
    # gather formulation, fixed summation order: a second launch gives the same bits
    dv2, dm2 = ops.masks_to_layout_bwd(vecs.detach().to(DEV), boxes.to(DEV), pm.detach().to(DEV), r, H, W, gout.to(DEV),
                                       need_dmasks=True)
    assert torch.equal(dv, dv2) and torch.equal(dm, dm2)
    # channel-restricted adjoint (only the appearance columns of a layout vector carry a gradient): same bits there,
    # zeros elsewhere
    dvr, _ = ops.masks_to_layout_bwd(vecs.detach().to(DEV), boxes.to(DEV), pm.detach().to(DEV), r, H, W, gout.to(DEV),
                                     channels=(24, 52))
    assert torch.equal(dvr[:, 24:], dv[:, 24:]) and float(dvr[:, :24].abs().max()) == 0.0
    # align_corners=True adjoint
    v2, p2 = vecs.detach().clone().requires_grad_(True), pm.detach().clone().requires_grad_(True)
    R.masks_to_layout(v2, boxes, p2, o2i, H, W, align_corners=True).backward(gout)
    dva, dma = ops.masks_to_layout_bwd(v2.detach().to(DEV), boxes.to(DEV), p2.detach().to(DEV), r, H, W, gout.to(DEV),
                                       align_corners=True, need_dmasks=True)
    close(dva, v2.grad, 1e-4)
    close(dma, p2.grad, 1e-4)


def test_layout_degenerate_boxes_and_empty_image():
    # zero-width box -> inf/NaN grid -> contributes nothing (grid_sample zero padding); image 1 has no objects
    vecs = cases.rand((3, 10), 1)
    boxes = torch.tensor([[0.2, 0.2, 0.2, 0.6], [0.0, 0.0, 1.0, 1.0], [0.5, 0.5, 0.9, 0.5]])
    masks = torch.ones(3, 8, 8)
    o2i = torch.tensor([0, 0, 2])
    r = torch.tensor([[0, 2], [2, 2], [2, 3]], dtype=torch.int32, device=DEV)
    out = ops.masks_to_layout_fwd(vecs.to(DEV), boxes.to(DEV), masks.to(DEV), r, 16, 16)
    assert torch.isfinite(out).all()
    assert out[1].abs().max().item() == 0.0
    ref1 = R.masks_to_layout(vecs[1:2], boxes[1:2], masks[1:2], torch.tensor([0]), 16)
    close(out[0], ref1[0])


def test_gconv_gather_pool_bit_exact_and_adjoints():
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = synthetic.make_batch(4, (64, 64), num_objs=30, kmin=1, kmax=12, seed=9)
    O, T = objs.numel(), triples.size(0)
    edges = triples[:, [0, 2]].contiguous()
    obj = cases.rand((O, 163), 1)
    pred = cases.rand((T, 128), 2)
    cur = ops.gconv_gather(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    ref = torch.cat([obj[edges[:, 0]], pred, obj[edges[:, 1]]], dim=1)
    assert torch.equal(cur.cpu(), ref)                                     # bit-exact index gather
    cur16 = ops.gconv_gather(obj.to(DEV), pred.to(DEV), edges.to(DEV), out_dtype=torch.bfloat16, ld_out=456)
    assert torch.equal(cur16[:, :454].cpu(), ref.to(torch.bfloat16)) and cur16[:, 454:].abs().max().item() == 0
    H, Dout = 512, 128
    new_t = cases.rand((T, 2 * H + Dout), 3)
    ptr, src = ops.build_incidence_csr(edges.numpy(), O)
    ptr_d, src_d = torch.from_numpy(ptr).to(DEV), torch.from_numpy(src).to(DEV)
    pooled = ops.gconv_pool(new_t.to(DEV), H + Dout, ptr_d, src_d, O, H)
    refp = torch.zeros(O, H).index_add(0, edges[:, 0], new_t[:, :H]).index_add(0, edges[:, 1], new_t[:, H + Dout:])
    cnt = torch.zeros(O).index_add(0, edges[:, 0], torch.ones(T)).index_add(0, edges[:, 1], torch.ones(T)).clamp(min=1)
    refp = refp / cnt.view(-1, 1)
    assert torch.equal(pooled.cpu(), refp)                                 # same summation order as CPU scatter_add
    # adjoints vs autograd
    nt = new_t.clone().requires_grad_(True)
    rp = (torch.zeros(O, H).index_add(0, edges[:, 0], nt[:, :H]).index_add(0, edges[:, 1], nt[:, H + Dout:])) / cnt.view(-1, 1)
    dpool, dnp = cases.rand((O, H), 4), cases.rand((T, Dout), 5)
    (rp * dpool).sum().backward()
    dnt = ops.gconv_pool_bwd(dpool.to(DEV), dnp.to(DEV), edges.to(DEV), ptr_d, T, H, Dout)
    close(dnt[:, :H], nt.grad[:, :H], 1e-6)
    close(dnt[:, H + Dout:], nt.grad[:, H + Dout:], 1e-6)
    assert torch.equal(dnt[:, H:H + Dout].cpu(), dnp)
    ob = obj.clone().requires_grad_(True)
    pr = pred.clone().requires_grad_(True)
    dcur = cases.rand((T, 454), 6)
    (torch.cat([ob[edges[:, 0]], pr, ob[edges[:, 1]]], dim=1) * dcur).sum().backward()
    dobj, dpred = ops.gconv_gather_bwd(dcur.to(DEV), ptr_d, src_d, O, T, 163, 128)
    close(dobj, ob.grad, 1e-6)
    assert torch.equal(dpred.cpu(), pr.grad)


def test_crop_golden_oracle_and_backward():
    g = torch.load(os.path.join(GOLD, 'ops.pt'))
    feats, bb, b2f = cases.crop_literals()
    close(ops.crop_bbox_fwd(feats.to(DEV), bb.to(DEV), b2f.to(DEV), 8, 6), g['crop_lit'])
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = cases.ragged_batch()
    close(ops.crop_bbox_fwd(imgs.to(DEV), boxes.to(DEV), o2i.to(DEV), 32, 32), g['crop_ragged'])
    c16 = ops.crop_bbox_fwd(imgs.to(DEV), boxes.to(DEV), o2i.to(DEV), 32, 32, out_format=ops.NHWC_BF16)
    close(c16[..., :3].permute(0, 3, 1, 2), g['crop_ragged'], 1e-2)
    assert c16[..., 3:].abs().max().item() == 0
    fr = imgs.clone().requires_grad_(True)
    ref = R.crop_bbox_batch(fr, boxes, o2i, 16)
    gout = cases.rand(tuple(ref.shape), 8)
    ref.backward(gout)
    df = ops.crop_bbox_bwd(gout.to(DEV), boxes.to(DEV), o2i.to(DEV), *imgs.shape)
    close(df, fr.grad, 1e-5)
    assert torch.equal(df, ops.crop_bbox_bwd(gout.to(DEV), boxes.to(DEV), o2i.to(DEV), *imgs.shape))   # gather: fixed order
    # non-identity box -> image mapping of the reference's demo (bilinear.py:289-295), bf16 channels-last gradient
    f2 = feats.clone().requires_grad_(True)
    ref2 = R.crop_bbox_batch(f2, bb, b2f, 8, 6)
    g2 = cases.rand(tuple(ref2.shape), 9)
    ref2.backward(g2)
    close(ops.crop_bbox_bwd(g2.to(DEV), bb.to(DEV), b2f.to(DEV), *feats.shape), f2.grad, 1e-5)
    g2n = torch.zeros(g2.shape[0], g2.shape[2], g2.shape[3], 8, dtype=torch.bfloat16)
    g2n[..., :3] = g2.permute(0, 2, 3, 1).to(torch.bfloat16)
    f3 = feats.clone().requires_grad_(True)
    R.crop_bbox_batch(f3, bb, b2f, 8, 6).backward(g2n[..., :3].float().permute(0, 3, 1, 2))
    close(ops.crop_bbox_bwd(g2n.to(DEV), bb.to(DEV), b2f.to(DEV), *feats.shape, grad_format=ops.NHWC_BF16), f3.grad, 1e-5)
    close(ops.crop_bbox_fwd(imgs.to(DEV), boxes.to(DEV), o2i.to(DEV), 16, 16, align_corners=True),
          R.crop_bbox_batch(imgs, boxes, o2i, 16, align_corners=True))
