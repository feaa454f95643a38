"""Drop-in proof: the reference's own callers run against the mirrors.

* the training loop body of train.py:193-215 — restated statement for statement in oracle/ref_harness.train_iteration,
  which bench.py's reference arm runs on the reference's own Trainer — here drives OUR Trainer with a plain collate
  batch (what data/coco.py's coco_collate_fn returns: no loader metadata), including the layout[:, :num_obj] slices;
* the checkpoint written by Trainer.save_checkpoint (trainer.py:152-203 keys) is rebuilt the way
  scripts/sample_images.py:133-144 does — Model(**checkpoint['model_kwargs']), load_state_dict, eval — and sampled with
  test_mode=True, GT / predicted boxes and masks and the per-object `features` override (sample_images.py:205-222);
* crop_bbox_batch at 224 pixels on the generated images (sample_images.py:226-229).
"""
import random
import tempfile

import pytest
import torch

from oracle import cases, ref_harness, restate as R
from scene_generation_b200 import args as sgargs, layout as L, synthetic
from scene_generation_b200.bilinear import crop_bbox_batch
from scene_generation_b200.model import Model
from scene_generation_b200.trainer import Trainer

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _trainer(cfg, sds, **over):
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'], output_dir=tempfile.mkdtemp(), **over)
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    for net, k in ((tr.model, 'g'), (tr.obj_discriminator, 'obj'), (tr.mask_discriminator, 'mask'), (tr.netD, 'img')):
        net.load_state_dict(sds[k])
    return tr, a


def test_reference_training_loop_body_runs_on_the_mirrors_and_matches_the_oracle():
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    tr, a = _trainer(cfg, sds)
    batch_cpu = cases.cfg1_batch()
    batch = [t.to(DEV) for t in batch_cpu]             # train.py:192: plain tensors, no metadata
    noise = cases.noise_for(21)
    oracle = R.OracleTrainer(sds, cfg)
    random.seed(21)
    oracle.step(batch_cpu, noise, use_gt=True)
    random.seed(21)
    orig = torch.randn
    torch.randn = lambda *a_, **k: noise.to(DEV).clone()
    try:
        out = ref_harness.train_iteration(tr, batch, use_gt=True)          # train.py:193-215 on OUR Trainer
    finally:
        torch.randn = orig
    imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
    N, D = 2, cfg['num_objs'] + 32
    # the reference's shapes: dense (N, D, H, W) layouts whose first num_obj channels are the class part (train.py:203)
    assert layout.shape == (N, D, 64, 64) and layout_pred.shape == layout.shape and layout_wrong.shape == layout.shape
    assert layout[:, :tr.num_obj, :, :].shape == (N, cfg['num_objs'], 64, 64)
    assert imgs_pred.shape == (N, 3, 64, 64) and masks_pred.shape == (batch[1].numel(), 32, 32)
    mine = {'g': tr.generator_losses.all_losses, 'mask': tr.d_mask_losses.all_losses, 'obj': tr.d_obj_losses.all_losses,
            'img': tr.d_img_losses.all_losses}
    for net, terms in oracle.losses.items():
        for name, r in terms.items():
            if name in mine[net]:
                assert abs(mine[net][name] - r) <= 0.015 * abs(r) + 1e-3, (net, name, mine[net][name], r)
    # the loader-metadata fast path returns channel-compacted layouts; expand_layout gives the reference's dense view
    meta = synthetic.HostMeta(batch_cpu)
    out_c = tr.model(*[meta.attach(tuple(t.to(DEV) for t in batch_cpu))[i] for i in (0, 1, 4, 5)],
                     boxes_gt=batch[2], masks_gt=batch[3], attributes=batch[7])
    assert getattr(out_c[3], '_sg_cmap', None) is not None
    dense = L.expand_layout(out_c[3], D)
    assert dense.shape == layout.shape
    assert torch.equal(dense[:, :tr.num_obj].float(), layout[:, :tr.num_obj].float())      # same class part as the dense run
    # checkpoint in the reference's format -> scripts/sample_images.py:133-144
    checkpoint = {}
    Trainer(a, tr.vocab, checkpoint)                      # the constructor fills checkpoint['*_kwargs'] (trainer.py:31-134)
    assert {'model_kwargs', 'd_obj_kwargs', 'd_mask_kwargs', 'd_img_kwargs'} <= set(checkpoint)
    path = tr.save_checkpoint(checkpoint, 1, a, 0)
    ck = torch.load(path, weights_only=False)
    for key in ('model_state', 'optim_state', 'd_obj_state', 'd_obj_optim_state', 'd_mask_state', 'd_mask_optim_state',
                'd_img_state', 'd_img_optim_state', 'counters', 'model_kwargs'):
        assert key in ck, key
    model = Model(**ck['model_kwargs'])                   # sample_images.py:134-135
    model.load_state_dict(ck['model_state'])
    model.eval()
    model.image_size = cfg['image_size']
    model.cuda()
    imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch
    feats = {int(c): torch.randn(5, 32).numpy() for c in objs.unique()}       # features_clustered_*.npy: class -> (k, rep)
    for use_gt_boxes, use_gt_masks, use_features in ((True, True, False), (False, False, True), (True, False, True)):
        all_features = None
        if use_features:                                   # sample_images.py:207-214
            all_features = []
            for obj_name in objs:
                f = feats[obj_name.item()]
                all_features.append(torch.from_numpy(f[random.randint(0, f.shape[0] - 1), :]).type(torch.float32).cuda())
        with torch.no_grad():
            model_out = model(imgs, objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks if use_gt_masks else None,
                              attributes=torch.zeros_like(attributes), test_mode=True, use_gt_box=use_gt_boxes,
                              features=all_features)
        imgs_p, boxes_p, masks_p, _, lay, _ = model_out    # sample_images.py:222
        assert imgs_p.shape == (N, 3, 64, 64) and torch.isfinite(imgs_p).all() and lay.shape == (N, D, 64, 64)
        assert float(imgs_p.abs().max()) <= 1.0           # tanh output
        crops = crop_bbox_batch(imgs_p, boxes if use_gt_boxes else boxes_p, obj_to_img, 224)     # sample_images.py:226-229
        assert crops.shape == (objs.numel(), 3, 224, 224) and torch.isfinite(crops).all()
    # the test-mode layout of the first setting vs the oracle (GT boxes and masks, eval BatchNorm)
    with torch.no_grad():
        torch.randn, keep = (lambda *a_, **k: noise.to(DEV).clone()), torch.randn
        try:
            mo = model(imgs, objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, attributes=attributes, test_mode=True,
                       use_gt_box=True)
        finally:
            torch.randn = keep
    ref = R.model_forward({k: v.cpu() for k, v in ck['model_state'].items()}, cfg, batch_cpu, noise, pool=None, test_mode=True,
                          use_gt_box=True, train=False)
    assert (mo[4].float().cpu() - ref[4]).abs().max() <= 2e-2 * ref[4].abs().max()
    assert (mo[0].float().cpu() - ref[0]).abs().mean() <= 3e-2
