"""CPU: the oracle restatement replays the committed goldens that were produced by the UNMODIFIED
reference (oracle/gen_golden.py).  This is what pins the oracle on machines without /root/reference."""
import os
import random

import pytest
import torch

from oracle import cases, restate as R

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def close(a, b, tol=1e-5):
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


@pytest.fixture(scope='module')
def g():
    return torch.load(os.path.join(GOLD, 'ops.pt'))


@pytest.fixture(scope='module')
def sds():
    return R.make_state_dicts(cases.CFG_SMALLG, seed=11)


def test_linspace_matches_torch():
    # ATen's CUDA kernel evaluates start + step*i from both ends (the formula restated in the oracle and in
    # the CUDA kernels); ATen's vectorised CPU kernel differs from it by at most 1 ulp for steps >= 64.
    for n in (1, 2, 5, 32, 33):
        assert torch.equal(R.linspace01(n), torch.linspace(0, 1, steps=n))
        assert torch.equal(R._linspace10(n), torch.linspace(1, 0, steps=n))
    for n in (64, 128, 256):
        assert (R.linspace01(n) - torch.linspace(0, 1, steps=n)).abs().max() <= 6e-8
        assert (R._linspace10(n) - torch.linspace(1, 0, steps=n)).abs().max() <= 6e-8


def test_graph(g, sds):
    for tag, batch in (('cfg1', cases.cfg1_batch()), ('ragged', cases.ragged_batch())):
        imgs, objs, boxes, masks, triples, o2i, t2i, attrs = batch
        o, p = R.scene_graph_to_vectors(sds['g'], objs, triples, attrs)
        close(o, g['gconv_%s_obj5' % tag])
        close(p, g['gconv_%s_pred5' % tag])


def test_layout_and_crop(g):
    vecs, boxes, masks, o2i = cases.layout_literals()
    close(R.masks_to_layout(vecs, boxes, masks, o2i, 24, 20), g['layout_lit'])
    close(R.masks_to_layout(vecs, boxes, masks, o2i, 24, 20, test_mode=True), g['layout_lit_test'])
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = cases.ragged_batch()
    lv = cases.rand((objs.numel(), 42), 3)
    pm = cases.rand((objs.numel(), 32, 32), 4, 0.0, 1.0)
    close(R.masks_to_layout(lv, boxes, masks, o2i, 32), g['layout_ragged_int'])
    close(R.masks_to_layout(lv, boxes, pm, o2i, 32), g['layout_ragged_float'])
    close(R.masks_to_layout(lv, boxes, masks, o2i, 32, test_mode=True), g['layout_ragged_test'])
    feats, bb, b2f = cases.crop_literals()
    close(R.crop_bbox_batch(feats, bb, b2f, 8, 6), g['crop_lit'])
    close(R.crop_bbox_batch(imgs, boxes, o2i, 32), g['crop_ragged'])


def test_conv_stacks(g, sds):
    sg = sds['g']
    x = cases.rand((2, 42, 64, 64), 5, 0.0, 1.0)
    close(R.global_generator(sg, x, n_blocks=2), g['generator_small'], 1e-4)
    sd = {k: v.clone() for k, v in sg.items()}
    close(R.mask_net(sd, cases.rand((8, 192), 6), update=True), g['mask_net'], 1e-4)
    close(sd['mask_net.2.running_var'], g['mask_net_running_var'])
    close(R.appearance_encoder(sg, cases.rand((8, 3, 64, 64), 8)), g['appearance_encoder'], 1e-4)
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = cases.ragged_batch()
    rs, ac, _ = R.ac_crop_discriminator(sds['obj'], imgs, objs, boxes, o2i)
    close(rs, g['objd_scores'], 1e-4)
    close(ac.view(1), g['objd_ac'].view(1), 1e-4)
    fd = R.multiscale_discriminator(sds['img'], cases.rand((2, 45, 64, 64), 9))
    for i in range(2):
        for j in range(5):
            close(fd[i][j], g['netD_%d_%d' % (i, j)], 1e-4)
    pm = cases.rand((objs.numel(), 32, 32), 4, 0.0, 1.0)
    fm = R.mask_discriminator(sds['mask'], pm.unsqueeze(1), R.one_hot(objs, 10))
    for j in range(4):
        close(fm[0][j], g['maskD_%d' % j], 1e-4)


def test_full_train_step_vs_reference_golden():
    """Model.forward + G step + three D steps + Adam, BASELINE configs[0] (fp32 CPU, ~30 s)."""
    g = torch.load(os.path.join(GOLD, 'step_cfg1.pt'))
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    ot = R.OracleTrainer(sds, cfg)
    random.seed(21)
    fwd = ot.step(cases.cfg1_batch(), cases.noise_for(21), use_gt=True)
    close(fwd[0].detach(), g['gt_imgs_pred'], 2e-4)
    close(fwd[2].detach(), g['gt_masks_pred'], 2e-5)
    for net, key in (('g', 'losses_g'), ('mask', 'losses_mask'), ('obj', 'losses_obj'), ('img', 'losses_img')):
        for name, val in g['gt_' + key].items():
            assert abs(ot.losses[net][name] - val) <= 2e-4 * max(1.0, abs(val)), (net, name)
    for k, v in g.items():
        if k.startswith('gt_after_'):
            net, name = k[len('gt_after_'):].split('.', 1)
            d = (ot.sd[net][name].detach().float() - v.float()).abs()
            assert d.max().item() <= 2.2e-4 + 1e-6, k


def test_three_iteration_trajectory_vs_reference_golden():
    """tests/golden/traj_cfg1.pt: loss terms of three consecutive iterations of the UNMODIFIED reference (use_gt
    alternating).  The oracle must follow it — this pins what carries over between iterations (Adam moments, BatchNorm
    running statistics, VectorPool contents and its python-random stream).  Tolerances per iteration as measured when
    the golden was written (two fp32 CPU programs with different summation orders separate under Adam's sign-like
    first steps): 2e-4, 1e-3, 1e-2 of max(1, |reference|)."""
    import random
    g = torch.load(os.path.join(GOLD, 'traj_cfg1.pt'))
    cfg = cases.CFG1
    ot = R.OracleTrainer(R.make_state_dicts(cfg, seed=5), cfg)
    batch = cases.cfg1_batch()
    random.seed(g['seed'])
    tol = (2e-4, 1e-3, 1e-2)
    for i in range(g['steps']):
        ot.step(batch, cases.noise_for(g['noise_seed']), use_gt=(i % 2 == 0))
        for net, terms in g['losses'][i].items():
            for name, val in terms.items():
                mine = ot.losses[net][name]
                assert abs(mine - val) <= tol[i] * max(1.0, abs(val)), (i, net, name, mine, val)
    # the box loss must have moved between the two use_gt iterations (it barely does when stale weights are used)
    assert g['losses'][2]['g']['bbox_pred'] < 0.95 * g['losses'][0]['g']['bbox_pred']


def test_vgg_feature_loss_vs_reference_golden():
    """losses.py:178-224 (Vgg19 slices relu1_1..relu5_1, VGGLoss weights 1/32..1) with seeded random weights: features,
    loss and d loss / d x of the reference's own classes (tests/golden/vgg.pt)."""
    g = torch.load(os.path.join(GOLD, 'vgg.pt'))
    sd = R.make_vgg_state_dict(seed=3)
    x = g['x'].clone().requires_grad_(True)
    feats = R.vgg19_features(sd, x)
    assert [f.shape[1] for f in feats] == [64, 128, 256, 512, 512]
    assert [f.shape[2] for f in feats] == [64, 32, 16, 8, 4]
    for i in range(3):
        assert torch.allclose(feats[i].mean(dim=(2, 3)), g['feat%d_mean_hw' % i], atol=1e-5)
        assert torch.allclose(feats[i].mean(dim=1), g['feat%d_mean_c' % i], atol=1e-5)
    assert torch.allclose(feats[3], g['feat3'], atol=1e-5) and torch.allclose(feats[4], g['feat4'], atol=1e-5)
    loss = R.vgg_loss(sd, x, g['y'])
    loss.backward()
    assert abs(float(loss) - float(g['loss'])) <= 1e-6
    assert torch.allclose(x.grad, g['dx'], atol=1e-7)
