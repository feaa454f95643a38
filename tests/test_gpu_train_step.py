"""Whole training iteration (Model.forward + G step + 3 D steps + Adam) on the GPU vs the committed
golden of the UNMODIFIED reference (tests/golden/step_cfg1.pt, BASELINE configs[0]).
Tolerance: bf16 tensor-core path vs fp32 reference -> 6 % on every loss term."""
import os
import random

import pytest
import torch

from oracle import cases, restate as R
from scene_generation_b200 import args as sgargs, synthetic
from scene_generation_b200.trainer import Trainer

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def make_trainer(cfg, sds):
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'])
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    tr.model.load_state_dict(sds['g'])
    tr.obj_discriminator.load_state_dict(sds['obj'])
    tr.mask_discriminator.load_state_dict(sds['mask'])
    tr.netD.load_state_dict(sds['img'])
    return tr


@pytest.mark.parametrize('use_gt,seed', [(True, 21), (False, 22)])
def test_train_step_losses_vs_reference_golden(use_gt, seed):
    cfg = cases.CFG1
    g = torch.load(os.path.join(GOLD, 'step_cfg1.pt'))
    sds = R.make_state_dicts(cfg, seed=5)
    tr = make_trainer(cfg, sds)
    batch = [t.to(DEV) for t in cases.cfg1_batch()]
    random.seed(seed)
    noise = cases.noise_for(seed).to(DEV)
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    try:
        tr.train_step(batch, use_gt=use_gt)
    finally:
        torch.randn = orig
    tag = 'gt' if use_gt else 'nogt'
    for lm, key in ((tr.generator_losses, 'losses_g'), (tr.d_mask_losses, 'losses_mask'), (tr.d_obj_losses, 'losses_obj'),
                    (tr.d_img_losses, 'losses_img')):
        mine = lm.all_losses
        for name, ref in g['%s_%s' % (tag, key)].items():
            assert name in mine, name
            assert abs(mine[name] - ref) <= 0.06 * abs(ref) + 5e-3, (tag, name, mine[name], ref)
    # one Adam step moves every element by ~lr*sign(grad): compare the update direction with the reference's
    lr = 1e-4
    nets = {'g': tr.model, 'obj': tr.obj_discriminator, 'mask': tr.mask_discriminator, 'img': tr.netD}
    for k, ref_after in g.items():
        if not k.startswith(tag + '_after_'):
            continue
        net, name = k[len(tag + '_after_'):].split('.', 1)
        before = sds[net][name]
        after = nets[net].state_dict()[name].detach().float().cpu()
        if 'running' in name:
            assert (after - ref_after).abs().max() <= 3e-2 * max(1.0, ref_after.abs().max()), k
            continue
        du, dr = (after - before).reshape(-1), (ref_after - before).reshape(-1)
        assert du.abs().max() <= 1.5 * lr, k
        strong = dr.abs() > 0.9 * lr          # elements whose reference update is a full, unambiguous step
        if strong.sum() > 0:
            agree = (torch.sign(du[strong]) == torch.sign(dr[strong])).float().mean().item()
            assert agree > 0.9, (k, agree)


def test_two_steps_run_and_stay_finite_at_128():
    a = sgargs.default_args(image_size=(128, 128), num_objs=172)
    torch.manual_seed(0)
    tr = Trainer(a, synthetic.make_vocab(172), {})
    for step in range(2):
        batch = synthetic.make_batch(2, (128, 128), 172, seed=step, device=DEV)
        out = tr.train_step(batch, use_gt=step == 0)
        assert out[0].shape == (2, 3, 128, 128)
        assert out[3].shape == (2, 204, 128, 128)
        for lm in (tr.generator_losses, tr.d_mask_losses, tr.d_obj_losses, tr.d_img_losses):
            for name, v in lm.items():
                assert v == v and abs(v) < 1e4, (name, v)
