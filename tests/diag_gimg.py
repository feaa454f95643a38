"""Per-loss-term gradient w.r.t. the generated image: CUDA path vs CPU oracle autograd (cosine, norm ratio).
Feeds the SAME image (the oracle's imgs_pred) to both, so only the discriminator forward/backward paths differ.
Run on the GPU box: python tests/diag_gimg.py   (test-side diagnostic; uses oracle/)"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, restate as R                                   # noqa: E402
from scene_generation_b200 import args as sgargs, synthetic               # noqa: E402
from scene_generation_b200.trainer import Trainer                         # noqa: E402

DEV = 'cuda'


def cosr(a, b):
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), float(a.norm() / (b.norm() + 1e-30))


def main(shapes):
    if shapes == 'cfg1':
        cfg, batch_cpu = cases.CFG1, cases.cfg1_batch()
    else:
        cfg = dict(cases.CFG1, image_size=(128, 128), num_objs=172)
        batch_cpu = synthetic.make_batch(2, (128, 128), 172, 3, 8, seed=1)
    sds = R.make_state_dicts(cfg, seed=5)
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'])
    a.cuda_graphs = False
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    for net, k in ((tr.model, 'g'), (tr.obj_discriminator, 'obj'), (tr.mask_discriminator, 'mask'), (tr.netD, 'img')):
        net.load_state_dict(sds[k])
    noise = cases.noise_for(21)
    sd = {k: {n: t.clone() for n, t in v.items()} for k, v in sds.items()}
    random.seed(21)
    with torch.no_grad():
        fwd = R.model_forward(sd['g'], cfg, batch_cpu, noise, pool=R.VectorPool(100), update=False)
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = batch_cpu
    img_ref = fwd[0].detach().clone().requires_grad_(True)
    layout = fwd[3].detach()
    gl = R.generator_losses(sd['g'], sd['obj'], sd['mask'], sd['img'], cfg, batch_cpu, (img_ref, fwd[1], fwd[2], layout), True)
    ref_grads = {k: torch.autograd.grad(gl[k], img_ref, retain_graph=True)[0] for k in
                 ('ac_loss', 'g_gan_obj_loss', 'g_gan_img_loss', 'g_gan_features_loss_img')}
    # CUDA: same image, dense f32 layout (fallback concat path) and the bf16 layout slot path
    B = [t.to(DEV) for t in batch_cpu]
    for p in list(tr.obj_discriminator.parameters()) + list(tr.netD.parameters()):
        p.requires_grad_(False)
    for lay_mode in ('f32 concat', 'bf16 slot'):
        x = fwd[0].detach().to(DEV).requires_grad_(True)
        if lay_mode == 'f32 concat':
            lay = layout.to(DEV)
        else:
            from scene_generation_b200 import layout as L
            ranges = torch.from_numpy(synthetic.image_ranges(o2i)).to(DEV)
            o2i_d = B[5]
            o2i_d._sg_ranges = ranges
            lv = torch.cat([R.one_hot(objs, cfg['num_objs']), torch.zeros(objs.numel(), 32)], 1)
            # rebuild the layout from the oracle's layout vectors is not available here: reuse the dense tensor through ToNhwc
            lay = layout.to(DEV)
        sf, ac, _ = tr.obj_discriminator(x, B[1], B[2], B[5])
        terms = {'ac_loss': ac * a.ac_loss_weight, 'g_gan_obj_loss': tr.gan_g_loss(sf) * a.d_obj_weight}
        with torch.no_grad():
            pr = tr.netD.forward_pair(lay, B[0])
        pf = tr.netD.forward_pair(lay, x)
        terms['g_gan_img_loss'] = tr.criterionGAN(pf, True) * a.d_img_weight
        terms['g_gan_features_loss_img'] = tr.calculate_features_loss(pf, pr) * a.d_img_features_weight
        print('== %s  layout %s' % (shapes, lay_mode))
        for k, v in terms.items():
            g = torch.autograd.grad(v, x, retain_graph=True)[0]
            c, r = cosr(g, ref_grads[k])
            print('   %-26s loss gpu %.5f oracle %.5f   d/dimg cos %.4f norm ratio %.4f' % (k, float(v), float(gl[k]), c, r))
        # feature-matching term split by scale and layer
        fw, dw = 4.0 / len(pf[0]), 1.0 / len(pf)
        ref_pf = R.multiscale_discriminator(sd['img'], torch.cat([layout, img_ref], dim=1))
        ref_pr = R.multiscale_discriminator(sd['img'], torch.cat([layout, imgs], dim=1))
        for i in range(len(pf)):
            for j in range(len(pf[i]) - 1):
                t = torch.nn.functional.l1_loss(pf[i][j].float(), pr[i][j].detach().float())
                tr_ = (ref_pf[i][j] - ref_pr[i][j].detach()).abs().mean()
                g = torch.autograd.grad(t, x, retain_graph=True)[0]
                gr = torch.autograd.grad(tr_, img_ref, retain_graph=True)[0]
                c, r = cosr(g, gr)
                same = float(((pf[i][j].float() - pr[i][j].float()) == 0).float().mean())
                print('   FM scale %d layer %d: L1 gpu %.5f oracle %.5f   d/dimg cos %.4f ratio %.4f   exact ties %.4f' % (
                    i, j, float(t), float(tr_), c, r, same))
        break


if __name__ == '__main__':
    for s in ('cfg1', 'cfg2'):
        main(s)
