"""Prints the measured deviation of the CUDA path from the CPU oracle (reproducible: the CUDA path has no
floating-point atomics) for the quantities the GPU tests bound.  Run on the GPU box: python tests/diag_parity.py
Only tests/, smoke() and bench.py's CPU arm may use oracle/: this is a test-side diagnostic."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, restate as R                                   # noqa: E402
from scene_generation_b200 import args as sgargs, synthetic               # noqa: E402
from scene_generation_b200.trainer import Trainer                         # noqa: E402

DEV = 'cuda'


def trainer(cfg, sds, **over):
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'], **over)
    a.cuda_graphs = False
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    for net, k in ((tr.model, 'g'), (tr.obj_discriminator, 'obj'), (tr.mask_discriminator, 'mask'), (tr.netD, 'img')):
        net.load_state_dict(sds[k])
    return tr


def one(cfg, n_img, kmin, kmax, compact, use_gt, seed=21):
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    hb = synthetic.make_batch(n_img, (H, H), cfg['num_objs'], kmin, kmax, seed=1)
    meta = synthetic.HostMeta(hb)
    tr = trainer(cfg, sds)
    tr.model.compact_layout = compact
    batch = meta.attach(tuple(t.to(DEV) for t in hb))
    noise = cases.noise_for(seed)
    oracle = R.OracleTrainer(sds, cfg)
    random.seed(seed)
    oracle.step(hb, noise, use_gt=use_gt)
    random.seed(seed)
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.to(DEV).clone()
    try:
        out = tr.train_step(batch, use_gt=use_gt)
    finally:
        torch.randn = orig
    mine = {'g': tr.generator_losses.all_losses, 'mask': tr.d_mask_losses.all_losses, 'obj': tr.d_obj_losses.all_losses,
            'img': tr.d_img_losses.all_losses}
    worst = 0.0
    for net, terms in oracle.losses.items():
        for name, r in terms.items():
            if name in mine[net]:
                rel = (mine[net][name] - r) / (abs(r) + 1e-9)
                worst = max(worst, abs(rel))
                print('   %-5s %-26s gpu %.5f oracle %.5f rel %+.4f' % (net, name, mine[net][name], r, rel))
    f = oracle.last_forward if hasattr(oracle, 'last_forward') else None
    print('   worst loss-term deviation %.4f' % worst)
    return out, oracle


if __name__ == '__main__':
    torch.manual_seed(0)
    for tag, cfg, n, kmin, kmax in (('cfg-1 64x64 D=42', cases.CFG1, 2, 3, 3),
                                    ('cfg-2 shapes 128x128 D=204', dict(cases.CFG1, image_size=(128, 128), num_objs=172), 2, 3, 8)):
        for compact in (False, True):
            for use_gt in (True, False):
                print('== %s  compact=%s use_gt=%s' % (tag, compact, use_gt))
                one(cfg, n, kmin, kmax, compact, use_gt)
