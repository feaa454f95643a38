"""Diagnostic: eager vs captured trajectories, parameter checksums before every step (GPU box only)."""
import random
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cases, restate as R
from scene_generation_b200 import synthetic
from tests.test_gpu_graph_step import make_trainer

DEV = 'cuda'
cfg = cases.CFG1
sds = R.make_state_dicts(cfg, seed=5)
H = 64
hbs = [tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], k, k, seed=s)) for k, s in ((3, 1), (4, 2), (3, 3))]
metas = [synthetic.HostMeta(hb) for hb in hbs]
noise = cases.noise_for(21).to(DEV)


def sums(tr):
    out = {}
    for n, p in tr.model.named_parameters():
        if n.startswith('box_net') or n.startswith('gconv.net1.0') or n.startswith('obj_emb') or n.startswith('layout_to_image.model.1.'):
            out[n] = p.detach().double().abs().sum().item()
    return out


def run(graphs, order):
    tr = make_trainer(cfg, sds, graphs)
    random.seed(77)
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    log = []
    try:
        for i, bi in enumerate(order):
            before = sums(tr)
            batch = metas[bi].attach(tuple(t.to(DEV) for t in hbs[bi]))
            out = tr.train_step(batch, use_gt=(i % 2 == 0))
            log.append((before, out[1].detach().float().cpu().clone(), dict(tr.generator_losses.all_losses)))
    finally:
        torch.randn = orig
    return log

for order in ([0, 1, 2, 0, 1], [0, 1, 0, 1, 0]):
    a, b = run(False, order), run(True, order)
    print('order', order)
    for i, ((sa, ba, la), (sb, bb, lb)) in enumerate(zip(a, b)):
        dmax = max(abs(sa[k] - sb[k]) / (abs(sa[k]) + 1e-12) for k in sa)
        worst = max(sa, key=lambda k: abs(sa[k] - sb[k]) / (abs(sa[k]) + 1e-12))
        print(' step %d: max rel param-checksum diff before step %.3e (%s); boxes_pred max diff %.3e; bbox loss %s vs %s' % (
            i, dmax, worst, (ba - bb).abs().max().item(), la.get('bbox_pred'), lb.get('bbox_pred')))
