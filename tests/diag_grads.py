"""Diagnostic (not a test): dump GPU-side gradients of one generator step at cfg-1 for offline
comparison with the CPU oracle."""
import random
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from oracle import cases, restate as R
from scene_generation_b200 import args as sgargs, synthetic
from scene_generation_b200.trainer import Trainer
import scene_generation_b200.model as M

cfg = cases.CFG1
sds = R.make_state_dicts(cfg, seed=5)
a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'])
tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
tr.model.load_state_dict(sds['g']); tr.obj_discriminator.load_state_dict(sds['obj'])
tr.mask_discriminator.load_state_dict(sds['mask']); tr.netD.load_state_dict(sds['img'])
batch = [t.cuda() for t in cases.cfg1_batch()]
imgs, objs, boxes, masks, triples, o2i, t2i, attrs = batch
noise = cases.noise_for(21).cuda()
orig = torch.randn
torch.randn = lambda *a, **k: noise.clone()
keep = {}
orig_ccv = M.Model.create_components_vecs
def ccv(self, *args, **kw):
    out = orig_ccv(self, *args, **kw)
    out[2].retain_grad(); keep['layout_vecs'] = out[2]
    return out
M.Model.create_components_vecs = ccv
random.seed(21)
out = tr.model(imgs, objs, triples, o2i, boxes_gt=boxes, masks_gt=masks, attributes=attrs)
torch.randn = orig
imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
imgs_pred.retain_grad(); layout._sg_nhwc.retain_grad()
tr.optimizer.step = lambda *a, **k: None
tr.train_generator(imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, o2i, True)
d = {'imgs_pred': imgs_pred.detach().cpu(), 'd_imgs_pred': imgs_pred.grad.cpu(), 'd_layout': layout._sg_nhwc.grad.float().cpu(),
     'layout': layout._sg_nhwc.detach().float().cpu(), 'd_layout_vecs': keep['layout_vecs'].grad.cpu(),
     'layout_vecs': keep['layout_vecs'].detach().cpu(), 'losses': tr.generator_losses.all_losses}
torch.save(d, 'gpurun_out/diag.pt')
print('saved', {k: (tuple(v.shape) if hasattr(v, 'shape') else v) for k, v in d.items()})
