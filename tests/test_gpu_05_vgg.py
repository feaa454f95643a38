"""VGG19 feature-matching loss (losses.py:178-224) on the tensor-core conv kernel + the 2x2 max-pool kernels vs the
reference golden (tests/golden/vgg.pt, written by the reference's own Vgg19 / VGGLoss with seeded weights)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from scene_generation_b200 import functional as Fn, losses

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_maxpool2x2_forward_backward_vs_torch():
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn((3, 16, 10, 14), generator=g)).to(torch.bfloat16)      # NCHW, many exact-zero ties
    xr = x.float().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    gy = torch.randn(yr.shape, generator=g).to(torch.bfloat16)
    yr.backward(gy.float())
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    y = Fn.MaxPool2Fn.apply(xn)
    y.backward(gy.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert torch.equal(y.detach().float().cpu().permute(0, 3, 1, 2), yr.detach())
    assert torch.equal(xn.grad.float().cpu().permute(0, 3, 1, 2), xr.grad)             # first-maximum tie rule included


def test_vgg_features_loss_and_input_gradient_vs_reference_golden():
    g = torch.load(os.path.join(GOLD, 'vgg.pt'))
    crit = losses.VGGLoss()
    sd = R.make_vgg_state_dict(seed=3)
    crit.vgg.load_torchvision_state_dict(sd)
    x = g['x'].to(DEV).requires_grad_(True)
    feats = crit.vgg(x)
    assert [tuple(f.shape[1:]) for f in feats] == [(64, 64, 64), (128, 32, 32), (256, 16, 16), (512, 8, 8), (512, 4, 4)]
    for i in range(3):       # bf16 activations through up to 13 conv layers: 3e-2 of the map's scale
        ref = g['feat%d_mean_hw' % i]
        assert (feats[i].float().mean(dim=(2, 3)).cpu() - ref).abs().max() <= 3e-2 * ref.abs().max()
    for i in (3, 4):
        ref = g['feat%d' % i]
        err = (feats[i].float().cpu() - ref).abs()
        assert err.max() <= 8e-2 * ref.abs().max() and err.mean() <= 1e-2 * ref.abs().max(), (i, err.max().item())
    loss = crit(x, g['y'].to(DEV))
    loss.backward()
    assert abs(float(loss) - float(g['loss'])) <= 2e-2 * float(g['loss'])
    cos = torch.dot(x.grad.flatten().cpu(), g['dx'].flatten()) / (x.grad.norm().cpu() * g['dx'].norm() + 1e-30)
    assert cos > 0.95, float(cos)
    assert all(p.grad is None for p in crit.vgg.parameters())          # frozen
