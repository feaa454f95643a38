"""GAN losses (mirror of scene_generation/losses.py:26-175) and the VGG19 feature-matching loss (losses.py:178-224).
The tiny loss reductions are kept in PyTorch — SURVEY.md §2 marks losses.py out of kernel scope (next row §8f-1); the
VGG19 stack itself (13 frozen 3x3 convolutions, 35.5 GFLOP / image) runs on the tensor-core conv kernel."""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from . import functional as Fn
from .functional import ConvSpec


def bce_loss(input, target):
    neg_abs = -input.abs()
    return (input.clamp(min=0) - input * target + (1 + neg_abs.exp()).log()).mean()


def gan_g_loss(scores_fake):
    scores_fake = scores_fake.view(-1)
    return bce_loss(scores_fake, torch.ones_like(scores_fake))


def gan_d_loss(scores_real, scores_fake):
    assert scores_real.size() == scores_fake.size()
    scores_real, scores_fake = scores_real.view(-1), scores_fake.view(-1)
    return bce_loss(scores_real, torch.ones_like(scores_real)) + bce_loss(scores_fake, torch.zeros_like(scores_fake))


def get_gan_losses(gan_type):
    if gan_type == 'gan':
        return gan_g_loss, gan_d_loss
    raise ValueError('GAN type "%s" is not on the hot path (reference default: gan)' % gan_type)


class GANLoss(nn.Module):
    """losses.py:135-175 (LSGAN: MSE of the last feature map of every scale against a constant label)."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, tensor=None):
        super().__init__()
        if not use_lsgan:
            raise NotImplementedError('no_lsgan is not the reference default')
        self.real_label, self.fake_label = target_real_label, target_fake_label

    def __call__(self, input, target_is_real):
        t = self.real_label if target_is_real else self.fake_label
        if isinstance(input[0], list):
            return sum(F.mse_loss(i[-1].float(), torch.full_like(i[-1], t, dtype=torch.float32)) for i in input)
        return F.mse_loss(input[-1].float(), torch.full_like(input[-1], t, dtype=torch.float32))


# ---------------------------------------------------------------------------------------------
# VGG19 feature matching (losses.py:178-224) — SURVEY.md §8f-2; parity: tests/test_gpu_05_vgg.py (reference golden)
# ---------------------------------------------------------------------------------------------
_VGG_SLICES = ((0,), (2, 'M', 5), (7, 'M', 10), (12, 14, 16, 'M', 19), (21, 23, 25, 'M', 28))      # losses.py:187-196
_VGG_CH = {0: (3, 64), 2: (64, 64), 5: (64, 128), 7: (128, 128), 10: (128, 256), 12: (256, 256), 14: (256, 256),
           16: (256, 256), 19: (256, 512), 21: (512, 512), 23: (512, 512), 25: (512, 512), 28: (512, 512)}


class Vgg19(nn.Module):
    """losses.py:179-209: the five slices of torchvision's vgg19().features ending at relu1_1 .. relu5_1, with the
    reference's parameter names ('slice<k>.<i>.weight').  Weights are frozen.  The reference downloads the ImageNet
    weights; here they come from `load_torchvision_state_dict` (a vgg19 state_dict file, keys 'features.<i>.*') or
    stay at a seeded He initialisation (throughput measurements, tests)."""

    def __init__(self, requires_grad=False, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        for k, items in enumerate(_VGG_SLICES):
            seq = nn.Sequential()
            for it in items:
                if it == 'M':
                    continue
                cin, cout = _VGG_CH[it]
                conv = nn.Conv2d(cin, cout, 3, padding=1)
                conv.weight.data.copy_(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
                conv.bias.data.copy_((torch.rand(cout, generator=g) - 0.5) * 0.1)
                seq.add_module(str(it), conv)
            setattr(self, 'slice%d' % (k + 1), seq)
        from .layers import channels_last_
        channels_last_(self)
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def load_torchvision_state_dict(self, sd):
        """sd: state_dict of torchvision.models.vgg19 (or of its .features): 'features.<i>.weight' / '<i>.weight'."""
        own = {}
        for k, items in enumerate(_VGG_SLICES):
            for it in items:
                if it == 'M':
                    continue
                for leaf in ('weight', 'bias'):
                    src = sd.get('features.%d.%s' % (it, leaf), sd.get('%d.%s' % (it, leaf)))
                    if src is None:
                        raise KeyError('vgg19 state_dict lacks features.%d.%s' % (it, leaf))
                    own['slice%d.%d.%s' % (k + 1, it, leaf)] = src
        return self.load_state_dict(own, strict=True)

    def forward(self, X):
        """X: (N,3,H,W) f32 -> [relu1_1, ..., relu5_1] as NCHW views of bf16 channels-last feature maps."""
        x = Fn.ToNhwcFn.apply(X, 8)
        feats = []
        for k, items in enumerate(_VGG_SLICES):
            seq = getattr(self, 'slice%d' % (k + 1))
            for it in items:
                if it == 'M':
                    x = Fn.MaxPool2Fn.apply(x)
                    continue
                conv = getattr(seq, str(it))
                x = Fn.conv(Fn.plain_fn(x), conv.weight, conv.bias, ConvSpec('s1', 3, 1, act=_lib.ACT_RELU))
            feats.append(Fn.FeatureViewFn.apply(x))
        return feats


class VGGLoss(nn.Module):
    """losses.py:212-224."""

    def __init__(self, weights_path=None):
        super().__init__()
        self.vgg = Vgg19().cuda()
        if weights_path:
            self.vgg.load_torchvision_state_dict(torch.load(weights_path, map_location='cpu'))
        self.criterion = nn.L1Loss()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]

    def forward(self, x, y):
        x_vgg = self.vgg(x)
        with torch.no_grad():
            y_vgg = self.vgg(y)
        loss = 0
        for w, a, b in zip(self.weights, x_vgg, y_vgg):
            loss = loss + w * self.criterion(a.float(), b.detach().float())
        return loss
