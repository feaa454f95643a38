"""GAN losses (mirror of scene_generation/losses.py:26-175).  Tiny reductions kept in PyTorch —
SURVEY.md §2 marks losses.py out of kernel scope (next row §8f-1)."""
import torch
import torch.nn.functional as F
from torch import nn


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
