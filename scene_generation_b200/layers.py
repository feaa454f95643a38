"""Parameter containers and building blocks (mirror of scene_generation/layers.py).

The torch.nn classes below are used as *parameter holders* so that module trees — and therefore
``state_dict`` keys — are identical to the reference's; their own ``forward`` is
never called.  All arithmetic goes through :mod:`scene_generation_b200.functional`.
Convolution weights are stored physically as [Cout][kh][kw][Cin] (the K-major operand layout of the
tensor-core kernels) while keeping the reference's logical shapes.
"""
import torch
import torch.nn as nn

from . import _lib
from . import functional as Fn
from .functional import ConvSpec, NapSpec


def channels_last_(module):
    """Re-lay the 4-D weights of every Conv2d / ConvTranspose2d below `module` in place."""
    for m in module.modules():
        if isinstance(m, nn.ConvTranspose2d):
            w = m.weight.data
            m.weight = nn.Parameter(w.permute(1, 2, 3, 0).contiguous().permute(3, 0, 1, 2))
        elif isinstance(m, nn.Conv2d):
            w = m.weight.data
            m.weight = nn.Parameter(w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    return module


def parse_activation(name):
    """get_activation (layers.py:34-47) incl. its quirk: every name maps to LeakyReLU; the slope
    comes from 'leakyrelu-X', else nn.LeakyReLU's default 0.01."""
    slope = 0.01
    if name.lower().startswith('leakyrelu') and '-' in name:
        slope = float(name.split('-')[1])
    return _lib.ACT_LEAKY, slope


class GlobalAvgPool(nn.Module):
    """layers.py:82-85 (holder; the pooling kernel is sg_gap_fwd)."""

    def forward(self, x):
        return Fn.GapFn.apply(x)


class Interpolate(nn.Module):
    """layers.py:304-314 (holder: nearest x2 upsampling is folded into the operand writer)."""

    def __init__(self, size=None, scale_factor=None, mode='nearest', align_corners=None):
        super().__init__()
        self.size, self.scale_factor, self.mode, self.align_corners = size, scale_factor, mode, align_corners


def get_norm_layer(norm_type='instance'):
    """layers.py:292-301."""
    import functools
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def build_mlp(dim_list, activation='relu', batch_norm='none', dropout=0, final_nonlinearity=True):
    """layers.py:215-231: Linear + ReLU after every layer (also the last).  Returns an MLP whose
    children are indexed like the reference nn.Sequential (Linear at 0, 2, ...)."""
    if batch_norm != 'none' or dropout > 0 or activation != 'relu':
        raise NotImplementedError('only the configuration used by the model (no norm, relu) is built')
    layers = []
    for i in range(len(dim_list) - 1):
        layers.append(nn.Linear(dim_list[i], dim_list[i + 1]))
        if i < len(dim_list) - 2 or final_nonlinearity:
            layers.append(nn.ReLU())
    return MLP(*layers)


class MLP(nn.Sequential):
    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            lin = mods[i]
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x = Fn.linear(x, lin.weight, lin.bias, _lib.ACT_RELU if relu else _lib.ACT_NONE)
            i += 2 if relu else 1
        return x


def build_cnn(arch, normalization='batch', activation='relu', padding='same', pooling='max', init='default'):
    """Arch-string CNN (layers.py:128-212) for the 'CK-X-S' conv layers the model uses
    ('C4-64-2,C4-128-2,C4-256-2'): norm + activation precede every conv except the first."""
    if isinstance(arch, str):
        arch = arch.split(',')
    cur_C, first, layers = 3, True, []
    if arch and arch[0][0] == 'I':
        cur_C, arch = int(arch[0][1:]), arch[1:]
    for s in arch:
        if s[0] != 'C':
            raise NotImplementedError('layer "%s": only conv layers are on the hot path' % s)
        vals = [int(v) for v in s[1:].split('-')]
        K, next_C, stride = (vals + [1])[:3] if len(vals) == 2 else vals
        if not first:
            if normalization != 'batch':
                # the reference also accepts 'instance' / 'none' (layers.py:164-167); only its default, the
                # BatchNorm2d + LeakyReLU crop CNN (args.py:68,96), is built on the CUDA kernels
                raise NotImplementedError("crop CNNs (appearance encoder, object discriminator) run with "
                                          "normalization='batch'; got %r" % (normalization,))
            layers.append(nn.BatchNorm2d(cur_C))
            layers.append(nn.LeakyReLU(parse_activation(activation)[1]))
        first = False
        P = 0 if padding == 'valid' else (K - 1) // 2
        layers.append(nn.Conv2d(cur_C, next_C, kernel_size=K, padding=P, stride=stride))
        cur_C = next_C
    return CropCNN(*layers), cur_C


class CropCNN(nn.Sequential):
    """conv4 s2 -> [BN, LeakyReLU, conv4 s2] x2 on box crops (AppearanceEncoder / AcDiscriminator)."""

    def forward(self, crops_nhwc):
        """crops_nhwc: bf16 (B, HH, WW, 8) operand from the crop kernel.  Returns raw bf16 NHWC."""
        mods = list(self)
        x, hw = crops_nhwc, crops_nhwc.shape[1:3]
        pend = None          # (stats, bn module, slope) waiting to be applied before the next conv
        for i, m in enumerate(mods):
            if not isinstance(m, nn.Conv2d):
                continue
            assert m.stride[0] == 2, 'only the stride-2 crop CNN is built'
            if pend is None:
                op = Fn.to_planes_fn(x)
            else:
                stats, bn, slope = pend
                op = Fn.nap(x, stats, bn.weight, bn.bias, None,
                            (bn.running_mean, bn.running_var),
                            NapSpec(norm='bn' if self.training else 'bn_eval', eps=bn.eps, momentum=bn.momentum, act=_lib.ACT_LEAKY, slope=slope, planes=True))
                if self.training:
                    bn.num_batches_tracked += 1
            nxt = next((j for j in range(i + 1, len(mods)) if isinstance(mods[j], nn.Conv2d)), None)
            has_bn = nxt is not None and isinstance(mods[i + 1], nn.BatchNorm2d)
            spec = ConvSpec('s2', m.kernel_size[0], m.padding[0], in_hw=tuple(hw), stats=has_bn)
            out = Fn.conv(op, m.weight, m.bias, spec)
            x, stats = out if has_bn else (out, None)
            hw = x.shape[1:3]
            pend = (stats, mods[i + 1], mods[i + 2].negative_slope) if has_bn else None
        return x


class ResnetBlock(nn.Module):
    """layers.py:234-273 (holder).  x + IN(conv3(reflpad(ReLU(IN(conv3(reflpad(x)))))))."""

    def __init__(self, dim, padding_type, norm_layer, activation=nn.ReLU(True), use_dropout=False):
        super().__init__()
        if padding_type != 'reflect' or use_dropout:
            raise NotImplementedError('only reflect padding without dropout is used by the model')
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0),
                                        norm_layer(dim), activation, nn.ReflectionPad2d(1),
                                        nn.Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim))
