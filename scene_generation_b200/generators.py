"""pix2pixHD-style generator, mask regressor and appearance encoder
(mirror of scene_generation/generators.py; forward/backward run on libsg_b200 kernels)."""
import torch
import torch.nn as nn

from . import _lib
from . import functional as Fn
from .functional import ConvSpec, NapSpec
from .layers import GlobalAvgPool, Interpolate, ResnetBlock, build_cnn, channels_last_, get_norm_layer


def weights_init(m):
    """generators.py:7-13."""
    name = m.__class__.__name__
    if name.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif name.find('BatchNorm2d') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def as_nhwc_operand(x, Cp=None):
    """Accept either a raw channels-last bf16 buffer (N,H,W,Cp) tagged on the tensor by Model
    (``_sg_nhwc``), or any logical (N,C,H,W) tensor (converted)."""
    raw = getattr(x, '_sg_nhwc', None)
    if raw is not None:
        return raw
    Cp = Cp or Fn.round_up(x.shape[1], 8)
    return Fn.ToNhwcFn.apply(x.float() if x.is_floating_point() else x, Cp)


class MaskNet(nn.Sequential):
    """mask_net (generators.py:16-28): log2(mask_size) x [nearest x2, conv3x3 pad 1, BatchNorm2d, ReLU],
    then conv1x1 -> 1.  The caller's ``.squeeze(1).sigmoid()`` (model.py:106-107) is folded into the last
    epilogue when ``fused_sigmoid`` is requested."""

    def forward(self, vecs, fused_sigmoid=False):
        O = vecs.shape[0]
        C = vecs.shape[1]
        x = Fn.ToNhwcFn.apply(vecs.reshape(O, C, 1, 1).float(), Fn.round_up(C, 8))      # (O,1,1,C) bf16
        mods = list(self)
        stats = bn = None
        i = 0
        while i < len(mods) and isinstance(mods[i], Interpolate):
            conv, nbn = mods[i + 1], mods[i + 2]
            if bn is None:
                op = Fn.nap(x, spec=NapSpec(up=2))
            else:
                op = Fn.nap(x, stats, bn.weight, bn.bias, None, (bn.running_mean, bn.running_var),
                            NapSpec(norm='bn' if self.training else 'bn_eval', eps=bn.eps, momentum=bn.momentum, act=_lib.ACT_RELU, up=2))
                if self.training:
                    bn.num_batches_tracked += 1
            x, stats = Fn.conv(op, conv.weight, conv.bias, ConvSpec('s1', 3, 1, stats=True))
            bn = nbn
            i += 4
        op = Fn.nap(x, stats, bn.weight, bn.bias, None, (bn.running_mean, bn.running_var),
                    NapSpec(norm='bn' if self.training else 'bn_eval', eps=bn.eps, momentum=bn.momentum, act=_lib.ACT_RELU))
        if self.training:
            bn.num_batches_tracked += 1
        last = mods[i]
        act = _lib.ACT_SIGMOID if fused_sigmoid else _lib.ACT_NONE
        return Fn.conv(op, last.weight, last.bias, ConvSpec('s1', 1, 0, act=act, out='f32_nchw'))   # (O,1,M,M)


def mask_net(dim, mask_size):
    layers, cur = [], 1
    while cur < mask_size:
        layers += [Interpolate(scale_factor=2, mode='nearest'), nn.Conv2d(dim, dim, kernel_size=3, padding=1),
                   nn.BatchNorm2d(dim), nn.ReLU()]
        cur *= 2
    if cur != mask_size:
        raise ValueError('Mask size must be a power of 2')
    layers.append(nn.Conv2d(dim, 1, kernel_size=1))
    return channels_last_(MaskNet(*layers))


class AppearanceEncoder(nn.Module):
    """generators.py:31-48: crop CNN -> GlobalAvgPool -> Linear."""

    def __init__(self, vocab, arch, normalization='none', activation='relu', padding='same', vecs_size=1024,
                 pooling='avg'):
        super().__init__()
        self.vocab = vocab
        cnn, channels = build_cnn(arch=arch, normalization=normalization, activation=activation, pooling=pooling,
                                  padding=padding)
        self.cnn = nn.Sequential(cnn, GlobalAvgPool(), nn.Linear(channels, vecs_size))
        channels_last_(self)

    def forward(self, crops):
        """crops: bf16 NHWC operand (B,HH,WW,8) from the crop kernel, or a (B,3,HH,WW) tensor."""
        if crops.dim() == 4 and crops.shape[1] == 3 and crops.shape[-1] != 8:
            crops = Fn.ToNhwcFn.apply(crops, 8)
        feat = self.cnn[0](crops)
        pooled = Fn.GapFn.apply(feat)
        lin = self.cnn[2]
        return Fn.linear(pooled, lin.weight, lin.bias)


def define_G(input_nc, output_nc, ngf, n_downsample_global=3, n_blocks_global=9, norm='instance'):
    """generators.py:51-57 (asserts CUDA like the reference: there is no CPU path)."""
    netG = GlobalGenerator(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, get_norm_layer(norm))
    assert torch.cuda.is_available()
    netG.cuda()
    netG.apply(weights_init)
    return netG


class GlobalGenerator(nn.Module):
    """generators.py:62-91.  The nn.Sequential only holds parameters at the reference's indices;
    forward() drives the fused conv / norm kernels:
      reflpad3+conv7 | IN+ReLU -> planes | conv3 s2 ... | IN+ReLU -> reflpad1 | 9 x resblock |
      convT phases ... | IN+ReLU -> reflpad3 | conv7 + tanh (f32 NCHW)."""

    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, norm_layer=nn.BatchNorm2d,
                 padding_type='reflect'):
        assert n_blocks >= 0
        super().__init__()
        act = nn.ReLU(True)
        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0), norm_layer(ngf), act]
        for i in range(n_downsampling):
            c = ngf * 2 ** i
            model += [nn.Conv2d(c, c * 2, kernel_size=3, stride=2, padding=1), norm_layer(c * 2), act]
        c = ngf * 2 ** n_downsampling
        for _ in range(n_blocks):
            model += [ResnetBlock(c, padding_type=padding_type, activation=act, norm_layer=norm_layer)]
        for i in range(n_downsampling):
            c = ngf * 2 ** (n_downsampling - i)
            model += [nn.ConvTranspose2d(c, c // 2, kernel_size=3, stride=2, padding=1, output_padding=1),
                      norm_layer(c // 2), act]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*model)
        self.n_downsampling, self.n_blocks = n_downsampling, n_blocks
        channels_last_(self)

    def forward(self, input):
        m = self.model
        IN_RELU = dict(norm='in', act=_lib.ACT_RELU)
        x = as_nhwc_operand(input)                                           # (N,H,W,Cp) bf16
        op = Fn.nap(x, spec=NapSpec(pad=3, pad_mode=1))
        # Model tags the layout with the channel range that carries a gradient (the appearance part; the
        # one-hot class channels are constants, model.py:165-168): the first conv's dgrad is restricted to it
        y, st = Fn.conv(op, m[1].weight, m[1].bias,
                        ConvSpec('s1', 7, 0, stats=True, dx_channels=getattr(input, '_sg_grad_channels', None)),
                        getattr(input, '_sg_cmap', None))      # channel-compacted layout: per-image gathered weights
        i = 4
        for d in range(self.n_downsampling):
            hw = tuple(y.shape[1:3])
            op = Fn.nap(y, st, spec=NapSpec(planes=True, **IN_RELU))
            y, st = Fn.conv(op, m[i].weight, m[i].bias, ConvSpec('s2', 3, 1, in_hw=hw, stats=True))
            i += 3
        last_plain = self.n_blocks == 0
        op = Fn.nap(y, st, spec=NapSpec(pad=0 if last_plain else 1, pad_mode=1, **IN_RELU))
        for b in range(self.n_blocks):
            blk = m[i].conv_block
            ya, sa = Fn.conv(op, blk[1].weight, blk[1].bias, ConvSpec('s1', 3, 0, stats=True))
            opa = Fn.nap(ya, sa, spec=NapSpec(pad=1, pad_mode=1, **IN_RELU))
            yb, sb = Fn.conv(opa, blk[5].weight, blk[5].bias, ConvSpec('s1', 3, 0, stats=True))
            pad_out = 0 if b == self.n_blocks - 1 else 1
            op = Fn.nap(yb, sb, None, None, op, None, NapSpec(norm='in', pad=pad_out, pad_mode=1, res_pad=1))
            i += 1
        for d in range(self.n_downsampling):
            y, st = Fn.conv(op, m[i].weight, m[i].bias, ConvSpec('T', 3, 1, stats=True))
            final = d == self.n_downsampling - 1
            op = Fn.nap(y, st, spec=NapSpec(pad=3 if final else 0, pad_mode=1, **IN_RELU))
            i += 3
        if self.n_downsampling == 0:
            raise NotImplementedError('n_downsampling == 0 is not used by the model')
        last = m[i + 1]
        return Fn.conv(op, last.weight, last.bias, ConvSpec('s1', 7, 0, act=_lib.ACT_TANH, out='f32_nchw'))
