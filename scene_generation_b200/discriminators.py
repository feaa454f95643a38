"""Object / image / mask discriminators (mirror of scene_generation/discriminators.py;
forward/backward run on libsg_b200 kernels)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import functional as Fn
from .bilinear import crop_bbox_batch
from .functional import ConvSpec, NapSpec
from .layers import GlobalAvgPool, build_cnn, channels_last_, get_norm_layer


def weights_init(m):
    """discriminators.py:57-63."""
    name = m.__class__.__name__
    if name.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif name.find('BatchNorm2d') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


class AcDiscriminator(nn.Module):
    """discriminators.py:10-36: crop CNN -> GAP -> Linear(1024) -> {real score, class scores + CE}."""

    def __init__(self, vocab, arch, normalization='none', activation='relu', padding='same', pooling='avg'):
        super().__init__()
        self.vocab = vocab
        cnn, D = build_cnn(arch=arch, normalization=normalization, activation=activation, pooling=pooling,
                           padding=padding)
        self.cnn = nn.Sequential(cnn, GlobalAvgPool(), nn.Linear(D, 1024))
        num_objects = len(vocab['object_to_idx'])
        self.real_classifier = nn.Linear(1024, 1)
        self.obj_classifier = nn.Linear(1024, num_objects)
        channels_last_(self)

    def forward(self, x, y):
        if x.dim() == 3:
            x = x[:, None]
        if not (x.dtype == torch.bfloat16 and x.shape[-1] == 8):
            x = Fn.ToNhwcFn.apply(x, 8)
        feat = self.cnn[0](x)
        lin = self.cnn[2]
        vecs = Fn.linear(Fn.GapFn.apply(feat), lin.weight, lin.bias)
        real_scores = Fn.linear(vecs, self.real_classifier.weight, self.real_classifier.bias)
        obj_scores = Fn.linear(vecs, self.obj_classifier.weight, self.obj_classifier.bias)
        ac_loss = F.cross_entropy(obj_scores, y)     # loss reduction: losses are out of kernel scope
        return real_scores, ac_loss


class AcCropDiscriminator(nn.Module):
    """discriminators.py:39-51."""

    def __init__(self, vocab, arch, normalization='none', activation='relu', object_size=64, padding='same',
                 pooling='avg'):
        super().__init__()
        self.vocab = vocab
        self.discriminator = AcDiscriminator(vocab, arch, normalization, activation, padding, pooling)
        self.object_size = object_size
        self.align_corners = False

    def forward(self, imgs, objs, boxes, obj_to_img):
        crops = crop_bbox_batch(imgs, boxes, obj_to_img, self.object_size, align_corners=self.align_corners,
                                operand=True)
        real_scores, ac_loss = self.discriminator(crops, objs)
        return real_scores, ac_loss, Fn.FeatureViewFn.apply(crops)[:, :3]


def define_D(input_nc, ndf, n_layers_D, norm='instance', use_sigmoid=False, num_D=1):
    netD = MultiscaleDiscriminator(input_nc, ndf, n_layers_D, get_norm_layer(norm), use_sigmoid, num_D)
    assert torch.cuda.is_available()
    netD.cuda()
    netD.apply(weights_init)
    return netD


def define_mask_D(input_nc, ndf, n_layers_D, norm='instance', use_sigmoid=False, num_D=1, num_objects=None):
    netD = MultiscaleMaskDiscriminator(input_nc, ndf, n_layers_D, get_norm_layer(norm), use_sigmoid, num_D, num_objects)
    assert torch.cuda.is_available()
    netD.cuda()
    netD.apply(weights_init)
    return netD


def _patch_layers(input_nc, ndf, n_layers, norm_layer, kw, extra_in=0):
    """Layer list shared by NLayerDiscriminator / NLayerMaskDiscriminator (discriminators.py:128-162,206-238)."""
    padw = int(np.ceil((kw - 1.0) / 2))
    seq = [[nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]]
    nf = ndf
    for _ in range(1, n_layers):
        nf_prev, nf = nf, min(nf * 2, 512)
        seq.append([nn.Conv2d(nf_prev, nf, kernel_size=kw, stride=2, padding=padw), norm_layer(nf), nn.LeakyReLU(0.2, True)])
    nf_prev, nf = nf, min(nf * 2, 512)
    seq.append([nn.Conv2d(nf_prev + extra_in, nf, kernel_size=kw, stride=1, padding=padw), norm_layer(nf),
                nn.LeakyReLU(0.2, True)])
    seq.append([nn.Conv2d(nf, 1, kernel_size=kw, stride=1, padding=padw)])
    return seq


def _run_patch_d(layers, x_plain, cond=None, n_cls=0, need_dx=True, dx_channels=None, cmap=None):
    """One PatchGAN column.  layers: list of nn.Sequential ([conv, (IN), (LeakyReLU)]).  x_plain: bf16 NHWC.
    Returns the list of feature maps (logical NCHW views; the last one f32)."""
    feats = []
    x = x_plain
    for j, seq in enumerate(layers):
        conv = seq[0]
        for mod in seq[1:]:
            if not isinstance(mod, (nn.InstanceNorm2d, nn.LeakyReLU)):
                raise NotImplementedError('PatchGAN columns run InstanceNorm2d + LeakyReLU (norm_D / norm_D_mask = '
                                          "'instance', the reference's default); got %s" % type(mod).__name__)
        k, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        has_norm = len(seq) > 1 and isinstance(seq[1], nn.InstanceNorm2d)
        has_act = any(isinstance(s, nn.LeakyReLU) for s in seq)
        last = j == len(layers) - 1
        if cond is not None and j == len(layers) - 2:
            x = Fn.ConcatCondFn.apply(x, cond, n_cls)
        first = j == 0
        if stride == 2:
            op = Fn.to_planes_fn(x)
            spec = dict(kind='s2', k=k, pad=pad, in_hw=tuple(x.shape[1:3]))
        else:
            op = Fn.plain_fn(x)
            spec = dict(kind='s1', k=k, pad=pad)
        if first:
            spec.update(need_dx=need_dx, dx_channels=dx_channels)
        cm = cmap if first else None          # channel-compacted input: per-image gathered weights in layer 0
        if last:
            y = Fn.conv(op, conv.weight, conv.bias, ConvSpec(out='f32_nchw', **spec), cm)
            feats.append(y)
            break
        if has_norm:
            y, st = Fn.conv(op, conv.weight, conv.bias, ConvSpec(stats=True, **spec), cm)
            x = Fn.nap(y, st, spec=NapSpec(norm='in', act=_lib.ACT_LEAKY, slope=0.2)).squeeze(1)
        else:
            x = Fn.conv(op, conv.weight, conv.bias,
                        ConvSpec(act=_lib.ACT_LEAKY if has_act else _lib.ACT_NONE, slope=0.2, **spec), cm)
        feats.append(Fn.FeatureViewFn.apply(x))
    return feats


class NLayerDiscriminator(nn.Module):
    """discriminators.py:206-245 (holder for model0..model{n+1})."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError('LSGAN (no sigmoid) is the configuration on the hot path')
        self.n_layers = n_layers
        for n, seq in enumerate(_patch_layers(input_nc, ndf, n_layers, norm_layer, kw=4)):
            setattr(self, 'model' + str(n), nn.Sequential(*seq))


class NLayerMaskDiscriminator(nn.Module):
    """discriminators.py:128-169 (holder)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False, num_objects=None):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError('LSGAN (no sigmoid) is the configuration on the hot path')
        self.n_layers = n_layers
        for n, seq in enumerate(_patch_layers(input_nc, ndf, n_layers, norm_layer, kw=3, extra_in=num_objects)):
            setattr(self, 'model' + str(n), nn.Sequential(*seq))


class MultiscaleDiscriminator(nn.Module):
    """discriminators.py:172-203.  ``forward(input)`` takes the reference's concatenated
    (N, layout+3, H, W) tensor; ``forward_pair(layout, img)`` avoids materialising the concat:
    the image is written into the spare channels of the channels-last layout buffer."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False, num_D=3):
        super().__init__()
        self.num_D, self.n_layers, self.input_nc = num_D, n_layers, input_nc
        for i in range(num_D):
            netD = NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid)
            for j in range(n_layers + 2):
                setattr(self, 'scale' + str(i) + '_layer' + str(j), getattr(netD, 'model' + str(j)))
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
        channels_last_(self)

    def _columns(self, x_plain, need_dx, dx_channels, cmap=None):
        result, cur = [], x_plain
        for i in range(self.num_D):
            layers = [getattr(self, 'scale' + str(self.num_D - 1 - i) + '_layer' + str(j)) for j in range(self.n_layers + 2)]
            result.append(_run_patch_d(layers, cur, need_dx=need_dx, dx_channels=dx_channels, cmap=cmap))
            if i != self.num_D - 1:
                cur = Fn.AvgPoolFn.apply(cur)
        return result

    def forward(self, input):
        x = Fn.ToNhwcFn.apply(input, Fn.round_up(input.shape[1], 8))
        return self._columns(x, input.requires_grad, None)

    def forward_pair(self, layout, img):
        """layout: (N,D,H,W) view produced by Model (tagged with its raw NHWC buffer) — treated as a
        constant, exactly like every gradient-carrying call site of the reference (trainer.py:249-250,
        309-319 pass detached layouts); img: f32 (N,3,H,W)."""
        raw = getattr(layout, '_sg_nhwc', None)
        cmap = getattr(layout, '_sg_cmap', None)      # channel-compacted layout: D = slots + appearance channels,
        D = layout.shape[1]                           # cmap already maps channels [D, D+3) to the image inputs
        if raw is None or raw.shape[3] < D + img.shape[1]:
            assert cmap is None
            # f32 layouts, or no spare channels behind a dense bf16 layout: materialise the concat — with the layout
            # DETACHED like on the fast path (trainer.py:249 match_layout = layout.detach(), train.py:211-215)
            return self.forward(torch.cat((layout.detach().float(), img), dim=1))
        x = Fn.ImageSlotFn.apply(raw.detach(), img, D)
        return self._columns(x, img.requires_grad, (D, D + img.shape[1]), cmap)

    def forward_pairs(self, pairs):
        """[(layout, img), ...] -> one result per pair from ONE batched pass.  Exact: every layer of a PatchGAN column
        (conv, InstanceNorm2d, LeakyReLU, AvgPool2d) acts per sample, so stacking the reference's separate
        discriminate() calls of a D step (trainer.py:309-319) along the batch changes no value; it divides the
        launch count of the step by the number of pairs and fills the small late-layer GEMMs.  Images must not
        require a gradient (the D step passes detached images)."""
        return _forward_pairs(self, pairs)


def _forward_pairs(netD, pairs):
    """MultiscaleDiscriminator.forward_pairs: see there."""
    from . import ops
    raws = [getattr(l, '_sg_nhwc', None) for l, _ in pairs]
    cmaps = [getattr(l, '_sg_cmap', None) for l, _ in pairs]
    D, C = pairs[0][0].shape[1], pairs[0][1].shape[1]
    same = all(r is not None and r.shape == raws[0].shape for r in raws) and all(l.shape[1] == D for l, _ in pairs)
    if not same or raws[0].shape[3] < D + C or any(img.requires_grad for _, img in pairs) \
            or len(set(c is None for c in cmaps)) != 1:
        return [netD.forward_pair(l, img) for l, img in pairs]
    N, H, W, Cp = raws[0].shape
    x = torch.empty((len(pairs) * N, H, W, Cp), dtype=torch.bfloat16, device=raws[0].device)
    for k, (raw, (_, img)) in enumerate(zip(raws, pairs)):
        xs = x[k * N:(k + 1) * N]
        xs.copy_(raw.detach())
        _lib.call('sg_nchw_to_nhwc', img.detach().contiguous().float().data_ptr(), 0, N, C, H, W, Cp, D, xs.data_ptr(),
                  ops._stream())
    cmap = None if cmaps[0] is None else torch.cat(cmaps, dim=0).contiguous()
    cols = netD._columns(x, False, None, cmap)
    return [[[f[k * N:(k + 1) * N] for f in col] for col in cols] for k in range(len(pairs))]


class MultiscaleMaskDiscriminator(nn.Module):
    """discriminators.py:87-125: the one-hot class vector is concatenated (broadcast over space) in front
    of the stride-1 layer."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False, num_D=3,
                 num_objects=None):
        super().__init__()
        self.num_D, self.n_layers, self.num_objects = num_D, n_layers, num_objects
        for i in range(num_D):
            netD = NLayerMaskDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid, num_objects)
            for j in range(n_layers + 2):
                setattr(self, 'scale' + str(i) + '_layer' + str(j), getattr(netD, 'model' + str(j)))
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
        channels_last_(self)

    def forward(self, input, cond):
        """input: (O,1,M,M) f32 / i64 masks; cond: (O,num_objects) one-hot or (O,) int64 class indices."""
        cls = cond if cond.dim() == 1 else cond.argmax(dim=1)
        x = Fn.ToNhwcFn.apply(input, 8)
        result, cur = [], x
        for i in range(self.num_D):
            layers = [getattr(self, 'scale' + str(self.num_D - 1 - i) + '_layer' + str(j)) for j in range(self.n_layers + 2)]
            result.append(_run_patch_d(layers, cur, cond=cls.contiguous(), n_cls=self.num_objects,
                                       need_dx=input.requires_grad, dx_channels=(0, 1)))
            if i != self.num_D - 1:
                cur = Fn.AvgPoolFn.apply(cur)
        return result
