"""CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

This module is the parity checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``scene_generation_b200/`` imports it.

Every function restates one piece of /root/reference/scene_generation (file:line in the
docstring) in plain fp32 torch-CPU / numpy arithmetic over a flat ``state_dict`` that uses
the reference's parameter names, so the same weights drive the reference, this oracle and
the CUDA modules.  Dense contractions (conv / linear) use torch's fp32 CPU kernels — the
"plain PyTorch fp32 reference" for the floating point kernels; index arithmetic, bilinear
sampling, pooling, normalisation and the losses are written out by hand.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against *the reference itself executed in the build container*
(``oracle/gen_golden.py`` → ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays
them without the reference).
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------


def linspace01(steps, dtype=torch.float32):
    """torch.linspace(0, 1, steps) written out: fp32 step, symmetric evaluation from both
    ends (what ATen does on CPU and CUDA); layout.py:114-115, bilinear.py:263-265."""
    if steps == 1:
        return torch.zeros(1, dtype=dtype)
    step = np.float32(1.0) / np.float32(steps - 1)
    out = np.empty(steps, dtype=np.float32)
    half = steps // 2
    for i in range(steps):
        if i < half:
            out[i] = np.float32(0.0) + step * np.float32(i)
        else:
            out[i] = np.float32(1.0) - step * np.float32(steps - i - 1)
    return torch.from_numpy(out).to(dtype)


def linear(x, sd, prefix):
    return x @ sd[prefix + '.weight'].t() + sd[prefix + '.bias']


def mlp2(x, sd, prefix):
    """build_mlp with two Linear layers and ReLU after BOTH (final_nonlinearity=True);
    layers.py:215-231.  Sequential indices 0 and 2."""
    h = torch.relu(linear(x, sd, prefix + '.0'))
    return torch.relu(linear(h, sd, prefix + '.2'))


def leaky(x, slope=0.2):
    return torch.where(x >= 0, x, x * slope)


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d(affine=False): per-(n,c) biased variance over H*W; layers.py:295."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def batch_norm_train(x, sd, prefix, eps=1e-5, momentum=0.1, update=False):
    """nn.BatchNorm2d in train mode: batch statistics (biased var) for the output; with
    `update` the running stats in `sd` are advanced in place with the UNBIASED var;
    generators.py:22, layers.py:26."""
    n = x.numel() // x.size(1)
    mean = x.mean(dim=(0, 2, 3))
    var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
    if update:
        with torch.no_grad():
            unbiased = var * (n / max(n - 1, 1))
            sd[prefix + '.running_mean'].mul_(1 - momentum).add_(momentum * mean)
            sd[prefix + '.running_var'].mul_(1 - momentum).add_(momentum * unbiased)
            sd[prefix + '.num_batches_tracked'].add_(1)
    xhat = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    return xhat * sd[prefix + '.weight'].view(1, -1, 1, 1) + sd[prefix + '.bias'].view(1, -1, 1, 1)


def batch_norm_eval(x, sd, prefix, eps=1e-5):
    m = sd[prefix + '.running_mean'].view(1, -1, 1, 1)
    v = sd[prefix + '.running_var'].view(1, -1, 1, 1)
    return (x - m) / torch.sqrt(v + eps) * sd[prefix + '.weight'].view(1, -1, 1, 1) + \
        sd[prefix + '.bias'].view(1, -1, 1, 1)


def reflect_pad(x, p):
    """nn.ReflectionPad2d(p) by index mirroring (no edge repeat); generators.py:69,87."""
    H, W = x.shape[-2:]
    ih = torch.tensor([abs(i) if i < H else 2 * (H - 1) - i for i in range(-p, H + p)])
    iw = torch.tensor([abs(i) if i < W else 2 * (W - 1) - i for i in range(-p, W + p)])
    return x[..., ih, :][..., iw]


def upsample_nearest2(x):
    """Interpolate(scale_factor=2, mode='nearest'); layers.py:304-314."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def avg_pool_3x3_s2(x):
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False); discriminators.py:99,184."""
    N, C, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    xp = F.pad(x, (1, 1, 1, 1))
    ones = F.pad(torch.ones(1, 1, H, W, dtype=x.dtype), (1, 1, 1, 1))
    s = torch.zeros(N, C, Ho, Wo, dtype=x.dtype)
    cnt = torch.zeros(1, 1, Ho, Wo, dtype=x.dtype)
    for dy in range(3):
        for dx in range(3):
            s = s + xp[:, :, dy:dy + 2 * Ho:2, dx:dx + 2 * Wo:2]
            cnt = cnt + ones[:, :, dy:dy + 2 * Ho:2, dx:dx + 2 * Wo:2]
    return s / cnt


# --------------------------------------------------------------------------------------
# graph.py
# --------------------------------------------------------------------------------------


def graph_triple_conv(sd, prefix, obj_vecs, pred_vecs, edges, hidden=512, dout=128):
    """GraphTripleConv.forward (graph.py:58-122), pooling='avg'.
    gather rows by edge index -> [s,p,o] concat -> net1 -> split -> per-object sum of the
    s- and o-messages -> divide by clamp(count,1) -> net2."""
    O, T = obj_vecs.size(0), pred_vecs.size(0)
    s_idx, o_idx = edges[:, 0], edges[:, 1]
    cur = torch.cat([obj_vecs[s_idx], pred_vecs, obj_vecs[o_idx]], dim=1)        # :79-84 bit-exact gather
    new_t = mlp2(cur, sd, prefix + '.net1')                                        # :85
    new_s, new_p, new_o = new_t[:, :hidden], new_t[:, hidden:hidden + dout], new_t[:, hidden + dout:]
    pooled = torch.zeros(O, hidden, dtype=obj_vecs.dtype)
    pooled = pooled.index_add(0, s_idx, new_s)                                     # :100 (all s first,
    pooled = pooled.index_add(0, o_idx, new_o)                                     # :101  then all o)
    counts = torch.zeros(O, dtype=obj_vecs.dtype)
    counts = counts.index_add(0, s_idx, torch.ones(T)).index_add(0, o_idx, torch.ones(T))
    pooled = pooled / counts.clamp(min=1).view(-1, 1)                              # :115-116
    return mlp2(pooled, sd, prefix + '.net2'), new_p                               # :120


def scene_graph_to_vectors(sd, objs, triples, attributes, num_gconv_layers=5):
    """Model.scene_graph_to_vectors (model.py:126-143)."""
    s, p, o = triples[:, 0], triples[:, 1], triples[:, 2]
    edges = torch.stack([s, o], dim=1)
    obj_vecs = sd['obj_embeddings.weight'][objs]
    pred_vecs = sd['pred_embeddings.weight'][p]
    if attributes is not None:
        obj_vecs = torch.cat([obj_vecs, attributes], dim=1)
    obj_vecs, pred_vecs = graph_triple_conv(sd, 'gconv', obj_vecs, pred_vecs, edges)
    for i in range(num_gconv_layers - 1):
        obj_vecs, pred_vecs = graph_triple_conv(sd, 'gconv_net.gconvs.%d' % i, obj_vecs, pred_vecs, edges)
    return obj_vecs, pred_vecs


# --------------------------------------------------------------------------------------
# bilinear sampling shared by layout.py and bilinear.py (F.grid_sample, zeros padding)
# --------------------------------------------------------------------------------------


def _unnormalize(coord, size, align_corners):
    if align_corners:
        return ((coord + 1) / 2) * (size - 1)
    return ((coord + 1) * size - 1) / 2


def grid_sample_bilinear(inp, gx, gy, align_corners=False):
    """F.grid_sample(inp, grid) restated: bilinear, padding_mode='zeros'.
    inp (C,Hin,Win); gx, gy (Hout,Wout) normalised coords.  Corner weights and the
    nw,ne,sw,se accumulation order follow ATen's grid_sampler_2d."""
    C, Hin, Win = inp.shape
    ix = _unnormalize(gx, Win, align_corners)
    iy = _unnormalize(gy, Hin, align_corners)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    out = torch.zeros(C, *gx.shape, dtype=inp.dtype)
    flat = inp.reshape(C, -1)
    for xx, yy, ww in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        ok = (xx >= 0) & (xx <= Win - 1) & (yy >= 0) & (yy <= Hin - 1) & torch.isfinite(ix) & torch.isfinite(iy)
        xi = torch.where(ok, xx, torch.zeros_like(xx)).long()
        yi = torch.where(ok, yy, torch.zeros_like(yy)).long()
        v = flat[:, (yi * Win + xi).reshape(-1)].reshape(C, *gx.shape)
        out = out + torch.where(ok, ww, torch.zeros_like(ww)).unsqueeze(0) * v
    return out


# --------------------------------------------------------------------------------------
# layout.py
# --------------------------------------------------------------------------------------


def boxes_to_grid(boxes, H, W):
    """_boxes_to_grid (layout.py:96-128): gx[o,w] = 2*(lin(W)[w]-x0)/(x1-x0)-1, gy likewise."""
    x0, y0, x1, y1 = boxes[:, 0:1], boxes[:, 1:2], boxes[:, 2:3], boxes[:, 3:4]
    X = (linspace01(W).view(1, W) - x0) / (x1 - x0)
    Y = (linspace01(H).view(1, H) - y0) / (y1 - y0)
    return X * 2 - 1, Y * 2 - 1      # (O,W), (O,H)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, test_mode=False, align_corners=False):
    """masks_to_layout + _pool_samples (layout.py:64-93, 131-184), pooling='sum'.
    Train branch: out[n] = sum_{o in n} grid_sample(vecs[o] (x) mask_o).  Test branch: objects
    sorted by total sampled mass (ascending, stable), painted where nothing was painted yet
    and the clean mask sample > 0.5 (layout.py:157-169)."""
    O, D = vecs.shape
    M = masks.size(1)
    W = H if W is None else W
    gx, gy = boxes_to_grid(boxes, H, W)
    o2i = obj_to_img.tolist()
    N = max(o2i) + 1
    out = torch.zeros(N, D, H, W, dtype=vecs.dtype)
    mf = masks.to(vecs.dtype)
    if not test_mode:
        for o in range(O):
            img_in = vecs[o].view(D, 1, 1) * mf[o].view(1, M, M)                   # :83
            GX = gx[o].view(1, W).expand(H, W)
            GY = gy[o].view(H, 1).expand(H, W)
            out[o2i[o]] = out[o2i[o]] + grid_sample_bilinear(img_in, GX, GY, align_corners)
        return out
    for n in range(N):
        idxs = [o for o in range(O) if o2i[o] == n]
        sampled, clean, mass = {}, {}, []
        for o in idxs:
            GX = gx[o].view(1, W).expand(H, W)
            GY = gy[o].view(H, 1).expand(H, W)
            sampled[o] = grid_sample_bilinear(vecs[o].view(D, 1, 1) * mf[o].view(1, M, M), GX, GY, align_corners)
            clean[o] = grid_sample_bilinear(mf[o].view(1, M, M), GX, GY, align_corners)[0]
            mass.append(float(sampled[o].sum()))
        order = np.argsort(mass)                                                   # :160-161
        occ = torch.zeros(H, W, dtype=vecs.dtype)
        for j in order:
            o = idxs[j]
            paint = (occ == 0).to(vecs.dtype) * (clean[o] > 0.5).to(vecs.dtype)   # :165
            occ = occ + paint
            out[n] = out[n] + sampled[o] * paint
    return out


# --------------------------------------------------------------------------------------
# bilinear.py
# --------------------------------------------------------------------------------------


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, align_corners=False):
    """crop_bbox_batch -> crop_bbox_batch_cudnn -> crop_bbox -> tensor_linspace
    (bilinear.py:26-130, 246-275).  crops[b] = grid_sample(feats[bbox_to_feats[b]], grid_b),
    grid x = (1-t)*(2*x0-1) + t*(2*x1-1), t = linspace(0,1,WW); result is in box order."""
    WW = HH if WW is None else WW
    B = bbox.size(0)
    C = feats.size(1)
    bb = 2 * bbox - 1                                                              # :121
    lin_w1, lin_w0 = linspace01(WW), linspace01(WW).flip(0)
    lin_h1, lin_h0 = linspace01(HH), linspace01(HH).flip(0)
    # torch.linspace(1,0,steps) is not bitwise flip(linspace(0,1)); restate it directly
    lin_w0 = _linspace10(WW)
    lin_h0 = _linspace10(HH)
    out = torch.zeros(B, C, HH, WW, dtype=feats.dtype)
    for b in range(B):
        X = lin_w0 * bb[b, 0] + lin_w1 * bb[b, 2]                                  # :274
        Y = lin_h0 * bb[b, 1] + lin_h1 * bb[b, 3]
        out[b] = grid_sample_bilinear(feats[int(bbox_to_feats[b])], X.view(1, WW).expand(HH, WW),
                                      Y.view(HH, 1).expand(HH, WW), align_corners)
    return out


def _linspace10(steps):
    """torch.linspace(1, 0, steps) (bilinear.py:263)."""
    if steps == 1:
        return torch.ones(1)
    step = (np.float32(0.0) - np.float32(1.0)) / np.float32(steps - 1)
    out = np.empty(steps, dtype=np.float32)
    half = steps // 2
    for i in range(steps):
        if i < half:
            out[i] = np.float32(1.0) + step * np.float32(i)
        else:
            out[i] = np.float32(0.0) - step * np.float32(steps - i - 1)
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------
# generators.py / layers.py
# --------------------------------------------------------------------------------------


def global_generator(sd, x, prefix='layout_to_image.model', n_down=4, n_blocks=9):
    """GlobalGenerator.forward (generators.py:62-91) with ResnetBlock (layers.py:234-273):
    reflpad3+conv7 -> IN -> ReLU; n_down x [conv3 s2 p1, IN, ReLU]; n_blocks x resblock;
    n_down x [convT k3 s2 p1 op1, IN, ReLU]; reflpad3+conv7 -> tanh."""
    i = 1
    h = F.conv2d(reflect_pad(x, 3), sd['%s.%d.weight' % (prefix, i)], sd['%s.%d.bias' % (prefix, i)])
    h = torch.relu(instance_norm(h))
    i = 4
    for _ in range(n_down):
        h = F.conv2d(h, sd['%s.%d.weight' % (prefix, i)], sd['%s.%d.bias' % (prefix, i)], stride=2, padding=1)
        h = torch.relu(instance_norm(h))
        i += 3
    for _ in range(n_blocks):
        p = '%s.%d.conv_block' % (prefix, i)
        r = F.conv2d(reflect_pad(h, 1), sd[p + '.1.weight'], sd[p + '.1.bias'])
        r = torch.relu(instance_norm(r))
        r = F.conv2d(reflect_pad(r, 1), sd[p + '.5.weight'], sd[p + '.5.bias'])
        h = h + instance_norm(r)
        i += 1
    for _ in range(n_down):
        h = F.conv_transpose2d(h, sd['%s.%d.weight' % (prefix, i)], sd['%s.%d.bias' % (prefix, i)],
                               stride=2, padding=1, output_padding=1)
        h = torch.relu(instance_norm(h))
        i += 3
    i += 1
    h = F.conv2d(reflect_pad(h, 3), sd['%s.%d.weight' % (prefix, i)], sd['%s.%d.bias' % (prefix, i)])
    return torch.tanh(h)


def mask_net(sd, vecs, prefix='mask_net', mask_size=32, update=False, train=True):
    """mask_net (generators.py:16-28): log2(mask_size) x [nearest x2, conv3 p1, BN, ReLU], conv1x1."""
    h = vecs.view(vecs.size(0), -1, 1, 1)
    i, cur = 0, 1
    while cur < mask_size:
        h = upsample_nearest2(h)
        h = F.conv2d(h, sd['%s.%d.weight' % (prefix, i + 1)], sd['%s.%d.bias' % (prefix, i + 1)], padding=1)
        bn = '%s.%d' % (prefix, i + 2)
        h = batch_norm_train(h, sd, bn, update=update) if train else batch_norm_eval(h, sd, bn)
        h = torch.relu(h)
        i += 4
        cur *= 2
    return F.conv2d(h, sd['%s.%d.weight' % (prefix, i)], sd['%s.%d.bias' % (prefix, i)])


def crop_cnn(sd, x, prefix, update=False, train=True, slope=0.2):
    """build_cnn('C4-64-2,C4-128-2,C4-256-2', norm='batch', act='leakyrelu-0.2', padding='valid')
    (layers.py:128-212): conv, then [BN, LeakyReLU, conv] x2 — norm/act precede convs 2 and 3.
    Sequential indices: conv 0, BN 1, act 2, conv 3, BN 4, act 5, conv 6."""
    h = F.conv2d(x, sd[prefix + '.0.weight'], sd[prefix + '.0.bias'], stride=2)
    for bn_i, conv_i in ((1, 3), (4, 6)):
        bn = '%s.%d' % (prefix, bn_i)
        h = batch_norm_train(h, sd, bn, update=update) if train else batch_norm_eval(h, sd, bn)
        h = leaky(h, slope)
        h = F.conv2d(h, sd['%s.%d.weight' % (prefix, conv_i)], sd['%s.%d.bias' % (prefix, conv_i)], stride=2)
    return h


def appearance_encoder(sd, crops, prefix='image_encoder', update=False, train=True):
    """AppearanceEncoder (generators.py:31-48): cnn -> GlobalAvgPool -> Linear."""
    h = crop_cnn(sd, crops, prefix + '.cnn.0', update=update, train=train)
    return linear(h.mean(dim=(2, 3)), sd, prefix + '.cnn.2')


# --------------------------------------------------------------------------------------
# discriminators.py
# --------------------------------------------------------------------------------------


def ac_discriminator(sd, crops, objs, prefix='discriminator', update=False, train=True):
    """AcDiscriminator.forward (discriminators.py:27-36)."""
    h = crop_cnn(sd, crops, prefix + '.cnn.0', update=update, train=train)
    vecs = linear(h.mean(dim=(2, 3)), sd, prefix + '.cnn.2')
    real = linear(vecs, sd, prefix + '.real_classifier')
    scores = linear(vecs, sd, prefix + '.obj_classifier')
    logp = scores - torch.logsumexp(scores, dim=1, keepdim=True)
    ac = -logp[torch.arange(objs.numel()), objs].mean()                             # F.cross_entropy
    return real, ac


def ac_crop_discriminator(sd, imgs, objs, boxes, obj_to_img, object_size=32, update=False, train=True,
                          align_corners=False):
    """AcCropDiscriminator.forward (discriminators.py:48-51)."""
    crops = crop_bbox_batch(imgs, boxes, obj_to_img, object_size, align_corners=align_corners)
    real, ac = ac_discriminator(sd, crops, objs, update=update, train=train)
    return real, ac, crops


def _nlayer_d(sd, scale, x, n_layers=3):
    """NLayerDiscriminator via the scale{i}_layer{j} aliases (discriminators.py:206-245):
    conv4 s2 p2 + LReLU; (n_layers-1) x [conv4 s2 p2, IN, LReLU]; conv4 s1 p2, IN, LReLU; conv4 s1 p2."""
    feats = []
    h = x
    for j in range(n_layers + 2):
        p = 'scale%d_layer%d.0' % (scale, j)
        stride = 2 if j < n_layers else 1
        h = F.conv2d(h, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=2)
        if 0 < j < n_layers + 1:
            h = instance_norm(h)
        if j < n_layers + 1:
            h = leaky(h, 0.2)
        feats.append(h)
    return feats


def multiscale_discriminator(sd, x, num_D=2, n_layers=3):
    """MultiscaleDiscriminator.forward (discriminators.py:192-203): scale num_D-1 sees the
    full-resolution input first; input is avg-pooled between scales."""
    res, cur = [], x
    for i in range(num_D):
        res.append(_nlayer_d(sd, num_D - 1 - i, cur, n_layers))
        if i != num_D - 1:
            cur = avg_pool_3x3_s2(cur)
    return res


def mask_discriminator(sd, x, cond, num_D=1, n_layers=2):
    """MultiscaleMaskDiscriminator (discriminators.py:87-169): conv3 s2 p1 + LReLU;
    [conv3 s2 p1, IN, LReLU] x (n_layers-1); concat one-hot cond broadcast over space;
    conv3 s1 p1, IN, LReLU; conv3 s1 p1."""
    res, cur = [], x
    for i in range(num_D):
        scale = num_D - 1 - i
        feats, h = [], cur
        for j in range(n_layers + 2):
            p = 'scale%d_layer%d.0' % (scale, j)
            if j == n_layers:
                a, _, c, d = h.shape
                h = torch.cat([h, cond.view(a, -1, 1, 1).expand(-1, -1, c, d)], dim=1)
            stride = 2 if j < n_layers else 1
            h = F.conv2d(h, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=1)
            if 0 < j < n_layers + 1:
                h = instance_norm(h)
            if j < n_layers + 1:
                h = leaky(h, 0.2)
            feats.append(h)
        res.append(feats)
        if i != num_D - 1:
            cur = avg_pool_3x3_s2(cur)
    return res


# --------------------------------------------------------------------------------------
# losses.py
# --------------------------------------------------------------------------------------


def bce_logits(x, target):
    """bce_loss (losses.py:26-44)."""
    return (x.clamp(min=0) - x * target + torch.log(1 + torch.exp(-x.abs()))).mean()


def gan_g_loss(scores_fake):
    return bce_logits(scores_fake.reshape(-1), 1.0)                                 # losses.py:58-69


def gan_d_loss(scores_real, scores_fake):
    return bce_logits(scores_real.reshape(-1), 1.0) + bce_logits(scores_fake.reshape(-1), 0.0)   # :72-88


def lsgan_loss(preds, target_is_real):
    """GANLoss.__call__ with use_lsgan (losses.py:135-175): sum over scales of MSE(last map, label)."""
    t = 1.0 if target_is_real else 0.0
    return sum(((p[-1] - t) ** 2).mean() for p in preds)


def features_loss(pred_fake, pred_real):
    """Trainer.calculate_features_loss (trainer.py:331-340)."""
    loss = 0
    fw, dw = 4.0 / len(pred_fake[0]), 1.0 / len(pred_fake)
    for i in range(len(pred_fake)):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + dw * fw * (pred_fake[i][j] - pred_real[i][j].detach()).abs().mean()
    return loss


def one_hot(objs, num_objs, dtype=torch.float32):
    oh = torch.zeros(objs.numel(), num_objs, dtype=dtype)
    oh[torch.arange(objs.numel()), objs] = 1.0
    return oh


# --------------------------------------------------------------------------------------
# utils.py VectorPool
# --------------------------------------------------------------------------------------


class VectorPool:
    """utils.py:62-90 restated (python `random` draws in the same order)."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.vectors = {}

    def query(self, objs, vectors):
        if self.pool_size == 0:
            return vectors
        out = []
        for obj, vec in zip(objs.tolist(), vectors):
            vec = vec.detach().clone()
            pool = self.vectors.setdefault(obj, [])
            if len(pool) == 0:
                out.append(vec)
                pool.append(vec)
            elif len(pool) < self.pool_size:
                rid = random.randint(0, len(pool) - 1)
                pool.append(vec)
                out.append(pool[rid])
            else:
                rid = random.randint(0, len(pool) - 1)
                out.append(pool[rid])
                pool[rid] = vec
        return torch.stack(out)


# --------------------------------------------------------------------------------------
# model.py / trainer.py
# --------------------------------------------------------------------------------------


def model_forward(sd, cfg, batch, noise, pool=None, update=False, test_mode=False, use_gt_box=False,
                  train=True, wrong_rep=None):
    """Model.forward (model.py:94-124) incl. scene_graph_to_vectors / create_components_vecs.
    `noise` is the (1, mask_noise_dim) draw of model.py:149 (injected so runs are comparable);
    `pool` a VectorPool (or `wrong_rep` given directly)."""
    imgs, objs, boxes, masks, triples, obj_to_img, _t2i, attributes = batch
    O = objs.numel()
    H, W = cfg['image_size']
    ac = cfg.get('align_corners', False)
    obj_vecs, _ = scene_graph_to_vectors(sd, objs, triples, attributes, cfg.get('gconv_num_layers', 5))
    mask_vecs = torch.cat([obj_vecs, noise.repeat(O, 1)], dim=1)                    # model.py:149-152
    crops = crop_bbox_batch(imgs, boxes, obj_to_img, 64, align_corners=ac)          # :156
    obj_repr = mlp2(appearance_encoder(sd, crops, update=update, train=train), sd, 'repr_net')   # :157
    oh = one_hot(objs, cfg['num_objs'])
    layout_vecs = torch.cat([oh, obj_repr], dim=1)                                  # :168
    if wrong_rep is None:
        wrong_rep = pool.query(objs, obj_repr) if pool is not None else obj_repr.detach()
    wrong_vecs = torch.cat([oh, wrong_rep], dim=1)                                  # :170-171
    boxes_pred = mlp2(obj_vecs, sd, 'box_net')                                      # :103
    scores = mask_net(sd, mask_vecs, mask_size=cfg.get('mask_size', 32), update=update, train=train)
    masks_pred = torch.sigmoid(scores.squeeze(1))                                   # :106-107
    if test_mode:
        bx = boxes if use_gt_box else boxes_pred
        mk = masks if masks is not None else masks_pred
        pred_layout = masks_to_layout(layout_vecs, bx, mk, obj_to_img, H, W, test_mode=True, align_corners=ac)
        imgs_pred = global_generator(sd, pred_layout, n_down=cfg.get('n_downsample_global', 4))
        return imgs_pred, boxes_pred, masks_pred, None, pred_layout, None
    gt_layout = masks_to_layout(layout_vecs, boxes, masks, obj_to_img, H, W, align_corners=ac)        # :119
    pred_layout = masks_to_layout(layout_vecs, boxes, masks_pred, obj_to_img, H, W, align_corners=ac)  # :120
    wrong_layout = masks_to_layout(wrong_vecs, boxes, masks, obj_to_img, H, W, align_corners=ac)       # :121
    imgs_pred = global_generator(sd, gt_layout, n_down=cfg.get('n_downsample_global', 4))             # :123
    return imgs_pred, boxes_pred, masks_pred, gt_layout, pred_layout, wrong_layout


DEFAULT_WEIGHTS = dict(bbox=10.0, ac=0.1, d_obj=0.1, d_mask=1.0, d_mask_feat=10.0, d_img=1.0, d_img_feat=10.0)


def generator_losses(sd_g, sd_obj, sd_mask, sd_img, cfg, batch, fwd, use_gt, update_obj=None, weights=None, vgg_sd=None):
    """Trainer.train_generator loss assembly (trainer.py:205-259), no L1; the VGG term (trainer.py:218-221) when
    vgg_sd (VGG19 feature weights, make_vgg_state_dict keys) is given, weight w['vgg'] (reference default 10)."""
    w = dict(DEFAULT_WEIGHTS, **(weights or {}))
    imgs, objs, boxes, masks, _tr, obj_to_img, _t2i, _attr = batch
    imgs_pred, boxes_pred, masks_pred, layout = fwd[0], fwd[1], fwd[2], fwd[3]
    ac_flag = cfg.get('align_corners', False)
    losses = {}
    if use_gt:
        losses['bbox_pred'] = ((boxes_pred - boxes) ** 2).mean() * w['bbox']                       # :215
    if vgg_sd is not None:
        losses['g_vgg'] = vgg_loss(vgg_sd, imgs_pred, imgs) * w.get('vgg', 10.0)                   # :218-221
    real, ac, _ = ac_crop_discriminator(sd_obj, imgs_pred, objs, boxes, obj_to_img,
                                        cfg.get('crop_size', 32), update=update_obj, align_corners=ac_flag)
    losses['ac_loss'] = ac * w['ac']                                                               # :224
    losses['g_gan_obj_loss'] = gan_g_loss(real) * w['d_obj']                                       # :226
    oh = one_hot(objs, cfg['num_objs'])
    sf = mask_discriminator(sd_mask, masks_pred.unsqueeze(1), oh)
    losses['g_gan_mask_obj_loss'] = lsgan_loss(sf, True) * w['d_mask']                             # :234-236
    sr = mask_discriminator(sd_mask, masks.float().unsqueeze(1), oh)
    losses['g_mask_features_loss'] = features_loss(sf, sr) * w['d_mask_feat']                      # :240-242
    pred_real = multiscale_discriminator(sd_img, torch.cat([layout, imgs], dim=1))                 # :246
    pred_fake = multiscale_discriminator(sd_img, torch.cat([layout.detach(), imgs_pred], dim=1))   # :250
    losses['g_gan_img_loss'] = lsgan_loss(pred_fake, True) * w['d_img']
    losses['g_gan_features_loss_img'] = features_loss(pred_fake, pred_real) * w['d_img_feat']     # :255
    order = (['bbox_pred'] if use_gt else []) + (['g_vgg'] if vgg_sd is not None else []) + ['ac_loss', 'g_gan_obj_loss', 'g_gan_mask_obj_loss',
                                                 'g_mask_features_loss', 'g_gan_img_loss',
                                                 'g_gan_features_loss_img']
    total = None
    for k in order:
        total = losses[k] if total is None else total + losses[k]
    losses['total_loss'] = total
    return losses


def obj_d_losses(sd_obj, cfg, imgs, imgs_pred, objs, boxes, obj_to_img, update=False):
    """Trainer.train_obj_discriminator (trainer.py:265-279); 'pred' boxes are the GT boxes (train.py:210)."""
    ac_flag = cfg.get('align_corners', False)
    sf, ac_f, _ = ac_crop_discriminator(sd_obj, imgs_pred, objs, boxes, obj_to_img, cfg.get('crop_size', 32),
                                        update=update, align_corners=ac_flag)
    sr, ac_r, _ = ac_crop_discriminator(sd_obj, imgs, objs, boxes, obj_to_img, cfg.get('crop_size', 32),
                                        update=update, align_corners=ac_flag)
    d = {'d_obj_gan_loss': gan_d_loss(sr, sf) * 0.5, 'd_ac_loss_real': ac_r, 'd_ac_loss_fake': ac_f}
    d['total_loss'] = d['d_obj_gan_loss'] + d['d_ac_loss_real'] + d['d_ac_loss_fake']
    return d


def mask_d_losses(sd_mask, cfg, masks, masks_pred, objs):
    """Trainer.train_mask_discriminator (trainer.py:281-300)."""
    oh = one_hot(objs, cfg['num_objs'])
    sf = mask_discriminator(sd_mask, masks_pred.unsqueeze(1), oh)
    sr = mask_discriminator(sd_mask, masks.float().unsqueeze(1), oh)
    d = {'fake_loss': lsgan_loss(sf, False) * 0.5, 'real_loss': lsgan_loss(sr, True) * 0.5}
    d['total_loss'] = d['fake_loss'] + d['real_loss']
    return d


def img_d_losses(sd_img, imgs, imgs_pred, layout, layout_wrong):
    """Trainer.train_image_discriminator (trainer.py:302-325)."""
    a = 0.25
    d = {
        'fake_image_loss': lsgan_loss(multiscale_discriminator(sd_img, torch.cat([layout, imgs_pred], 1)), False) * a,
        'wrong_texture_loss': lsgan_loss(multiscale_discriminator(sd_img, torch.cat([layout_wrong, imgs], 1)), False) * a,
        'd_img_gan_real_loss': lsgan_loss(multiscale_discriminator(sd_img, torch.cat([layout, imgs], 1)), True) * 0.5,
    }
    d['total_loss'] = d['fake_image_loss'] + d['wrong_texture_loss'] + d['d_img_gan_real_loss']
    return d


class OracleTrainer:
    """One full train iteration of train.py:198-215 on CPU: Model.forward, G step and the three
    D steps with Adam(lr, betas=(0.5, 0.999)) each (trainer.py:60,80,106,133).  State is four
    flat dicts of leaf tensors keyed like the reference state_dicts."""

    def __init__(self, sds, cfg, lr=1e-4, mask_lr=1e-5, beta1=0.5, pool_size=100, vgg_sd=None):
        self.cfg = cfg
        self.vgg_sd = vgg_sd          # frozen VGG19 feature weights: adds the g_vgg term (vgg_features_weight 10)
        self.sd = {k: {n: (t.clone().requires_grad_(t.is_floating_point())) for n, t in sd.items()}
                   for k, sd in sds.items()}
        self.opt = {}
        for k, sd in self.sd.items():
            params = [t for n, t in sd.items() if t.requires_grad and not n.endswith(('running_mean', 'running_var'))]
            for n, t in sd.items():
                if n.endswith(('running_mean', 'running_var')):
                    t.requires_grad_(False)
            self.opt[k] = torch.optim.Adam(params, lr=(mask_lr if k == 'mask' else lr), betas=(beta1, 0.999))
        self.pool = VectorPool(pool_size)
        self.losses = {}

    def step(self, batch, noise, use_gt=True):
        cfg = self.cfg
        imgs, objs, boxes, masks, triples, obj_to_img, t2i, attributes = batch
        if not use_gt:
            attributes = torch.zeros_like(attributes)                                # train.py:196-197
            batch = (imgs, objs, boxes, masks, triples, obj_to_img, t2i, attributes)
        fwd = model_forward(self.sd['g'], cfg, batch, noise, pool=self.pool, update=True)
        gl = generator_losses(self.sd['g'], self.sd['obj'], self.sd['mask'], self.sd['img'], cfg, batch, fwd,
                              use_gt, update_obj=True, vgg_sd=self.vgg_sd)
        for k in ('g', 'obj', 'mask', 'img'):
            self.opt[k].zero_grad()
        gl['total_loss'].backward()
        self.opt['g'].step()
        imgs_pred, masks_pred = fwd[0].detach(), fwd[2].detach()
        layout, layout_wrong = fwd[3].detach(), fwd[5].detach()
        ml = mask_d_losses(self.sd['mask'], cfg, masks, masks_pred, objs)
        self.opt['mask'].zero_grad()
        ml['total_loss'].backward()
        self.opt['mask'].step()
        ol = obj_d_losses(self.sd['obj'], cfg, imgs, imgs_pred, objs, boxes, obj_to_img, update=True)
        self.opt['obj'].zero_grad()
        ol['total_loss'].backward()
        self.opt['obj'].step()
        il = img_d_losses(self.sd['img'], imgs, imgs_pred, layout, layout_wrong)
        self.opt['img'].zero_grad()
        il['total_loss'].backward()
        self.opt['img'].step()
        self.losses = {'g': {k: float(v) for k, v in gl.items()}, 'mask': {k: float(v) for k, v in ml.items()},
                       'obj': {k: float(v) for k, v in ol.items()}, 'img': {k: float(v) for k, v in il.items()}}
        return fwd


# --------------------------------------------------------------------------------------
# losses.py:178-224  VGG19 feature-matching loss
# --------------------------------------------------------------------------------------
# torchvision vgg19().features[0:30]: (layer index of the conv, Cin, Cout); 'M' = MaxPool2d(2, 2).  The reference cuts
# it into five slices ending at relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 (losses.py:187-196).
VGG19_LAYERS = ((0, 3, 64), (2, 64, 64), 'M', (5, 64, 128), (7, 128, 128), 'M', (10, 128, 256), (12, 256, 256),
                (14, 256, 256), (16, 256, 256), 'M', (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512), 'M',
                (28, 512, 512))
VGG19_TAPS = (0, 5, 10, 19, 28)          # conv indices whose ReLU output is a feature map
VGG_LOSS_WEIGHTS = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)      # losses.py:216


def make_vgg_state_dict(seed=0):
    """Seeded random VGG19 feature weights under torchvision's keys ('features.<i>.weight/bias'), He-scaled so that
    activations keep their magnitude through 13 ReLU layers (the pretrained weights cannot be downloaded here; real
    ones load through the same keys)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for l in VGG19_LAYERS:
        if l == 'M':
            continue
        i, cin, cout = l
        sd['features.%d.weight' % i] = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        sd['features.%d.bias' % i] = (torch.rand(cout, generator=g) - 0.5) * 0.1
    return sd


def vgg19_features(sd, x):
    """Vgg19.forward (losses.py:202-209): [relu1_1, relu2_1, relu3_1, relu4_1, relu5_1] of x (N,3,H,W)."""
    feats = []
    h = x
    for l in VGG19_LAYERS:
        if l == 'M':
            h = F.max_pool2d(h, 2, 2)
            continue
        i = l[0]
        h = F.relu(F.conv2d(h, sd['features.%d.weight' % i], sd['features.%d.bias' % i], padding=1))
        if i in VGG19_TAPS:
            feats.append(h)
    return feats


def vgg_loss(sd, x, y):
    """VGGLoss.forward (losses.py:218-224): sum_i w_i * L1(vgg(x)_i, vgg(y)_i.detach())."""
    fx, fy = vgg19_features(sd, x), vgg19_features(sd, y)
    loss = 0
    for w, a, b in zip(VGG_LOSS_WEIGHTS, fx, fy):
        loss = loss + w * (a - b.detach()).abs().mean()
    return loss


# --------------------------------------------------------------------------------------
# deterministic weights (shared by golden generation, tests, smoke and bench)
# --------------------------------------------------------------------------------------


def make_state_dicts(cfg, seed=0):
    """Seeded random weights with the reference's names/shapes and init
    families: conv/convT N(0,0.02) for G / image D / mask D (generators.py:7-13,
    discriminators.py:57-63), kaiming-normal gconv Linears (graph.py:27-30), PyTorch-default
    style uniform elsewhere.  Values come from one torch.Generator so every machine gets the
    same tensors; the reference modules load them through load_state_dict."""
    g = torch.Generator().manual_seed(seed)
    num_objs, D = cfg['num_objs'], cfg['num_objs'] + cfg.get('rep_size', 32)
    A = cfg.get('num_attributes', 35)
    ngf, n_down, n_blocks = cfg.get('ngf', 64), cfg.get('n_downsample_global', 4), cfg.get('n_blocks', 9)
    H = 512
    randn = lambda *s: torch.randn(*s, generator=g)

    def lin(sd, name, dout, din, kaiming=False):
        if kaiming:
            sd[name + '.weight'] = randn(dout, din) * math.sqrt(2.0 / din)
        else:
            sd[name + '.weight'] = (torch.rand(dout, din, generator=g) * 2 - 1) / math.sqrt(din)
        sd[name + '.bias'] = (torch.rand(dout, generator=g) * 2 - 1) / math.sqrt(din)

    def conv(sd, name, cout, cin, k, std=0.02, transpose=False):
        shape = (cin, cout, k, k) if transpose else (cout, cin, k, k)
        if std is None:
            bound = 1.0 / math.sqrt(cin * k * k)
            sd[name + '.weight'] = (torch.rand(*shape, generator=g) * 2 - 1) * bound
        else:
            sd[name + '.weight'] = randn(*shape) * std
        sd[name + '.bias'] = (torch.rand(cout, generator=g) * 2 - 1) / math.sqrt(cin * k * k)

    def bn(sd, name, c, std=None):
        sd[name + '.weight'] = torch.ones(c) if std is None else 1.0 + randn(c) * std
        sd[name + '.bias'] = torch.zeros(c)
        sd[name + '.running_mean'] = torch.zeros(c)
        sd[name + '.running_var'] = torch.ones(c)
        sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    sg = {}
    sg['obj_embeddings.weight'] = randn(num_objs, 128)
    sg['pred_embeddings.weight'] = randn(7, 128)
    din1 = 3 * 128 + 2 * A
    for prefix, d_in in [('gconv', din1)] + [('gconv_net.gconvs.%d' % i, 384) for i in range(cfg.get('gconv_num_layers', 5) - 1)]:
        lin(sg, prefix + '.net1.0', H, d_in, kaiming=True)
        lin(sg, prefix + '.net1.2', 2 * H + 128, H, kaiming=True)
        lin(sg, prefix + '.net2.0', H, H, kaiming=True)
        lin(sg, prefix + '.net2.2', 128, H, kaiming=True)
    lin(sg, 'box_net.0', H, 128)
    lin(sg, 'box_net.2', 4, H)
    i, cur, gm = 0, 1, 192
    while cur < cfg.get('mask_size', 32):
        conv(sg, 'mask_net.%d' % (i + 1), gm, gm, 3, std=None)
        bn(sg, 'mask_net.%d' % (i + 2), gm)
        i += 4
        cur *= 2
    conv(sg, 'mask_net.%d' % i, 1, gm, 1, std=None)
    lin(sg, 'repr_net.0', 64, gm)
    lin(sg, 'repr_net.2', cfg.get('rep_size', 32), 64)
    for name, (co, ci) in zip(('0', '3', '6'), ((64, 3), (128, 64), (256, 128))):
        conv(sg, 'image_encoder.cnn.0.' + name, co, ci, 4, std=None)
    bn(sg, 'image_encoder.cnn.0.1', 64)
    bn(sg, 'image_encoder.cnn.0.4', 128)
    lin(sg, 'image_encoder.cnn.2', gm, 256)
    p = 'layout_to_image.model'
    conv(sg, p + '.1', ngf, D, 7)
    idx, c = 4, ngf
    for _ in range(n_down):
        conv(sg, '%s.%d' % (p, idx), c * 2, c, 3)
        c *= 2
        idx += 3
    for _ in range(n_blocks):
        conv(sg, '%s.%d.conv_block.1' % (p, idx), c, c, 3)
        conv(sg, '%s.%d.conv_block.5' % (p, idx), c, c, 3)
        idx += 1
    for _ in range(n_down):
        conv(sg, '%s.%d' % (p, idx), c // 2, c, 3, transpose=True)
        c //= 2
        idx += 3
    conv(sg, '%s.%d' % (p, idx + 1), 3, ngf, 7)

    so = {}
    for name, (co, ci) in zip(('0', '3', '6'), ((64, 3), (128, 64), (256, 128))):
        conv(so, 'discriminator.cnn.0.' + name, co, ci, 4, std=None)
    bn(so, 'discriminator.cnn.0.1', 64)
    bn(so, 'discriminator.cnn.0.4', 128)
    lin(so, 'discriminator.cnn.2', 1024, 256)
    lin(so, 'discriminator.real_classifier', 1, 1024)
    lin(so, 'discriminator.obj_classifier', num_objs, 1024)

    sm = {}
    conv(sm, 'scale0_layer0.0', 64, 1, 3)
    conv(sm, 'scale0_layer1.0', 128, 64, 3)
    conv(sm, 'scale0_layer2.0', 256, 128 + num_objs, 3)
    conv(sm, 'scale0_layer3.0', 1, 256, 3)

    si = {}
    ndf = cfg.get('ndf', 64)
    for s in range(2):
        chans = [(ndf, D + 3), (ndf * 2, ndf), (ndf * 4, ndf * 2), (ndf * 8, ndf * 4), (1, ndf * 8)]
        for j, (co, ci) in enumerate(chans):
            conv(si, 'scale%d_layer%d.0' % (s, j), co, ci, 4)
    return {'g': sg, 'obj': so, 'mask': sm, 'img': si}
