"""masks/boxes -> layout (mirror of scene_generation/layout.py; kernel: sg_masks_to_layout_*)."""
import torch

from . import functional as Fn
from . import ops, synthetic

COMPACT_CC = 64     # channels of a channel-compacted layout (one 64-wide K block of the tensor-core convolutions)


def _ranges(obj_to_img, N=None):
    """Per-image object ranges, computed once on the host (replaces the per-object .item() loop of
    layout.py:143-155; the object->image map is host data produced by the collate function)."""
    r = getattr(obj_to_img, '_sg_ranges', None)
    if r is not None:
        return r
    return torch.from_numpy(synthetic.image_ranges(obj_to_img, N)).to(obj_to_img.device)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, pooling='sum', test_mode=False, align_corners=False,
                    nhwc_bf16=False, N=None, cmap=None, grad_channels=None):
    """layout.py:64-93.  Returns (N,D,H,W); with nhwc_bf16 the result is a bf16 view of a channels-last
    buffer (N,H,W,Cp) which is also attached as ``._sg_nhwc`` for the conv operands.

    cmap (int32 (N, COMPACT_CC), nhwc_bf16 only): ``vecs`` are channel-compacted layout vectors (class slots of
    the image instead of the vocabulary, see Model._compact_plan); the result then has D = vecs.shape[1]
    compact channels, carries the map as ``._sg_cmap`` and is expanded by expand_layout().
    grad_channels=(c0, c1): columns of ``vecs`` that carry a gradient (default: all)."""
    if pooling != 'sum':
        raise NotImplementedError('only pooling="sum" is used by the model')
    W = H if W is None else W
    ranges = _ranges(obj_to_img, N)
    D = vecs.shape[1]
    if test_mode:
        with torch.no_grad():
            fmt = ops.NHWC_BF16 if nhwc_bf16 else ops.NCHW_F32
            raw = ops.masks_to_layout_fwd(vecs, boxes, masks, ranges, H, W, align_corners, fmt, test_mode=True, raw=True)
    else:
        raw = Fn.LayoutFn.apply(vecs, boxes, masks, ranges, H, W, align_corners, nhwc_bf16,
                                COMPACT_CC if cmap is not None else None, grad_channels)
    if not nhwc_bf16:
        return raw
    out = raw.permute(0, 3, 1, 2)[:, :D]
    out._sg_nhwc = raw
    if cmap is not None:
        assert not test_mode and cmap.shape == (raw.shape[0], COMPACT_CC)
        out._sg_cmap = cmap
    return out


def expand_layout(layout, num_channels):
    """Dense (N, num_channels, H, W) f32 tensor of a channel-compacted layout (the tensor the reference's
    masks_to_layout returns); dense layouts pass through."""
    cmap = getattr(layout, '_sg_cmap', None)
    if cmap is None:
        return layout
    N, Dc, H, W = layout.shape
    idx = cmap[:, :Dc].long()
    ok = (idx >= 0) & (idx < num_channels)
    dense = torch.zeros((N, num_channels + 1, H, W), dtype=torch.float32, device=layout.device)
    tgt = torch.where(ok, idx, torch.full_like(idx, num_channels))          # unused slots land in a scratch channel
    dense.scatter_add_(1, tgt.view(N, Dc, 1, 1).expand(N, Dc, H, W), layout.float())
    return dense[:, :num_channels]


def boxes_to_layout(*args, **kwargs):
    raise NotImplementedError('boxes_to_layout is dead code in the reference (layout.py:59 raises TypeError)')
