"""Differentiable box crops (mirror of scene_generation/bilinear.py; kernel: sg_crop_bbox_*)."""
from . import functional as Fn


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, backend='cudnn', align_corners=False, operand=False):
    """bilinear.py:26-130.  operand=True returns the bf16 NHWC (B,HH,WW,8) tensor the crop CNNs read;
    otherwise the reference's (B,C,HH,WW) f32 tensor."""
    WW = HH if WW is None else WW
    return Fn.CropFn.apply(feats, bbox, bbox_to_feats, HH, WW, align_corners, operand)


def crop_bbox(feats, bbox, HH, WW=None, backend='cudnn', align_corners=False):
    """bilinear.py:101-130: one box per feature map."""
    import torch
    idx = torch.arange(feats.size(0), device=feats.device)
    return crop_bbox_batch(feats, bbox, idx, HH, WW, align_corners=align_corners)
