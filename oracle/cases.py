"""Deterministic inputs for the golden cases.  TEST INFRASTRUCTURE ONLY.

Shared by ``oracle/gen_golden.py`` (runs the reference on them, here in the build container)
and by the tests (replay the oracle / the CUDA path on the same inputs anywhere).
"""
import torch

from scene_generation_b200 import synthetic

CFG1 = dict(image_size=(64, 64), num_objs=10, rep_size=32, mask_size=32, n_downsample_global=4,
            gconv_num_layers=5, crop_size=32, ngf=64, n_blocks=9)          # BASELINE.json configs[0]
CFG_SMALLG = dict(CFG1, ngf=8, n_blocks=2)                                 # reduced generator for op goldens


def cfg1_batch(seed=1):
    """configs[0]: 4-object graphs (3 real + __image__), 64x64, batch 2."""
    return synthetic.make_batch(2, image_size=(64, 64), num_objs=10, kmin=3, kmax=3, seed=seed)


def ragged_batch(seed=2):
    """ragged graph sizes 1..6 real objects, 3 images."""
    return synthetic.make_batch(3, image_size=(64, 64), num_objs=10, kmin=1, kmax=6, seed=seed)


def layout_literals():
    """Input literals of the reference's layout.py __main__ demo (layout.py:188-254): 6 objects,
    3-d vecs, 5x5 masks, two images.  The demo gives no expected output."""
    vecs = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=torch.float32)
    boxes = torch.tensor([[0.25, 0.125, 0.5, 0.875], [0, 0, 1, 0.25], [0.6125, 0, 0.875, 1],
                          [0, 0.8, 1, 1.0], [0.25, 0.125, 0.5, 0.875], [0.6125, 0, 0.875, 1]])
    diamond = torch.tensor([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]],
                           dtype=torch.float32)
    ring = torch.tensor([[0, 0, 1, 0, 0], [0, 1, 0, 1, 0], [1, 0, 0, 0, 1], [0, 1, 0, 1, 0], [0, 0, 1, 0, 0]],
                        dtype=torch.float32)
    masks = torch.stack([diamond, ring, diamond, diamond, diamond, diamond])
    obj_to_img = torch.tensor([0, 0, 0, 1, 1, 1])
    return vecs, boxes, masks, obj_to_img


def crop_literals():
    """Boxes / mapping of the reference's bilinear.py __main__ demo (bilinear.py:289-295):
    box_to_feats = [1, 0, 1] exercises the non-identity permutation branch (bilinear.py:94-98)."""
    g = torch.Generator().manual_seed(7)
    feats = torch.rand(2, 3, 16, 20, generator=g)
    boxes = torch.tensor([[0, 0, 1, 1], [0.25, 0.25, 0.75, 0.75], [0, 0, 0.5, 0.5]], dtype=torch.float32)
    box_to_feats = torch.tensor([1, 0, 1])
    return feats, boxes, box_to_feats


def rand(shape, seed, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def noise_for(seed):
    """The (1,64) layout-noise draw of model.py:149 under torch.manual_seed(seed)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn((1, 64), generator=g)
