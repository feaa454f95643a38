"""Imports the UNMODIFIED reference — from /root/reference in the build container, else from the byte-identical
staging copy oracle/_ref/ (oracle/build_ref.py) that travels to the GPU box.  TEST / BASELINE INFRASTRUCTURE ONLY.

Used by ``oracle/gen_golden.py`` to generate ``tests/golden``, by ``tests/test_oracle_golden.py`` and by bench.py's
reference arm.  On CPU the four stubs SURVEY.md §8c lists are installed: a dummy ``tensorboardX``, ``.cuda()`` /
``.to('cuda')`` as identity, ``torch.cuda.is_available`` patched only while modules are constructed,
``torch.cuda.FloatTensor`` -> ``torch.FloatTensor``.  ``install(device='cuda')`` (the informational reference-on-GPU
arm) only needs the ``tensorboardX`` stub.
"""
import contextlib
import os
import sys
import types

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REFERENCE_ROOT = os.environ.get('SG_REFERENCE_ROOT') or \
    ('/root/reference' if os.path.isdir('/root/reference/scene_generation') else _STAGED)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'scene_generation'))


def which():
    """'reference tree' or 'staged copy (oracle/_ref, sha256-verified)'"""
    if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED):
        from oracle import build_ref
        return 'staged copy oracle/_ref (sha256 %s)' % ('verified' if build_ref.verify() else 'MISMATCH')
    return 'reference tree %s' % REFERENCE_ROOT


_installed = False


def install(device='cpu'):
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if 'tensorboardX' not in sys.modules:
        tb = types.ModuleType('tensorboardX')

        class SummaryWriter:                     # noqa: D401 - stub
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None
        tb.SummaryWriter = SummaryWriter
        sys.modules['tensorboardX'] = tb
    if device != 'cpu':          # a real GPU: the reference's own .cuda() / 'cuda' strings just work
        _installed = True
        return
    torch.nn.Module.cuda = lambda self, device=None: self
    _orig_to = torch.nn.Module.to

    def _to(self, *args, **kwargs):
        args = tuple(a for a in args if not (isinstance(a, str) and a.startswith('cuda')))
        if kwargs.get('device', None) is not None and str(kwargs['device']).startswith('cuda'):
            kwargs.pop('device')
        if not args and not kwargs:
            return self
        return _orig_to(self, *args, **kwargs)
    torch.nn.Module.to = _to
    torch.cuda.FloatTensor = torch.FloatTensor
    _installed = True


@contextlib.contextmanager
def pretend_cuda():
    """define_G / define_D assert torch.cuda.is_available() (generators.py:54, discriminators.py:70,81)."""
    orig = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        yield
    finally:
        torch.cuda.is_available = orig


def modules():
    install()
    import scene_generation.bilinear as bilinear
    import scene_generation.discriminators as discriminators
    import scene_generation.generators as generators
    import scene_generation.graph as graph
    import scene_generation.layers as layers
    import scene_generation.layout as layout
    import scene_generation.losses as losses
    import scene_generation.model as model
    import scene_generation.utils as utils
    return types.SimpleNamespace(bilinear=bilinear, discriminators=discriminators, generators=generators,
                                 graph=graph, layers=layers, layout=layout, losses=losses, model=model,
                                 utils=utils)


def make_args(image_size=(64, 64), output_dir='/tmp/sg_ref_out', **over):
    install()
    from scene_generation.args import parser
    args = parser.parse_args(['--output_dir', output_dir])
    args.image_size = tuple(image_size)
    args.vgg_features_weight = 0.0      # pretrained VGG19 cannot be downloaded (no network)
    for k, v in over.items():
        setattr(args, k, v)
    return args


def make_trainer(vocab, image_size=(64, 64), device='cpu', **over):
    """Reference Trainer (trainer.py:15-134) on CPU (stubs) or on a real GPU (device='cuda')."""
    install(device)
    ctx = pretend_cuda() if device == 'cpu' else contextlib.nullcontext()
    with ctx:
        from scene_generation.trainer import Trainer
        args = make_args(image_size=image_size, **over)
        trainer = Trainer(args, vocab, {})
    return trainer, args


def train_iteration(trainer, batch, use_gt):
    """the body of the reference's training loop, train.py:193-215, statement for statement (the batch is already on
    the trainer's device; the use_gt coin of train.py:195 is the caller's)"""
    imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch
    if not use_gt:
        attributes = torch.zeros_like(attributes)
    model_out = trainer.model(imgs, objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, attributes=attributes)
    imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = model_out
    layout_one_hot = layout[:, :trainer.num_obj, :, :]                 # noqa: F841  (train.py:203-204, unused there too)
    layout_pred_one_hot = layout_pred[:, :trainer.num_obj, :, :]       # noqa: F841
    trainer.train_generator(imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, obj_to_img, use_gt)
    imgs_pred_detach = imgs_pred.detach()
    masks_pred_detach = masks_pred.detach()
    boxes_pred_detach = boxes.detach()
    layout_detach = layout.detach()
    layout_wrong_detach = layout_wrong.detach()
    trainer.train_mask_discriminator(masks, masks_pred_detach, objs)
    trainer.train_obj_discriminator(imgs, imgs_pred_detach, objs, boxes, boxes_pred_detach, obj_to_img)
    trainer.train_image_discriminator(imgs, imgs_pred_detach, layout_detach, layout_wrong_detach)
    return model_out


def load(module, sd):
    """load_state_dict with strict key checking — proves oracle.make_state_dicts names/shapes
    match the reference modules."""
    missing, unexpected = module.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    return module
