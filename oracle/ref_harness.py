"""Imports the UNMODIFIED reference from /root/reference on CPU.  TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/gen_golden.py`` to generate ``tests/golden`` and by ``tests/test_oracle_vs_reference.py``
(skipped when the reference tree is absent).  The four stubs are the ones SURVEY.md §8c lists:
a dummy ``tensorboardX``, ``.cuda()`` / ``.to('cuda')`` as identity, ``torch.cuda.is_available``
patched only while modules are constructed, ``torch.cuda.FloatTensor`` -> ``torch.FloatTensor``.
"""
import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get('SG_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'scene_generation'))


_installed = False


def install():
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


def make_trainer(vocab, image_size=(64, 64), **over):
    """Reference Trainer on CPU (trainer.py:15-134)."""
    install()
    with pretend_cuda():
        from scene_generation.trainer import Trainer
        args = make_args(image_size=image_size, **over)
        trainer = Trainer(args, vocab, {})
    return trainer, args


def load(module, sd):
    """load_state_dict with strict key checking — proves oracle.make_state_dicts names/shapes
    match the reference modules."""
    missing, unexpected = module.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    return module
