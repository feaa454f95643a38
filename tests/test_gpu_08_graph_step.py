"""Captured training iterations (Trainer.train_step with CUDA graphs) vs the same iterations launched eagerly.

Both arms start from the same weights, see the same batches, the same mask noise (torch.randn patched to a
constant) and the same python `random` stream (VectorPool policy).  Every reduction of the CUDA path has a fixed
order, so the two arms must agree BIT FOR BIT: one iteration from identical state, and whole trajectories."""
import random

import pytest
import torch

from oracle import cases, restate as R
from scene_generation_b200 import _lib, args as sgargs, synthetic
from scene_generation_b200.trainer import Trainer, _StepGraph

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def make_trainer(cfg, sds, graphs):
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'])
    a.cuda_graphs = graphs
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    tr.model.load_state_dict(sds['g'])
    tr.obj_discriminator.load_state_dict(sds['obj'])
    tr.mask_discriminator.load_state_dict(sds['mask'])
    tr.netD.load_state_dict(sds['img'])
    return tr


def run(tr, host_batches, metas, steps, noise, from_host=False):
    random.seed(77)
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    losses = []
    try:
        for i in range(steps):
            hb, m = host_batches[i % len(host_batches)], metas[i % len(host_batches)]
            batch = m.attach(hb) if from_host else m.attach(tuple(t.to(DEV) for t in hb))
            tr.train_step(batch, use_gt=(i % 2 == 0))
            losses.append({k: dict(lm.all_losses) for k, lm in (('g', tr.generator_losses), ('mask', tr.d_mask_losses),
                                                                  ('obj', tr.d_obj_losses), ('img', tr.d_img_losses))})
    finally:
        torch.randn = orig
    return losses


def _state_tensors(tr):
    """every tensor a training iteration updates: parameters, buffers, Adam moments / step counters, the VectorPool"""
    ts = []
    for net in (tr.model, tr.netD, tr.obj_discriminator, tr.mask_discriminator):
        ts += [('%s.%s' % (type(net).__name__, k), v) for k, v in net.state_dict().items()]
    for oname, opt in (('g', tr.optimizer), ('img', tr.optimizer_d_img), ('obj', tr.optimizer_d_obj), ('mask', tr.optimizer_d_mask)):
        for gi, group in enumerate(opt.param_groups):
            for pi, p in enumerate(group['params']):
                for k, v in opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        ts.append(('opt_%s.%d.%d.%s' % (oname, gi, pi, k), v))
    ts.append(('pool.store', tr.model.fake_pool.store))
    return ts


def _snapshot(tr):
    import copy
    pool = tr.model.fake_pool
    return [(n, t, t.clone()) for n, t in _state_tensors(tr)], copy.deepcopy(pool.slots), pool.used, random.getstate()


def _restore(tr, snap):
    import copy
    tensors, slots, used, rstate = snap
    with torch.no_grad():
        for _, t, saved in tensors:
            t.copy_(saved)                       # in place: the captured graphs keep pointing at the same memory
    tr.model.fake_pool.slots, tr.model.fake_pool.used = copy.deepcopy(slots), used
    random.setstate(rstate)


def _one_step(tr, batch, noise, graph):
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    try:
        tr.train_step(batch, use_gt=True, graph=graph)
    finally:
        torch.randn = orig
    losses = {'%s.%s' % (k, n): v for k, lm in (('g', tr.generator_losses), ('mask', tr.d_mask_losses), ('obj', tr.d_obj_losses),
                                               ('img', tr.d_img_losses)) for n, v in lm.all_losses.items()}
    params = {n: t.detach().float().clone() for n, t in _state_tensors(tr) if t.is_floating_point()}
    return losses, params


@pytest.mark.parametrize('cfg_name', ['CFG1', 'mid'])
def test_one_replayed_iteration_equals_one_eager_iteration(cfg_name):
    """From the SAME weights / Adam state / pool contents a graph replay and an eagerly launched iteration run the same
    kernels on the same inputs; no reduction on the path depends on scheduling (no floating-point atomics), although
    the captured iteration runs independent sub-steps as parallel graph branches.  So every loss term, every
    parameter, every Adam moment and the VectorPool must come out BIT-IDENTICAL — for the iteration that is captured
    and for later replays."""
    if cfg_name == 'CFG1':
        cfg, n_img, kmin, kmax = cases.CFG1, 2, 3, 3
    else:                # 128x128, full vocabulary, ragged object counts
        cfg, n_img, kmin, kmax = dict(cases.CFG1, image_size=(128, 128), num_objs=172), 4, 3, 8
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    hb = tuple(t.pin_memory() for t in synthetic.make_batch(n_img, (H, H), cfg['num_objs'], kmin, kmax, seed=1))
    meta = synthetic.HostMeta(hb)
    batch = meta.attach(tuple(t.to(DEV) for t in hb))
    noise = cases.noise_for(21).to(DEV)
    tr = make_trainer(cfg, sds, graphs=True)
    random.seed(5)
    _one_step(tr, batch, noise, graph=None)      # first sight of the geometry: eager
    for phase in ('capture', 'replay'):          # the iteration that is captured (+ replayed once), then a pure replay
        snap = _snapshot(tr)
        lg, pg = _one_step(tr, batch, noise, graph=None)
        assert tr.use_graphs and any(isinstance(v, _StepGraph) for v in tr._graphs.values())
        _restore(tr, snap)
        le, pe = _one_step(tr, batch, noise, graph=False)
        assert le == lg, (phase, {k: (le[k], lg[k]) for k in le if le[k] != lg[k]})
        bad = [k for k in pe if not torch.equal(pe[k], pg[k])]
        assert not bad, (phase, bad[:10])
        _one_step(tr, batch, noise, graph=None)  # move on by one (replayed) iteration


@pytest.mark.parametrize('from_host', [False, True])
def test_captured_steps_match_eager_steps(from_host):
    """9 iterations over three batches (two geometries, both values of the use_gt coin): the trajectory with captured /
    replayed iterations equals the eagerly launched one bit for bit (losses of every step, final state)."""
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    host_batches = [tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=1)),
                    tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 4, 4, seed=2)),
                    tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=3))]
    metas = [synthetic.HostMeta(hb) for hb in host_batches]
    noise = cases.noise_for(21).to(DEV)
    steps = 9
    eager = make_trainer(cfg, sds, graphs=False)
    le = run(eager, host_batches, metas, steps, noise)
    assert not eager._graphs
    graphed = make_trainer(cfg, sds, graphs=True)
    _lib.reset_launch_count()
    lg = run(graphed, host_batches, metas, steps, noise, from_host=from_host)
    assert graphed.use_graphs, 'capture failed and the trainer fell back to eager launches'
    captured = [v for v in graphed._graphs.values() if isinstance(v, _StepGraph)]
    assert len(captured) >= 2 and all(c.launches > 100 for c in captured), [getattr(c, 'launches', c) for c in graphed._graphs.values()]
    assert _lib.launch_count() > steps * 100          # replayed launches are accounted
    for i, (a, b) in enumerate(zip(le, lg)):
        assert a == b, (i, {n: {k: (a[n][k], b[n][k]) for k in a[n] if a[n][k] != b[n][k]} for n in a})
    for (name, te), (_, tg) in zip(_state_tensors(eager), _state_tensors(graphed)):
        assert torch.equal(te, tg), name


def test_graph_cache_is_bounded_and_restore_checkpoint_drops_graphs():
    """at most args.graph_cache captured geometries stay alive (least recently used first); restoring a checkpoint
    forgets the captured iterations (their Adam launches point at the replaced optimizer state)"""
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'], graph_cache=2)
    tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
    noise = cases.noise_for(3).to(DEV)
    hbs = [tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], k, k, seed=k)) for k in (2, 3, 4)]
    metas = [synthetic.HostMeta(hb) for hb in hbs]
    random.seed(1)
    for rep in range(2):                         # every geometry twice: eager sighting, then capture
        for hb, m in zip(hbs, metas):
            _one_step(tr, m.attach(tuple(t.to(DEV) for t in hb)), noise, graph=None)
    assert tr.use_graphs
    assert sum(isinstance(v, _StepGraph) for v in tr._graphs.values()) == 2
    ck = {}
    tr.save_checkpoint(ck, 0, argparse_ns(a), 0)
    tr.restore_checkpoint(ck)
    assert not tr._graphs
    _one_step(tr, metas[0].attach(tuple(t.to(DEV) for t in hbs[0])), noise, graph=None)


def test_flat_gradient_buffers_give_the_plain_gradients():
    """data-parallel plumbing on one GPU (args.flat_grads): .grad tensors are views of one flat buffer per network and
    the weight / bias gradient kernels write into them directly — the iteration must end in the same bits as with
    ordinary gradients, eagerly launched and replayed from the captured graphs"""
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    hb = tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=1))
    meta = synthetic.HostMeta(hb)
    noise = cases.noise_for(21).to(DEV)
    finals = []
    for flat in (False, True, True):
        a = sgargs.default_args(image_size=cfg['image_size'], num_objs=cfg['num_objs'], flat_grads=flat)
        a.cuda_graphs = len(finals) == 2
        tr = Trainer(a, synthetic.make_vocab(cfg['num_objs']), {})
        for net, k in ((tr.model, 'g'), (tr.obj_discriminator, 'obj'), (tr.mask_discriminator, 'mask'), (tr.netD, 'img')):
            net.load_state_dict(sds[k])
        assert bool(tr.reducers) == flat
        random.seed(5)
        for i in range(3):               # use_gt every time: every parameter has a gradient (Adam skips None gradients)
            _one_step(tr, meta.attach(tuple(t.to(DEV) for t in hb)), noise, graph=None)
        finals.append({n: t.detach().clone() for n, t in _state_tensors(tr)})
    for other in finals[1:]:
        bad = [n for n in finals[0] if not torch.equal(finals[0][n], other[n])]
        assert not bad, bad[:10]


def argparse_ns(a):
    import copy, tempfile
    b = copy.copy(a)
    b.output_dir = tempfile.mkdtemp()
    return b


def test_vector_pool_plan_is_consumed_once_per_step():
    """graph replay runs the VectorPool policy on the host exactly once per iteration (same python-random stream
    as eager launches): after the same steps both pools hold the same bookkeeping."""
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    hb = tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=1))
    meta = synthetic.HostMeta(hb)
    noise = cases.noise_for(3).to(DEV)
    a, b = make_trainer(cfg, sds, graphs=False), make_trainer(cfg, sds, graphs=True)
    run(a, [hb], [meta], 5, noise)
    run(b, [hb], [meta], 5, noise)
    assert a.model.fake_pool.slots == b.model.fake_pool.slots and a.model.fake_pool.used == b.model.fake_pool.used
    assert random.random() is not None
