"""Captured training iterations (Trainer.train_step with CUDA graphs) vs the same iterations launched eagerly.

Both arms start from the same weights, see the same batches, the same mask noise (torch.randn patched to a
constant) and the same python `random` stream (VectorPool policy).  The only differences left are the order of
fp32 atomics (IN/BN statistics, split-K weight gradients), i.e. run-to-run noise — which Adam's sign-like first
steps amplify along a trajectory.  So (1) ONE iteration from identical state must agree like two eager iterations
do, and (2) along 9 iterations the loss terms must stay within the eager run-to-run spread (+5 %) and no
parameter may be further apart than a few sign flips allow (update <= lr per element and step)."""
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
    """From the SAME weights / Adam state / pool contents: a graph replay and an eagerly launched iteration must
    agree like two eager iterations agree with each other (atomics-order noise only)."""
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
    report, bad = {}, []
    for phase in ('capture', 'replay'):          # the iteration that is captured (+ replayed once), then a pure replay
        snap = _snapshot(tr)
        lg, pg = _one_step(tr, batch, noise, graph=None)
        assert tr.use_graphs and any(isinstance(v, _StepGraph) for v in tr._graphs.values())
        _restore(tr, snap)
        le, pe = _one_step(tr, batch, noise, graph=False)
        _restore(tr, snap)
        le2, pe2 = _one_step(tr, batch, noise, graph=False)
        rep = report[phase] = {'losses': {}, 'params_max': {}}
        for k, ref in le.items():
            spread, d = abs(le2[k] - ref), abs(lg[k] - ref)
            rep['losses'][k] = (ref, d, spread)
            # one sample of the eager spread: keep a floor of 0.5 % (2 % for the chaotic image-discriminator terms, whose
            # eager-vs-eager spread was measured up to 0.45 % in a single iteration)
            floor = 2e-2 if ('img' in k or k.endswith('total_loss')) else 5e-3
            if d > 4 * spread + floor * abs(ref) + 1e-5:
                bad.append((phase, k, ref, lg[k], le2[k]))
        for k, ref in pe.items():
            if ref.dim() < 2:
                # biases in front of a norm layer have an exactly-zero true gradient: what is computed is the
                # cancellation residue of a long sum, which depends on the order of the fp32 atomics (deterministic
                # under eager launches, different — equally valid — when the kernels run back to back in a graph)
                continue
            spread, d = (pe2[k] - ref).abs().max().item(), (pg[k] - ref).abs().max().item()
            scale = ref.abs().max().item()
            rep['params_max'][k] = (scale, d, spread)
            if spread > 0.5 * scale:
                continue          # noise-dominated tensor (two eager iterations already disagree by half its scale)
            if d > 8 * spread + 5e-2 * scale + 1e-6:
                bad.append((phase, k, scale, d, spread))
        _one_step(tr, batch, noise, graph=None)  # move on by one (replayed) iteration
    try:
        import json, os
        os.makedirs('gpurun_out', exist_ok=True)
        json.dump(report, open('gpurun_out/graph_step_spread_%s.json' % cfg_name, 'w'), indent=0)
    except OSError:
        pass
    assert not bad, bad[:10]


@pytest.mark.parametrize('from_host', [False, True])
def test_captured_steps_match_eager_steps(from_host):
    cfg = cases.CFG1
    sds = R.make_state_dicts(cfg, seed=5)
    H = cfg['image_size'][0]
    # two batch geometries (different object counts), each seen with both values of the use_gt coin
    host_batches = [tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=1)),
                    tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 4, 4, seed=2)),
                    tuple(t.pin_memory() for t in synthetic.make_batch(2, (H, H), cfg['num_objs'], 3, 3, seed=3))]
    metas = [synthetic.HostMeta(hb) for hb in host_batches]
    noise = cases.noise_for(21).to(DEV)
    steps = 9
    eager = make_trainer(cfg, sds, graphs=False)
    le = run(eager, host_batches, metas, steps, noise)
    assert not eager._graphs
    eager2 = make_trainer(cfg, sds, graphs=False)
    le2 = run(eager2, host_batches, metas, steps, noise)      # run-to-run spread of the eager trajectory
    graphed = make_trainer(cfg, sds, graphs=True)
    _lib.reset_launch_count()
    lg = run(graphed, host_batches, metas, steps, noise, from_host=from_host)
    assert graphed.use_graphs, 'capture failed and the trainer fell back to eager launches'
    captured = [v for v in graphed._graphs.values() if isinstance(v, _StepGraph)]
    assert len(captured) >= 2 and all(c.launches > 100 for c in captured), [getattr(c, 'launches', c) for c in graphed._graphs.values()]
    assert _lib.launch_count() > steps * 100          # replayed launches are accounted
    for i, (a, b, a2) in enumerate(zip(le, lg, le2)):
        for net in a:
            for name, ref in a[net].items():
                spread = abs(a2[net][name] - ref)
                # the image-discriminator game is chaotic at batch 2 (two eager runs drift apart by 5-10 % within a few
                # iterations); every other term follows its eager trajectory closely
                chaotic = net == 'img' or 'img' in name or name == 'total_loss'
                tol = 0.30 if chaotic else 0.05
                assert abs(b[net][name] - ref) <= 4 * spread + tol * abs(ref) + 5e-3, (i, net, name, b[net][name], ref, a2[net][name])
    lr = 1e-4
    nets = lambda t: (t.model, t.netD, t.obj_discriminator, t.mask_discriminator)
    for ne, ng, n2 in zip(nets(eager), nets(graphed), nets(eager2)):
        for (name, pe), (_, pg), (_, p2) in zip(ne.state_dict().items(), ng.state_dict().items(), n2.state_dict().items()):
            if not pe.is_floating_point():
                assert torch.equal(pe, pg), name
                continue
            d, d2 = (pe.float() - pg.float()).abs(), (pe.float() - p2.float()).abs()
            if 'running' in name:
                assert d.max() <= 2e-2 * max(1.0, pe.abs().max().item()) + 3 * d2.max(), name
                continue
            # Adam moves an element by ~lr per step (a little more while the second-moment estimate lags a growing
            # gradient); the graphed trajectory may be as far from an eager one as a second eager trajectory is (x3),
            # and never much further than opposite steps every iteration
            assert d.max() <= 3 * steps * lr + 1e-6, (name, d.max().item())
            assert d.mean() <= 3 * d2.mean() + 0.1 * steps * lr, (name, d.mean().item(), d2.mean().item())


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
