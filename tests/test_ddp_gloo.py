"""CPU, world_size 2, gloo: the data-parallel plumbing (flat gradient buffer all-reduce, parameter
broadcast) gives every rank the mean gradient of the concatenated batch, also when a parameter got no
gradient on a rank (the use_gt coin flip, train.py:195)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scene_generation_b200 import ddp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)          # different initial weights per rank: broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
    conv = torch.nn.Conv2d(3, 5, 3)
    conv.weight = torch.nn.Parameter(conv.weight.data.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))   # channels-last master
    ddp.broadcast_parameters(conv)
    ret['conv%d' % rank] = conv.weight.detach().clone()
    ddp.broadcast_parameters(net)
    red = ddp.FlatGradReducer(net)
    full = torch.arange(32, dtype=torch.float32).view(8, 4) / 10
    x = full[rank * 4:(rank + 1) * 4]
    red.zero()
    h = net[1](net[0](x))
    loss = h.pow(2).mean()           # net[2] is unused -> its gradient slots stay zero
    loss.backward()
    red.allreduce()
    ret[rank] = [p.grad.clone() for p in net.parameters()] + [p.detach().clone() for p in net.parameters()]
    # ranks with DIFFERENT local batch geometries (rank 0: 3 rows, rank 1: 5 rows) and several networks reduced in the
    # trainer's fixed order (generator first, then the discriminators): every rank issues the same collectives in the
    # same order whatever its local shapes are, and gets the mean of the per-rank gradients
    nets = [torch.nn.Linear(4, 2) for _ in range(3)]
    for n_ in nets:
        ddp.broadcast_parameters(n_)
    reds = [ddp.FlatGradReducer(n_) for n_ in nets]
    xs = torch.arange(4 * (3 + 2 * rank), dtype=torch.float32).view(-1, 4) / 7 + rank
    local = []
    for r_, n_ in zip(reds, nets):
        r_.zero()
        n_(xs).pow(2).mean().backward()
        local.append([p.grad.clone() for p in n_.parameters()])
    works = [reds[0].allreduce(async_op=True)]          # gloo: finished on return (None), NCCL: a work handle
    for r_ in reds[1:]:
        r_.allreduce()
    for w in works:
        if w is not None:
            w.wait()
    ret['geo_local%d' % rank] = local
    ret['geo%d' % rank] = [[p.grad.clone() for p in n_.parameters()] for n_ in nets]
    dist.destroy_process_group()


def test_flat_grad_allreduce_equals_big_batch_gradient():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g0, g1 = ret[0], ret[1]
    for k in range(3):
        for j in range(2):
            mean = (ret['geo_local0'][k][j] + ret['geo_local1'][k][j]) / 2
            assert torch.allclose(ret['geo0'][k][j], mean, atol=1e-6) and torch.equal(ret['geo0'][k][j], ret['geo1'][k][j])
    assert torch.equal(ret['conv0'], ret['conv1']) and ret['conv0'].stride() == ret['conv1'].stride()
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                     # identical grads and params on both ranks
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
    full = torch.arange(32, dtype=torch.float32).view(8, 4) / 10
    net[1](net[0](full)).pow(2).mean().backward()
    n = len(list(net.parameters()))
    for p, g in zip(net.parameters(), g0[:n]):
        ref = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(g, ref, atol=1e-6)
