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
    # bucketed variant: all-reduces launched from post-accumulate-grad hooks while backward is still running; two
    # steps (the second with a different loss) must give exactly the gradients of the monolithic reducer
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
    net2.load_state_dict(net.state_dict())
    red2 = ddp.FlatGradReducer(net2, bucket_mb=40 / (1 << 20))      # ~10 floats per bucket -> several buckets
    assert red2.buckets is not None and len(red2.buckets) >= 3
    outs = []
    for step in range(2):
        for r_, n_ in ((red, net), (red2, net2)):
            r_.zero()
            h = n_[1](n_[0](x))
            loss = h.pow(2).mean() if step == 0 else (h.sum() + n_[2](h).pow(2).sum())
            loss.backward()
            r_.allreduce()
        outs.append(all(torch.equal(a.grad, b.grad) for a, b in zip(net.parameters(), net2.parameters())))
    ret['bucketed%d' % rank] = outs
    dist.destroy_process_group()


def test_flat_grad_allreduce_equals_big_batch_gradient():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g0, g1 = ret[0], ret[1]
    assert ret['bucketed0'] == [True, True] and ret['bucketed1'] == [True, True]
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
