"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo on
CPU for tests).  The path shards over the batch of scene graphs; the only exchange is
the gradient all-reduce per network per optimizer step, done on ONE flat f32 buffer whose slices are the
parameters' .grad tensors (no packing copies)."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def flat_view(t):
    """1-D view of the memory of a dense (possibly permuted, e.g. channels-last) tensor — collectives need
    contiguous tensors, the conv weights are stored [Cout][kh][kw][Cin] under their logical (Cout,Cin,kh,kw) shape."""
    if t.is_contiguous():
        return t.view(-1)
    return t.as_strided((t.numel(),), (1,), t.storage_offset())


def broadcast_parameters(module, src=0):
    """identical replicas at start (and BN buffers from rank 0, PyTorch-DDP style)."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            if t.numel():
                dist.broadcast(flat_view(t.data), src)


class FlatGradReducer:
    """Owns a flat f32 gradient buffer; every parameter's .grad is a (stride-preserving) view into it.
    Parameters that received no gradient in a step contribute zeros (covers the use_gt coin flip,
    train.py:195, where box_net gets no gradient).  allreduce() averages the buffer over the ranks with ONE
    collective (NCCL: ReduceOp.AVG, no separate division pass over the 762 MB generator buffer)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        al = lambda n: (n + 3) // 4 * 4          # 16-byte aligned slices: the Adam kernel takes its float4 path
        total = sum(al(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            # same physical layout as the parameter (dense, possibly permuted)
            p.grad = torch.as_strided(self.flat, p.shape, p.stride(), off)
            off += al(p.numel())

    def zero(self):
        self.flat.zero_()

    def allreduce(self, async_op=False):
        """average over the ranks.  async_op: returns the work handle (None on a single rank); the caller must wait()
        on it — a stream-side wait for NCCL — before reading the gradients."""
        ws = world_size()
        if ws <= 1:
            return None
        if dist.get_backend() == 'nccl':
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, async_op=async_op) if async_op else \
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        # gloo (CPU tests) has no AVG
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(ws)
        return None
