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
    train.py:195, where box_net gets no gradient).

    bucket_mb (opt-in, SG_DDP_BUCKET_MB; validated with gloo on CPU, not yet on NCCL / inside captured iterations):
    the buffer is cut into contiguous buckets of about that many megabytes along the parameter order; a
    post-accumulate-grad hook per parameter counts its bucket down and launches the bucket's all-reduce
    asynchronously as soon as the backward pass has produced all of its gradients, so the collective overlaps the
    rest of the backward pass (DDP-style).  allreduce() then launches what is left (buckets holding parameters
    without a gradient this step), waits for every handle and divides once."""

    def __init__(self, module, bucket_mb=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        al = lambda n: (n + 3) // 4 * 4          # 16-byte aligned slices: the Adam kernel takes its float4 path
        total = sum(al(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        offsets = []
        for p in self.params:
            n = p.numel()
            # same physical layout as the parameter (dense, possibly permuted)
            g = torch.as_strided(self.flat, p.shape, p.stride(), off)
            p.grad = g
            offsets.append(off)
            off += al(n)
        self.buckets = None
        if bucket_mb:
            cap = max(1, int(bucket_mb * (1 << 20) / 4))
            self.buckets = []                    # [start, end, parameter indices]
            start, members = 0, []
            for i, p in enumerate(self.params):
                members.append(i)
                end = offsets[i] + al(p.numel())
                if end - start >= cap or i == len(self.params) - 1:
                    self.buckets.append((start, end, members))
                    start, members = end, []
            self._bucket_of = {}
            for b, (_, _, members) in enumerate(self.buckets):
                for i in members:
                    self._bucket_of[i] = b
            self._pending = [0] * len(self.buckets)
            self._handles = [None] * len(self.buckets)
            self._armed = False
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(param):
            if not self._armed:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self._handles[b] is None and world_size() > 1:
            start, end, _ = self.buckets[b]
            self._handles[b] = dist.all_reduce(self.flat[start:end], op=dist.ReduceOp.SUM, async_op=True)

    def zero(self):
        self.flat.zero_()
        if self.buckets is not None:             # arm the hooks for the backward pass that follows
            self._pending = [len(m) for _, _, m in self.buckets]
            self._handles = [None] * len(self.buckets)
            self._armed = True

    def allreduce(self):
        ws = world_size()
        if self.buckets is not None and self._armed:
            self._armed = False
            if ws > 1:
                for b in range(len(self.buckets)):
                    self._launch(b)              # buckets with parameters that got no gradient this step
                for h in self._handles:
                    if h is not None:
                        h.wait()
                self.flat.div_(ws)
            return
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(ws)
