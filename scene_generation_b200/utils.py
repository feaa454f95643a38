"""LossManager / VectorPool / flag parsers (mirror of scene_generation/utils.py).

Host-side bookkeeping — listed "next" in SURVEY.md §8f.  LossManager keeps loss terms on the device
and only syncs when values are read (the reference calls .item() per term, utils.py:56)."""
import random

import torch


def int_tuple(s):
    return tuple(int(i) for i in s.split(','))


def float_tuple(s):
    return tuple(float(i) for i in s.split(','))


def str_tuple(s):
    return tuple(s.split(','))


def bool_flag(s):
    if s == '1':
        return True
    if s == '0':
        return False
    raise ValueError('Invalid value "%s" for bool flag (should be 0 or 1)' % s)


class LossManager(object):
    def __init__(self):
        self.total_loss = None
        self._terms = {}

    def add_loss(self, loss, name, weight=1.0, use_loss=True):
        cur = loss * weight
        if use_loss:
            self.total_loss = cur if self.total_loss is None else self.total_loss + cur
        self._terms[name] = cur.detach()

    @property
    def all_losses(self):
        return {k: float(v) for k, v in self._terms.items()}

    def items(self):
        return self.all_losses.items()


class VectorPool:
    """utils.py:62-90: per-class pool of appearance vectors with the reference's replacement policy and
    the same python `random` draws in the same order.  The policy only depends on the class ids and on
    how full each class pool is — never on the vector values — so the bookkeeping runs on the host from the
    (host-side) object list while the vectors themselves stay in one device tensor: no device->host copy,
    no pipeline drain (the reference moves every vector to the CPU and back, utils.py:71-89).

    The call is split in two so that the device half can sit inside a captured CUDA graph:
      plan(objs)          host: the policy -> one int64 index vector of FIXED length 3*O
                          [read rows (O) | written pool rows (O, padded) | batch rows written there (O, padded)]
      apply(idx, vectors) device: one gather from cat(store, vectors) and one index_copy_ into the store.
    Rows >= capacity address the batch vectors; the padding writes go to a scratch row nobody reads."""

    def __init__(self, pool_size, max_rows=None):
        self.pool_size = pool_size
        self.max_rows = max_rows     # upper bound of pool rows (classes x pool_size) when known: no regrowth
        self.slots = {}          # class id -> list of global row ids
        self.store = None        # (capacity + 1, R) device tensor; the last row is the scratch row
        self.used = 0

    @property
    def capacity(self):
        return 0 if self.store is None else self.store.shape[0] - 1

    def reserve(self, need, like):
        """make room for `need` more rows (never called inside a graph capture once max_rows is allocated)"""
        cap = self.capacity
        if self.store is not None and self.used + need <= cap:
            return
        new_cap = max(1024, 2 * cap, self.used + need, self.max_rows or 0)
        store = torch.zeros((new_cap + 1, like.shape[1]), dtype=like.dtype, device=like.device)
        if self.store is not None:
            if self.slots and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError('VectorPool would have to grow inside a CUDA graph capture; pass max_rows')
            store[:cap] = self.store[:cap]
            # row ids >= the old capacity meant "batch vector": none are kept between calls, nothing to remap
        self.store = store

    def plan(self, objs_l):
        """host half of utils.py:67-90 for the class list of one batch.  Needs reserve() first (the indices of
        the batch rows depend on the capacity)."""
        cap = self.capacity
        content = {}             # row id -> index of the batch vector written into it during this call
        src = []
        slots, pool_size, randint, used = self.slots, self.pool_size, random.randint, self.used
        for i, obj in enumerate(objs_l):
            ids = slots.get(obj)
            if ids is None:
                ids = slots[obj] = []
            n = len(ids)
            if n == 0:
                src.append(cap + 1 + i)
                ids.append(used)
                content[used] = i
                used += 1
            elif n < pool_size:
                g = ids[randint(0, n - 1)]
                ids.append(used)
                content[used] = i
                used += 1
                src.append(cap + 1 + content[g] if g in content else g)
            else:
                g = ids[randint(0, n - 1)]
                src.append(cap + 1 + content[g] if g in content else g)
                content[g] = i
        self.used = used
        no, pad = len(src), len(src) - len(content)
        return torch.tensor(src + list(content.keys()) + [cap] * pad + list(content.values()) + [0] * pad, dtype=torch.long)

    def apply(self, idx, vectors):
        """device half: idx is plan()'s vector on the device of `vectors`."""
        vecs = vectors.detach()
        no = vecs.shape[0]
        out = torch.cat([self.store, vecs], dim=0).index_select(0, idx[:no])
        self.store.index_copy_(0, idx[no:2 * no], vecs.index_select(0, idx[2 * no:]))
        return out

    def query(self, objs, vectors, planned=None):
        """planned: the index vector of plan() already on the device (graph replay: the host half ran outside)."""
        if self.pool_size == 0:
            return vectors
        if planned is None:
            objs_l = getattr(objs, '_sg_host', None)
            if objs_l is None:
                objs_l = objs.tolist()
            self.reserve(len(objs_l), vectors)
            planned = self.plan(objs_l).to(vectors.device, non_blocking=True)
        return self.apply(planned, vectors)
