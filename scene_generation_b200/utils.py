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
    no pipeline drain (the reference moves every vector to the CPU and back, utils.py:71-89)."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.slots = {}          # class id -> list of global row ids
        self.store = None        # (capacity, R) device tensor
        self.used = 0

    def _grow(self, need, like):
        cap = 0 if self.store is None else self.store.shape[0]
        if self.used + need <= cap:
            return
        new_cap = max(1024, 2 * cap, self.used + need)
        store = torch.zeros((new_cap, like.shape[1]), dtype=like.dtype, device=like.device)
        if self.store is not None:
            store[:cap] = self.store
        self.store = store

    def query(self, objs, vectors):
        if self.pool_size == 0:
            return vectors
        objs_l = getattr(objs, '_sg_host', None)
        if objs_l is None:
            objs_l = objs.tolist()
        vecs = vectors.detach()
        self._grow(len(objs_l), vecs)
        cap = self.store.shape[0]
        content = {}             # row id -> index of the batch vector written into it during this call
        src = []
        slots, pool_size, randint, used = self.slots, self.pool_size, random.randint, self.used
        for i, obj in enumerate(objs_l):
            ids = slots.get(obj)
            if ids is None:
                ids = slots[obj] = []
            n = len(ids)
            if n == 0:
                src.append(cap + i)
                ids.append(used)
                content[used] = i
                used += 1
            elif n < pool_size:
                g = ids[randint(0, n - 1)]
                ids.append(used)
                content[used] = i
                used += 1
                src.append(cap + content[g] if g in content else g)
            else:
                g = ids[randint(0, n - 1)]
                src.append(cap + content[g] if g in content else g)
                content[g] = i
        self.used = used
        no = len(src)
        allidx = torch.tensor(src + list(content.keys()) + list(content.values()), dtype=torch.long).to(vecs.device, non_blocking=True)
        nw = len(content)
        out = torch.cat([self.store, vecs], dim=0).index_select(0, allidx[:no])
        self.store.index_copy_(0, allidx[no:no + nw], vecs.index_select(0, allidx[no + nw:]))
        return out
