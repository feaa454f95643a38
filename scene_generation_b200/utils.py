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
    """utils.py:62-90: per-class pool of appearance vectors (host side, python `random`)."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.vectors = {}

    def query(self, objs, vectors):
        if self.pool_size == 0:
            return vectors
        objs_l = objs.tolist()
        vecs = vectors.detach().float().cpu()
        out = []
        for obj, vec in zip(objs_l, vecs):
            vec = vec.clone()
            pool = self.vectors.setdefault(obj, [])
            if len(pool) == 0:
                out.append(vec)
                pool.append(vec)
            elif len(pool) < self.pool_size:
                rid = random.randint(0, len(pool) - 1)
                pool.append(vec)
                out.append(pool[rid])
            else:
                rid = random.randint(0, len(pool) - 1)
                out.append(pool[rid])
                pool[rid] = vec
        return torch.stack(out).to(vectors.device)
