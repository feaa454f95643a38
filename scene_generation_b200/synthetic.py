"""Synthetic COCO-Stuff-shaped scene-graph batches.

Mirrors the output contract of the reference collate function
(/root/reference/scene_generation/data/coco.py:501-547): a batch is the 8-tuple
``(imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes)`` with the
objects of one image contiguous, the ``__image__`` object (class 0, box (0,0,1,1),
all-ones mask) last in every image (coco.py:312-317) and one ``__in_image__`` triple
(predicate 0) per real object (coco.py:358-413).  The generator itself is new code:
numpy ``RandomState`` so the same seed gives the same batch on every machine.
"""
import numpy as np
import torch

NUM_PREDS = 7          # __in_image__ + 6 spatial relations (coco.py:236-246)
NUM_ATTRIBUTES = 35    # 10 size bins + 25 location bins (coco.py:25-26,98)


def make_vocab(num_objs=172):
    """Minimal vocab dict with the keys Model / AcDiscriminator read
    (model.py:30-37, discriminators.py:24)."""
    return {
        'object_to_idx': {str(i): i for i in range(num_objs)},
        'pred_idx_to_name': ['__in_image__', 'left of', 'right of', 'above', 'below',
                             'inside', 'surrounding'],
        'num_attributes': NUM_ATTRIBUTES,
    }


def make_batch(n_imgs, image_size=(128, 128), num_objs=172, kmin=3, kmax=8, mask_size=32,
               seed=0, device='cpu', float_masks=False):
    """Build one synthetic batch (see module docstring).  Returns a tuple of tensors."""
    rng = np.random.RandomState(seed)
    H, W = image_size
    objs, boxes, masks, triples, obj_to_img, triple_to_img, attrs = [], [], [], [], [], [], []
    offset = 0
    for i in range(n_imgs):
        k = int(rng.randint(kmin, kmax + 1))
        cls = rng.randint(1, num_objs, size=k)
        x0 = rng.uniform(0.0, 0.6, size=k)
        y0 = rng.uniform(0.0, 0.6, size=k)
        ww = rng.uniform(0.15, 0.40, size=k)
        hh = rng.uniform(0.15, 0.40, size=k)
        x1 = np.minimum(x0 + ww, 1.0)
        y1 = np.minimum(y0 + hh, 1.0)
        for j in range(k):
            objs.append(int(cls[j]))
            boxes.append([x0[j], y0[j], x1[j], y1[j]])
            # filled ellipse with a little noise: closer to a real instance mask than iid bits
            yy, xx = np.mgrid[0:mask_size, 0:mask_size]
            cy, cx = (mask_size - 1) / 2.0, (mask_size - 1) / 2.0
            ry, rx = rng.uniform(0.3, 0.5) * mask_size, rng.uniform(0.3, 0.5) * mask_size
            m = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0)
            flip = rng.uniform(size=m.shape) < 0.02
            masks.append(np.logical_xor(m, flip).astype(np.int64))
            a = np.zeros(NUM_ATTRIBUTES, dtype=np.float32)
            a[rng.randint(0, 10)] = 1.0
            a[10 + rng.randint(0, 25)] = 1.0
            attrs.append(a)
            obj_to_img.append(i)
        # __image__ object last
        objs.append(0)
        boxes.append([0.0, 0.0, 1.0, 1.0])
        masks.append(np.ones((mask_size, mask_size), dtype=np.int64))
        a = np.zeros(NUM_ATTRIBUTES, dtype=np.float32)
        a[9] = 1.0
        a[10 + 12] = 1.0
        attrs.append(a)
        obj_to_img.append(i)
        img_idx = offset + k
        # one random relation per real object, then one __in_image__ per real object
        for c in range(k):
            others = [o for o in range(k) if o != c]
            if others:
                other = int(others[rng.randint(0, len(others))])
                p = int(rng.randint(1, NUM_PREDS))
                if rng.uniform() > 0.5:
                    s, o = c, other
                else:
                    s, o = other, c
                triples.append([offset + s, p, offset + o])
                triple_to_img.append(i)
        for c in range(k):
            triples.append([offset + c, 0, img_idx])
            triple_to_img.append(i)
        offset += k + 1
    imgs = rng.uniform(-1.0, 1.0, size=(n_imgs, 3, H, W)).astype(np.float32)
    t = lambda a, dt: torch.from_numpy(np.asarray(a, dtype=dt)).to(device)
    masks_t = t(np.stack(masks), np.int64)
    if float_masks:
        masks_t = masks_t.float()
    return (t(imgs, np.float32), t(objs, np.int64), t(boxes, np.float32), masks_t,
            t(triples, np.int64), t(obj_to_img, np.int64), t(triple_to_img, np.int64),
            t(np.stack(attrs), np.float32))


def image_ranges(obj_to_img, n_imgs=None):
    """Per-image contiguous object ranges [start, end) from obj_to_img, computed on the host
    once per batch (replaces the per-object ``.item()`` loop of layout.py:143-155)."""
    o2i = obj_to_img.detach().cpu().numpy() if torch.is_tensor(obj_to_img) else np.asarray(obj_to_img)
    if n_imgs is None:
        n_imgs = int(o2i.max()) + 1 if o2i.size else 0
    starts = np.searchsorted(o2i, np.arange(n_imgs), side='left')
    ends = np.searchsorted(o2i, np.arange(n_imgs), side='right')
    if o2i.size and np.any(np.diff(o2i) < 0):
        raise ValueError('objects of one image must be contiguous and images in order '
                         '(layout.py:152-155)')
    return np.stack([starts, ends], axis=1).astype(np.int32)


MAX_CLASS_SLOTS = 32


def class_slots(objs, obj_to_img, n_imgs):
    """Per-image class slots for channel-compacted layouts (csrc/compact.cu): the distinct classes of an image
    in order of first appearance.  Returns (obj_slot int64 (O,), slot_cls int32 (N, MAX_CLASS_SLOTS) padded
    with -1, the largest number of slots any image uses).  An image with more than MAX_CLASS_SLOTS classes
    reports its true count (the model then falls back to dense layouts) and its overflow objects get slot 0."""
    tables = [dict() for _ in range(n_imgs)]
    obj_slot = []
    for c, n in zip(objs, obj_to_img):
        t = tables[n]
        if c not in t:
            t[c] = len(t)
        obj_slot.append(t[c] if t[c] < MAX_CLASS_SLOTS else 0)
    slot_cls = np.full((n_imgs, MAX_CLASS_SLOTS), -1, dtype=np.int32)
    for n, t in enumerate(tables):
        for c, k in t.items():
            if k < MAX_CLASS_SLOTS:
                slot_cls[n, k] = c
    used = max((len(t) for t in tables), default=0)
    return torch.tensor(obj_slot, dtype=torch.int64), torch.from_numpy(slot_cls), used


class HostMeta:
    """Host-side index structures of one collated batch, computed from the CPU tensors the loader already
    holds (object ranges per image, CSR of triple incidences per object, the class list).  attach() tags the
    device copies of the batch with them so that Model.forward needs no device->host synchronisation
    (the reference syncs per object: layout.py:143-155, utils.py:71)."""

    def __init__(self, batch_cpu):
        from . import ops
        imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch_cpu
        self.n_imgs = imgs.shape[0]
        self.ranges = torch.from_numpy(image_ranges(obj_to_img, self.n_imgs))
        ptr, src = ops.build_incidence_csr(triples[:, [0, 2]].numpy(), objs.numel())
        self.seg_ptr, self.seg_src = torch.from_numpy(ptr), torch.from_numpy(src)
        self.objs = objs.tolist()
        self.obj_slot, self.slot_cls, self.slots_used = class_slots(self.objs, obj_to_img.tolist(), self.n_imgs)
        if torch.cuda.is_available():
            self.ranges, self.seg_ptr, self.seg_src, self.obj_slot, self.slot_cls = (
                t.pin_memory() for t in (self.ranges, self.seg_ptr, self.seg_src, self.obj_slot, self.slot_cls))

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.ranges, self.seg_ptr, self.seg_src, self.obj_slot, self.slot_cls))

    def attach(self, batch_dev):
        imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch_dev
        dev = objs.device
        obj_to_img._sg_ranges = self.ranges.to(dev, non_blocking=True)
        triples._sg_csr = (self.seg_ptr.to(dev, non_blocking=True), self.seg_src.to(dev, non_blocking=True))
        objs._sg_host = self.objs
        objs._sg_compact = (self.obj_slot.to(dev, non_blocking=True), self.slot_cls.to(dev, non_blocking=True), self.slots_used)
        return batch_dev
