"""CPU: the C-ABI library loads and exports every symbol of include/sg_b200.h; host-side index logic
(tap / phase tables, CSR, image ranges, synthetic batches) against plain PyTorch."""
import ctypes
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from scene_generation_b200 import _lib, convspec, ops, synthetic


def test_library_builds_and_exports_declared_symbols():
    path = _lib.build()
    h = ctypes.CDLL(path)
    decl = _lib.declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if not hasattr(h, s)]
    assert not missing, missing
    bound = set(_lib._SIGS) | set(_lib._RESTYPES)
    assert set(decl) == bound, (set(decl) ^ bound)
    assert _lib.lib().sg_arch() == 100
    assert b'sm_100a' in _lib.lib().sg_version()


def test_struct_sizes_match_header_layout():
    # sg_tap_t 8 B, sg_phase_t 16 B, sg_wtap_t 16 B: the descriptors are passed by pointer across the ABI
    assert ctypes.sizeof(_lib.Tap) == 8 and ctypes.sizeof(_lib.Phase) == 16 and ctypes.sizeof(_lib.WTap) == 16
    assert ctypes.sizeof(_lib.ConvDesc) % 8 == 0 and ctypes.sizeof(_lib.WgradDesc) % 8 == 0


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.crop_bbox_fwd(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4), torch.zeros(1, dtype=torch.long), 2, 2)


# ---- emulate the documented semantics of sg_conv_tc / sg_wgrad_tc to validate the tap tables -----------
def emulate_conv(x5, w3, Hout, Wout, taps, phases=None, oh_mul=1, ow_mul=1, in_h0=0, in_w0=0):
    N, P, H, W, C = x5.shape
    Cout = w3.shape[0]
    phases = phases or [(0, len(taps), 0, 0)]
    y = torch.zeros(N, Hout * oh_mul, Wout * ow_mul, Cout)
    for (tb, nt, a, b) in phases:
        for (dh, dw, pl, wt) in taps[tb:tb + nt]:
            for h in range(Hout):
                for w in range(Wout):
                    hh, ww = h + dh + in_h0, w + dw + in_w0
                    if 0 <= hh < H and 0 <= ww < W:
                        y[:, h * oh_mul + a, w * ow_mul + b] += x5[:, pl, hh, ww] @ w3[:, wt].t()
    return y


def planes(x):     # (N,C,H,W) -> (N,4,ceil,ceil,C)
    N, C, H, W = x.shape
    out = torch.zeros(N, 4, (H + 1) // 2, (W + 1) // 2, C)
    for ph, pw in itertools.product(range(2), range(2)):
        s = x[:, :, ph::2, pw::2]
        out[:, ph * 2 + pw, :s.shape[2], :s.shape[3]] = s.permute(0, 2, 3, 1)
    return out


def w3_of(w):      # (Cout,Cin,k,k) -> (Cout,k*k,Cin)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1, w.shape[1])


@pytest.mark.parametrize('H,k,p', [(7, 4, 2), (8, 3, 1), (6, 4, 0)])
def test_tap_tables_stride2_and_adjoint(H, k, p):
    torch.manual_seed(0)
    x, w = torch.randn(2, 3, H, H), torch.randn(5, 3, k, k)
    ref = F.conv2d(x, w, stride=2, padding=p)
    Ho = ref.shape[2]
    y = emulate_conv(planes(x), w3_of(w), Ho, Ho, convspec.conv_s2(k, p))
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-4)
    # adjoint: phases write the parity planes of dx
    dy = torch.randn_like(ref)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, stride=2, padding=p).backward(dy)
    taps, phases = convspec.dgrad_s2(k, p)
    wT = w3_of(w.permute(1, 0, 2, 3))
    Hh = (H + 1) // 2
    dxp = torch.zeros(2, 4, Hh, Hh, 3)
    for (tb, nt, a, b) in phases:
        dxp[:, a * 2 + b] = emulate_conv(dy.permute(0, 2, 3, 1).unsqueeze(1), wT, Hh, Hh, taps[tb:tb + nt])
    assert torch.allclose(dxp, planes(xr.grad), atol=1e-4) or torch.allclose(dxp[planes(torch.ones_like(x)) > 0], planes(xr.grad)[planes(torch.ones_like(x)) > 0], atol=1e-4)


def test_tap_tables_transposed_conv_and_adjoint():
    torch.manual_seed(1)
    x, w = torch.randn(2, 4, 5, 5), torch.randn(4, 3, 3, 3)       # ConvT weight (Cin_t, Cout_t, k, k)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    taps, phases = convspec.convT_s2(3, 1)
    y = emulate_conv(x.permute(0, 2, 3, 1).unsqueeze(1), w3_of(w.permute(1, 0, 2, 3)), 5, 5, taps, phases, 2, 2)
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-4)
    dy = torch.randn_like(ref)
    xr = x.clone().requires_grad_(True)
    F.conv_transpose2d(xr, w, stride=2, padding=1, output_padding=1).backward(dy)
    dx = emulate_conv(planes(dy), w3_of(w), 5, 5, convspec.dgrad_convT(3, 1))   # B = [Cin_t][taps][Cout_t]
    assert torch.allclose(dx.permute(0, 3, 1, 2), xr.grad, atol=1e-4)


def test_tap_tables_stride1_and_dgrad():
    torch.manual_seed(2)
    x, w = torch.randn(1, 3, 6, 6), torch.randn(4, 3, 3, 3)
    taps, off = convspec.conv_s1(3, 1)
    y = emulate_conv(x.permute(0, 2, 3, 1).unsqueeze(1), w3_of(w), 6, 6, taps, in_h0=off, in_w0=off)
    assert torch.allclose(y.permute(0, 3, 1, 2), F.conv2d(x, w, padding=1), atol=1e-4)
    dy = torch.randn(1, 4, 6, 6)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, padding=1).backward(dy)
    dx = emulate_conv(dy.permute(0, 2, 3, 1).unsqueeze(1), w3_of(w.permute(1, 0, 2, 3)), 6, 6, convspec.dgrad_s1(3, 1))
    assert torch.allclose(dx.permute(0, 3, 1, 2), xr.grad, atol=1e-4)


def test_incidence_csr_order_matches_reference_scatter_order():
    edges = np.array([[0, 2], [1, 2], [2, 0], [0, 1], [3, 3]])
    ptr, src = ops.build_incidence_csr(edges, 5)
    assert ptr.tolist() == [0, 3, 5, 8, 10, 10]
    # object 0: subject of t0, t3 then object of t2
    assert src[ptr[0]:ptr[1]].tolist() == [0, 6, 5]
    assert src[ptr[2]:ptr[3]].tolist() == [4, 1, 3]
    assert src[ptr[3]:ptr[4]].tolist() == [8, 9]
    with pytest.raises(IndexError):
        ops.build_incidence_csr(np.array([[0, 7]]), 3)


def test_synthetic_batch_contract_and_ranges():
    b = synthetic.make_batch(5, (32, 32), 20, 1, 6, seed=3)
    imgs, objs, boxes, masks, triples, o2i, t2i, attrs = b
    assert imgs.shape == (5, 3, 32, 32) and objs.dtype == torch.int64 and masks.dtype == torch.int64
    r = synthetic.image_ranges(o2i)
    assert r[0, 0] == 0 and r[-1, 1] == objs.numel() and (r[1:, 0] == r[:-1, 1]).all()
    for i, (s, e) in enumerate(r):
        assert objs[e - 1] == 0 and (o2i[s:e] == i).all()                    # __image__ last, contiguous
        assert boxes[e - 1].tolist() == [0, 0, 1, 1] and masks[e - 1].min() == 1
    assert ((triples[:, 0] >= 0) & (triples[:, 2] < objs.numel())).all()
    same = synthetic.make_batch(5, (32, 32), 20, 1, 6, seed=3)
    assert all(torch.equal(x, y) for x, y in zip(b, same))
    with pytest.raises(ValueError):
        synthetic.image_ranges(torch.tensor([0, 1, 0]))


def test_vector_pool_matches_reference_policy_and_rng_order():
    """device-resident VectorPool (host index logic) == the reference's CPU pool (utils.py:62-90, restated in the
    oracle) for the same python `random` stream, incl. reads of vectors written earlier in the same batch."""
    import random
    from oracle.restate import VectorPool as RefPool
    from scene_generation_b200.utils import VectorPool
    a, b = VectorPool(3), RefPool(3)
    for step in range(50):
        n = random.Random(step).randint(1, 12)
        objs = torch.tensor([random.Random(step * 31 + i).randint(0, 4) for i in range(n)])
        vecs = torch.randn(n, 5)
        random.seed(step)
        ra = a.query(objs, vecs)
        random.seed(step)
        rb = b.query(objs, vecs)
        assert torch.equal(ra, rb)
    assert VectorPool(0).query(objs, vecs) is vecs


def emulate_wgrad(dy5, x5, Hred, Wred, taps, Cout, Cin, w_taps):
    """documented semantics of sg_wgrad_tc: dw[co, wtap, ci] = sum dy[img, pa, h+dha, w+dwa, co] * x[img, pb, h+dhb, w+dwb, ci]"""
    dw = torch.zeros(Cout, w_taps, Cin)
    for (dha, dwa, pa, dhb, dwb, pb, wt) in taps:
        for h in range(Hred):
            for w in range(Wred):
                ha, wa, hb, wb = h + dha, w + dwa, h + dhb, w + dwb
                if 0 <= ha < dy5.shape[2] and 0 <= wa < dy5.shape[3] and 0 <= hb < x5.shape[2] and 0 <= wb < x5.shape[3]:
                    dw[:, wt] += torch.einsum('nc,nd->cd', dy5[:, pa, ha, wa, :Cout], x5[:, pb, hb, wb, :Cin])
    return dw


def test_wgrad_tap_tables():
    torch.manual_seed(3)
    # stride-1 (pad 1), stride-2 (k4 p2, odd size) and transposed conv
    x, w = torch.randn(2, 3, 6, 6), torch.randn(4, 3, 3, 3, requires_grad=True)
    out = F.conv2d(x, w, padding=1)
    dy = torch.randn_like(out)
    out.backward(dy)
    got = emulate_wgrad(dy.permute(0, 2, 3, 1).unsqueeze(1), x.permute(0, 2, 3, 1).unsqueeze(1), 6, 6, convspec.wgrad_s1(3, 1), 4, 3, 9)
    assert torch.allclose(got, w.grad.permute(0, 2, 3, 1).reshape(4, 9, 3), atol=1e-4)
    x, w = torch.randn(2, 3, 7, 7), torch.randn(4, 3, 4, 4, requires_grad=True)
    out = F.conv2d(x, w, stride=2, padding=2)
    dy = torch.randn_like(out)
    out.backward(dy)
    got = emulate_wgrad(dy.permute(0, 2, 3, 1).unsqueeze(1), planes(x), out.shape[2], out.shape[3], convspec.wgrad_s2(4, 2), 4, 3, 16)
    assert torch.allclose(got, w.grad.permute(0, 2, 3, 1).reshape(4, 16, 3), atol=1e-4)
    x, w = torch.randn(2, 4, 5, 5), torch.randn(4, 3, 3, 3, requires_grad=True)      # ConvT weight (Cin_t, Cout_t, k, k)
    out = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    dy = torch.randn_like(out)
    out.backward(dy)
    got = emulate_wgrad(planes(dy), x.permute(0, 2, 3, 1).unsqueeze(1), 5, 5, convspec.wgrad_convT(3, 1), 3, 4, 9)
    assert torch.allclose(got, w.grad.permute(1, 2, 3, 0).reshape(3, 9, 4), atol=1e-4)


def test_host_meta_matches_device_free_derivation():
    b = synthetic.make_batch(4, (32, 32), 20, 1, 5, seed=8)
    m = synthetic.HostMeta(b)
    assert torch.equal(m.ranges, torch.from_numpy(synthetic.image_ranges(b[5])))
    ptr, src = ops.build_incidence_csr(b[4][:, [0, 2]].numpy(), b[1].numel())
    assert m.seg_ptr.tolist() == ptr.tolist() and m.seg_src.tolist() == src.tolist() and m.objs == b[1].tolist()
    assert m.nbytes() == 4 * (m.ranges.numel() + m.seg_ptr.numel() + m.seg_src.numel() + m.slot_cls.numel()) + 8 * m.obj_slot.numel()
    # class slots of the channel-compacted layouts: slot_cls[img, obj_slot[o]] is the object's class
    o2i, objs = b[5].tolist(), b[1].tolist()
    for o, (n, c) in enumerate(zip(o2i, objs)):
        assert int(m.slot_cls[n, int(m.obj_slot[o])]) == c
    for n in range(4):
        row = [c for c in m.slot_cls[n].tolist() if c >= 0]
        assert sorted(row) == sorted(set(objs[o] for o in range(len(objs)) if o2i[o] == n))
    assert m.slots_used == max(len(set(objs[o] for o in range(len(objs)) if o2i[o] == n)) for n in range(4))


def test_class_slots_overflow_and_expand_layout_cpu():
    """More classes than slots -> slots_used reports it (the model then keeps dense layouts); expand_layout
    scatters compact channels back to the vocabulary."""
    from scene_generation_b200 import layout as L
    objs = list(range(40)) + [3, 3]
    o2i = [0] * 40 + [1, 1]
    slot, table, used = synthetic.class_slots(objs, o2i, 2)
    assert used == 40 and table.shape == (2, synthetic.MAX_CLASS_SLOTS) and int(slot.max()) < synthetic.MAX_CLASS_SLOTS
    assert table[1].tolist()[:2] == [3, -1] and slot[-2:].tolist() == [0, 0]
    cmap = torch.full((1, 64), -1, dtype=torch.int32)
    cmap[0, :2] = torch.tensor([4, 1], dtype=torch.int32)
    cmap[0, 24:27] = torch.tensor([6, 7, 8], dtype=torch.int32)
    lay = torch.zeros(1, 27, 2, 2)
    lay[0, 0], lay[0, 1], lay[0, 24:27] = 1.0, 2.0, 3.0
    lay._sg_cmap = cmap
    dense = L.expand_layout(lay, 9)
    assert dense.shape == (1, 9, 2, 2)
    assert dense[0, :, 0, 0].tolist() == [0, 2, 0, 0, 1, 0, 3, 3, 3]
    plain = torch.ones(1, 9, 2, 2)
    assert L.expand_layout(plain, 9) is plain


def test_vector_pool_plan_has_fixed_length_and_pads_to_scratch_row():
    """graph replay feeds VectorPool.plan() as a static-shape input: [O read rows | O written rows | O sources], the
    padding writes target the scratch row (index = capacity) which no read ever addresses; max_rows sizes the store
    once so that nothing grows inside a capture."""
    import random
    from scene_generation_b200.utils import VectorPool
    pool = VectorPool(2, max_rows=5 * 2)
    vecs = torch.arange(12, dtype=torch.float32).view(6, 2)
    objs = [3, 3, 1, 3, 4, 1]
    random.seed(0)
    pool.reserve(len(objs), vecs)
    cap = pool.capacity
    assert cap >= 10 and pool.store.shape == (cap + 1, 2)
    idx = pool.plan(objs)
    assert idx.shape == (3 * len(objs),) and idx.dtype == torch.long
    src, rows, vals = idx[:6], idx[6:12], idx[12:]
    assert ((src < cap) | (src > cap)).all()                  # never the scratch row
    n_written = int((rows != cap).sum())
    assert n_written == len(set(rows[rows != cap].tolist()))  # real writes address distinct pool rows
    assert (rows[n_written:] == cap).all() and (vals[n_written:] == 0).all()
    out = pool.apply(idx, vecs)
    assert out.shape == vecs.shape
    # first sighting of a class returns the object's own vector (utils.py:73-75)
    assert torch.equal(out[0], vecs[0]) and torch.equal(out[2], vecs[2]) and torch.equal(out[4], vecs[4])
    store_before = pool.store
    for step in range(20):                                    # fill every class pool: the store must never be reallocated
        pool.reserve(len(objs), vecs)
        pool.apply(pool.plan(objs), vecs)
    assert pool.store is store_before and pool.used <= 10


def test_batch_meta_geometry_keys_the_iteration_graphs():
    """Trainer keys its captured iterations on the shapes of the batch and of the loader's index tensors."""
    from scene_generation_b200.trainer import _BatchMeta
    b1 = synthetic.make_batch(3, (32, 32), 20, 2, 2, seed=1)
    b2 = synthetic.make_batch(3, (32, 32), 20, 2, 2, seed=2)      # same geometry, different content
    b3 = synthetic.make_batch(3, (32, 32), 20, 4, 4, seed=1)      # more objects
    assert _BatchMeta.of(b1) is None                              # no loader metadata attached -> eager path
    metas = [synthetic.HostMeta(b) for b in (b1, b2, b3)]
    m1, m2, m3 = (_BatchMeta.of(m.attach(b)) for m, b in zip(metas, (b1, b2, b3)))
    assert m1.geometry() == m2.geometry() != m3.geometry()
    assert m1.objs_host == b1[1].tolist() and m1.slots_used >= 1
    assert [tuple(t.shape) for t in m1.tensors()] == [(3, 2), (b1[1].numel() + 1,), (2 * b1[4].shape[0],), (b1[1].numel(),),
                                                      (3, synthetic.MAX_CLASS_SLOTS)]


def test_entry_points_validate_arguments_before_any_launch():
    """Argument errors come back as an error code + message through the C ABI (no kernel is launched, so this runs
    without a GPU): size limits of the norm/act/pad writers, malformed convolution descriptors, probe arguments."""
    d = _lib.NapDesc()
    d.src, d.N, d.H, d.W, d.C = 1 << 20, 2, 8, 8, 4096          # never dereferenced: validation comes first
    d.up, d.pad = 1, 0
    with pytest.raises(RuntimeError, match='2048 channels'):
        _lib.call('sg_norm_act_pad_fwd', ctypes.byref(d), ctypes.c_void_p(1 << 21), None)
    d.C = 12                                                     # not a multiple of 8
    with pytest.raises(RuntimeError, match='multiple of 8'):
        _lib.call('sg_norm_act_pad_fwd', ctypes.byref(d), ctypes.c_void_p(1 << 21), None)
    d.C, d.N = 64, 20000                                         # 4 N must stay below the grid limit
    with pytest.raises(RuntimeError, match='grid limits'):
        _lib.call('sg_norm_act_pad_fwd', ctypes.byref(d), ctypes.c_void_p(1 << 21), None)
    assert _lib.lib().sg_norm_act_pad_bwd_parts(2, 8, 8, 12) == 0                    # invalid sizes: no parts
    parts = _lib.lib().sg_norm_act_pad_bwd_parts(32, 128, 128, 64)
    assert 1 <= parts <= 32
    c = _lib.ConvDesc()
    c.nphases = 7
    with pytest.raises(RuntimeError, match='nphases'):
        _lib.call('sg_conv_tc', ctypes.byref(c), None)
    with pytest.raises(RuntimeError, match='sg_probe_mma_rate'):
        _lib.call('sg_probe_mma_rate', 48, 16, 0, ctypes.c_void_p(1 << 20), None)
