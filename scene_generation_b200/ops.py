"""Low-level tensor -> C-ABI wrappers (PyTorch tensors in, PyTorch tensors out).

Only plumbing lives here: pointer extraction, output allocation, descriptor filling.  All arithmetic
happens in libsg_b200.so.  Every wrapper enqueues on the current torch CUDA stream.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, Phase, Tap, WgradDesc, WTap

F32, BF16 = 0, 1
NCHW_F32, NHWC_BF16 = 0, 1


_stream_cache = [None]
CACHE_STREAM = [True]      # False: ask torch for the current stream at every launch (multi-stream captures, where the
                           # autograd engine switches streams between backward nodes)


def _stream():
    """cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs several microseconds, which
    adds up over ~1100 launches per step: the handle is cached until refresh_stream() is called (Model.forward
    and every Trainer sub-step call it on entry, i.e. after any `with torch.cuda.stream(...)` switch)."""
    if not CACHE_STREAM[0]:
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    s = _stream_cache[0]
    if s is None:
        s = _stream_cache[0] = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    return s


def refresh_stream():
    _stream_cache[0] = None


class on_stream:
    """`with torch.cuda.stream(s)` that also drops the cached stream handle on entry and exit, so the library launches
    follow torch's current stream into a side stream and back."""

    def __init__(self, stream):
        self.ctx = torch.cuda.stream(stream)

    def __enter__(self):
        self.ctx.__enter__()
        refresh_stream()
        return self

    def __exit__(self, *exc):
        r = self.ctx.__exit__(*exc)
        refresh_stream()
        return r


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# Per-stream scratch of the fixed-order reductions (include/sg_b200.h: sg_wgrad_desc_t.ws, sg_colsum_bf16): f32 partial
# rows / split slabs that a second kernel adds in order.  Launches of one stream are serialised, so they share a
# buffer; every stream (graph branch) gets its own.
COLSUM_MAX_BLOCKS = 296
SCRATCH_FLOATS = 16 << 20          # 64 MB per stream: split-K slabs of the weight gradients, bias-gradient partial rows
_scratch = {}


def stream_scratch(device):
    """f32 partial-sum scratch of torch's current stream on `device`"""
    key = (device.index, _stream().value)
    ws = _scratch.get(key)
    if ws is None:
        ws = _scratch[key] = torch.empty(SCRATCH_FLOATS, dtype=torch.float32, device=device)
    return ws


def colsum(x2, C, out=None):
    """f32 (C,) column sums of a bf16 (rows, ld) matrix: bias gradients (fixed summation order)"""
    rows, ld = x2.shape
    if out is None:
        out = torch.empty((C,), dtype=torch.float32, device=x2.device)
    assert out.is_contiguous() and out.numel() == C and out.dtype == torch.float32
    ws = stream_scratch(x2.device)
    _lib.call('sg_colsum_bf16', _ptr(x2), rows, C, ld, _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('libsg_b200 ops need CUDA tensors (there is no CPU fallback)')


def round_up(x, m):
    return (x + m - 1) // m * m


_MASK_DT = {torch.float32: 0, torch.int64: 1, torch.uint8: 2}


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------
def masks_to_layout_fwd(vecs, boxes, masks, ranges, H, W, align_corners=False, out_format=NCHW_F32, test_mode=False,
                        raw=False, Cp=None):
    """ranges: int32 (N,2) device tensor.  Returns (N,D,H,W) f32, or a (N,D,H,W) *view* of a
    channels-last bf16 buffer with Cp = round_up(D, 8) physical channels."""
    _need_cuda(vecs, boxes, masks, ranges)
    O, D = vecs.shape
    M = masks.shape[1]
    N = ranges.shape[0]
    vecs, boxes, masks = vecs.contiguous().float(), boxes.contiguous().float(), masks.contiguous()
    Cp = Cp or round_up(D, 8)
    assert Cp >= D and Cp % 8 == 0
    if out_format == NHWC_BF16:
        buf = torch.empty((N, H, W, Cp), dtype=torch.bfloat16, device=vecs.device)
    else:
        buf = torch.empty((N, D, H, W), dtype=torch.float32, device=vecs.device)
    args = [_ptr(vecs), _ptr(boxes), _ptr(masks), _MASK_DT[masks.dtype], _ptr(ranges), O, D, M, N, H, W,
            int(align_corners), out_format, Cp]
    if test_mode:
        ws = torch.empty(max(O, 1), dtype=torch.float32, device=vecs.device)
        _lib.call('sg_masks_to_layout_test', *args, _ptr(ws), _ptr(buf), _stream())
    else:
        _lib.call('sg_masks_to_layout_fwd', *args, _ptr(buf), _stream())
    if out_format == NHWC_BF16 and not raw:
        return buf.permute(0, 3, 1, 2)[:, :D]
    return buf


def masks_to_layout_bwd(vecs, boxes, masks, ranges, H, W, grad, align_corners=False, need_dmasks=False, channels=None):
    """grad: (N,D,H,W) f32 contiguous, or a channels-last bf16 tensor whose storage is (N,H,W,Cp).
    channels=(c0, c1): only these columns of d vecs are computed (the others are zero; c0 is rounded down to 8)."""
    O, D = vecs.shape
    M = masks.shape[1]
    N = ranges.shape[0]
    vecs, boxes, masks = vecs.contiguous().float(), boxes.contiguous().float(), masks.contiguous()
    if grad.dtype == torch.bfloat16:
        if grad.dim() == 4 and grad.shape[1] == H and grad.shape[2] == W and grad.is_contiguous():
            g = grad                                   # raw (N,H,W,Cp) buffer
        else:                                          # logical (N,D,H,W) view / tensor
            g = torch.zeros((N, H, W, round_up(D, 8)), dtype=torch.bfloat16, device=grad.device)
            g[..., :D] = grad.permute(0, 2, 3, 1)
        Cp = g.shape[3]
        fmt = NHWC_BF16
    else:
        g = grad.contiguous().float()
        Cp = round_up(D, 8)
        fmt = NCHW_F32
    dvecs = torch.empty((O, D), dtype=torch.float32, device=vecs.device)
    dmasks = torch.empty((O, M, M), dtype=torch.float32, device=vecs.device) if need_dmasks else None
    c0, c1 = channels or (0, D)
    ws = stream_scratch(vecs.device)
    _lib.call('sg_masks_to_layout_bwd', _ptr(vecs), _ptr(boxes), _ptr(masks), _MASK_DT[masks.dtype], _ptr(ranges),
              O, D, M, N, H, W, int(align_corners), fmt, Cp, _ptr(g), c0 // 8 * 8, c1, _ptr(dvecs), _ptr(dmasks), _ptr(ws),
              ws.numel(), _stream())
    return dvecs, dmasks


# ---------------------------------------------------------------------------------------------
# graph
# ---------------------------------------------------------------------------------------------
def build_incidence_csr(edges_cpu, O):
    """CSR of (triple, role) incidences per object in the reference's accumulation order
    (graph.py:100-101: all subject uses in triple order, then all object uses).  Host side, numpy."""
    e = np.asarray(edges_cpu, dtype=np.int64).reshape(-1, 2)
    T = e.shape[0]
    if T and (e.min() < 0 or e.max() >= O):
        raise IndexError('edge index out of range [0, %d)' % O)
    obj = np.concatenate([e[:, 0], e[:, 1]])
    src = np.concatenate([2 * np.arange(T), 2 * np.arange(T) + 1])
    order = np.argsort(obj, kind='stable')
    counts = np.bincount(obj, minlength=O)
    ptr = np.zeros(O + 1, dtype=np.int32)
    np.cumsum(counts, out=ptr[1:])
    return ptr, src[order].astype(np.int32)


def gconv_gather(obj_vecs, pred_vecs, edges, out_dtype=torch.float32, ld_out=None):
    _need_cuda(obj_vecs, pred_vecs, edges)
    O, Do = obj_vecs.shape
    T, Dp = pred_vecs.shape
    ld = ld_out or (2 * Do + Dp)
    out = torch.empty((T, ld), dtype=out_dtype, device=obj_vecs.device)
    _lib.call('sg_gconv_gather_fwd', _ptr(obj_vecs.contiguous()), _ptr(pred_vecs.contiguous()), _ptr(edges.contiguous()),
              O, T, Do, Dp, BF16 if out_dtype == torch.bfloat16 else F32, ld, _ptr(out), _stream())
    return out


def gconv_pool(new_t, col_o, seg_ptr, seg_src, O, H, avg=True, out_dtype=torch.float32, ld_out=None):
    ld = ld_out or H
    out = torch.empty((O, ld), dtype=out_dtype, device=new_t.device)
    assert new_t.dtype == torch.float32 and new_t.stride(1) == 1
    _lib.call('sg_gconv_pool_fwd', _ptr(new_t), new_t.stride(0), col_o, _ptr(seg_ptr), _ptr(seg_src), O, H, int(avg),
              BF16 if out_dtype == torch.bfloat16 else F32, ld, _ptr(out), _stream())
    return out


def gconv_pool_bwd(dpooled, dnew_p, edges, seg_ptr, T, H, Dout, avg=True):
    out = torch.empty((T, 2 * H + Dout), dtype=torch.float32, device=dpooled.device)
    _lib.call('sg_gconv_pool_bwd', _ptr(dpooled.contiguous()), _ptr(None if dnew_p is None else dnew_p.contiguous()),
              _ptr(edges.contiguous()), _ptr(seg_ptr), T, H, Dout, int(avg), 2 * H + Dout, _ptr(out), _stream())
    return out


def gconv_gather_bwd(dcur, seg_ptr, seg_src, O, T, Do, Dp):
    dcur = dcur.contiguous()
    dobj = torch.empty((O, Do), dtype=torch.float32, device=dcur.device)
    dpred = torch.empty((T, Dp), dtype=torch.float32, device=dcur.device)
    _lib.call('sg_gconv_gather_bwd', _ptr(dcur), dcur.stride(0), _ptr(seg_ptr), _ptr(seg_src), O, T, Do, Dp,
              _ptr(dobj), _ptr(dpred), _stream())
    return dobj, dpred


# ---------------------------------------------------------------------------------------------
# crop
# ---------------------------------------------------------------------------------------------
def crop_bbox_fwd(feats, boxes, box_to_feats, HH, WW, align_corners=False, out_format=NCHW_F32):
    _need_cuda(feats, boxes, box_to_feats)
    N, C, H, W = feats.shape
    B = boxes.shape[0]
    feats, boxes, m = feats.contiguous().float(), boxes.contiguous().float(), box_to_feats.contiguous()
    Cp = round_up(C, 8)
    if out_format == NHWC_BF16:
        out = torch.empty((B, HH, WW, Cp), dtype=torch.bfloat16, device=feats.device)
    else:
        out = torch.empty((B, C, HH, WW), dtype=torch.float32, device=feats.device)
    _lib.call('sg_crop_bbox_fwd', _ptr(feats), _ptr(boxes), _ptr(m), N, C, H, W, B, HH, WW, int(align_corners),
              out_format, Cp, _ptr(out), _stream())
    return out


def crop_bbox_bwd(grad, boxes, box_to_feats, N, C, H, W, align_corners=False, grad_format=NCHW_F32):
    B = boxes.shape[0]
    if grad_format == NHWC_BF16:
        HH, WW, Cp = grad.shape[1], grad.shape[2], grad.shape[3]
    else:
        HH, WW, Cp = grad.shape[2], grad.shape[3], round_up(C, 8)
    dfeats = torch.empty((N, C, H, W), dtype=torch.float32, device=grad.device)
    _lib.call('sg_crop_bbox_bwd', _ptr(boxes.contiguous().float()), _ptr(box_to_feats.contiguous()), N, C, H, W, B,
              HH, WW, int(align_corners), grad_format, Cp, _ptr(grad.contiguous()), _ptr(dfeats), _stream())
    return dfeats


# ---------------------------------------------------------------------------------------------
# tensor-core conv / gemm
# ---------------------------------------------------------------------------------------------
_desc_cache = {}


def conv_tc(x5, w3, y, y_strides, Hout, Wout, taps, phases=None, oh_mul=1, ow_mul=1, in_h0=0, in_w0=0,
            bias=None, act=_lib.ACT_NONE, slope=0.0, stats=False, w_rows=None, mn_cols=None):
    """x5: bf16 (N,P,H,W,C) contiguous; w3: bf16 (Cout,taps,C) contiguous — or (N,R,taps,C) per-image weights
    (channel-compacted operands), of which rows [w_rows[0], w_rows[1]) of every image are the output channels;
    y: f32/bf16 output storage
    addressed as img*os_img + (h*oh_mul+oh_off)*os_h + (w*ow_mul+ow_off)*os_w + co*os_c with
    y_strides=(os_img, os_h, os_w[, os_c]) in elements.  taps: sequence of (dh, dw, plane, wtap).
    phases: sequence of (tap_begin, ntaps, oh_off, ow_off) or None for a single phase.
    mn_cols=(c0, c1): "transposed" use of an fprop weight tensor (rows, taps, C) — the contraction runs over its rows
    and the output channels are its columns [c0, c1) (dgrad with the same bf16 copy of the weights as fprop).
    stats=True: also returns the (N, slots, Cout, 2) f32 partial sum / sum-of-squares tensor of the pre-activation
    output (sg_conv_desc_t.stats; sg_norm_finalize adds the slots).
    The filled descriptor is cached per call-site geometry; only the pointers change per call."""
    key = ('c', x5.shape, w3.shape, y.dtype, tuple(y_strides), Hout, Wout, id(taps), id(phases), oh_mul, ow_mul,
           in_h0, in_w0, act, slope, w_rows, mn_cols)
    ent = _desc_cache.get(key)
    if ent is None or ent[1] is not taps or ent[2] is not phases:
        _need_cuda(x5, w3, y)
        assert x5.dtype == torch.bfloat16 and w3.dtype == torch.bfloat16 and x5.is_contiguous() and w3.is_contiguous()
        d = ConvDesc()
        d.x_N, d.x_P, d.x_H, d.x_W, d.x_C = x5.shape
        if w3.dim() == 4:
            assert w3.shape[0] == x5.shape[0]
            _, d.w_img_rows, d.w_taps, d.w_C = w3.shape
            r0, r1 = w_rows or (0, w3.shape[1])
            d.w_Cout, d.w_row0 = r1 - r0, r0
        elif mn_cols is not None:
            assert w_rows is None and mn_cols[0] % 8 == 0 and 0 <= mn_cols[0] < mn_cols[1] <= w3.shape[2]
            d.w_rows, d.w_taps, d.w_C = w3.shape
            d.w_mn, d.w_col0, d.w_Cout = 1, mn_cols[0], mn_cols[1] - mn_cols[0]
            d.w_img_rows = d.w_row0 = 0
        else:
            assert w_rows is None
            d.w_Cout, d.w_taps, d.w_C = w3.shape
            d.w_img_rows = d.w_row0 = 0
        d.y_dtype = BF16 if y.dtype == torch.bfloat16 else F32
        d.y_os_img, d.y_os_h, d.y_os_w = y_strides[:3]
        d.y_os_c = y_strides[3] if len(y_strides) > 3 else 1
        d.Hout, d.Wout, d.oh_mul, d.ow_mul, d.in_h0, d.in_w0 = Hout, Wout, oh_mul, ow_mul, in_h0, in_w0
        ph = phases if phases is not None else [(0, len(taps), 0, 0)]
        d.nphases = len(ph)
        for i, p in enumerate(ph):
            d.phases[i] = Phase(*p)
        d.ntaps = len(taps)
        for i, tp in enumerate(taps):
            d.taps[i] = Tap(*tp)
        d.act, d.slope = act, slope
        slots = ctypes.c_int(0)
        _lib.call('sg_conv_stats_slots', ctypes.byref(d), ctypes.byref(slots))
        d.stats_slots = slots.value
        ent = (d, taps, phases, ctypes.byref(d))
        _desc_cache[key] = ent
    d = ent[0]
    d.x, d.w, d.y = x5.data_ptr(), w3.data_ptr(), y.data_ptr()
    d.bias = None if bias is None else bias.data_ptr()
    st = None
    if stats:
        st = torch.empty((d.x_N, d.stats_slots, d.w_Cout, 2), dtype=torch.float32, device=x5.device)
    d.stats = None if st is None else st.data_ptr()
    _lib.call('sg_conv_tc', ent[3], _stream())
    return (y, st) if stats else y


def wgrad_tc(dy5, x5, dw, Hred, Wred, taps, Cout, Cin, ksplit=0):
    """dy5: bf16 (N,P,H,W,Cd); x5: bf16 (N,P,H,W,Cx); dw: f32 (Cout, w_taps, dw_C), overwritten — or
    (N, Cout, w_taps, dw_C): per-image weight gradients (no reduction over images).
    taps: sequence of (dha, dwa, pa, dhb, dwb, pb, wtap)."""
    key = ('w', dy5.shape, x5.shape, dw.shape, Hred, Wred, id(taps), Cout, Cin, ksplit)
    ent = _desc_cache.get(key)
    if ent is None or ent[1] is not taps:
        _need_cuda(dy5, x5, dw)
        assert dy5.dtype == torch.bfloat16 and x5.dtype == torch.bfloat16 and dw.dtype == torch.float32
        assert dy5.is_contiguous() and x5.is_contiguous() and dw.is_contiguous()
        d = WgradDesc()
        d.N, d.dy_P, d.dy_H, d.dy_W, d.dy_C = dy5.shape
        _, d.x_P, d.x_H, d.x_W, d.x_C = x5.shape
        d.Hred, d.Wred = Hred, Wred
        d.Cout, d.Cin, d.w_taps, d.dw_C = Cout, Cin, dw.shape[-2], dw.shape[-1]
        d.per_image = int(dw.dim() == 4)
        assert dw.dim() == 3 or dw.shape[0] == dy5.shape[0]
        d.ntaps = len(taps)
        for i, tp in enumerate(taps):
            d.taps[i] = WTap(*tp, 0)
        d.ksplit = ksplit
        ent = (d, taps, None, ctypes.byref(d))
        _desc_cache[key] = ent
    d = ent[0]
    d.dy, d.x, d.dw = dy5.data_ptr(), x5.data_ptr(), dw.data_ptr()
    ws = stream_scratch(dw.device)
    d.ws, d.ws_floats = ws.data_ptr(), ws.numel()
    _lib.call('sg_wgrad_tc', ent[3], _stream())
    return dw
