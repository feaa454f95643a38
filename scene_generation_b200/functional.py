"""torch.autograd glue over the C ABI: every Function's forward AND backward call libsg_b200 kernels.

PyTorch is used for what the tier allows — device memory, streams, the autograd tape — not for
arithmetic.  Activations between tensor-core convolutions live in two bf16 formats:
  * "raw"     : plain NHWC (N,H,W,C), the output of a convolution;
  * "operand" : (N,P,Hp,Wp,C), what the next convolution's TMA boxes read (P=1 padded plane, or
                P=4 parity planes for stride-2 access).
"""
import ctypes
import os
from dataclasses import dataclass

import torch

from . import _lib, convspec, ops
from ._lib import NapDesc
from .ops import _ptr, _stream, round_up

BF = torch.bfloat16
ONE_TAP = ((0, 0, 0, 0),)
ONE_WTAP = ((0, 0, 0, 0, 0, 0, 0),)


# ---------------------------------------------------------------------------------------------
# weights: f32 masters are stored physically as [Cout][taps][Cin]; bf16 operands are repacked
# whenever the master's version counter changes (i.e. after an optimizer step)
# ---------------------------------------------------------------------------------------------
_pack_cache = {}


# Data parallel: every parameter's .grad is a view into its network's flat all-reduce buffer (ddp.FlatGradReducer),
# zero-filled before backward.  AccumulateGrad would ADD each fresh gradient into it (257 add kernels per iteration);
# instead, while a set is installed here (Trainer._backward), the first weight / bias gradient of a parameter in a
# backward pass is written by its kernel straight into the view and the Function returns None for it; later
# contributions to the same parameter (a network applied twice) go through autograd's accumulation as usual.
DIRECT_GRADS = [None]


def _direct_target(param, like):
    """param.grad if this gradient may be written into it directly (see DIRECT_GRADS), else None"""
    done = DIRECT_GRADS[0]
    g = param.grad
    if done is None or g is None or id(param) in done or g.dtype != torch.float32 or g.shape != param.shape \
            or g.stride() != param.stride() or not g.is_cuda:
        return None
    done.add(id(param))
    return g


def master3(weight, kind):
    """f32 view (Cout, taps, Cin) of a Conv2d / ConvTranspose2d / Linear weight (no copy)."""
    if weight.dim() == 2:
        assert weight.is_contiguous()
        return weight.view(weight.shape[0], 1, weight.shape[1])
    kh, kw = weight.shape[2], weight.shape[3]
    v = weight.permute(1, 2, 3, 0) if kind == 'T' else weight.permute(0, 2, 3, 1)
    if not v.is_contiguous():
        raise RuntimeError('conv weights must be stored channels-last ([Cout][kh][kw][Cin]); use '
                           'scene_generation_b200.layers to build the modules')
    return v.reshape(v.shape[0], kh * kw, v.shape[3])


def grad_like_weight(g3, weight, kind):
    """(Cout,taps,Cin) f32 gradient -> tensor with the logical shape of `weight` (no copy)."""
    if weight.dim() == 2:
        return g3.view(weight.shape)
    kh, kw = weight.shape[2], weight.shape[3]
    v = g3.view(g3.shape[0], kh, kw, g3.shape[2])
    return v.permute(3, 0, 1, 2) if kind == 'T' else v.permute(0, 3, 1, 2)


# SG_DGRAD_WT=1: dgrad reads a second, transposed bf16 copy [Cin][taps][Cout] (K-major B operand) instead of the
# fprop copy through the MN-major-B kernel variant (A/B switch; costs one more pack kernel per weight and step)
DGRAD_WT = os.environ.get('SG_DGRAD_WT', '0') == '1'
# SG_SMALL_COUT_TC=0: adjoints of tiny-Cout convolutions as direct CUDA-core kernels instead of tap-unrolled
# tensor-core GEMMs (A/B switch)
SMALL_COUT_TC = os.environ.get('SG_SMALL_COUT_TC', '1') != '0'


def packed_weights(weight, kind, need_t=False):
    """bf16 operand(s) of a master weight: wk (Cout, taps, Cin_p) — read K-major by fprop and MN-major by dgrad —
    and, only on request, the transposed copy wt (Cin, taps, Cout_p).  Re-packed when the master's version counter
    moved or after invalidate_packed()."""
    key = id(weight)
    ent = _pack_cache.get(key)
    if ent is not None and ent[0] == weight._version and ent[3] is weight and (ent[2] is not None or not need_t):
        return ent[1], ent[2]
    m3 = master3(weight.detach(), kind)
    Cout, taps, Cin = m3.shape
    stale = ent is not None and ent[3] is weight and ent[1].shape == (Cout, taps, round_up(Cin, 8))
    # A stale entry of the same weight is re-packed IN PLACE: captured CUDA graphs (and PackedAdam launches inside
    # them) hold the operand's address, which must survive a load_state_dict / any in-place edit of the master.
    wk = ent[1] if stale else torch.empty((Cout, taps, round_up(Cin, 8)), dtype=BF, device=weight.device)
    wt = None
    if need_t:
        wt = ent[2] if (stale and ent[2] is not None) else torch.empty((Cin, taps, round_up(Cout, 8)), dtype=BF, device=weight.device)
    _lib.call('sg_pack_weight', _ptr(m3), Cout, taps, Cin, wk.shape[2], 0 if wt is None else wt.shape[2], _ptr(wk), _ptr(wt),
              _stream())
    _pack_cache[key] = (weight._version, wk, wt, weight, bool(stale and ent[4] is True and wt is None), Cin)
    return wk, wt


def operand_entry(weight):
    """cache entry (version, wk, wt, weight, maintained, Cin) of a master whose bf16 operand is current, else None"""
    ent = _pack_cache.get(id(weight))
    if ent is not None and ent[3] is weight and ent[0] == weight._version:
        return ent
    return None


def mark_operands_maintained(params):
    """optim.PackedAdam rewrote the operands of these masters together with the masters themselves: the entries stay
    valid (the fused update does not move Tensor._version) and survive drop_unmaintained()."""
    for p in params:
        ent = _pack_cache.get(id(p))
        if ent is not None and ent[3] is p:
            _pack_cache[id(p)] = ent[:4] + (True,) + ent[5:]


def drop_unmaintained():
    """Forget every operand that no optimizer keeps in step with its master (per-image gathered weights, frozen
    networks, weights updated by a foreign optimizer).  Called around CUDA-graph captures / replays, which change
    masters behind the version counters."""
    for k in [k for k, ent in _pack_cache.items() if not (len(ent) > 4 and ent[4] is True)]:
        del _pack_cache[k]


def packed_weights_cmap(weight, kind, cmap):
    """Per-image bf16 operands of a channel-compacted input (csrc/compact.cu): wk (N,Cout,taps,Cc) and
    wt (N,Cc,taps,Cout_p) with channel j of image n taken from dense input channel cmap[n,j].  One entry
    per weight: the layouts of one step share their channel map."""
    key = (id(weight), 'cmap')
    ent = _pack_cache.get(key)
    if ent is not None and ent[0] == weight._version and ent[3] is weight and ent[4] is cmap:
        return ent[1], ent[2]
    m3 = master3(weight.detach(), kind)
    Cout, taps, Cin = m3.shape
    N, Cc = cmap.shape
    wk = torch.empty((N, Cout, taps, Cc), dtype=BF, device=weight.device)
    wt = torch.empty((N, Cc, taps, round_up(Cout, 8)), dtype=BF, device=weight.device)
    _lib.call('sg_pack_weight_cmap', _ptr(m3), Cout, taps, Cin, _ptr(cmap), N, Cc, wt.shape[3], _ptr(wk), _ptr(wt), _stream())
    _pack_cache[key] = (weight._version, wk, wt, weight, cmap)
    return wk, wt


def clear_weight_cache():
    _pack_cache.clear()


def invalidate_packed(params):
    """Drop the bf16 operands of these masters.  The version counter alone is not enough: fused optimizers
    (torch._fused_adam_) update parameters WITHOUT bumping Tensor._version, so every optimizer step has to say
    which weights it changed (Trainer._step does)."""
    for p in params:
        _pack_cache.pop(id(p), None)
        _pack_cache.pop((id(p), 'cmap'), None)


# ---------------------------------------------------------------------------------------------
# small kernel wrappers
# ---------------------------------------------------------------------------------------------
def cast_pad(x, ld_dst=None, mask_y=None, slope=0.0):
    """f32 (rows, cols) -> bf16 (rows, ld_dst) zero padded; optional relu'/leaky' mask from the output."""
    x = x.contiguous()
    rows, cols = x.shape
    ld = ld_dst or round_up(cols, 8)
    out = torch.empty((rows, ld), dtype=BF, device=x.device)
    _lib.call('sg_cast_pad_bf16', _ptr(x), rows, cols, x.stride(0), ld, _ptr(mask_y), float(slope), _ptr(out), _stream())
    return out


_parts_cache = {}


def _nap_parts(N, H, W, C):
    """partial-sum slots per image of the norm-backward reduction (sg_norm_act_pad_bwd_parts)"""
    key = (N, H, W, C)
    v = _parts_cache.get(key)
    if v is None:
        v = _parts_cache[key] = int(_lib.lib().sg_norm_act_pad_bwd_parts(N, H, W, C))
        assert v > 0
    return v


def _nap_desc(src, scale, shift, act, slope, res, res_strides, up, pad, pad_mode, planes):
    N, H, W, C = src.shape
    d = NapDesc()
    d.src, d.N, d.H, d.W, d.C = src.data_ptr(), N, H, W, C
    d.scale = None if scale is None else scale.data_ptr()
    d.shift = None if shift is None else shift.data_ptr()
    d.act, d.slope = act, slope
    d.res = None if res is None else res
    d.res_os_img, d.res_os_h, d.res_os_w = res_strides
    d.up, d.pad, d.pad_mode, d.planes = up, pad, pad_mode, int(planes)
    return d


def operand_shape(N, H, W, C, up=1, pad=0, planes=False):
    Hp, Wp = H * up + 2 * pad, W * up + 2 * pad
    return (N, 4, (Hp + 1) // 2, (Wp + 1) // 2, C) if planes else (N, 1, Hp, Wp, C)


def nap_forward(src, scale=None, shift=None, act=_lib.ACT_NONE, slope=0.0, res_ptr=None, res_strides=(0, 0, 0),
                up=1, pad=0, pad_mode=0, planes=False):
    src = src.contiguous()
    N, H, W, C = src.shape
    out = torch.empty(operand_shape(N, H, W, C, up, pad, planes), dtype=BF, device=src.device)
    d = _nap_desc(src, scale, shift, act, slope, res_ptr, res_strides, up, pad, pad_mode, planes)
    _lib.call('sg_norm_act_pad_fwd', ctypes.byref(d), _ptr(out), _stream())
    return out


def to_planes(raw):
    """plain bf16 NHWC -> 4 parity planes (N,4,ceil(H/2),ceil(W/2),C)."""
    return nap_forward(raw, planes=True)


def act_bwd_from_output(dy, y, act, slope, out_planes=False):
    """dz = dy * act'(.) for relu / leaky fused in a conv epilogue (sign of the output = sign of the
    pre-activation).  dy, y: plain bf16 NHWC."""
    N, H, W, C = y.shape
    d = _nap_desc(y, None, None, act, slope, None, (0, 0, 0), 1, 0, 0, False)
    shape = (N, 4, (H + 1) // 2, (W + 1) // 2, C) if out_planes else (N, H, W, C)
    out = torch.empty(shape, dtype=BF, device=y.device)
    _lib.call('sg_norm_act_pad_bwd', ctypes.byref(d), _ptr(dy), None, None, 0, 1.0, None, int(out_planes), _ptr(out), None,
              _stream())
    return out


# ---------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvSpec:
    kind: str = 's1'          # 's1' stride-1, 's2' stride-2 over parity planes, 'T' ConvTranspose2d(k3,s2,p1,op1)
    k: int = 3
    pad: int = 0              # zero padding realised by the TMA out-of-bounds fill
    in_hw: tuple = None       # logical (H, W) of the un-planed input (needed for 's2')
    act: int = _lib.ACT_NONE  # epilogue activation
    slope: float = 0.0
    stats: bool = False       # emit per-(image,channel) sum / sum-sq for a following norm
    out: str = 'bf16'         # 'bf16' NHWC | 'f32_nchw' | 'f32_nhwc'
    need_dx: bool = True
    dx_channels: tuple = None  # restrict dgrad to input channels [c0, c1)


def _out_hw(spec, x5):
    _, _, Hx, Wx, _ = x5.shape
    if spec.kind == 's1':
        return Hx + 2 * spec.pad - spec.k + 1, Wx + 2 * spec.pad - spec.k + 1
    if spec.kind == 's2':
        H, W = spec.in_hw
        return (H + 2 * spec.pad - spec.k) // 2 + 1, (W + 2 * spec.pad - spec.k) // 2 + 1
    return 2 * Hx, 2 * Wx


def _conv_forward(x5, wk, bias, spec, Cout):
    N = x5.shape[0]
    Ho, Wo = _out_hw(spec, x5)
    dev = x5.device
    if spec.out == 'bf16':
        assert Cout % 8 == 0
        y = torch.empty((N, Ho, Wo, Cout), dtype=BF, device=dev)
        strides = (Ho * Wo * Cout, Wo * Cout, Cout, 1)
    elif spec.out == 'f32_nhwc':
        y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device=dev)
        strides = (Ho * Wo * Cout, Wo * Cout, Cout, 1)
    else:
        y = torch.empty((N, Cout, Ho, Wo), dtype=torch.float32, device=dev)
        strides = (Cout * Ho * Wo, Wo, 1, Ho * Wo)
    kw = dict(bias=bias, act=spec.act, slope=spec.slope, stats=spec.stats)
    if spec.kind == 's1':
        taps, off = convspec.conv_s1(spec.k, spec.pad)
        r = ops.conv_tc(x5, wk, y, strides, Ho, Wo, taps, in_h0=off, in_w0=off, **kw)
    elif spec.kind == 's2':
        r = ops.conv_tc(x5, wk, y, strides, Ho, Wo, convspec.conv_s2(spec.k, spec.pad), **kw)
    else:
        taps, phases = convspec.convT_s2(spec.k, 1)
        r = ops.conv_tc(x5, wk, y, strides, Ho // 2, Wo // 2, taps, phases=phases, oh_mul=2, ow_mul=2, **kw)
    return r if spec.stats else (y, None)       # stats: (N, slots, Cout, 2) partial sums of the epilogue


class ConvFn(torch.autograd.Function):
    """nn.Conv2d / nn.ConvTranspose2d fprop + dgrad + wgrad on tcgen05 (generators.py:69-87,
    layers.py:251,266, discriminators.py:134-156,211-233)."""

    @staticmethod
    def forward(ctx, x5, weight, bias, spec, cmap=None):
        """cmap: int32 (N, Cc) channel map of a channel-compacted operand (x5's Cc channels of image n are the
        dense input channels cmap[n]); the convolution then runs with per-image gathered weights."""
        if cmap is None:
            wk, _ = packed_weights(weight, spec.kind)
        else:
            assert spec.kind in ('s1', 's2') and cmap.dtype == torch.int32 and cmap.is_contiguous()
            assert cmap.shape == (x5.shape[0], x5.shape[4])
            wk, _ = packed_weights_cmap(weight, spec.kind, cmap)
        Cout = wk.shape[-3]
        y, stats = _conv_forward(x5, wk, bias, spec, Cout)
        ctx.spec = spec
        ctx.cmap = cmap
        ctx.save_for_backward(x5, weight, bias, y if spec.act != _lib.ACT_NONE else None)
        if stats is None:
            stats = torch.empty(0, device=x5.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        spec = ctx.spec
        x5, weight, bias, y = ctx.saved_tensors
        cmap = ctx.cmap
        m3 = master3(weight, spec.kind)
        Cout = m3.shape[0]
        wk = wt = None
        if cmap is not None:
            _, wt = packed_weights_cmap(weight, spec.kind, cmap)
            Cin, taps_n, Coutp = wt.shape[-3:]      # Cin: channels of the operand (compacted: Cc)
        else:
            if DGRAD_WT:
                _, wt = packed_weights(weight, spec.kind, need_t=True)
            else:
                wk, _ = packed_weights(weight, spec.kind)
            Cin, taps_n, Coutp = m3.shape[2], m3.shape[1], round_up(Cout, 8)
        N = x5.shape[0]
        want_planes = spec.kind == 'T'
        # ---- dz: gradient w.r.t. the pre-activation conv output, as a bf16 operand -------------
        if spec.out == 'f32_nchw':
            _, _, Ho, Wo = dy.shape
            dz = torch.empty((N, Ho, Wo, Coutp), dtype=BF, device=dy.device)
            _lib.call('sg_act_bwd_nchw', _ptr(dy.contiguous()), _ptr(y if y is not None else dy), N, Cout, Ho, Wo,
                      spec.act, Coutp, _ptr(dz), _stream())
        elif spec.out == 'f32_nhwc':
            _, Ho, Wo, _ = dy.shape
            assert spec.act in (_lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_LEAKY)
            mask = y.view(-1, Cout) if spec.act != _lib.ACT_NONE else None
            dz = cast_pad(dy.reshape(-1, Cout), Coutp, mask_y=mask, slope=spec.slope).view(N, Ho, Wo, Coutp)
        else:
            _, Ho, Wo, _ = dy.shape
            dy = dy.contiguous()
            if spec.act != _lib.ACT_NONE:
                dz = act_bwd_from_output(dy, y, spec.act, spec.slope, out_planes=want_planes)
            else:
                dz = to_planes(dy) if want_planes else dy
        if dz.dim() == 4:
            dz5 = to_planes(dz) if want_planes else dz.unsqueeze(1)
        else:
            dz5 = dz
        # ---- bias gradient ----------------------------------------------------------------------
        db = None
        if bias is not None and ctx.needs_input_grad[2]:
            tgt = _direct_target(bias, None)
            db = ops.colsum(dz5.reshape(-1, Coutp), Cout, out=tgt)
            if tgt is not None:
                db = None
        # ---- weight gradient ---------------------------------------------------------------------
        dw = None
        col = None                # tap-unrolled dz of a tiny-Cout convolution, shared by its wgrad and dgrad
        if ctx.needs_input_grad[1]:
            tgt = _direct_target(weight, m3)
            g3 = torch.empty_like(m3) if tgt is None else master3(tgt, spec.kind)
            g3_dense = g3
            if cmap is not None:    # per-image gradients of the gathered weights, scattered back below
                g3 = torch.empty((N, Cout, taps_n, Cin), dtype=torch.float32, device=dy.device)
            small = spec.kind == 's1' and spec.pad == 0 and Cout <= 4 and Cin == x5.shape[4] and Cin % 8 == 0 and cmap is None \
                and dz5.shape[1] == 1 and SMALL_COUT_TC
            if small:
                # tiny Cout (the generator's 64 -> 3 output conv): tap-unrolled dz, then ordinary tensor-core GEMMs over
                # the padded pixel grid for both adjoints (csrc/smallconv.cu)
                kk = spec.k * spec.k
                Kp = round_up(Cout * kk, 8)
                Hp, Wp = x5.shape[2], x5.shape[3]
                col = torch.empty((N, 1, Hp, Wp, Kp), dtype=BF, device=dy.device)
                _lib.call('sg_im2col_dz', _ptr(dz5), dz5.shape[4], Cout, spec.k, N, Ho, Wo, Kp, _ptr(col), _stream())
                ops.wgrad_tc(col, x5, g3.view(Cout * kk, 1, Cin), Hp, Wp, ONE_WTAP, Cout * kk, Cin)
            elif spec.kind == 's1' and spec.pad == 0 and Cout <= 3 and Cin == 64 and x5.shape[4] == 64 and spec.k in (3, 7) \
                    and dz5.shape[1] == 1:
                ws = torch.empty(296 * g3.numel(), dtype=torch.float32, device=dy.device)     # SG_WGRAD_SMALL_BLOCKS partials
                _lib.call('sg_wgrad_small_cout', _ptr(dz5), dz5.shape[4], _ptr(x5), Cout, spec.k, Cin, N, Ho, Wo, _ptr(g3),
                          _ptr(ws), ws.numel(), _stream())
            elif spec.kind == 's1':
                ops.wgrad_tc(dz5, x5, g3, Ho, Wo, convspec.wgrad_s1(spec.k, spec.pad), Cout, Cin)
            elif spec.kind == 's2':
                ops.wgrad_tc(dz5, x5, g3, Ho, Wo, convspec.wgrad_s2(spec.k, spec.pad), Cout, Cin)
            else:
                ops.wgrad_tc(dz5, x5, g3, x5.shape[2], x5.shape[3], convspec.wgrad_convT(spec.k, 1), Cout, Cin)
            if cmap is not None:
                _lib.call('sg_wgrad_cmap_scatter', _ptr(g3), _ptr(cmap), N, Cout, taps_n, Cin, m3.shape[2], _ptr(g3_dense),
                          _stream())
                g3 = g3_dense
            dw = grad_like_weight(g3, weight, spec.kind) if tgt is None else None
        # ---- input gradient (in the operand's own format) ---------------------------------------
        dx = None
        if spec.need_dx and ctx.needs_input_grad[0]:
            Cx = x5.shape[4]
            c0, c1 = spec.dx_channels or (0, Cin)
            c0 = c0 // 8 * 8          # keep the output pointer 16-byte aligned for the vector-store epilogue
            partial = (c0, c1) != (0, Cin) or Cx != Cin
            dx = (torch.zeros if partial else torch.empty)(x5.shape, dtype=BF, device=x5.device)
            # the weight operand of the adjoint: rows [c0, c1) of a transposed copy, or columns [c0, c1) of the fprop copy
            if wk is not None:
                wsub, wsel = wk, dict(mn_cols=(c0, c1))
            elif cmap is None:
                wsub, wsel = wt[c0:c1], {}
            else:
                wsub, wsel = wt, dict(w_rows=(c0, c1))
            _, P, Hx, Wx, _ = x5.shape
            base = dx.view(-1)[c0:]
            if spec.kind == 's1' and spec.pad == 0 and Cout <= 4 and Cx == Cin and Cin % 8 == 0 and not partial and cmap is None \
                    and dz5.shape[1] == 1 and SMALL_COUT_TC:
                kk = spec.k * spec.k
                Kp = round_up(Cout * kk, 8)
                if col is None:
                    col = torch.empty((N, 1, Hx, Wx, Kp), dtype=BF, device=dy.device)
                    _lib.call('sg_im2col_dz', _ptr(dz5), dz5.shape[4], Cout, spec.k, N, Ho, Wo, Kp, _ptr(col), _stream())
                # wcol[ci][co*kk + tap] = w[co][tap][ci]: the (Cin, Cout*kk) transpose of the master, as a 1-tap operand
                wcol = cast_pad(m3.detach().permute(2, 0, 1).reshape(Cin, Cout * kk), Kp).view(Cin, 1, Kp)
                ops.conv_tc(col, wcol, base, (Hx * Wx * Cx, Wx * Cx, Cx, 1), Hx, Wx, ONE_TAP)
            elif spec.kind == 's1' and spec.pad == 0 and Cout <= 4 and Cin in (32, 64) and Cx == Cin and not partial \
                    and dz5.shape[1] == 1:
                # tiny Cout (the 64 -> 3 output conv): direct CUDA-core dgrad instead of a K=3 tensor-core GEMM
                _lib.call('sg_dgrad_small_cout', _ptr(dz5), dz5.shape[4], _ptr(m3), Cout, spec.k, Cin, N, Ho, Wo, _ptr(dx),
                          _stream())
            elif spec.kind == 's1':
                ops.conv_tc(dz5, wsub, base, (Hx * Wx * Cx, Wx * Cx, Cx, 1), Hx, Wx, convspec.dgrad_s1(spec.k, spec.pad), **wsel)
            elif spec.kind == 's2':
                for (a, b, ptaps) in convspec.dgrad_s2_phase_taps(spec.k, spec.pad):
                    plane = base[(a * 2 + b) * Hx * Wx * Cx:]
                    ops.conv_tc(dz5, wsub, plane, (4 * Hx * Wx * Cx, Wx * Cx, Cx, 1), Hx, Wx, ptaps, **wsel)
            else:
                ops.conv_tc(dz5, wsub, base, (Hx * Wx * Cx, Wx * Cx, Cx, 1), Hx, Wx, convspec.dgrad_convT(spec.k, 1), **wsel)
        return dx, dw, db, None, None


def conv(x5, weight, bias, spec, cmap=None):
    y, stats = ConvFn.apply(x5, weight, bias, spec, cmap)
    return (y, stats) if spec.stats else y


class LinearFn(torch.autograd.Function):
    """nn.Linear (+ fused ReLU) as a 1-tap tcgen05 GEMM (layers.py:215-231, graph.py:85,120).  x: f32 (M, K), or a bf16
    operand (M, Kp >= K) whose pad columns are zero (the graph gather writes that directly)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        wk, _ = packed_weights(weight, 's1')
        M, Kx = x.shape
        Nout, K = weight.shape
        xb = cast_pad(x) if x.dtype != BF else x
        assert xb.shape[1] == wk.shape[2] and Kx in (K, wk.shape[2])
        y = torch.empty((M, Nout), dtype=torch.float32, device=x.device)
        if M > 0:
            ops.conv_tc(xb.view(1, 1, 1, M, xb.shape[1]), wk, y, (0, 0, Nout, 1), 1, M, ONE_TAP, bias=bias, act=act)
        ctx.act = act
        ctx.Kx = Kx
        ctx.bias_param = bias           # identity only (direct gradient writes); not needed for the arithmetic
        ctx.save_for_backward(xb, weight, y if act != _lib.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, weight, y = ctx.saved_tensors
        M, (Nout, K), Kx = dy.shape[0], weight.shape, ctx.Kx
        Np = round_up(Nout, 8)
        dz = cast_pad(dy, Np, mask_y=y, slope=0.0)
        dz5 = dz.view(1, 1, 1, M, Np)
        dx = dw = db = None
        if M == 0:
            return (torch.zeros((0, Kx), device=dy.device) if ctx.needs_input_grad[0] else None,
                    torch.zeros_like(weight), torch.zeros(Nout, device=dy.device), None)
        if ctx.needs_input_grad[0]:
            # Kx > K (padded bf16 input): the pad columns of the operand weights are zero, so are those of dx
            dx = torch.empty((M, Kx), dtype=torch.float32, device=dy.device)
            if DGRAD_WT:
                assert Kx == K
                ops.conv_tc(dz5, packed_weights(weight, 's1', need_t=True)[1], dx, (0, 0, K, 1), 1, M, ONE_TAP)
            else:
                ops.conv_tc(dz5, packed_weights(weight, 's1')[0], dx, (0, 0, Kx, 1), 1, M, ONE_TAP, mn_cols=(0, Kx))
        if ctx.needs_input_grad[1]:
            tgt = _direct_target(weight, None)
            dw = torch.empty((Nout, 1, K), dtype=torch.float32, device=dy.device) if tgt is None else tgt.view(Nout, 1, K)
            ops.wgrad_tc(dz5, xb.view(1, 1, 1, M, xb.shape[1]), dw, 1, M, ONE_WTAP, Nout, K)
            dw = dw.view(Nout, K) if tgt is None else None
        if ctx.needs_input_grad[2]:
            tgt = _direct_target(ctx.bias_param, None) if ctx.bias_param is not None else None
            db = ops.colsum(dz, Nout, out=tgt)
            if tgt is not None:
                db = None
        return dx, dw, db, None


def linear(x, weight, bias, act=_lib.ACT_NONE):
    return LinearFn.apply(x, weight, bias, act)


# ---------------------------------------------------------------------------------------------
# normalisation + activation + padding operand writer
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class NapSpec:
    norm: str = None          # None | 'in' | 'bn'
    eps: float = 1e-5
    momentum: float = 0.1
    act: int = _lib.ACT_NONE
    slope: float = 0.0
    up: int = 1
    pad: int = 0
    pad_mode: int = 0         # 0 zeros, 1 reflection
    planes: bool = False
    res_pad: int = 0          # the residual operand carries this much halo around the plain tensor


class NapFn(torch.autograd.Function):
    """InstanceNorm2d / BatchNorm2d(train) + ReLU/LeakyReLU + residual add + ReflectionPad2d / nearest
    upsample / parity-plane split in one pass (layers.py:251-267,292-314, generators.py:20-23)."""

    @staticmethod
    def forward(ctx, src, stats, gamma, beta, residual, running, spec):
        src = src.contiguous()
        N, H, W, C = src.shape
        dev = src.device
        scale = shift = mean = rstd = None
        if spec.norm == 'bn_eval':      # BatchNorm2d in eval mode: fixed affine from the running statistics
            rm, rv = running
            sc = gamma.detach() * torch.rsqrt(rv + spec.eps)
            scale, shift = sc.repeat(N).contiguous(), (beta.detach() - rm * sc).repeat(N).contiguous()
        elif spec.norm is not None:
            scale, shift, mean, rstd = torch.empty((4, N * C), dtype=torch.float32, device=dev).unbind(0)   # one allocation
            rm, rv = (running if running is not None else (None, None))
            assert stats.dim() == 4 and stats.shape[0] == N and stats.shape[2] == C     # (N, slots, C, 2)
            _lib.call('sg_norm_finalize', _ptr(stats), stats.shape[1], 0 if spec.norm == 'in' else 1, N, C, float(H * W), spec.eps,
                      _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), spec.momentum, _ptr(scale), _ptr(shift), _ptr(mean),
                      _ptr(rstd), _stream())
        res_ptr, res_strides = None, (0, 0, 0)
        if residual is not None:
            _, _, Hr, Wr, Cr = residual.shape
            p = spec.res_pad
            assert residual.is_contiguous() and Cr == C and Hr == H + 2 * p and Wr == W + 2 * p
            res_strides = (Hr * Wr * Cr, Wr * Cr, Cr)
            res_ptr = residual.data_ptr() + 2 * (p * Wr * Cr + p * Cr)
        out = nap_forward(src, scale, shift, spec.act, spec.slope, res_ptr, res_strides, spec.up, spec.pad, spec.pad_mode,
                          spec.planes)
        ctx.spec = spec
        ctx.res_shape = None if residual is None else tuple(residual.shape)
        ctx.res_strides = res_strides
        ctx.save_for_backward(src, scale, shift, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        src, scale, shift, mean, rstd = ctx.saved_tensors
        N, H, W, C = src.shape
        dev = src.device
        g = g.contiguous()
        d = _nap_desc(src, scale, shift, spec.act, spec.slope, None, ctx.res_strides, spec.up, spec.pad, spec.pad_mode,
                      spec.planes)
        dsrc = torch.empty((N, H, W, C), dtype=BF, device=dev)
        dres = dres_ptr = None
        if ctx.res_shape is not None and ctx.needs_input_grad[4]:
            assert spec.act == _lib.ACT_NONE
            dres = torch.zeros(ctx.res_shape, dtype=BF, device=dev)
            p = spec.res_pad
            dres_ptr = dres.data_ptr() + 2 * (p * ctx.res_shape[3] * C + p * C)
        sums = None
        bn = spec.norm == 'bn'
        parts = 0
        if spec.norm in ('in', 'bn'):
            parts = _nap_parts(N, H, W, C)
            # [parts][N][C][2] partials | [N][C][2] per-image sums | [C][2] batch totals (a single part: no partials)
            sums = torch.empty((((parts + 1 if parts > 1 else 1) * N + 1) * C, 2), dtype=torch.float32, device=dev)
        count = float(H * W * (N if bn else 1))
        _lib.call('sg_norm_act_pad_bwd', ctypes.byref(d), _ptr(g), _ptr(mean), _ptr(rstd), int(bn), count, _ptr(sums), 0,
                  _ptr(dsrc), dres_ptr, _stream())
        dgamma = dbeta = None
        if bn:
            dgamma, dbeta = sums[-C:, 1].contiguous(), sums[-C:, 0].contiguous()
        return dsrc, None, dgamma, dbeta, dres, None, None


def nap(src, stats=None, gamma=None, beta=None, residual=None, running=None, spec=NapSpec()):
    return NapFn.apply(src, stats, gamma, beta, residual, running, spec)


def to_planes_fn(raw):
    """differentiable plain NHWC -> parity planes"""
    return nap(raw, spec=NapSpec(planes=True))


def plain_fn(raw, pad=0):
    """differentiable plain NHWC -> (N,1,H+2p,W+2p,C) zero-haloed operand (a free view when pad=0)"""
    if pad == 0:
        return raw.unsqueeze(1)
    return nap(raw, spec=NapSpec(pad=pad))


# ---------------------------------------------------------------------------------------------
# layout / crop / graph
# ---------------------------------------------------------------------------------------------
class LayoutFn(torch.autograd.Function):
    """masks_to_layout, train branch (layout.py:64-93,149-155).  nhwc_bf16=True returns the raw
    channels-last bf16 buffer (N,H,W,Cp) that feeds the generator / discriminator operands; otherwise
    the reference's (N,D,H,W) f32 tensor."""

    @staticmethod
    def forward(ctx, vecs, boxes, masks, ranges, H, W, align_corners, nhwc_bf16, Cp=None, grad_channels=None):
        """grad_channels=(c0, c1): only these columns of vecs carry a gradient (model.py:165-168: the one-hot part of a
        layout vector is a constant); the adjoint computes just them, the other columns of d vecs are zero."""
        fmt = ops.NHWC_BF16 if nhwc_bf16 else ops.NCHW_F32
        out = ops.masks_to_layout_fwd(vecs, boxes, masks, ranges, H, W, align_corners, fmt, raw=True, Cp=Cp)
        ctx.save_for_backward(vecs, boxes, masks, ranges)
        ctx.cfg = (H, W, align_corners, grad_channels)
        return out

    @staticmethod
    def backward(ctx, g):
        vecs, boxes, masks, ranges = ctx.saved_tensors
        H, W, ac, chans = ctx.cfg
        need_dm = ctx.needs_input_grad[2] and masks.is_floating_point()
        dv, dm = ops.masks_to_layout_bwd(vecs, boxes, masks, ranges, H, W, g.contiguous(), ac, need_dmasks=need_dm,
                                         channels=chans)
        return dv, None, dm, None, None, None, None, None, None, None


class CropFn(torch.autograd.Function):
    """crop_bbox_batch (bilinear.py:26-130); output NHWC bf16 operand (Cp=8) or NCHW f32."""

    @staticmethod
    def forward(ctx, feats, boxes, box_to_feats, HH, WW, align_corners, nhwc_bf16):
        fmt = ops.NHWC_BF16 if nhwc_bf16 else ops.NCHW_F32
        out = ops.crop_bbox_fwd(feats, boxes, box_to_feats, HH, WW, align_corners, fmt)
        ctx.save_for_backward(boxes, box_to_feats)
        ctx.cfg = (tuple(feats.shape), align_corners, fmt)
        return out

    @staticmethod
    def backward(ctx, g):
        boxes, b2f = ctx.saved_tensors
        shape, ac, fmt = ctx.cfg
        return ops.crop_bbox_bwd(g.contiguous(), boxes, b2f, *shape, align_corners=ac, grad_format=fmt), None, None, None, None, None, None


class GatherFn(torch.autograd.Function):
    """[obj[s] | pred | obj[o]] row gather (graph.py:79-84), bf16 operand for net1's first GEMM."""

    @staticmethod
    def forward(ctx, obj_vecs, pred_vecs, edges, seg_ptr, seg_src, as_bf16):
        O, Do = obj_vecs.shape
        T, Dp = pred_vecs.shape
        ld = round_up(2 * Do + Dp, 8) if as_bf16 else 2 * Do + Dp
        out = ops.gconv_gather(obj_vecs.float(), pred_vecs.float(), edges, BF if as_bf16 else torch.float32, ld)
        ctx.save_for_backward(seg_ptr, seg_src)
        ctx.dims = (O, T, Do, Dp)
        return out

    @staticmethod
    def backward(ctx, g):
        seg_ptr, seg_src = ctx.saved_tensors
        O, T, Do, Dp = ctx.dims
        dobj, dpred = ops.gconv_gather_bwd(g.float(), seg_ptr, seg_src, O, T, Do, Dp)
        return dobj, dpred, None, None, None, None


class PoolFn(torch.autograd.Function):
    """per-object average of the subject / object messages (graph.py:94-116) + pass-through of new_p."""

    @staticmethod
    def forward(ctx, new_t, edges, seg_ptr, seg_src, O, H, Dout, avg):
        pooled = ops.gconv_pool(new_t, H + Dout, seg_ptr, seg_src, O, H, avg)
        new_p = new_t[:, H:H + Dout].contiguous()
        ctx.save_for_backward(edges, seg_ptr)
        ctx.dims = (new_t.shape[0], H, Dout, avg)
        return pooled, new_p

    @staticmethod
    def backward(ctx, dpooled, dnew_p):
        edges, seg_ptr = ctx.saved_tensors
        T, H, Dout, avg = ctx.dims
        return ops.gconv_pool_bwd(dpooled, dnew_p, edges, seg_ptr, T, H, Dout, avg), None, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
# pooling / concat / format conversion
# ---------------------------------------------------------------------------------------------
class GapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):           # bf16 (N,H,W,C) -> f32 (N,C)
        N, H, W, C = x.shape
        y = torch.empty((N, C), dtype=torch.float32, device=x.device)
        _lib.call('sg_gap_fwd', _ptr(x.contiguous()), N, H * W, C, _ptr(y), _stream())
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = ctx.shape
        gx = torch.empty(ctx.shape, dtype=BF, device=g.device)
        _lib.call('sg_gap_bwd', _ptr(g.contiguous().float()), N, H * W, C, _ptr(gx), _stream())
        return gx


class AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):           # bf16 (N,H,W,C)
        N, H, W, C = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, Ho, Wo, C), dtype=BF, device=x.device)
        _lib.call('sg_avgpool3x3s2_fwd', _ptr(x.contiguous()), N, H, W, C, _ptr(y), _stream())
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = ctx.shape
        gx = torch.empty(ctx.shape, dtype=BF, device=g.device)
        _lib.call('sg_avgpool3x3s2_bwd', _ptr(g.contiguous()), N, H, W, C, _ptr(gx), _stream())
        return gx


class MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) of torchvision's VGG19 (losses.py:187-196) on bf16 (N,H,W,C); gradient to the first maximum
    of every window (ATen's rule)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty((N, H // 2, W // 2, C), dtype=BF, device=x.device)
        _lib.call('sg_maxpool2x2_fwd', _ptr(x), N, H, W, C, _ptr(y), _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        N, H, W, C = x.shape
        gx = torch.empty_like(x)
        _lib.call('sg_maxpool2x2_bwd', _ptr(g.contiguous()), _ptr(x), N, H, W, C, _ptr(gx), _stream())
        return gx


class ConcatCondFn(torch.autograd.Function):
    """concat of a broadcast one-hot class vector behind the feature channels (discriminators.py:107-109)."""

    @staticmethod
    def forward(ctx, x, cls, n_cls):   # x bf16 (N,H,W,Cs)
        N, H, W, Cs = x.shape
        Cd = round_up(Cs + n_cls, 8)
        out = torch.empty((N, H, W, Cd), dtype=BF, device=x.device)
        _lib.call('sg_concat_cond', _ptr(x.contiguous()), H * W, N, Cs, Cd, _ptr(cls), n_cls, _ptr(out), _stream())
        ctx.dims = (N, H, W, Cs, Cd)
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W, Cs, Cd = ctx.dims
        gx = torch.empty((N, H, W, Cs), dtype=BF, device=g.device)
        _lib.call('sg_slice_channels', _ptr(g.contiguous()), N * H * W, Cd, Cs, _ptr(gx), _stream())
        return gx, None, None


class ImageSlotFn(torch.autograd.Function):
    """cat((layout, img), dim=1) (trainer.py:246,250,328): writes the f32 NCHW image into channels
    [c0, c0+3) of a COPY of the channels-last bf16 layout buffer.  Only the image gets a gradient —
    the layout operand is detached at every call site whose gradient is used (trainer.py:249-250)."""

    @staticmethod
    def forward(ctx, layout_nhwc, img, c0):
        N, H, W, Cp = layout_nhwc.shape
        out = layout_nhwc.clone()
        C = img.shape[1]
        _lib.call('sg_nchw_to_nhwc', _ptr(img.contiguous().float()), 0, N, C, H, W, Cp, c0, _ptr(out), _stream())
        ctx.dims = (N, C, H, W, Cp, c0)
        return out

    @staticmethod
    def backward(ctx, g):
        N, C, H, W, Cp, c0 = ctx.dims
        gi = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
        _lib.call('sg_nhwc_to_nchw', _ptr(g.contiguous()), N, C, H, W, Cp, c0, _ptr(gi), _stream())
        return None, gi, None


class ToNhwcFn(torch.autograd.Function):
    """f32 / i64 NCHW -> zero-padded bf16 NHWC (Cp) operand, gradient back to f32 NCHW."""

    @staticmethod
    def forward(ctx, x, Cp):
        N, C, H, W = x.shape
        ctx.dims = (N, C, H, W, Cp)
        dt = 1 if x.dtype == torch.int64 else 0
        xs = x.contiguous() if dt else x.contiguous().float()
        if H * W == 1 and dt == 0:
            # vectors: NCHW == NHWC, a padded cast (the pixel-per-thread kernel below would run one block)
            return cast_pad(xs.reshape(N, C), Cp).view(N, 1, 1, Cp)
        out = torch.zeros((N, H, W, Cp), dtype=BF, device=x.device)
        _lib.call('sg_nchw_to_nhwc', _ptr(xs), dt, N, C, H, W, Cp, 0, _ptr(out), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        N, C, H, W, Cp = ctx.dims
        if H * W == 1:
            return g.reshape(N, Cp)[:, :C].float().reshape(N, C, 1, 1), None
        gi = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
        _lib.call('sg_nhwc_to_nchw', _ptr(g.contiguous()), N, C, H, W, Cp, 0, _ptr(gi), _stream())
        return gi, None


class FeatureViewFn(torch.autograd.Function):
    """Logical NCHW view of a plain bf16 NHWC feature map (what the reference returns from its
    discriminators); the gradient comes back channels-last and is made contiguous NHWC."""

    @staticmethod
    def forward(ctx, x):
        return x.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        return g.permute(0, 2, 3, 1).contiguous().to(BF)
