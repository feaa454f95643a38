"""Tap / phase tables that express every dense contraction of the hot path as one call of the
tcgen05 implicit-GEMM kernel (see include/sg_b200.h, sg_conv_tc / sg_wgrad_tc).

All tables are pure index arithmetic on the host.  Weight tap index is always kh*KW + kw of the
[Cout][KH*KW][Cin] operand.  "planes" are the 4 parity planes of an NHWC tensor:
plane[2*ph+pw][i][j] = x[2*i+ph][2*j+pw], which turn a stride-2 access into stride-1 TMA boxes.
"""
from functools import lru_cache


@lru_cache(maxsize=None)
def conv_s1(k, pad=0):
    """nn.Conv2d(k, stride=1): tap (kh,kw) reads x[h+kh-pad, w+kw-pad].  With a pre-padded input
    (reflection padding materialised by the producer) call with pad=0; zero padding comes from the
    TMA out-of-bounds fill with in_h0 = in_w0 = -pad (returned second)."""
    taps = tuple((kh, kw, 0, kh * k + kw) for kh in range(k) for kw in range(k))
    return taps, -pad


@lru_cache(maxsize=None)
def conv_s2(k, pad):
    """nn.Conv2d(k, stride=2, padding=pad) over parity planes of the UNPADDED input:
    y[i,j] = sum x[2i+kh-pad, 2j+kw-pad]  ->  plane ((kh-pad)%2, (kw-pad)%2), offset floor((kh-pad)/2)."""
    taps = []
    for kh in range(k):
        for kw in range(k):
            r, c = kh - pad, kw - pad
            taps.append((r // 2, c // 2, (r % 2) * 2 + (c % 2), kh * k + kw))
    return tuple(taps)


@lru_cache(maxsize=None)
def convT_s2(k, pad):
    """nn.ConvTranspose2d(k, stride=2, padding=pad) as 4 output-parity phases:
    y[2i+a, 2j+b] = sum_{kh: (a+pad-kh) even} x[i + (a+pad-kh)/2, ...] * W[kh, kw].
    Returns (taps, phases) with phases = (tap_begin, ntaps, oh_off=a, ow_off=b)."""
    taps, phases = [], []
    for a in range(2):
        for b in range(2):
            begin = len(taps)
            for kh in range(k):
                if (a + pad - kh) % 2:
                    continue
                for kw in range(k):
                    if (b + pad - kw) % 2:
                        continue
                    taps.append(((a + pad - kh) // 2, (b + pad - kw) // 2, 0, kh * k + kw))
            phases.append((begin, len(taps) - begin, a, b))
    return tuple(taps), tuple(phases)


@lru_cache(maxsize=None)
def dgrad_s1(k, pad):
    """adjoint of conv_s1 w.r.t. its input: dx[h,w] = sum dy[h+pad-kh, w+pad-kw] * W[.,kh,kw,.]
    (weights in the transposed [Cin][taps][Cout] operand)."""
    return tuple((pad - kh, pad - kw, 0, kh * k + kw) for kh in range(k) for kw in range(k))


@lru_cache(maxsize=None)
def dgrad_s2(k, pad):
    """adjoint of conv_s2 w.r.t. its input = transposed conv of dy: phases over the input parity.
    dx[2i+a, 2j+b] = sum_{kh: (a+pad-kh) even} dy[i + (a+pad-kh)/2, ...]"""
    return convT_s2(k, pad)


@lru_cache(maxsize=None)
def dgrad_convT(k, pad):
    """adjoint of convT_s2 w.r.t. its input = stride-2 conv of dy (parity planes of dy):
    dx[i,j] = sum dy[2i+kh-pad, 2j+kw-pad] * W[kh,kw]."""
    return conv_s2(k, pad)


@lru_cache(maxsize=None)
def wgrad_s1(k, pad):
    """dW[co,kh,kw,ci] = sum dy[h,w,co] * x[h+kh-pad, w+kw-pad, ci]; entries (dha,dwa,pa,dhb,dwb,pb,wtap)."""
    return tuple((0, 0, 0, kh - pad, kw - pad, 0, kh * k + kw) for kh in range(k) for kw in range(k))


@lru_cache(maxsize=None)
def wgrad_s2(k, pad):
    """stride-2 conv: dW[co,kh,kw,ci] = sum dy[i,j,co] * x[2i+kh-pad, 2j+kw-pad, ci] (x in parity planes)."""
    out = []
    for kh in range(k):
        for kw in range(k):
            r, c = kh - pad, kw - pad
            out.append((0, 0, 0, r // 2, c // 2, (r % 2) * 2 + (c % 2), kh * k + kw))
    return tuple(out)


@lru_cache(maxsize=None)
def wgrad_convT(k, pad):
    """transposed conv: dW[ci,co,kh,kw] = sum x[i,j,ci] * dy[2i+kh-pad, 2j+kw-pad, co] (dy in parity planes)."""
    out = []
    for kh in range(k):
        for kw in range(k):
            r, c = kh - pad, kw - pad
            out.append((r // 2, c // 2, (r % 2) * 2 + (c % 2), 0, 0, 0, kh * k + kw))
    return tuple(out)


@lru_cache(maxsize=None)
def dgrad_s2_phase_taps(k, pad):
    """per-phase tap tuples of dgrad_s2 (one launch per parity plane of dx): ((a, b, taps), ...)"""
    taps, phases = convT_s2(k, pad)
    return tuple((a, b, taps[tb:tb + nt]) for (tb, nt, a, b) in phases)
