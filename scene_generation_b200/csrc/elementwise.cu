// Element-wise / normalisation / re-layout kernels around the tensor-core convolutions.
//
// The reference runs InstanceNorm2d / BatchNorm2d / ReLU / LeakyReLU / ReflectionPad2d / nearest
// upsampling / AvgPool2d / GlobalAvgPool / concat as separate ATen kernels (generators.py:16-91,
// layers.py:82-85,234-314, discriminators.py:99-110,184).  Here the conv epilogue already produced the
// per-(image,channel) sum / sum-of-squares, so one "operand writer" pass applies
//     out = pad/planes/upsample( act( src * scale + shift ) + residual )
// and writes the bf16 NHWC operand of the next convolution; its adjoint folds the padding halo,
// applies act' and the norm backward in a reduce + apply pair.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void load8(const bf16* p, float* f) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 v = __bfloat1622float2(h2[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}
__device__ __forceinline__ void load8f(const float* p, float* f) {   // p 32-byte aligned
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pk[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(pk);
}

// ---------------------------------------------------------------------------------------------
// cast / pack
// ---------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, long rows, int cols, long ld_src, int ld_dst,
                                const float* __restrict__ mask_y, float slope, bf16* __restrict__ dst) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = rows * ld_dst;
  if (idx >= total) return;
  long r = idx / ld_dst;
  int c = (int)(idx - r * ld_dst);
  float v = 0.f;
  if (c < cols) {
    v = src[r * ld_src + c];
    if (mask_y != nullptr && !(mask_y[r * ld_src + c] > 0.f)) v *= slope;   // relu'/leaky' from the output
  }
  dst[idx] = __float2bfloat16(v);
}

// f32 [Cout][taps][Cin] -> bf16 [Cout][taps][Cin_p]  and (optionally) bf16 [Cin][taps][Cout_p]
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int taps, int Cin, int Cin_p, int Cout_p,
                                   bf16* __restrict__ wk, bf16* __restrict__ wt) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n1 = (long)Cout * taps * Cin_p;
  long n2 = wt ? (long)Cin * taps * Cout_p : 0;
  if (idx < n1) {
    int ci = idx % Cin_p;
    long r = idx / Cin_p;   // co*taps + t
    wk[idx] = __float2bfloat16(ci < Cin ? w[r * Cin + ci] : 0.f);
  } else if (idx < n1 + n2) {
    long j = idx - n1;
    int co = j % Cout_p;
    long r = j / Cout_p;    // ci*taps + t
    int t = r % taps, ci = r / taps;
    wt[j] = __float2bfloat16(co < Cout ? w[((long)co * taps + t) * Cin + ci] : 0.f);
  }
}

// transposed operand: wt[ci][t][co] = w[co][t][ci] through a 32x32 shared-memory tile (both sides coalesced)
__global__ void pack_weight_t_kernel(const float* __restrict__ w, int Cout, int taps, int Cin, int Cout_p,
                                     bf16* __restrict__ wt) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int co = co0 + r, ci = ci0 + threadIdx.x;
    tile[r][threadIdx.x] = (co < Cout && ci < Cin) ? w[((long)co * taps + t) * Cin + ci] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < Cin && co < Cout_p) wt[((long)ci * taps + t) * Cout_p + co] = __float2bfloat16(tile[threadIdx.x][r]);
  }
}

// ---------------------------------------------------------------------------------------------
// norm finalize: conv-epilogue sums -> per-(image,channel) scale/shift (+ saved mean/rstd)
// ---------------------------------------------------------------------------------------------
__global__ void norm_finalize_kernel(const float* __restrict__ stats, int n_slots, int n_img, int C, float count, float eps,
                                     float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  // InstanceNorm2d(affine=False): the conv epilogue left n_slots partial (sum, sum of squares) pairs per
  // (image, channel); they are added in slot order (fixed order: bit-reproducible)
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * C) return;
  const int n = idx / C, c = idx - n * C;
  const float2* ps = reinterpret_cast<const float2*>(stats) + (long)n * n_slots * C + c;
  float s = 0.f, ss = 0.f;
  for (int k = 0; k < n_slots; ++k) {
    const float2 v = ps[(long)k * C];
    s += v.x;
    ss += v.y;
  }
  float mean = s / count;
  float var = fmaxf(ss / count - mean * mean, 0.f);
  float rstd = rsqrtf(var + eps);
  scale[idx] = rstd;
  shift[idx] = -mean * rstd;
  save_mean[idx] = mean;
  save_rstd[idx] = rstd;
}

// BatchNorm2d (train): statistics over all images.  One block per channel: the per-(image, slot) partial sums are
// reduced across the block in a fixed order (thread-strided serial sums, shuffle tree, warps in order), then the
// (identical) per-image scale / shift rows are written in parallel.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int n_slots, int n_img, int C, float count, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, float momentum,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float red[2][32];
  __shared__ float res[4];
  const int c = blockIdx.x;
  float s = 0.f, ss = 0.f;
  for (int r = threadIdx.x; r < n_img * n_slots; r += blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(stats)[(long)r * C + c];
    s += v.x;
    ss += v.y;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = ss = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[0][w]; ss += red[1][w]; }
    float tot = count * n_img;
    float mean = s / tot;
    float var = fmaxf(ss / tot - mean * mean, 0.f);
    float rstd = rsqrtf(var + eps);
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (tot / fmaxf(tot - 1.f, 1.f));
    }
    res[0] = g * rstd; res[1] = b - mean * g * rstd; res[2] = mean; res[3] = rstd;
  }
  __syncthreads();
  const float sc = res[0], sh = res[1], mu = res[2], rs = res[3];
  for (int n = threadIdx.x; n < n_img; n += blockDim.x) {
    const long i = (long)n * C + c;
    scale[i] = sc;
    shift[i] = sh;
    save_mean[i] = mu;
    save_rstd[i] = rs;
  }
}

// ---------------------------------------------------------------------------------------------
// operand writer (forward)
// ---------------------------------------------------------------------------------------------
struct NapArgs {
  const bf16* src;   // plain NHWC [N][H][W][C]
  int N, H, W, C;
  const float* scale;   // (N*C) or null
  const float* shift;
  int act;
  float slope;
  const bf16* res;      // residual, addressed in source coordinates; null if none
  long long res_os_img, res_os_h, res_os_w;
  int up, pad, pad_mode, planes;
};

__device__ __forceinline__ int src_coord(int hp, int pad, int pad_mode, int Hu) {
  int hu = hp - pad;
  if (hu < 0) return pad_mode ? -hu : -1;
  if (hu >= Hu) return pad_mode ? 2 * (Hu - 1) - hu : -1;
  return hu;
}

template <int ACT>
__device__ __forceinline__ float act_fwd_t(float v, float slope) {
  if (ACT == SG_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == SG_ACT_LEAKY) return v >= 0.f ? v : v * slope;
  return v;
}
template <int ACT>
__device__ __forceinline__ float act_grad_t(float z, float slope) {
  if (ACT == SG_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (ACT == SG_ACT_LEAKY) return z >= 0.f ? 1.f : slope;
  return 1.f;
}
__device__ __forceinline__ int div_up_factor(int v, int up) { return up == 1 ? v : (up == 2 ? v >> 1 : v / up); }
__device__ __forceinline__ void unpack8(const uint4& raw, float* f) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 v = __bfloat1622float2(h2[q]); f[2 * q] = v.x; f[2 * q + 1] = v.y; }
}

// These kernels stream 2 B in and 2 B out per element; what bounds them is the instruction issue rate, so the index
// arithmetic is kept out of the per-item path: blockDim.x = the C/8 channel chunks (a thread's chunk, its scale /
// shift vectors and its row base pointers are loop constants), blockDim.y = pixel lanes of ONE output row,
// blockIdx = (pixel group, output row i, image * planes + plane): no division by a run-time value except the
// upsampling factor.  Each thread issues the 16-byte loads of its ITEMS pixels before the first use.
template <int ACT, int ITEMS>
__global__ void __launch_bounds__(256) nap_fwd_kernel(NapArgs a, bf16* __restrict__ out) {
  const int Hu = a.H * a.up, Wu = a.W * a.up;
  const int Hp = Hu + 2 * a.pad, Wp = Wu + 2 * a.pad;
  const int Ho = a.planes ? (Hp + 1) / 2 : Hp, Wo = a.planes ? (Wp + 1) / 2 : Wp;
  const int ch8 = threadIdx.x * 8;
  const int i = blockIdx.y;
  const int pl = a.planes ? (int)(blockIdx.z & 3) : 0;
  const int n = a.planes ? (int)(blockIdx.z >> 2) : (int)blockIdx.z;
  const int hp = a.planes ? 2 * i + (pl >> 1) : i;
  const int hu = hp < Hp ? src_coord(hp, a.pad, a.pad_mode, Hu) : -1;
  const bool row_ok = hu >= 0 && hu < Hu;
  const int h = row_ok ? div_up_factor(hu, a.up) : 0;
  const bf16* srow = a.src + ((long)n * a.H + h) * a.W * a.C + ch8;
  const bf16* rrow = a.res ? a.res + (long)n * a.res_os_img + (long)h * a.res_os_h + ch8 : nullptr;
  bf16* orow = out + ((long)blockIdx.z * Ho + i) * Wo * a.C + ch8;
  const int j0 = blockIdx.x * (blockDim.y * ITEMS) + threadIdx.y;
  uint4 raw[ITEMS], rres[ITEMS];
  bool ok[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int j = j0 + k * blockDim.y;
    const int wp = a.planes ? 2 * j + (pl & 1) : j;
    const int wu = (j < Wo && wp < Wp) ? src_coord(wp, a.pad, a.pad_mode, Wu) : -1;
    ok[k] = row_ok && wu >= 0 && wu < Wu;
    raw[k] = make_uint4(0u, 0u, 0u, 0u);
    rres[k] = make_uint4(0u, 0u, 0u, 0u);
    if (ok[k]) {
      const int w = div_up_factor(wu, a.up);
      raw[k] = *reinterpret_cast<const uint4*>(srow + w * a.C);
      if (rrow) rres[k] = *reinterpret_cast<const uint4*>(rrow + (long)w * a.res_os_w);
    }
  }
  float sc[8], sh[8];
  if (a.scale) {
    load8f(a.scale + (long)n * a.C + ch8, sc);
    load8f(a.shift + (long)n * a.C + ch8, sh);
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int j = j0 + k * blockDim.y;
    if (j >= Wo) continue;
    float f[8];
    unpack8(raw[k], f);
    if (ok[k]) {
      if (a.scale) {
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = fmaf(f[q], sc[q], sh[q]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = act_fwd_t<ACT>(f[q], a.slope);
      if (rrow) {
        float r[8];
        unpack8(rres[k], r);
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] += r[q];
      }
    }
    store8(orow + j * a.C, f);
  }
}

// ---------------------------------------------------------------------------------------------
// operand writer (backward): fold halo/planes/upsample, act', norm backward
// ---------------------------------------------------------------------------------------------
struct NapBwdArgs {
  NapArgs f;
  const bf16* g;          // grad in the forward operand's format
  const float* save_mean; // (N*C) or null (no norm)
  const float* save_rstd;
  int bn;                 // 1: statistics shared over images (BatchNorm)
  float count;            // elements per statistic
  float* sums;            // [parts][N][C*2] partial S1 = sum g', S2 = sum g' * xhat per CTA of the reduce kernel, then
                          // [N][C*2] their sums in part order, then (BatchNorm) [C*2] batch totals
  int parts;              // partial slots per image
  const float* sums_final;  // [N][C*2] per-image sums (+ [C*2] batch totals): what the apply kernel reads
  int out_planes;         // dsrc written as parity planes (for transposed-conv producers)
  bf16* dsrc;             // [N][H][W][C] plain, or planes [N][4][ceil(H/2)][ceil(W/2)][C]
  bf16* dres;             // optional: folded grad (no act') in plain source layout
};

// accumulate the folded output-gradient for source pixel (h,w) of one image; gimg = the image's operand gradient
// already offset to the thread's 8 channels; offsets inside an image fit 32 bits (checked by the host)
__device__ __forceinline__ int operand_off(const NapArgs& a, int Ho, int Wo, int hp, int wp) {
  return a.planes ? ((((hp & 1) * 2 + (wp & 1)) * Ho + (hp >> 1)) * Wo + (wp >> 1)) * a.C : (hp * Wo + wp) * a.C;
}
__device__ __forceinline__ void add8(const bf16* p, float* acc) {
  float t[8];
  load8(p, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] += t[k];
}
// The taps are always added in the same order (operand rows: own, top mirror, bottom mirror; columns likewise; the
// up x up block row-major), so every variant below produces the same bits.  The first tap is loaded unconditionally:
// in the common cases it is the only one, and its load can be issued early.
__device__ __forceinline__ void fold_grad_img(const NapArgs& a, const bf16* gimg, int Ho, int Wo, int h, int w, float* acc) {
  const int Hu = a.H * a.up, Wu = a.W * a.up;
  const bool reflect = a.pad_mode && a.pad > 0;
  if (a.up == 1) {
    load8(gimg + operand_off(a, Ho, Wo, h + a.pad, w + a.pad), acc);
    if (!reflect) return;
    // a source pixel has mirror images in the halo only within pad of a border (and not ON the border)
    const int r1 = (h >= 1 && h <= a.pad) ? a.pad - h : -1;
    const int r2 = (h <= Hu - 2 && h >= Hu - 1 - a.pad) ? a.pad + 2 * (Hu - 1) - h : -1;
    const int c1 = (w >= 1 && w <= a.pad) ? a.pad - w : -1;
    const int c2 = (w <= Wu - 2 && w >= Wu - 1 - a.pad) ? a.pad + 2 * (Wu - 1) - w : -1;
    if ((r1 & r2 & c1 & c2) < 0) return;          // all four are -1: no mirror
    const int r0 = h + a.pad, c0 = w + a.pad;
#pragma unroll
    for (int ri = 0; ri < 3; ++ri) {
      const int hp = ri == 0 ? r0 : (ri == 1 ? r1 : r2);
      if (hp < 0) continue;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const int wp = ci == 0 ? c0 : (ci == 1 ? c1 : c2);
        if (wp < 0 || (ri == 0 && ci == 0)) continue;
        add8(gimg + operand_off(a, Ho, Wo, hp, wp), acc);
      }
    }
    return;
  }
  if (a.up == 2 && !reflect) {
    // nearest x2: the 2 x 2 operand pixels of the source pixel
    const int hp = 2 * h + a.pad, wp = 2 * w + a.pad;
    float t1[8], t2[8], t3[8];
    load8(gimg + operand_off(a, Ho, Wo, hp, wp), acc);
    load8(gimg + operand_off(a, Ho, Wo, hp, wp + 1), t1);
    load8(gimg + operand_off(a, Ho, Wo, hp + 1, wp), t2);
    load8(gimg + operand_off(a, Ho, Wo, hp + 1, wp + 1), t3);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = ((acc[k] + t1[k]) + t2[k]) + t3[k];
    return;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int uh = 0; uh < a.up; ++uh) {
    const int hu = h * a.up + uh;
    const int r0 = hu + a.pad;
    const int r1 = (reflect && hu >= 1 && hu <= a.pad) ? a.pad - hu : -1;
    const int r2 = (reflect && hu <= Hu - 2 && hu >= Hu - 1 - a.pad) ? a.pad + 2 * (Hu - 1) - hu : -1;
    for (int uw = 0; uw < a.up; ++uw) {
      const int wu = w * a.up + uw;
      const int c0 = wu + a.pad;
      const int c1 = (reflect && wu >= 1 && wu <= a.pad) ? a.pad - wu : -1;
      const int c2 = (reflect && wu <= Wu - 2 && wu >= Wu - 1 - a.pad) ? a.pad + 2 * (Wu - 1) - wu : -1;
#pragma unroll
      for (int ri = 0; ri < 3; ++ri) {
        const int hp = ri == 0 ? r0 : (ri == 1 ? r1 : r2);
        if (hp < 0) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const int wp = ci == 0 ? c0 : (ci == 1 ? c1 : c2);
          if (wp < 0) continue;
          add8(gimg + operand_off(a, Ho, Wo, hp, wp), acc);
        }
      }
    }
  }
}

// SINGLE: no upsampling and no reflection halo, every source pixel has exactly one operand pixel
template <bool SINGLE>
__device__ __forceinline__ uint4 fold_load(const NapArgs& a, const bf16* gimg, int Ho, int Wo, int h, int w) {
  return *reinterpret_cast<const uint4*>(gimg + operand_off(a, Ho, Wo, h + a.pad, w + a.pad));
}

__device__ __forceinline__ void operand_dims(const NapArgs& a, int* Ho, int* Wo) {
  const int Hp = a.H * a.up + 2 * a.pad, Wp = a.W * a.up + 2 * a.pad;
  *Ho = a.planes ? (Hp + 1) / 2 : Hp;
  *Wo = a.planes ? (Wp + 1) / 2 : Wp;
}

// per-thread constants of the backward kernels (the thread's 8 channels of its image)
struct NapChan {
  float sc[8], sh[8], mu[8], rs[8];
};
template <int ACT>
__device__ __forceinline__ void nap_chan_load(const NapBwdArgs& b, int n, int ch8, NapChan& c) {
  const NapArgs& a = b.f;
  const long pc = (long)n * a.C + ch8;
  if (a.scale && ACT != SG_ACT_NONE) {
    load8f(a.scale + pc, c.sc);
    load8f(a.shift + pc, c.sh);
  }
  if (b.save_mean) {
    load8f(b.save_mean + pc, c.mu);
    load8f(b.save_rstd + pc, c.rs);
  }
}
// gp = folded gradient (in) -> g' = gp * act'(z) (out); xh = xhat (0 without a norm)
template <int ACT>
__device__ __forceinline__ void gprime_apply(const NapBwdArgs& b, const NapChan& c, const float* x, float* gp, float* xh) {
  const NapArgs& a = b.f;
  if (ACT != SG_ACT_NONE) {
    if (a.scale) {
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] *= act_grad_t<ACT>(fmaf(x[k], c.sc[k], c.sh[k]), a.slope);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] *= act_grad_t<ACT>(x[k], a.slope);
    }
  }
  if (b.save_mean) {
#pragma unroll
    for (int k = 0; k < 8; ++k) xh[k] = (x[k] - c.mu[k]) * c.rs[k];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) xh[k] = 0.f;
  }
}

// grid (parts, N), block (C/8 chunks, pixel lanes): a CTA sums its pixel range of one image; (h, w) of a lane's
// pixels advance incrementally (one division per thread)
template <int ACT, bool SINGLE>
__global__ void __launch_bounds__(256, 2) nap_bwd_reduce_kernel(NapBwdArgs b, int pix_per_block) {
  extern __shared__ float red[];                // [lanes][nC*16]: (S1,S2) pairs per channel
  const NapArgs& a = b.f;
  const int nC = blockDim.x, lanes = blockDim.y;
  const int ch = threadIdx.x, lane = threadIdx.y;
  const int n = blockIdx.y;
  const int HW = a.H * a.W;
  int Ho, Wo;
  operand_dims(a, &Ho, &Wo);
  const bf16* gimg = b.g + (long)n * (a.planes ? 4 : 1) * Ho * Wo * a.C + ch * 8;
  const bf16* simg = a.src + (long)n * HW * a.C + ch * 8;
  NapChan c;
  nap_chan_load<ACT>(b, n, ch * 8, c);
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(p_begin + pix_per_block, HW);
  const int dh = lanes / a.W, dw = lanes - dh * a.W;
  int p = p_begin + lane;
  int h = p / a.W, w = p - h * a.W;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
#pragma unroll 2
  for (; p < p_end; p += lanes) {
    float gp[8], xh[8], x[8];
    if (SINGLE) unpack8(fold_load<true>(a, gimg, Ho, Wo, h, w), gp);
    else fold_grad_img(a, gimg, Ho, Wo, h, w, gp);
    load8(simg + p * a.C, x);
    gprime_apply<ACT>(b, c, x, gp, xh);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += gp[k];
      s2[k] += gp[k] * xh[k];
    }
    h += dh;
    w += dw;
    if (w >= a.W) { w -= a.W; ++h; }
  }
  const int tid = lane * nC + ch;
  float* mine = red + tid * 16;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mine[2 * k] = s1[k];
    mine[2 * k + 1] = s2[k];
  }
  __syncthreads();
  // one partial per (CTA, image, channel, sum), lanes added in lane order: no atomics, fixed order; sg_sum_parts adds
  // the parts of an image in part order (a single part is written straight to the final slot)
  float* dst = b.sums + ((long)(b.parts > 1 ? blockIdx.x : 0) * a.N + n) * a.C * 2;
  for (int t = tid; t < nC * 16; t += nC * lanes) {
    float v = 0.f;
    for (int l = 0; l < lanes; ++l) v += red[l * nC * 16 + t];
    dst[t] = v;
  }
}

// BatchNorm: totals[c*2+j] = sum over images of the per-image sums, stored behind them.  Block = 32 columns x 32 image
// lanes: lane r adds images r, r + 32, ... in order, the 32 lane sums are then added in lane order (fixed order).
__global__ void __launch_bounds__(1024) bn_total_kernel(float* sums, int N, int C) {
  __shared__ float part[32][33];
  const int t = blockIdx.x * 32 + threadIdx.x;
  float v = 0.f;
  if (t < C * 2)
    for (int n = threadIdx.y; n < N; n += 32) v += sums[(long)n * C * 2 + t];
  part[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0 && t < C * 2) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) a += part[r][threadIdx.x];
    sums[(long)N * C * 2 + t] = a;
  }
}

// grid (pixel group, h, n), block (C/8 chunks, pixel lanes of the source row); a thread loads the gradient taps and the
// source of its ITEMS pixels before the first store
template <int ACT, bool SINGLE, int ITEMS>
__global__ void __launch_bounds__(256, 2) nap_bwd_apply_kernel(NapBwdArgs b) {
  const NapArgs& a = b.f;
  const int ch8 = threadIdx.x * 8;
  const int h = blockIdx.y, n = blockIdx.z;
  int Ho, Wo;
  operand_dims(a, &Ho, &Wo);
  const bf16* gimg = b.g + (long)n * (a.planes ? 4 : 1) * Ho * Wo * a.C + ch8;
  const bf16* srow = a.src + ((long)n * a.H + h) * a.W * a.C + ch8;
  const int w0 = blockIdx.x * (blockDim.y * ITEMS) + threadIdx.y;
  uint4 graw[ITEMS], xraw[ITEMS];
  float gp[ITEMS][8];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int w = w0 + j * blockDim.y;
    graw[j] = make_uint4(0u, 0u, 0u, 0u);
    xraw[j] = make_uint4(0u, 0u, 0u, 0u);
    if (w < a.W) {
      if (SINGLE) graw[j] = fold_load<true>(a, gimg, Ho, Wo, h, w);
      else fold_grad_img(a, gimg, Ho, Wo, h, w, gp[j]);
      xraw[j] = *reinterpret_cast<const uint4*>(srow + w * a.C);
    }
  }
  NapChan c;
  nap_chan_load<ACT>(b, n, ch8, c);
  float osc[8], m1[8], m2[8];
  if (a.scale) load8f(a.scale + (long)n * a.C + ch8, osc);          // scale = rstd (* gamma)
  if (b.save_mean) {
    float s01[8], s23[8];
    const float* sm = b.sums_final + ((long)(b.bn ? a.N : n) * a.C + ch8) * 2;   // (S1,S2) pairs of 8 channels
    load8f(sm, s01);
    load8f(sm + 8, s23);
    const float inv = 1.f / b.count;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      m1[k] = (k < 4 ? s01[2 * k] : s23[2 * k - 8]) * inv;
      m2[k] = (k < 4 ? s01[2 * k + 1] : s23[2 * k - 7]) * inv;
    }
  }
  bf16* drow;
  if (b.out_planes) {
    const int Hh = (a.H + 1) / 2, Wh = (a.W + 1) / 2;
    drow = b.dsrc + (((long)n * 4 + (h & 1) * 2) * Hh + (h >> 1)) * Wh * a.C + ch8;
  } else {
    drow = b.dsrc + ((long)n * a.H + h) * a.W * a.C + ch8;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int w = w0 + j * blockDim.y;
    if (w >= a.W) continue;
    if (SINGLE) unpack8(graw[j], gp[j]);
    if (b.dres) store8(b.dres + (long)n * a.res_os_img + (long)h * a.res_os_h + (long)w * a.res_os_w + ch8, gp[j]);
    float x[8], xh[8], o[8];
    unpack8(xraw[j], x);
    gprime_apply<ACT>(b, c, x, gp[j], xh);
    if (b.save_mean) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = osc[k] * (gp[j][k] - m1[k] - xh[k] * m2[k]);
    } else if (a.scale) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = gp[j][k] * osc[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = gp[j][k];
    }
    int off;
    if (b.out_planes) {
      const int Hh = (a.H + 1) / 2, Wh = (a.W + 1) / 2;
      off = ((w & 1) * Hh * Wh + (w >> 1)) * a.C;
    } else {
      off = w * a.C;
    }
    store8(drow + off, o);
  }
}

// InstanceNorm backward of SMALL maps (H*W <= 256) in one kernel (SG_NAP_FUSED=0 falls back to reduce + apply).  nap_bwd_reduce + nap_bwd_apply read the gradient operand and the source twice, need a memset and — on the
// 8x8 resblock maps — are latency-bound (20 + 12 us for 4 MB).  Here one CTA owns (image, slab of SC 8-channel chunks):
// every (pixel, chunk) item is loaded once into registers, the per-channel sums S1 = sum g', S2 = sum g' xhat are
// reduced inside the CTA (deterministic, no atomics, no workspace) and the gradient is applied from the registers.
constexpr int NAPF_THREADS = 256;
constexpr int NAPF_MAX_ITEMS = 4;      // items per thread: H*W*SC <= 1024

template <int ACT, bool SINGLE>
__global__ void __launch_bounds__(NAPF_THREADS, 2) nap_bwd_fused_kernel(NapBwdArgs b, int SC, int sc_shift) {
  __shared__ float part[NAPF_THREADS][17];        // per-thread partial (S1,S2) x 8 channels (+1: bank spread)
  __shared__ float stat[32][16];                  // per chunk of the slab: m1[8], m2[8]
  const NapArgs& a = b.f;
  const int nC = a.C / 8;
  const int HW = a.H * a.W;
  const int n = blockIdx.y;
  const int ch0 = blockIdx.x * SC;
  const int c = threadIdx.x & (SC - 1);           // SC is a power of two dividing NAPF_THREADS: one chunk per thread
  const int ch = ch0 + c;
  const bool live = ch < nC;
  int Ho, Wo;
  operand_dims(a, &Ho, &Wo);
  const bf16* gimg = b.g + (long)n * (a.planes ? 4 : 1) * Ho * Wo * a.C + ch * 8;
  const bf16* simg = a.src + (long)n * HW * a.C + ch * 8;
  NapChan cst;
  if (live) nap_chan_load<ACT>(b, n, ch * 8, cst);
  float gp[NAPF_MAX_ITEMS][8], xh[NAPF_MAX_ITEMS][8];
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NAPF_MAX_ITEMS; ++j) {
    const int p = (threadIdx.x + j * NAPF_THREADS) >> sc_shift;
    if (p < HW && live) {
      const int h = p / a.W, w = p - h * a.W;
      float x[8];
      if (SINGLE) unpack8(fold_load<true>(a, gimg, Ho, Wo, h, w), gp[j]);
      else fold_grad_img(a, gimg, Ho, Wo, h, w, gp[j]);
      load8(simg + p * a.C, x);
      if (b.dres) store8(b.dres + (long)n * a.res_os_img + (long)h * a.res_os_h + (long)w * a.res_os_w + ch * 8, gp[j]);
      gprime_apply<ACT>(b, cst, x, gp[j], xh[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1[k] += gp[j][k];
        s2[k] += gp[j][k] * xh[j][k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    part[threadIdx.x][2 * k] = s1[k];
    part[threadIdx.x][2 * k + 1] = s2[k];
  }
  __syncthreads();
  if (threadIdx.x < SC * 16) {
    const int cc = threadIdx.x / 16, e = threadIdx.x % 16;
    float v = 0.f;
    for (int t = cc; t < NAPF_THREADS; t += SC) v += part[t][e];     // fixed order: deterministic
    stat[cc][e] = v / b.count;
  }
  __syncthreads();
  if (!live) return;
  float sc[8];
  load8f(a.scale + (long)n * a.C + ch * 8, sc);
  const int Hh = (a.H + 1) / 2, Wh = (a.W + 1) / 2;
  bf16* dimg = b.dsrc + (long)n * (b.out_planes ? 4 * Hh * Wh : HW) * a.C + ch * 8;
#pragma unroll
  for (int j = 0; j < NAPF_MAX_ITEMS; ++j) {
    const int p = (threadIdx.x + j * NAPF_THREADS) >> sc_shift;
    if (p >= HW) continue;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = sc[k] * (gp[j][k] - stat[c][2 * k] - xh[j][k] * stat[c][2 * k + 1]);
    int off;
    if (b.out_planes) {
      const int h = p / a.W, w = p - h * a.W;
      off = ((((h & 1) * 2 + (w & 1)) * Hh + (h >> 1)) * Wh + (w >> 1)) * a.C;
    } else {
      off = p * a.C;
    }
    store8(dimg + off, o);
  }
}

// ---------------------------------------------------------------------------------------------
// small layout / pooling kernels
// ---------------------------------------------------------------------------------------------
// f32 NCHW grad * act'(y) -> bf16 NHWC [N][H][W][Cp]   (tanh / sigmoid heads)
__global__ void act_bwd_nchw_kernel(const float* __restrict__ dy, const float* __restrict__ y, int N, int C, int H, int W,
                                    int act, int Cp, bf16* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * H * W;
  if (idx >= total) return;
  int n = idx / ((long)H * W);
  long p = idx % ((long)H * W);
  for (int c = 0; c < Cp; ++c) {
    float v = 0.f;
    if (c < C) {
      long s = ((long)n * C + c) * H * W + p;
      float yy = y[s], g = dy[s];
      v = act == SG_ACT_TANH ? g * (1.f - yy * yy) : (act == SG_ACT_SIGMOID ? g * yy * (1.f - yy) : g);
    }
    out[idx * Cp + c] = __float2bfloat16(v);
  }
}

// f32 NCHW (or f32/i64 single-channel) -> bf16 NHWC [N][H][W][Cp] at channel offset c0 (other channels untouched)
__global__ void nchw_to_nhwc_kernel(const void* __restrict__ src, int src_dtype, int N, int C, int H, int W, int Cp, int c0,
                                    bf16* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * H * W;
  if (idx >= total) return;
  int n = idx / ((long)H * W);
  long p = idx % ((long)H * W);
  for (int c = 0; c < C; ++c) {
    long s = ((long)n * C + c) * H * W + p;
    float v = src_dtype == 0 ? ((const float*)src)[s] : (float)((const long long*)src)[s];
    out[idx * Cp + c0 + c] = __float2bfloat16(v);
  }
}
// bf16 NHWC channels [c0, c0+C) -> f32 NCHW
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, int N, int C, int H, int W, int Cp, int c0,
                                    float* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * H * W;
  if (idx >= total) return;
  int n = idx / ((long)H * W);
  long p = idx % ((long)H * W);
  for (int c = 0; c < C; ++c) out[((long)n * C + c) * H * W + p] = __bfloat162float(src[idx * Cp + c0 + c]);
}

// copy bf16 NHWC [rows][Cs] into channels [0,Cs) of [rows][Cd] and zero-fill / one-hot the rest
__global__ void concat_cond_kernel(const bf16* __restrict__ src, long rows_per_img, int n_img, int Cs, int Cd,
                                   const long long* __restrict__ cls, int n_cls, bf16* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)n_img * rows_per_img * Cd;
  if (idx >= total) return;
  int c = idx % Cd;
  long r = idx / Cd;
  int n = r / rows_per_img;
  bf16 v;
  if (c < Cs) v = src[r * Cs + c];
  else v = __float2bfloat16((cls != nullptr && c - Cs < n_cls && cls[n] == c - Cs) ? 1.f : 0.f);
  out[idx] = v;
}
// adjoint: channels [0,Cs) of [rows][Cd] -> [rows][Cs]
__global__ void slice_channels_kernel(const bf16* __restrict__ src, long rows, int Cd, int Cs, bf16* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Cs) return;
  int c = idx % Cs;
  long r = idx / Cs;
  out[idx] = src[r * Cd + c];
}

// MaxPool2d(2, 2) on bf16 NHWC (torchvision VGG19 features 4/9/18/27, losses.py:187-196) and its adjoint.  The gradient
// goes to the FIRST maximum of the window in scan order (h, then w), like ATen's max_pool2d_with_indices — after a ReLU
// all-zero windows are common, so the tie rule matters.  Odd trailing rows / columns are dropped (floor mode).
__global__ void maxpool2_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, bf16* __restrict__ y) {
  const int nC = C / 8;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * Ho * Wo * nC;
  if (idx >= total) return;
  int ch = idx % nC;
  long r = idx / nC;
  int j = r % Wo; r /= Wo;
  int i = r % Ho;
  int n = r / Ho;
  float m[8];
  load8(x + (((long)n * H + 2 * i) * W + 2 * j) * C + ch * 8, m);
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    float t[8];
    load8(x + (((long)n * H + 2 * i + (q >> 1)) * W + 2 * j + (q & 1)) * C + ch * 8, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = (t[k] > m[k] || t[k] != t[k]) ? t[k] : m[k];     // NaN propagates like ATen
  }
  store8(y + idx * 8, m);
}

// one thread per (input pixel, 8-channel chunk): recompute the window's arg-max from x and take gy if it is this pixel
__global__ void maxpool2_bwd_kernel(const bf16* __restrict__ gy, const bf16* __restrict__ x, int N, int H, int W, int C, int Ho,
                                    int Wo, bf16* __restrict__ gx) {
  const int nC = C / 8;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * H * W * nC;
  if (idx >= total) return;
  int ch = idx % nC;
  long r = idx / nC;
  int w = r % W; r /= W;
  int h = r % H;
  int n = r / H;
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = 0.f;
  const int i = h >> 1, j = w >> 1;
  if (i < Ho && j < Wo) {
    const int me = (h & 1) * 2 + (w & 1);
    float m[8], g[8];
    int arg[8];
    load8(x + (((long)n * H + 2 * i) * W + 2 * j) * C + ch * 8, m);
#pragma unroll
    for (int k = 0; k < 8; ++k) arg[k] = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      float t[8];
      load8(x + (((long)n * H + 2 * i + (q >> 1)) * W + 2 * j + (q & 1)) * C + ch * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (t[k] > m[k] || t[k] != t[k]) { m[k] = t[k]; arg[k] = q; }
    }
    load8(gy + (((long)n * Ho + i) * Wo + j) * C + ch * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (arg[k] == me) ? g[k] : 0.f;
  }
  store8(gx + idx * 8, o);
}

// AvgPool2d(3, stride 2, pad 1, count_include_pad=False) on bf16 NHWC
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, bf16* __restrict__ y) {
  const int nC = C / 8;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * Ho * Wo * nC;
  if (idx >= total) return;
  int ch = idx % nC;
  long r = idx / nC;
  int j = r % Wo; r /= Wo;
  int i = r % Ho;
  int n = r / Ho;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int cnt = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      int h = 2 * i + dy, w = 2 * j + dx;
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      float t[8];
      load8(x + (((long)n * H + h) * W + w) * C + ch * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += t[k];
      ++cnt;
    }
  float inv = 1.f / (float)cnt;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] *= inv;
  store8(y + idx * 8, acc);
}
__global__ void avgpool_bwd_kernel(const bf16* __restrict__ gy, int N, int H, int W, int C, int Ho, int Wo, bf16* __restrict__ gx) {
  const int nC = C / 8;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)N * H * W * nC;
  if (idx >= total) return;
  int ch = idx % nC;
  long r = idx / nC;
  int w = r % W; r /= W;
  int h = r % H;
  int n = r / H;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int i = max((h - 1) / 2, 0); i <= (h + 1) / 2; ++i) {
    if (i >= Ho || abs(2 * i - h) > 1) continue;
    for (int j = max((w - 1) / 2, 0); j <= (w + 1) / 2; ++j) {
      if (j >= Wo || abs(2 * j - w) > 1) continue;
      int ch_cnt = (min(2 * i + 1, H - 1) - max(2 * i - 1, 0) + 1) * (min(2 * j + 1, W - 1) - max(2 * j - 1, 0) + 1);
      float t[8];
      load8(gy + (((long)n * Ho + i) * Wo + j) * C + ch * 8, t);
      float inv = 1.f / (float)ch_cnt;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += t[k] * inv;
    }
  }
  store8(gx + idx * 8, acc);
}

// GlobalAvgPool: bf16 [N][HW][C] -> f32 [N][C]   (layers.py:82-85)
__global__ void gap_fwd_kernel(const bf16* __restrict__ x, int N, int HW, int C, float* __restrict__ y) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * C) return;
  int c = idx % C, n = idx / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += __bfloat162float(x[((long)n * HW + p) * C + c]);
  y[idx] = s / (float)HW;
}
__global__ void gap_bwd_kernel(const float* __restrict__ gy, int N, int HW, int C, bf16* __restrict__ gx) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * HW * C) return;
  int c = idx % C;
  int n = idx / ((long)HW * C);
  gx[idx] = __float2bfloat16(gy[(long)n * C + c] / (float)HW);
}

// column sums of a bf16 [rows][ld] matrix (ld % 8 == 0) (bias gradient).  Deterministic: every CTA leaves one partial
// row dst[blockIdx.x][C] (lanes added in lane order); sg_sum_parts adds the partial rows in CTA order.
// thread = (8-channel chunk, row lane); 16-byte loads.
__global__ void colsum_kernel(const bf16* __restrict__ x, long rows, int C, int ld, int rows_per_block, float* __restrict__ dst) {
  extern __shared__ float red[];     // [lanes][nC*8]
  const int nC = ld / 8;
  const int lanes = blockDim.x / nC;
  const int ch = threadIdx.x % nC, lane = threadIdx.x / nC;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (lane < lanes) {
    long r0 = (long)blockIdx.x * rows_per_block;
    long r1 = min(r0 + rows_per_block, rows);
    for (long r = r0 + lane; r < r1; r += lanes) {
      float t[8];
      load8(x + r * ld + ch * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[lane * nC * 8 + ch * 8 + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * nC * 8 + c];
    dst[(long)blockIdx.x * C + c] = s;
  }
}

}  // namespace

#define LAUNCH_1D(kernel, total, stream, ...) kernel<<<sg_cdiv((total), 256), 256, 0, stream>>>(__VA_ARGS__)

extern "C" int sg_cast_pad_bf16(const float* src, long long rows, int cols, long long ld_src, int ld_dst,
                                const float* mask_y, float slope, void* dst, sg_stream_t stream) {
  SG_CHECK_ARG(rows >= 0 && cols > 0 && ld_dst >= cols && ld_src >= cols, "cast_pad_bf16: bad sizes");
  if (rows == 0) return SG_OK;
  LAUNCH_1D(cast_pad_kernel, rows * ld_dst, stream, src, rows, cols, ld_src, ld_dst, mask_y, slope, (bf16*)dst);
  SG_CHECK_LAUNCH("sg_cast_pad_bf16");
  return SG_OK;
}

extern "C" int sg_pack_weight(const float* w, int Cout, int taps, int Cin, int Cin_p, int Cout_p, void* wk, void* wt,
                              sg_stream_t stream) {
  SG_CHECK_ARG(Cout > 0 && taps > 0 && Cin > 0 && Cin_p >= Cin && Cin_p % 8 == 0, "pack_weight: bad sizes");
  SG_CHECK_ARG(wt == nullptr || (Cout_p >= Cout && Cout_p % 8 == 0), "pack_weight: bad Cout_p");
  long total = (long)Cout * taps * Cin_p;
  LAUNCH_1D(pack_weight_kernel, total, stream, w, Cout, taps, Cin, Cin_p, Cout_p, (bf16*)wk, (bf16*)nullptr);
  SG_CHECK_LAUNCH("sg_pack_weight");
  if (wt) {
    dim3 grid(sg_cdiv(Cout_p, 32), sg_cdiv(Cin, 32), taps);
    pack_weight_t_kernel<<<grid, dim3(32, 8), 0, stream>>>(w, Cout, taps, Cin, Cout_p, (bf16*)wt);
    SG_CHECK_LAUNCH("sg_pack_weight(transposed)");
  }
  return SG_OK;
}

extern "C" int sg_norm_finalize(const float* stats, int n_slots, int mode, int n_img, int C, float count, float eps, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float* scale,
                                float* shift, float* save_mean, float* save_rstd, sg_stream_t stream) {
  SG_CHECK_ARG(stats && scale && shift && save_mean && save_rstd, "norm_finalize: null pointer");
  SG_CHECK_ARG((mode == 0 || mode == 1) && n_img > 0 && C > 0 && count > 0 && n_slots > 0, "norm_finalize: bad arguments");
  if (mode == 0) {
    LAUNCH_1D(norm_finalize_kernel, n_img * C, stream, stats, n_slots, n_img, C, count, eps, scale, shift, save_mean, save_rstd);
  } else {
    const long rows = (long)n_img * n_slots;
    const int threads = rows >= 128 ? 128 : (rows > 32 ? 64 : 32);
    bn_finalize_kernel<<<C, threads, 0, stream>>>(stats, n_slots, n_img, C, count, eps, gamma, beta, running_mean, running_var,
                                                  momentum, scale, shift, save_mean, save_rstd);
  }
  SG_CHECK_LAUNCH("sg_norm_finalize");
  return SG_OK;
}

// SG_NAP_FUSED=0 disables the single-kernel InstanceNorm backward of small maps (A/B switch)
static bool nap_fused_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SG_NAP_FUSED");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// launch shape of the norm-backward reduction: threads per CTA, pixels per CTA, CTAs (= partial slots) per image
static void nap_reduce_shape(int N, int H, int W, int C, int* lanes_out, int* pix_per_block_out, int* parts_out) {
  const int nC = C / 8;
  const long HW = (long)H * W;
  int lanes = nC >= 256 ? 1 : 256 / nC;
  if (lanes > HW) lanes = (int)HW;
  // ~16 CTAs per SM over all images, at most SG_NAP_MAX_PARTS per image, at least 2 pixels per thread
  long parts = (2368 + N - 1) / N;
  if (parts > SG_NAP_MAX_PARTS) parts = SG_NAP_MAX_PARTS;
  if (parts < 1) parts = 1;
  long ppb = (HW + parts - 1) / parts;
  if (ppb < lanes * 2) ppb = lanes * 2;
  ppb = ((ppb + lanes - 1) / lanes) * lanes;
  *lanes_out = lanes;
  *pix_per_block_out = (int)ppb;
  *parts_out = (int)((HW + ppb - 1) / ppb);
}

extern "C" int sg_norm_act_pad_bwd_parts(int N, int H, int W, int C) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return 0;
  int lanes, ppb, parts;
  nap_reduce_shape(N, H, W, C, &lanes, &ppb, &parts);
  return parts;
}

static int nap_check(const sg_nap_desc_t* d) {
  SG_CHECK_ARG(d && d->src, "norm_act_pad: null pointer");
  SG_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->C % 8 == 0, "norm_act_pad: bad sizes (C must be a multiple of 8)");
  SG_CHECK_ARG(d->C <= 2048, "norm_act_pad: at most 2048 channels (one thread per 8 channels, 256 threads)");
  SG_CHECK_ARG(d->up >= 1 && d->pad >= 0, "norm_act_pad: bad up / pad");
  {
    const long Hp = (long)d->H * d->up + 2 * d->pad + 1, Wp = (long)d->W * d->up + 2 * d->pad + 1;
    SG_CHECK_ARG(Hp * Wp * d->C < (1L << 31), "norm_act_pad: one image's operand must stay below 2^31 elements");
    SG_CHECK_ARG(Hp < 65536 && (long)d->N * 4 < 65536, "norm_act_pad: grid limits (rows, images * 4 < 65536)");
  }
  SG_CHECK_ARG(d->up == 1 || d->up == 2, "norm_act_pad: up must be 1 or 2");
  SG_CHECK_ARG(d->pad >= 0 && (d->pad_mode == 0 || d->pad_mode == 1), "norm_act_pad: bad padding");
  SG_CHECK_ARG(!(d->pad_mode == 1 && d->pad >= d->H * d->up), "norm_act_pad: reflection pad must be smaller than the input");
  SG_CHECK_ARG(d->act == SG_ACT_NONE || d->act == SG_ACT_RELU || d->act == SG_ACT_LEAKY, "norm_act_pad: unsupported activation");
  return SG_OK;
}
static NapArgs nap_args(const sg_nap_desc_t* d) {
  NapArgs a;
  a.src = (const bf16*)d->src; a.N = d->N; a.H = d->H; a.W = d->W; a.C = d->C;
  a.scale = d->scale; a.shift = d->shift; a.act = d->act; a.slope = d->slope;
  a.res = (const bf16*)d->res; a.res_os_img = d->res_os_img; a.res_os_h = d->res_os_h; a.res_os_w = d->res_os_w;
  a.up = d->up; a.pad = d->pad; a.pad_mode = d->pad_mode; a.planes = d->planes;
  return a;
}

extern "C" int sg_norm_act_pad_fwd(const sg_nap_desc_t* d, void* out, sg_stream_t stream) {
  if (int e = nap_check(d)) return e;
  SG_CHECK_ARG(out != nullptr, "norm_act_pad_fwd: null output");
  NapArgs a = nap_args(d);
  const int Hp = d->H * d->up + 2 * d->pad, Wp = d->W * d->up + 2 * d->pad;
  const int Ho = d->planes ? (Hp + 1) / 2 : Hp, Wo = d->planes ? (Wp + 1) / 2 : Wp;
  const int nC = d->C / 8;
  int lanes = 256 / nC;
  if (lanes > Wo) lanes = Wo;
  const int per_lane = sg_cdiv(Wo, lanes);
  const int items = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
  const dim3 block(nC, lanes), grid(sg_cdiv(Wo, lanes * items), Ho, d->N * (d->planes ? 4 : 1));
#define NAP_FWD_LAUNCH(ACT)                                                                      \
  do {                                                                                           \
    if (items == 4) nap_fwd_kernel<ACT, 4><<<grid, block, 0, stream>>>(a, (bf16*)out);           \
    else if (items == 2) nap_fwd_kernel<ACT, 2><<<grid, block, 0, stream>>>(a, (bf16*)out);      \
    else nap_fwd_kernel<ACT, 1><<<grid, block, 0, stream>>>(a, (bf16*)out);                      \
  } while (0)
  if (d->act == SG_ACT_RELU) NAP_FWD_LAUNCH(SG_ACT_RELU);
  else if (d->act == SG_ACT_LEAKY) NAP_FWD_LAUNCH(SG_ACT_LEAKY);
  else NAP_FWD_LAUNCH(SG_ACT_NONE);
#undef NAP_FWD_LAUNCH
  SG_CHECK_LAUNCH("sg_norm_act_pad_fwd");
  return SG_OK;
}

extern "C" int sg_norm_act_pad_bwd(const sg_nap_desc_t* d, const void* grad, const float* save_mean, const float* save_rstd,
                                   int bn, float count, float* sums, int out_planes, void* dsrc, void* dres,
                                   sg_stream_t stream) {
  if (int e = nap_check(d)) return e;
  SG_CHECK_ARG(grad && dsrc, "norm_act_pad_bwd: null pointer");
  SG_CHECK_ARG(save_mean == nullptr || (save_rstd && sums && count > 0), "norm_act_pad_bwd: norm backward needs rstd/sums/count");
  NapBwdArgs b;
  b.f = nap_args(d);
  b.g = (const bf16*)grad; b.save_mean = save_mean; b.save_rstd = save_rstd; b.bn = bn; b.count = count; b.sums = sums;
  b.parts = 1;
  b.sums_final = sums;
  b.out_planes = out_planes; b.dsrc = (bf16*)dsrc; b.dres = (bf16*)dres;
  const int nC = d->C / 8;
  const bool single = d->up == 1 && !(d->pad_mode && d->pad > 0);
  if (save_mean && !bn && nap_fused_enabled() && (long)d->H * d->W <= 256) {
    // experimental single-kernel path for small InstanceNorm maps: SC chunks per CTA with H*W*SC <= 1024 items
    int SC = 8;
    while (SC > 1 && (long)d->H * d->W * SC > (long)NAPF_THREADS * NAPF_MAX_ITEMS) SC >>= 1;
    if ((long)d->H * d->W * SC <= (long)NAPF_THREADS * NAPF_MAX_ITEMS) {
      if (out_planes && ((d->H & 1) || (d->W & 1)))
        cudaMemsetAsync(dsrc, 0, sizeof(bf16) * 4 * (size_t)d->N * ((d->H + 1) / 2) * ((d->W + 1) / 2) * d->C, stream);
      int sc_shift = 0;
      while ((1 << sc_shift) < SC) ++sc_shift;
      const dim3 grid(sg_cdiv(nC, SC), d->N);
#define NAP_FUSED_LAUNCH(ACT)                                                                                  \
  do {                                                                                                         \
    if (single) nap_bwd_fused_kernel<ACT, true><<<grid, NAPF_THREADS, 0, stream>>>(b, SC, sc_shift);           \
    else nap_bwd_fused_kernel<ACT, false><<<grid, NAPF_THREADS, 0, stream>>>(b, SC, sc_shift);                 \
  } while (0)
      if (d->act == SG_ACT_RELU) NAP_FUSED_LAUNCH(SG_ACT_RELU);
      else if (d->act == SG_ACT_LEAKY) NAP_FUSED_LAUNCH(SG_ACT_LEAKY);
      else NAP_FUSED_LAUNCH(SG_ACT_NONE);
#undef NAP_FUSED_LAUNCH
      SG_CHECK_LAUNCH("sg_norm_act_pad_bwd(fused)");
      return SG_OK;
    }
  }
  if (save_mean) {
    int lanes, pix_per_block, parts;
    nap_reduce_shape(d->N, d->H, d->W, d->C, &lanes, &pix_per_block, &parts);
    const long per = (long)d->N * d->C * 2;
    float* fin = sums + (parts > 1 ? (long)parts * per : 0);       // per-image sums: behind the partials
    b.parts = parts;
    b.sums_final = fin;
    const dim3 grid(parts, d->N), block(nC, lanes);
    const size_t smem = sizeof(float) * 16 * nC * lanes;
#define NAP_RED_LAUNCH(ACT)                                                                                     \
  do {                                                                                                          \
    if (single) nap_bwd_reduce_kernel<ACT, true><<<grid, block, smem, stream>>>(b, pix_per_block);              \
    else nap_bwd_reduce_kernel<ACT, false><<<grid, block, smem, stream>>>(b, pix_per_block);                    \
  } while (0)
    if (d->act == SG_ACT_RELU) NAP_RED_LAUNCH(SG_ACT_RELU);
    else if (d->act == SG_ACT_LEAKY) NAP_RED_LAUNCH(SG_ACT_LEAKY);
    else NAP_RED_LAUNCH(SG_ACT_NONE);
#undef NAP_RED_LAUNCH
    SG_CHECK_LAUNCH("sg_norm_act_pad_bwd(reduce)");
    if (parts > 1)
      if (int e = sg_sum_parts(sums, per, parts, per, fin, stream, "sg_norm_act_pad_bwd(sum parts)")) return e;
    if (bn) {
      bn_total_kernel<<<sg_cdiv(2 * d->C, 32), dim3(32, 32), 0, stream>>>(fin, d->N, d->C);
      SG_CHECK_LAUNCH("sg_norm_act_pad_bwd(bn totals)");
    }
  }
  if (out_planes) {
    // odd sizes leave unwritten slots in the parity planes: zero them first
    if ((d->H & 1) || (d->W & 1))
      cudaMemsetAsync(dsrc, 0, sizeof(bf16) * 4 * (size_t)d->N * ((d->H + 1) / 2) * ((d->W + 1) / 2) * d->C, stream);
  }
  {
    int lanes = 256 / nC;
    if (lanes > d->W) lanes = d->W;
    const int items = sg_cdiv(d->W, lanes) >= 2 ? 2 : 1;
    const dim3 block(nC, lanes), grid(sg_cdiv(d->W, lanes * items), d->H, d->N);
#define NAP_APPLY_LAUNCH(ACT)                                                                     \
  do {                                                                                            \
    if (single && items == 2) nap_bwd_apply_kernel<ACT, true, 2><<<grid, block, 0, stream>>>(b);  \
    else if (single) nap_bwd_apply_kernel<ACT, true, 1><<<grid, block, 0, stream>>>(b);           \
    else if (items == 2) nap_bwd_apply_kernel<ACT, false, 2><<<grid, block, 0, stream>>>(b);      \
    else nap_bwd_apply_kernel<ACT, false, 1><<<grid, block, 0, stream>>>(b);                      \
  } while (0)
    if (d->act == SG_ACT_RELU) NAP_APPLY_LAUNCH(SG_ACT_RELU);
    else if (d->act == SG_ACT_LEAKY) NAP_APPLY_LAUNCH(SG_ACT_LEAKY);
    else NAP_APPLY_LAUNCH(SG_ACT_NONE);
#undef NAP_APPLY_LAUNCH
  }
  SG_CHECK_LAUNCH("sg_norm_act_pad_bwd(apply)");
  return SG_OK;
}

extern "C" int sg_act_bwd_nchw(const float* dy, const float* y, int N, int C, int H, int W, int act, int Cp, void* out,
                               sg_stream_t stream) {
  SG_CHECK_ARG(dy && y && out && Cp >= C && Cp % 8 == 0, "act_bwd_nchw: bad arguments");
  LAUNCH_1D(act_bwd_nchw_kernel, (long)N * H * W, stream, dy, y, N, C, H, W, act, Cp, (bf16*)out);
  SG_CHECK_LAUNCH("sg_act_bwd_nchw");
  return SG_OK;
}

extern "C" int sg_nchw_to_nhwc(const void* src, int src_dtype, int N, int C, int H, int W, int Cp, int c0, void* out,
                               sg_stream_t stream) {
  SG_CHECK_ARG(src && out && c0 >= 0 && c0 + C <= Cp && (src_dtype == 0 || src_dtype == 1), "nchw_to_nhwc: bad arguments");
  LAUNCH_1D(nchw_to_nhwc_kernel, (long)N * H * W, stream, src, src_dtype, N, C, H, W, Cp, c0, (bf16*)out);
  SG_CHECK_LAUNCH("sg_nchw_to_nhwc");
  return SG_OK;
}

extern "C" int sg_nhwc_to_nchw(const void* src, int N, int C, int H, int W, int Cp, int c0, float* out, sg_stream_t stream) {
  SG_CHECK_ARG(src && out && c0 >= 0 && c0 + C <= Cp, "nhwc_to_nchw: bad arguments");
  LAUNCH_1D(nhwc_to_nchw_kernel, (long)N * H * W, stream, (const bf16*)src, N, C, H, W, Cp, c0, out);
  SG_CHECK_LAUNCH("sg_nhwc_to_nchw");
  return SG_OK;
}

extern "C" int sg_concat_cond(const void* src, long long rows_per_img, int n_img, int Cs, int Cd, const long long* cls,
                              int n_cls, void* out, sg_stream_t stream) {
  SG_CHECK_ARG(src && out && Cd >= Cs && rows_per_img > 0 && n_img > 0, "concat_cond: bad arguments");
  LAUNCH_1D(concat_cond_kernel, (long)n_img * rows_per_img * Cd, stream, (const bf16*)src, rows_per_img, n_img, Cs, Cd, cls,
            n_cls, (bf16*)out);
  SG_CHECK_LAUNCH("sg_concat_cond");
  return SG_OK;
}

extern "C" int sg_slice_channels(const void* src, long long rows, int Cd, int Cs, void* out, sg_stream_t stream) {
  SG_CHECK_ARG(src && out && Cd >= Cs && rows > 0, "slice_channels: bad arguments");
  LAUNCH_1D(slice_channels_kernel, rows * Cs, stream, (const bf16*)src, rows, Cd, Cs, (bf16*)out);
  SG_CHECK_LAUNCH("sg_slice_channels");
  return SG_OK;
}

extern "C" int sg_avgpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* y, sg_stream_t stream) {
  SG_CHECK_ARG(x && y && C % 8 == 0, "avgpool_fwd: bad arguments");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  LAUNCH_1D(avgpool_fwd_kernel, (long)N * Ho * Wo * (C / 8), stream, (const bf16*)x, N, H, W, C, Ho, Wo, (bf16*)y);
  SG_CHECK_LAUNCH("sg_avgpool3x3s2_fwd");
  return SG_OK;
}

extern "C" int sg_avgpool3x3s2_bwd(const void* gy, int N, int H, int W, int C, void* gx, sg_stream_t stream) {
  SG_CHECK_ARG(gy && gx && C % 8 == 0, "avgpool_bwd: bad arguments");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  LAUNCH_1D(avgpool_bwd_kernel, (long)N * H * W * (C / 8), stream, (const bf16*)gy, N, H, W, C, Ho, Wo, (bf16*)gx);
  SG_CHECK_LAUNCH("sg_avgpool3x3s2_bwd");
  return SG_OK;
}

extern "C" int sg_maxpool2x2_fwd(const void* x, int N, int H, int W, int C, void* y, sg_stream_t stream) {
  SG_CHECK_ARG(x && y && C % 8 == 0 && H >= 2 && W >= 2 && N > 0, "maxpool2x2_fwd: bad arguments");
  int Ho = H / 2, Wo = W / 2;
  LAUNCH_1D(maxpool2_fwd_kernel, (long)N * Ho * Wo * (C / 8), stream, (const bf16*)x, N, H, W, C, Ho, Wo, (bf16*)y);
  SG_CHECK_LAUNCH("sg_maxpool2x2_fwd");
  return SG_OK;
}

extern "C" int sg_maxpool2x2_bwd(const void* gy, const void* x, int N, int H, int W, int C, void* gx, sg_stream_t stream) {
  SG_CHECK_ARG(gy && x && gx && C % 8 == 0 && H >= 2 && W >= 2 && N > 0, "maxpool2x2_bwd: bad arguments");
  int Ho = H / 2, Wo = W / 2;
  LAUNCH_1D(maxpool2_bwd_kernel, (long)N * H * W * (C / 8), stream, (const bf16*)gy, (const bf16*)x, N, H, W, C, Ho, Wo, (bf16*)gx);
  SG_CHECK_LAUNCH("sg_maxpool2x2_bwd");
  return SG_OK;
}

extern "C" int sg_gap_fwd(const void* x, int N, int HW, int C, float* y, sg_stream_t stream) {
  SG_CHECK_ARG(x && y && N > 0 && HW > 0 && C > 0, "gap_fwd: bad arguments");
  LAUNCH_1D(gap_fwd_kernel, (long)N * C, stream, (const bf16*)x, N, HW, C, y);
  SG_CHECK_LAUNCH("sg_gap_fwd");
  return SG_OK;
}

extern "C" int sg_gap_bwd(const float* gy, int N, int HW, int C, void* gx, sg_stream_t stream) {
  SG_CHECK_ARG(gy && gx && N > 0 && HW > 0 && C > 0, "gap_bwd: bad arguments");
  LAUNCH_1D(gap_bwd_kernel, (long)N * HW * C, stream, gy, N, HW, C, (bf16*)gx);
  SG_CHECK_LAUNCH("sg_gap_bwd");
  return SG_OK;
}

extern "C" int sg_colsum_bf16(const void* x, long long rows, int C, int ld, float* out, float* ws, long long ws_floats,
                              sg_stream_t stream) {
  SG_CHECK_ARG(x && out && rows > 0 && C > 0 && ld >= C, "colsum: bad arguments");
  SG_CHECK_ARG(ld % 8 == 0 && ld <= 8192, "colsum: ld must be a multiple of 8 (<= 8192)");
  const int nC = ld / 8;
  int threads = nC >= 256 ? nC : (256 / nC) * nC;
  SG_CHECK_ARG(threads <= 1024, "colsum: too many channels");
  const int lanes = threads / nC;
  int rpb = (int)((rows + SG_COLSUM_MAX_BLOCKS - 1) / SG_COLSUM_MAX_BLOCKS);
  if (rpb < lanes * 4) rpb = lanes * 4;
  const int blocks = sg_cdiv(rows, rpb);
  SG_CHECK_ARG(blocks == 1 || (ws != nullptr && ws_floats >= (long long)blocks * C),
               "colsum: needs %lld floats of workspace", (long long)blocks * C);
  size_t smem = sizeof(float) * (size_t)lanes * nC * 8;
  colsum_kernel<<<blocks, threads, smem, stream>>>((const bf16*)x, rows, C, ld, rpb, blocks == 1 ? out : ws);
  SG_CHECK_LAUNCH("sg_colsum_bf16");
  if (blocks == 1) return SG_OK;
  return sg_sum_parts(ws, C, blocks, C, out, stream, "sg_colsum_bf16(sum)");
}
