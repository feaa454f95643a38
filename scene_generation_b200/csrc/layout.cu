// masks/boxes -> spatial layout composition (reference: scene_generation/layout.py:64-184).
//
//   out[n, c, h, w] = sum_{o in image n} vecs[o, c] * S_o(h, w)
//   S_o(h, w)       = bilinear, zero-padded sample of mask_o at the box grid of layout.py:96-128
//
// The reference materialises (O,D,M,M) and (O,D,H,W) tensors; here S_o is evaluated once per
// (object, pixel) into shared memory and the D channels are produced by FMAs from it, so the only
// HBM traffic is the output write (fwd) / the gradient read (bwd): N*Cp*H*W*s bytes.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

constexpr int TP = 64;        // pixels per tile (contiguous in the flattened H*W index)
constexpr int MAXO = 32;      // objects processed per shared-memory chunk
constexpr int THREADS = 256;

struct LayoutArgs {
  const float* vecs;      // (O, D) fp32
  const float* boxes;     // (O, 4) fp32 xyxy
  const void* masks;      // (O, M, M) f32 / i64 / u8
  const int* ranges;      // (N, 2) object range per image
  int mask_dtype;         // 0 f32, 1 i64, 2 u8
  int O, D, M, N, H, W;
  int Cp;                 // physical channel count of NHWC tensors (>= D, multiple of 8)
  int align_corners;
};

__device__ __forceinline__ float load_mask(const void* masks, int dtype, long idx) {
  if (dtype == 0) return ((const float*)masks)[idx];
  if (dtype == 1) return (float)((const long long*)masks)[idx];
  return (float)((const unsigned char*)masks)[idx];
}

// S_o(h,w) and (optionally) the 4 corner indices/weights
__device__ __forceinline__ float sample_mask(const LayoutArgs& a, int o, int h, int w, const float* bx) {
  float x0 = bx[0], y0 = bx[1];
  float ww = __fsub_rn(bx[2], x0), hh = __fsub_rn(bx[3], y0);
  float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(sg_linspace01(w, a.W), x0), ww), 2.f), 1.f);
  float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(sg_linspace01(h, a.H), y0), hh), 2.f), 1.f);
  SgBilin ax = sg_axis(gx, a.M, a.align_corners);
  SgBilin ay = sg_axis(gy, a.M, a.align_corners);
  const long base = (long)o * a.M * a.M;
  float s = 0.f;
  if (ay.ok0 && ax.ok0) s = __fmaf_rn(__fmul_rn(ax.w0, ay.w0), load_mask(a.masks, a.mask_dtype, base + (long)ay.i0 * a.M + ax.i0), s);
  if (ay.ok0 && ax.ok1) s = __fmaf_rn(__fmul_rn(ax.w1, ay.w0), load_mask(a.masks, a.mask_dtype, base + (long)ay.i0 * a.M + ax.i0 + 1), s);
  if (ay.ok1 && ax.ok0) s = __fmaf_rn(__fmul_rn(ax.w0, ay.w1), load_mask(a.masks, a.mask_dtype, base + (long)(ay.i0 + 1) * a.M + ax.i0), s);
  if (ay.ok1 && ax.ok1) s = __fmaf_rn(__fmul_rn(ax.w1, ay.w1), load_mask(a.masks, a.mask_dtype, base + (long)(ay.i0 + 1) * a.M + ax.i0 + 1), s);
  return s;
}

// ------------------------------------------------------------------------------------------------
// forward, train branch (layout.py:149-155): NHWC bf16 (Cp channels, zero padded) or NCHW fp32
// ------------------------------------------------------------------------------------------------
template <bool NHWC_BF16>
__global__ void __launch_bounds__(THREADS) layout_fwd_kernel(LayoutArgs a, void* out) {
  extern __shared__ float smem[];
  float* sS = smem;                       // [MAXO][TP]
  float* sV = smem + MAXO * TP;           // [MAXO][Cp]
  __shared__ int sActive[MAXO];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * TP;
  const int HW = a.H * a.W;
  const int o_begin = a.ranges[2 * n], o_end = a.ranges[2 * n + 1];
  const int chunks = a.Cp / 8;
  constexpr int MAX_ITEMS = 8;            // TP*chunks/THREADS <= 8  (Cp <= 256)
  float acc[MAX_ITEMS][8];
#pragma unroll
  for (int i = 0; i < MAX_ITEMS; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int ob = o_begin; ob < o_end; ob += MAXO) {
    const int nobj = min(MAXO, o_end - ob);
    __syncthreads();
    if (threadIdx.x < MAXO) sActive[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < nobj * a.Cp; i += THREADS) {
      int o = i / a.Cp, c = i % a.Cp;
      sV[o * a.Cp + c] = (c < a.D) ? a.vecs[(long)(ob + o) * a.D + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nobj * TP; i += THREADS) {
      int o = i / TP, px = i % TP;
      int p = p0 + px;
      float s = 0.f;
      if (p < HW) s = sample_mask(a, ob + o, p / a.W, p % a.W, a.boxes + 4 * (ob + o));
      sS[o * TP + px] = s;
      if (s != 0.f) sActive[o] = 1;
    }
    __syncthreads();
    if (NHWC_BF16) {
#pragma unroll
      for (int it = 0; it < MAX_ITEMS; ++it) {
        int idx = it * THREADS + threadIdx.x;
        if (idx >= TP * chunks) break;
        int px = idx / chunks, ch = idx % chunks;
        for (int o = 0; o < nobj; ++o) {
          if (!sActive[o]) continue;
          float s = sS[o * TP + px];
          if (s == 0.f) continue;
          const float4* v = reinterpret_cast<const float4*>(sV + o * a.Cp + ch * 8);
          float4 v0 = v[0], v1 = v[1];
          acc[it][0] = __fmaf_rn(v0.x, s, acc[it][0]); acc[it][1] = __fmaf_rn(v0.y, s, acc[it][1]);
          acc[it][2] = __fmaf_rn(v0.z, s, acc[it][2]); acc[it][3] = __fmaf_rn(v0.w, s, acc[it][3]);
          acc[it][4] = __fmaf_rn(v1.x, s, acc[it][4]); acc[it][5] = __fmaf_rn(v1.y, s, acc[it][5]);
          acc[it][6] = __fmaf_rn(v1.z, s, acc[it][6]); acc[it][7] = __fmaf_rn(v1.w, s, acc[it][7]);
        }
      }
    } else {
      // NCHW: item = (channel group of 8 strided channels?) -> keep it simple: one (c, px) per slot
#pragma unroll
      for (int it = 0; it < MAX_ITEMS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int idx = (it * 8 + j) * THREADS + threadIdx.x;   // over D*TP, px fastest
          if (idx >= a.D * TP) break;
          int c = idx / TP, px = idx % TP;
          float r = acc[it][j];
          for (int o = 0; o < nobj; ++o) {
            float s = sS[o * TP + px];
            if (s != 0.f) r = __fmaf_rn(sV[o * a.Cp + c], s, r);
          }
          acc[it][j] = r;
        }
      }
    }
  }
  if (NHWC_BF16) {
    __nv_bfloat16* o16 = (__nv_bfloat16*)out;
#pragma unroll
    for (int it = 0; it < MAX_ITEMS; ++it) {
      int idx = it * THREADS + threadIdx.x;
      if (idx >= TP * chunks) break;
      int px = idx / chunks, ch = idx % chunks;
      int p = p0 + px;
      if (p >= HW) continue;
      __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[j] = __floats2bfloat162_rn(acc[it][2 * j], acc[it][2 * j + 1]);
      *reinterpret_cast<uint4*>(o16 + ((long)n * HW + p) * a.Cp + ch * 8) = *reinterpret_cast<uint4*>(pk);
    }
  } else {
    float* o32 = (float*)out;
#pragma unroll
    for (int it = 0; it < MAX_ITEMS; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int idx = (it * 8 + j) * THREADS + threadIdx.x;
        if (idx >= a.D * TP) break;
        int c = idx / TP, px = idx % TP;
        int p = p0 + px;
        if (p < HW) o32[((long)n * a.D + c) * HW + p] = acc[it][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward, NHWC bf16, bandwidth-oriented variant.  Per 16x8-pixel tile only the objects whose (slightly
// dilated) box intersects the tile are sampled and accumulated — typically 2-3 of the image's 4-9 —
// and the compacted list keeps the reference's object order (deterministic summation).
// ------------------------------------------------------------------------------------------------
constexpr int MAXA = 12;      // active objects per pass (12: three CTAs of the Cp = 208 layout per SM)

constexpr int TW = 16, TH = 8;       // pixel tile of the NHWC kernel: TH row segments of TW * Cp * 2 contiguous bytes
constexpr int TPX = TW * TH;

__device__ __forceinline__ bool box_touches_rect(const LayoutArgs& a, const float* bx, int h_lo, int h_hi, int w_lo, int w_hi) {
  float x0 = bx[0], y0 = bx[1], x1 = bx[2], y1 = bx[3];
  float ww = x1 - x0, hh = y1 - y0;
  if (!(isfinite(ww) && isfinite(hh) && isfinite(x0) && isfinite(y0))) return true;   // let the sampler decide
  if (ww == 0.f || hh == 0.f) return false;           // inf / NaN grid -> grid_sample yields 0 everywhere
  // bilinear taps reach one mask texel beyond the box: dilate by |extent|/(M-1) (covers both align_corners modes)
  float mx = fabsf(ww) / (float)max(a.M - 1, 1), my = fabsf(hh) / (float)max(a.M - 1, 1);
  float xa = fminf(x0, x1) - mx, xb = fmaxf(x0, x1) + mx, ya = fminf(y0, y1) - my, yb = fmaxf(y0, y1) + my;
  float sx = (float)max(a.W - 1, 1), sy = (float)max(a.H - 1, 1);
  float wl = xa * sx - 1.f, wh = xb * sx + 1.f, hl = ya * sy - 1.f, hu = yb * sy + 1.f;
  return !((float)w_hi < wl || (float)w_lo > wh || (float)h_hi < hl || (float)h_lo > hu);
}

// one axis of the sampling grid of layout.py:96-128 for pixel coordinate `i` of `n` (same operations, same order
// as sample_mask)
__device__ __forceinline__ SgBilin grid_axis(const LayoutArgs& a, int i, int n, float b0, float b1) {
  float ext = __fsub_rn(b1, b0);
  float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(sg_linspace01(i, n), b0), ext), 2.f), 1.f);
  return sg_axis(g, a.M, a.align_corners);
}

// Tile kernel.  A TW x TH pixel tile of the NHWC output is TH contiguous row segments of TW * Cp * 2 bytes: it is
// composed in shared memory and leaves with TH bulk (TMA) stores, so the FMA phase issues no global stores and no
// per-pixel address arithmetic.  The bilinear sampling grid is separable: the x taps of an object are shared by
// the TH rows of the tile and the y taps by its TW columns, so the (IEEE) divisions of the grid are evaluated
// TW + TH times per object and tile instead of once per sample.  A layout vector is cat(one_hot(class),
// appearance) (model.py:165-168): per object only ~5 of the Cp/8 eight-channel chunks are non-zero.  The tile is
// zero-filled once and only the UNION of non-zero chunks of the active objects is accumulated, one
// (pixel, chunk) item per thread, objects in the reference's order.
__global__ void __launch_bounds__(THREADS) layout_fwd_tile_kernel(LayoutArgs a, __nv_bfloat16* __restrict__ out, int tiles_w) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_raw);                   // [TH][TW][Cp]
  float* sS = reinterpret_cast<float*>(smem_raw + (size_t)TPX * a.Cp * 2);            // [MAXA][TPX]
  float* sV = sS + MAXA * TPX;                                                        // [MAXA][Cp]
  __shared__ int sAct[MAXA];
  __shared__ unsigned sNz[MAXA];
  __shared__ SgBilin sAx[MAXA * TW];
  __shared__ SgBilin sAy[MAXA * TH];
  __shared__ int sNact, sNext, sNuni;
  __shared__ int sUni[32];
  __shared__ unsigned sUmask[32];
  const int n = blockIdx.y;
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
  const int w0 = tw * TW, h0 = th * TH;
  const int nw = min(TW, a.W - w0), nh = min(TH, a.H - h0);
  const int o_begin = a.ranges[2 * n], o_end = a.ranges[2 * n + 1];
  const int chunks = a.Cp / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int scan = o_begin;        // next object to test
  bool first_pass = true;
  for (;;) {
    __syncthreads();         // the previous pass has consumed sAct / sS / sV
    if (warp == 0) {
      // ---- warp 0 compacts the next <= MAXA objects whose (dilated) box touches the tile, in object order ----
      int nact = 0, pos = scan;
      while (pos < o_end && nact < MAXA) {
        int o = pos + lane;
        bool act = (o < o_end) && box_touches_rect(a, a.boxes + 4 * o, h0, h0 + nh - 1, w0, w0 + nw - 1);
        unsigned m = __ballot_sync(0xffffffffu, act);
        int before = __popc(m & ((1u << lane) - 1));
        int room = MAXA - nact;
        if (act && before < room) sAct[nact + before] = o;
        int total = __popc(m);
        if (total > room) {
          int last = -1, cnt = 0;          // stop right after the last object that fitted
          for (int b = 0; b < 32; ++b)
            if (m & (1u << b)) { if (cnt < room) last = b; ++cnt; }
          nact = MAXA;
          pos += last + 1;
          break;
        }
        nact += total;
        pos += 32;
      }
      if (lane == 0) { sNact = nact; sNext = min(pos, o_end); }
    } else {
      if (first_pass) {      // ---- meanwhile the other warps clear the tile ----
        uint4* t4 = reinterpret_cast<uint4*>(tile);
        const int n16 = TPX * a.Cp / 8;
        for (int i = threadIdx.x - 32; i < n16; i += THREADS - 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (threadIdx.x - 32 < MAXA) sNz[threadIdx.x - 32] = 0u;
    }
    __syncthreads();
    const int nact = sNact;
    scan = sNext;
    if (nact == 0) break;
    // ---- grid taps per (object, tile column) and (object, tile row); the objects' vectors -------------------
    for (int i = threadIdx.x; i < nact * (TW + TH); i += THREADS) {
      const int k = i / (TW + TH), j = i - k * (TW + TH);
      const float* bx = a.boxes + 4 * sAct[k];
      if (j < TW) sAx[k * TW + j] = grid_axis(a, min(w0 + j, a.W - 1), a.W, bx[0], bx[2]);
      else sAy[k * TH + (j - TW)] = grid_axis(a, min(h0 + j - TW, a.H - 1), a.H, bx[1], bx[3]);
    }
    for (int i = threadIdx.x; i < nact * a.Cp; i += THREADS) {
      int k = i / a.Cp, c = i - k * a.Cp;
      float v = (c < a.D) ? a.vecs[(long)sAct[k] * a.D + c] : 0.f;
      sV[i] = v;
      if (v != 0.f) atomicOr(&sNz[k], 1u << (c >> 3));     // which 8-channel chunks of this object are non-zero
    }
    __syncthreads();
    // ---- warp 0: union of non-zero chunks (+ per chunk the objects that have it); all: sample the masks -----
    if (warp == 0) {
      unsigned m = 0u;
      if (lane < chunks)
        for (int k = 0; k < nact; ++k) m |= ((sNz[k] >> lane) & 1u) << k;
      unsigned b = __ballot_sync(0xffffffffu, m != 0u);
      if (m != 0u) {
        int pos = __popc(b & ((1u << lane) - 1));
        sUni[pos] = lane;
        sUmask[pos] = m;
      }
      if (lane == 0) sNuni = __popc(b);
    }
    for (int i = threadIdx.x; i < nact * TPX; i += THREADS) {
      const int k = i / TPX, px = i - k * TPX;
      const int r = px / TW, c = px - r * TW;
      float s = 0.f;
      if (r < nh && c < nw) {
        const SgBilin ax = sAx[k * TW + c], ay = sAy[k * TH + r];
        const long base = (long)sAct[k] * a.M * a.M;
        if (ay.ok0 && ax.ok0) s = __fmaf_rn(__fmul_rn(ax.w0, ay.w0), load_mask(a.masks, a.mask_dtype, base + (long)ay.i0 * a.M + ax.i0), s);
        if (ay.ok0 && ax.ok1) s = __fmaf_rn(__fmul_rn(ax.w1, ay.w0), load_mask(a.masks, a.mask_dtype, base + (long)ay.i0 * a.M + ax.i0 + 1), s);
        if (ay.ok1 && ax.ok0) s = __fmaf_rn(__fmul_rn(ax.w0, ay.w1), load_mask(a.masks, a.mask_dtype, base + (long)(ay.i0 + 1) * a.M + ax.i0), s);
        if (ay.ok1 && ax.ok1) s = __fmaf_rn(__fmul_rn(ax.w1, ay.w1), load_mask(a.masks, a.mask_dtype, base + (long)(ay.i0 + 1) * a.M + ax.i0 + 1), s);
      }
      sS[i] = s;
    }
    __syncthreads();
    // ---- one (pixel, non-zero chunk) item per thread: lanes = consecutive pixels, the chunk is warp-uniform ----
    const int nuni = sNuni;
    for (int i = threadIdx.x; i < nuni * TPX; i += THREADS) {
      const int ui = i / TPX, px = i - ui * TPX;
      const int chunk = sUni[ui];
      __nv_bfloat16* dst = tile + (size_t)px * a.Cp + chunk * 8;
      float acc[8];
      if (first_pass) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      } else {   // > MAXA objects overlap this tile: continue from the partial sum
        uint4 raw = *reinterpret_cast<const uint4*>(dst);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h2[j]); acc[2 * j] = f.x; acc[2 * j + 1] = f.y; }
      }
      for (unsigned m = sUmask[ui]; m != 0u; m &= m - 1u) {
        const int k = __ffs(m) - 1;
        const float s = sS[k * TPX + px];
        if (s == 0.f) continue;
        const float4* v = reinterpret_cast<const float4*>(sV + k * a.Cp + chunk * 8);
        const float4 v0 = v[0], v1 = v[1];
        acc[0] = __fmaf_rn(v0.x, s, acc[0]); acc[1] = __fmaf_rn(v0.y, s, acc[1]);
        acc[2] = __fmaf_rn(v0.z, s, acc[2]); acc[3] = __fmaf_rn(v0.w, s, acc[3]);
        acc[4] = __fmaf_rn(v1.x, s, acc[4]); acc[5] = __fmaf_rn(v1.y, s, acc[5]);
        acc[6] = __fmaf_rn(v1.z, s, acc[6]); acc[7] = __fmaf_rn(v1.w, s, acc[7]);
      }
      __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(pk);
    }
    first_pass = false;
    if (scan >= o_end) break;
  }
  // ---- the finished tile leaves as TH bulk stores, one per row segment ---------------------------------------
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the async proxy
  __syncthreads();
  if ((int)threadIdx.x < nh) {
    const int r = threadIdx.x;
    const unsigned long long gdst =
        reinterpret_cast<unsigned long long>(out + (((size_t)n * a.H + h0 + r) * a.W + w0) * a.Cp);
    const unsigned ssrc = (unsigned)__cvta_generic_to_shared(tile + (size_t)r * TW * a.Cp);
    const unsigned bytes = (unsigned)nw * a.Cp * 2;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must stay valid until it has been read
  }
}

// ------------------------------------------------------------------------------------------------
// backward (train branch): dvecs[o,c] = sum_p S_o(p) g[n,c,p];  dmasks through the bilinear taps.
// No atomics: every output element is produced by one thread after reductions in a fixed order.
// ------------------------------------------------------------------------------------------------
// pixel rectangle outside of which S_o is exactly zero (the dilated box of box_touches_rect); the whole image for
// boxes the estimate cannot handle (non-finite corners), nothing for zero extents (grid_sample yields 0 everywhere)
__device__ __forceinline__ void object_rect(const LayoutArgs& a, const float* bx, int& h_lo, int& h_hi, int& w_lo, int& w_hi) {
  h_lo = 0; h_hi = a.H - 1; w_lo = 0; w_hi = a.W - 1;
  float x0 = bx[0], y0 = bx[1], x1 = bx[2], y1 = bx[3];
  float ww = x1 - x0, hh = y1 - y0;
  if (!(isfinite(ww) && isfinite(hh) && isfinite(x0) && isfinite(y0))) return;
  if (ww == 0.f || hh == 0.f) { h_hi = -1; w_hi = -1; return; }
  float mx = fabsf(ww) / (float)max(a.M - 1, 1), my = fabsf(hh) / (float)max(a.M - 1, 1);
  float xa = fminf(x0, x1) - mx, xb = fmaxf(x0, x1) + mx, ya = fminf(y0, y1) - my, yb = fmaxf(y0, y1) + my;
  float sx = (float)max(a.W - 1, 1), sy = (float)max(a.H - 1, 1);
  float wl = floorf(xa * sx - 1.f), wh = ceilf(xb * sx + 1.f), hl = floorf(ya * sy - 1.f), hu = ceilf(yb * sy + 1.f);
  w_lo = (int)fminf(fmaxf(wl, 0.f), (float)a.W);
  w_hi = (int)fmaxf(fminf(wh, (float)(a.W - 1)), -1.f);
  h_lo = (int)fminf(fmaxf(hl, 0.f), (float)a.H);
  h_hi = (int)fmaxf(fminf(hu, (float)(a.H - 1)), -1.f);
}

constexpr int BWD_CW = 32;    // channels per CTA of the dvecs kernel (64 B of an NHWC bf16 pixel)
constexpr int BWD_BANDS = 8;  // horizontal bands of an image, one CTA each (partial sums added in band order)

// dvecs.  grid = (channel slabs of BWD_CW inside [c_begin, c_end), images, bands).  The CTA walks the objects of its
// image in order; for each, its 256 threads stride over the pixels of the object's rectangle inside the band (rows,
// then columns), each keeping BWD_CW partial sums; these are combined by a shuffle tree inside each warp and then
// over the 8 warps in warp order, and stored to part[band][o][c].  sg_sum_parts adds the bands in band order.
template <bool NHWC_BF16>
__global__ void __launch_bounds__(THREADS) layout_bwd_vecs_kernel(LayoutArgs a, const void* grad, int c_begin, int bands,
                                                                  float* part) {
  __shared__ float red[THREADS / 32][BWD_CW];
  const int n = blockIdx.y;
  const int c0 = c_begin + blockIdx.x * BWD_CW;
  const int HW = a.H * a.W;
  const int band_h = (a.H + bands - 1) / bands;
  const int b_lo = blockIdx.z * band_h, b_hi = min(a.H, b_lo + band_h) - 1;
  const int o_begin = a.ranges[2 * n], o_end = a.ranges[2 * n + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dst = part + (long)blockIdx.z * a.O * a.D;
  for (int o = o_begin; o < o_end; ++o) {
    const float* bx = a.boxes + 4 * o;
    int h_lo, h_hi, w_lo, w_hi;
    object_rect(a, bx, h_lo, h_hi, w_lo, w_hi);
    h_lo = max(h_lo, b_lo);
    h_hi = min(h_hi, b_hi);
    const int rw = w_hi - w_lo + 1, rh = h_hi - h_lo + 1;
    float acc[BWD_CW];
#pragma unroll
    for (int k = 0; k < BWD_CW; ++k) acc[k] = 0.f;
    if (rw > 0 && rh > 0) {
      for (int i = threadIdx.x; i < rw * rh; i += THREADS) {
        const int h = h_lo + i / rw, w = w_lo + i % rw;
        const float sv = sample_mask(a, o, h, w, bx);
        if (sv == 0.f) continue;
        if (NHWC_BF16) {
          const __nv_bfloat16* g = (const __nv_bfloat16*)grad + ((long)n * HW + (long)h * a.W + w) * a.Cp + c0;
#pragma unroll
          for (int q = 0; q < BWD_CW / 8; ++q) {
            if (c0 + 8 * q >= a.Cp) break;
            const uint4 r0 = *reinterpret_cast<const uint4*>(g + 8 * q);
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f0 = __bfloat1622float2(h0[j]);
              acc[8 * q + 2 * j] = __fmaf_rn(sv, f0.x, acc[8 * q + 2 * j]);
              acc[8 * q + 2 * j + 1] = __fmaf_rn(sv, f0.y, acc[8 * q + 2 * j + 1]);
            }
          }
        } else {
          const float* g = (const float*)grad + (long)n * a.D * HW + (long)h * a.W + w;
#pragma unroll
          for (int k = 0; k < BWD_CW; ++k)
            if (c0 + k < a.D) acc[k] = __fmaf_rn(sv, g[(long)(c0 + k) * HW], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < BWD_CW; ++k) {
      float v = acc[k];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[k] = v;
    }
    __syncthreads();                                     // the previous object's readers are done
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < BWD_CW; ++k) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < BWD_CW && c0 + threadIdx.x < a.D) {
      float v = 0.f;
      for (int wq = 0; wq < THREADS / 32; ++wq) v += red[wq][threadIdx.x];
      dst[(long)o * a.D + c0 + threadIdx.x] = v;
    }
  }
}

// dmasks[o] = Ay^T * dS * Ax with dS[h,w] = sum_c vecs[o,c] g[n,c,h,w] and Ay / Ax the (at most two taps per pixel)
// bilinear row / column operators.  One CTA per object; pixel rows are visited in order: the 256 threads form
// dS[h, .] in shared memory, then thread mx < M owns column mx of the result: it adds the row's pixels whose x taps
// hit mx in ascending w and deposits the sum through the row's (at most two) y taps.
template <bool NHWC_BF16>
__global__ void __launch_bounds__(THREADS) layout_bwd_masks_kernel(LayoutArgs a, const void* grad, float* dmasks) {
  extern __shared__ float smem[];
  float* sV = smem;                          // [D]
  float* sDs = sV + a.D;                     // [W]
  float* sDm = sDs + a.W;                    // [M][M]
  SgBilin* sAx = reinterpret_cast<SgBilin*>(sDm + a.M * a.M);   // [W]
  __shared__ int sImg;
  const int o = blockIdx.x;
  if (threadIdx.x == 0) {
    int img = -1;
    for (int n = 0; n < a.N; ++n)
      if (o >= a.ranges[2 * n] && o < a.ranges[2 * n + 1]) { img = n; break; }
    sImg = img;
  }
  for (int c = threadIdx.x; c < a.D; c += THREADS) sV[c] = a.vecs[(long)o * a.D + c];
  for (int i = threadIdx.x; i < a.M * a.M; i += THREADS) sDm[i] = 0.f;
  const float* bx = a.boxes + 4 * o;
  for (int w = threadIdx.x; w < a.W; w += THREADS) sAx[w] = grid_axis(a, w, a.W, bx[0], bx[2]);
  __syncthreads();
  const int n = sImg;
  const int HW = a.H * a.W;
  int h_lo, h_hi, w_lo, w_hi;
  object_rect(a, bx, h_lo, h_hi, w_lo, w_hi);
  if (n >= 0 && w_hi >= w_lo) {
    for (int h = h_lo; h <= h_hi; ++h) {
      const SgBilin ay = grid_axis(a, h, a.H, bx[1], bx[3]);
      if (!ay.ok0 && !ay.ok1) continue;                 // block-uniform
      for (int w = w_lo + threadIdx.x; w <= w_hi; w += THREADS) {
        float ds = 0.f;
        if (NHWC_BF16) {
          const __nv_bfloat16* g = (const __nv_bfloat16*)grad + ((long)n * HW + (long)h * a.W + w) * a.Cp;
          for (int c = 0; c < a.D; ++c) ds = __fmaf_rn(sV[c], __bfloat162float(g[c]), ds);
        } else {
          const float* g = (const float*)grad + (long)n * a.D * HW + (long)h * a.W + w;
          for (int c = 0; c < a.D; ++c) ds = __fmaf_rn(sV[c], g[(long)c * HW], ds);
        }
        sDs[w] = ds;
      }
      __syncthreads();
      if ((int)threadIdx.x < a.M) {
        const int mx = threadIdx.x;
        float t = 0.f;
        for (int w = w_lo; w <= w_hi; ++w) {
          const SgBilin ax = sAx[w];
          if (ax.ok0 && ax.i0 == mx) t = __fmaf_rn(ax.w0, sDs[w], t);
          else if (ax.ok1 && ax.i0 + 1 == mx) t = __fmaf_rn(ax.w1, sDs[w], t);
        }
        if (ay.ok0) sDm[ay.i0 * a.M + mx] = __fmaf_rn(ay.w0, t, sDm[ay.i0 * a.M + mx]);
        if (ay.ok1) sDm[(ay.i0 + 1) * a.M + mx] = __fmaf_rn(ay.w1, t, sDm[(ay.i0 + 1) * a.M + mx]);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < a.M * a.M; i += THREADS) dmasks[(long)o * a.M * a.M + i] = sDm[i];
}

// ------------------------------------------------------------------------------------------------
// test-mode compositing (layout.py:157-169)
//   mass[o] = sum_{c,h,w} vecs[o,c] S_o(h,w);  objects of an image painted in ascending mass order,
//   a pixel is owned by the first object (in that order) whose clean mask sample S_o > 0.5.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) layout_mass_kernel(LayoutArgs a, float* mass) {
  __shared__ float red[THREADS / 32];
  const int o = blockIdx.x;
  float vs = 0.f;
  for (int c = threadIdx.x; c < a.D; c += THREADS) vs += a.vecs[(long)o * a.D + c];
  float ss = 0.f;
  for (int p = threadIdx.x; p < a.H * a.W; p += THREADS) ss += sample_mask(a, o, p / a.W, p % a.W, a.boxes + 4 * o);
  // block reduce both
  for (int pass = 0; pass < 2; ++pass) {
    float v = pass == 0 ? vs : ss;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < THREADS / 32; ++i) t += red[i];
    if (pass == 0) vs = t; else ss = t;
  }
  if (threadIdx.x == 0) mass[o] = vs * ss;
}

template <bool NHWC_BF16>
__global__ void __launch_bounds__(THREADS) layout_test_kernel(LayoutArgs a, const float* mass, void* out) {
  extern __shared__ float smem[];
  int* sOwner = (int*)smem;               // [TP]
  float* sOwnS = smem + TP;               // [TP]
  __shared__ int sOrder[1024];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * TP;
  const int HW = a.H * a.W;
  const int o_begin = a.ranges[2 * n], o_end = a.ranges[2 * n + 1];
  const int nobj = min(o_end - o_begin, 1024);
  // rank by (mass, index): stable ascending order
  for (int i = threadIdx.x; i < nobj; i += THREADS) {
    float mi = mass[o_begin + i];
    int rank = 0;
    for (int j = 0; j < nobj; ++j) {
      float mj = mass[o_begin + j];
      rank += (mj < mi) || (mj == mi && j < i);
    }
    sOrder[rank] = i;
  }
  __syncthreads();
  for (int px = threadIdx.x; px < TP; px += THREADS) {
    int p = p0 + px;
    int owner = -1;
    float os = 0.f;
    if (p < HW) {
      for (int r = 0; r < nobj; ++r) {
        int o = o_begin + sOrder[r];
        float s = sample_mask(a, o, p / a.W, p % a.W, a.boxes + 4 * o);
        if (s > 0.5f) { owner = o; os = s; break; }
      }
    }
    sOwner[px] = owner;
    sOwnS[px] = os;
  }
  __syncthreads();
  if (NHWC_BF16) {
    __nv_bfloat16* o16 = (__nv_bfloat16*)out;
    for (int idx = threadIdx.x; idx < TP * a.Cp; idx += THREADS) {
      int px = idx / a.Cp, c = idx % a.Cp;
      int p = p0 + px;
      if (p >= HW) continue;
      int o = sOwner[px];
      float v = (o >= 0 && c < a.D) ? a.vecs[(long)o * a.D + c] * sOwnS[px] : 0.f;
      o16[((long)n * HW + p) * a.Cp + c] = __float2bfloat16(v);
    }
  } else {
    float* o32 = (float*)out;
    for (int idx = threadIdx.x; idx < TP * a.D; idx += THREADS) {
      int c = idx / TP, px = idx % TP;
      int p = p0 + px;
      if (p >= HW) continue;
      int o = sOwner[px];
      o32[((long)n * a.D + c) * HW + p] = (o >= 0) ? a.vecs[(long)o * a.D + c] * sOwnS[px] : 0.f;
    }
  }
}

int check_args(const LayoutArgs& a, int out_format) {
  SG_CHECK_ARG(a.O >= 0 && a.D > 0 && a.M > 0 && a.N >= 0 && a.H > 0 && a.W > 0, "masks_to_layout: bad sizes");
  SG_CHECK_ARG(a.mask_dtype >= 0 && a.mask_dtype <= 2, "masks_to_layout: mask_dtype must be 0 (f32), 1 (i64) or 2 (u8)");
  SG_CHECK_ARG(out_format == 0 || out_format == 1, "masks_to_layout: out_format must be 0 (NCHW f32) or 1 (NHWC bf16)");
  SG_CHECK_ARG(a.Cp >= a.D && a.Cp % 8 == 0 && a.Cp <= 256, "masks_to_layout: Cp must be a multiple of 8 in [D, 256] (got D=%d Cp=%d)", a.D, a.Cp);
  return SG_OK;
}

}  // namespace

extern "C" int sg_masks_to_layout_fwd(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                                      const int* img_ranges, int O, int D, int M, int N, int H, int W,
                                      int align_corners, int out_format, int Cp, void* out, cudaStream_t stream) {
  LayoutArgs a{vecs, boxes, masks, img_ranges, mask_dtype, O, D, M, N, H, W, Cp, align_corners};
  if (int e = check_args(a, out_format)) return e;
  if (N == 0) return SG_OK;
  dim3 grid(sg_cdiv((long)H * W, TP), N);
  size_t smem = sizeof(float) * (MAXO * TP + MAXO * Cp);
  if (out_format == 1) {
    SG_CHECK_ARG(((uintptr_t)out & 15) == 0, "masks_to_layout: NHWC output must be 16-byte aligned");
    size_t smem2 = (size_t)TPX * Cp * 2 + sizeof(float) * (MAXA * TPX + MAXA * Cp);
    static size_t smem_set = 0;
    if (smem2 > smem_set) {
      cudaFuncSetAttribute(layout_fwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      smem_set = smem2;
    }
    const int tiles_w = sg_cdiv(W, TW), tiles_h = sg_cdiv(H, TH);
    dim3 grid2(tiles_w * tiles_h, N);
    layout_fwd_tile_kernel<<<grid2, THREADS, smem2, stream>>>(a, (__nv_bfloat16*)out, tiles_w);
  } else {
    cudaFuncSetAttribute(layout_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    layout_fwd_kernel<false><<<grid, THREADS, smem, stream>>>(a, out);
  }
  SG_CHECK_LAUNCH("sg_masks_to_layout_fwd");
  return SG_OK;
}

extern "C" int sg_masks_to_layout_bwd(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                                      const int* img_ranges, int O, int D, int M, int N, int H, int W,
                                      int align_corners, int grad_format, int Cp, const void* grad_out, int c_begin, int c_end,
                                      float* dvecs, float* dmasks, float* ws, long long ws_floats, cudaStream_t stream) {
  LayoutArgs a{vecs, boxes, masks, img_ranges, mask_dtype, O, D, M, N, H, W, Cp, align_corners};
  if (int e = check_args(a, grad_format)) return e;
  SG_CHECK_ARG(dvecs != nullptr, "masks_to_layout_bwd: dvecs is null");
  SG_CHECK_ARG(!dmasks || M <= THREADS, "masks_to_layout_bwd: mask size must be <= %d for the mask gradient", THREADS);
  if (O == 0) return SG_OK;
  SG_CHECK_ARG(c_begin >= 0 && c_begin % 8 == 0 && c_begin < c_end && c_end <= D,
               "masks_to_layout_bwd: channel range [%d, %d) must start at a multiple of 8 inside [0, %d]", c_begin, c_end, D);
  const long per = (long)O * D;
  // image height bands: more CTAs for the few, large images of a batch (each band leaves one partial row set)
  int bands = H >= 64 ? BWD_BANDS : 1;
  if (ws == nullptr || ws_floats < (long long)bands * per) bands = 1;
  float* part = bands > 1 ? ws : dvecs;
  // channels outside [c_begin, c_end) and objects outside every image range (none in a well-formed batch): zero
  cudaMemsetAsync(part, 0, sizeof(float) * (size_t)bands * per, stream);
  if (N > 0) {
    dim3 grid(sg_cdiv(c_end - c_begin, BWD_CW), N, bands);
    if (grad_format == 1) layout_bwd_vecs_kernel<true><<<grid, THREADS, 0, stream>>>(a, grad_out, c_begin, bands, part);
    else layout_bwd_vecs_kernel<false><<<grid, THREADS, 0, stream>>>(a, grad_out, c_begin, bands, part);
    SG_CHECK_LAUNCH("sg_masks_to_layout_bwd");
  }
  if (bands > 1)
    if (int e = sg_sum_parts(part, per, bands, per, dvecs, stream, "sg_masks_to_layout_bwd(sum bands)")) return e;
  if (dmasks) {
    size_t smem = sizeof(float) * ((size_t)D + W + (size_t)M * M) + sizeof(SgBilin) * (size_t)W + 16;
    if (grad_format == 1) {
      cudaFuncSetAttribute(layout_bwd_masks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      layout_bwd_masks_kernel<true><<<O, THREADS, smem, stream>>>(a, grad_out, dmasks);
    } else {
      cudaFuncSetAttribute(layout_bwd_masks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      layout_bwd_masks_kernel<false><<<O, THREADS, smem, stream>>>(a, grad_out, dmasks);
    }
    SG_CHECK_LAUNCH("sg_masks_to_layout_bwd(masks)");
  }
  return SG_OK;
}

extern "C" int sg_masks_to_layout_test(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                                       const int* img_ranges, int O, int D, int M, int N, int H, int W,
                                       int align_corners, int out_format, int Cp, float* mass_ws, void* out,
                                       cudaStream_t stream) {
  LayoutArgs a{vecs, boxes, masks, img_ranges, mask_dtype, O, D, M, N, H, W, Cp, align_corners};
  if (int e = check_args(a, out_format)) return e;
  SG_CHECK_ARG(mass_ws != nullptr, "masks_to_layout_test: mass workspace (O floats) is null");
  if (N == 0) return SG_OK;
  if (O > 0) {
    layout_mass_kernel<<<O, THREADS, 0, stream>>>(a, mass_ws);
    SG_CHECK_LAUNCH("sg_masks_to_layout_test(mass)");
  }
  dim3 grid(sg_cdiv((long)H * W, TP), N);
  size_t smem = sizeof(float) * 2 * TP;
  if (out_format == 1) layout_test_kernel<true><<<grid, THREADS, smem, stream>>>(a, mass_ws, out);
  else layout_test_kernel<false><<<grid, THREADS, smem, stream>>>(a, mass_ws, out);
  SG_CHECK_LAUNCH("sg_masks_to_layout_test");
  return SG_OK;
}
