// Direct (CUDA-core) adjoints for convolutions with a tiny output-channel count (the generator's last
// 7x7 conv, 64 -> 3, generators.py:87).  On the tensor-core path their GEMMs have N (dgrad: K) = 3 and waste
// > 90 % of every MMA; as direct convolutions they are FMA / bandwidth bound (SURVEY.md §7 "last conv").
//
//   dgrad :  dx[n, u, v, ci] = sum_{kh,kw} sum_{co} dz[n, u-kh, v-kw, co] * w[co, kh, kw, ci]     (u < H+k-1)
//            i.e. the gradient w.r.t. the PRE-PADDED operand of a stride-1 "valid" convolution.
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

constexpr int TILE = 16;          // output pixels per block edge
constexpr int MAXK = 7;
constexpr int MAXCO = 4;

template <int CIN>
__global__ void __launch_bounds__(TILE * TILE)
dgrad_small_cout_kernel(const __nv_bfloat16* __restrict__ dz, int dzC, const float* __restrict__ w, int Cout, int k, int N,
                        int H, int W, __nv_bfloat16* __restrict__ dx) {
  extern __shared__ float smem[];
  const int taps = k * k;
  float* sW = smem;                                   // [Cout][taps][CIN]
  float* sZ = smem + Cout * taps * CIN;               // [Cout][TILE+k-1][TILE+k-1]
  const int P = TILE + k - 1;
  const int Hp = H + k - 1, Wp = W + k - 1;
  const int n = blockIdx.z, u0 = blockIdx.y * TILE, v0 = blockIdx.x * TILE;
  for (int i = threadIdx.x; i < Cout * taps * CIN; i += blockDim.x) sW[i] = w[i];
  for (int i = threadIdx.x; i < Cout * P * P; i += blockDim.x) {
    int co = i / (P * P), r = (i / P) % P, c = i % P;
    int h = u0 - (k - 1) + r, x = v0 - (k - 1) + c;       // dz row/col feeding output (u0.., v0..)
    float v = 0.f;
    if (h >= 0 && h < H && x >= 0 && x < W) v = __bfloat162float(dz[(((long)n * H + h) * W + x) * dzC + co]);
    sZ[i] = v;
  }
  __syncthreads();
  const int tu = threadIdx.x / TILE, tv = threadIdx.x % TILE;
  const int u = u0 + tu, v = v0 + tv;
  float acc[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw)
      for (int co = 0; co < Cout; ++co) {
        // dz[u-kh, v-kw] sits at sZ[co][tu + (k-1) - kh][tv + (k-1) - kw]
        float z = sZ[(co * P + tu + (k - 1) - kh) * P + tv + (k - 1) - kw];
        if (z == 0.f) continue;
        const float4* wr = reinterpret_cast<const float4*>(sW + (co * taps + kh * k + kw) * CIN);
#pragma unroll
        for (int c4 = 0; c4 < CIN / 4; ++c4) {
          float4 q = wr[c4];
          acc[4 * c4] = fmaf(z, q.x, acc[4 * c4]);
          acc[4 * c4 + 1] = fmaf(z, q.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(z, q.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(z, q.w, acc[4 * c4 + 3]);
        }
      }
  if (u < Hp && v < Wp) {
    __nv_bfloat16* dst = dx + (((long)n * Hp + u) * Wp + v) * CIN;
#pragma unroll
    for (int c8 = 0; c8 < CIN / 8; ++c8) {
      __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[j] = __floats2bfloat162_rn(acc[8 * c8 + 2 * j], acc[8 * c8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(dst + 8 * c8) = *reinterpret_cast<uint4*>(pk);
    }
  }
}

// wgrad: dw[co, kh*k+kw, ci] = sum_{n,h,w} dz[n,h,w,co] * xop[n, h+kh, w+kw, ci]   (xop pre-padded: (H+k-1, W+k-1))
// Persistent blocks walk over 8x32-pixel tiles; the bf16 input patch (tile + halo, 64 channels) is staged in
// shared memory once and reused by all k*k taps; thread = (channel, tap group) keeps its 3 x 13 partial sums in
// registers across tiles and flushes them with one atomic per output at the end.
constexpr int WT_H = 8, WT_W = 32;
template <int K>
__global__ void __launch_bounds__(256)
wgrad_small_cout_kernel(const __nv_bfloat16* __restrict__ dz, int dzC, const __nv_bfloat16* __restrict__ xop, int Cout, int N,
                        int H, int W, float* __restrict__ ws) {
  constexpr int CIN = 64, TAPS = K * K, PH = WT_H + K - 1, PW = WT_W + K - 1, TPG = (TAPS + 3) / 4;
  extern __shared__ __nv_bfloat16 sm16[];
  __nv_bfloat16* sX = sm16;                                                  // [PH][PW][CIN]
  float* sZ = reinterpret_cast<float*>(sm16 + PH * PW * CIN);                // [WT_H*WT_W][MAXCO]
  const int ci = threadIdx.x & 63, tq = threadIdx.x >> 6;
  const int Hp = H + K - 1, Wp = W + K - 1;
  const int tiles_w = (W + WT_W - 1) / WT_W, tiles_h = (H + WT_H - 1) / WT_H;
  const int n_tiles = N * tiles_h * tiles_w;
  float acc[TPG][MAXCO - 1];
#pragma unroll
  for (int t = 0; t < TPG; ++t)
#pragma unroll
    for (int c = 0; c < MAXCO - 1; ++c) acc[t][c] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * WT_H, w0 = tw * WT_W;
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * (CIN / 8); i += 256) {           // 16-byte copies of the patch
      int c8 = i % (CIN / 8), pc = (i / (CIN / 8)) % PW, pr = i / ((CIN / 8) * PW);
      int hh = h0 + pr, wc = w0 + pc;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (hh < Hp && wc < Wp) v = *reinterpret_cast<const uint4*>(xop + (((long)n * Hp + hh) * Wp + wc) * CIN + c8 * 8);
      *reinterpret_cast<uint4*>(sX + (pr * PW + pc) * CIN + c8 * 8) = v;
    }
    for (int i = threadIdx.x; i < WT_H * WT_W * (MAXCO - 1); i += 256) {
      int co = i % (MAXCO - 1), px = i / (MAXCO - 1);
      int hh = h0 + px / WT_W, wc = w0 + px % WT_W;
      float v = 0.f;
      if (co < Cout && hh < H && wc < W) v = __bfloat162float(dz[(((long)n * H + hh) * W + wc) * dzC + co]);
      sZ[px * (MAXCO - 1) + co] = v;
    }
    __syncthreads();
    for (int px = 0; px < WT_H * WT_W; ++px) {
      const float z0 = sZ[px * 3], z1 = sZ[px * 3 + 1], z2 = sZ[px * 3 + 2];
      if (z0 == 0.f && z1 == 0.f && z2 == 0.f) continue;
      const int pr = px / WT_W, pc = px % WT_W;
#pragma unroll
      for (int t = 0; t < TPG; ++t) {
        const int tap = tq + 4 * t;
        if (tap < TAPS) {
          const int kh = tap / K, kw = tap % K;
          float x = __bfloat162float(sX[((pr + kh) * PW + pc + kw) * CIN + ci]);
          acc[t][0] = fmaf(z0, x, acc[t][0]);
          acc[t][1] = fmaf(z1, x, acc[t][1]);
          acc[t][2] = fmaf(z2, x, acc[t][2]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TPG; ++t) {
    const int tap = tq + 4 * t;
    if (tap < TAPS)
      for (int co = 0; co < Cout && co < MAXCO - 1; ++co)
        ws[((long)blockIdx.x * Cout + co) * TAPS * CIN + (long)tap * CIN + ci] = acc[t][co];   // this CTA's partial sums
  }
}

// dw[e] = sum over the CTAs of wgrad_small_cout_kernel of their partial sums, in CTA order (fixed order, no atomics)
__global__ void wgrad_small_reduce_kernel(const float* __restrict__ ws, int n_blocks, int elems, float* __restrict__ dw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  float v = 0.f;
  for (int b = 0; b < n_blocks; ++b) v += ws[(long)b * elems + e];
  dw[e] = v;
}

// Tap-unrolled copy of the output gradient of a tiny-Cout convolution ("im2col along K"):
//   col[n, u, v, co * k*k + kh * k + kw] = dz[n, u - kh, v - kw, co]     (zero outside, u < H+k-1, v < W+k-1)
// With it both adjoints of the layer are ordinary tensor-core GEMMs over the PADDED pixel grid:
//   dgrad  dx[n,u,v,ci]        = sum_K col[n,u,v,K] * wcol[ci][K]                (sg_conv_tc, one tap, K = Cout*k*k)
//   wgrad  dw[co, tap, ci]     = sum_{n,u,v} col[n,u,v,co*k*k+tap] * xop[n,u,v,ci]   (sg_wgrad_tc, one tap)
// instead of GEMMs whose contraction (dgrad) or output (wgrad) dimension is 3.  thread = (pixel, 8-channel chunk).
// KS > 0: kernel size known at compile time (the tap decomposition of a column is then a few multiplies and is
// computed once per thread, not per pixel); KS = 0: any k.
template <int KS>
__global__ void __launch_bounds__(256) im2col_dz_kernel(const __nv_bfloat16* __restrict__ dz, int dzC, int Cout, int k_rt, int N, int H,
                                                        int W, int Kp, __nv_bfloat16* __restrict__ col) {
  // grid.x = padded output row (n, u); block = (8-column chunks of K, lanes over the row's pixels v)
  const int k = KS > 0 ? KS : k_rt;
  const int taps = k * k;
  const int Hp = H + k - 1, Wp = W + k - 1;
  const int n = blockIdx.x / Hp, u = blockIdx.x - n * Hp;
  const int ch = threadIdx.x;
  int kw[8];
  bool row_ok[8];
  long rbase[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int K = ch * 8 + j;
    const int c = K / taps, tap = K - c * taps;
    const int kh = tap / k;
    kw[j] = tap - kh * k;
    const int h = u - kh;
    row_ok[j] = K < Cout * taps && h >= 0 && h < H;
    rbase[j] = ((long)n * H + (row_ok[j] ? h : 0)) * W * dzC + c;
  }
  __nv_bfloat16* crow = col + ((long)blockIdx.x * Wp) * Kp + ch * 8;
  for (int v = threadIdx.y; v < Wp; v += blockDim.y) {
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = v - kw[j];
      o[j] = (row_ok[j] && x >= 0 && x < W) ? dz[rbase[j] + (long)x * dzC] : __float2bfloat16(0.f);
    }
    *reinterpret_cast<uint4*>(crow + (long)v * Kp) = *reinterpret_cast<const uint4*>(o);
  }
}

}  // namespace

extern "C" int sg_im2col_dz(const void* dz, int dzC, int Cout, int k, int N, int H, int W, int Kp, void* col, sg_stream_t stream) {
  SG_CHECK_ARG(dz && col && Cout >= 1 && dzC >= Cout && k >= 1 && N > 0 && H > 0 && W > 0, "im2col_dz: bad arguments");
  SG_CHECK_ARG(Kp % 8 == 0 && Kp >= Cout * k * k, "im2col_dz: Kp must be a multiple of 8 >= Cout*k*k");
  const int chunks = Kp / 8;
  SG_CHECK_ARG(chunks <= 256, "im2col_dz: Kp must be <= 2048");
  int lanes = 256 / chunks;
  if (lanes > W + k - 1) lanes = W + k - 1;
  const dim3 block(chunks, lanes), grid((unsigned)(N * (H + k - 1)));
  if (k == 7) im2col_dz_kernel<7><<<grid, block, 0, stream>>>((const __nv_bfloat16*)dz, dzC, Cout, k, N, H, W, Kp, (__nv_bfloat16*)col);
  else if (k == 3) im2col_dz_kernel<3><<<grid, block, 0, stream>>>((const __nv_bfloat16*)dz, dzC, Cout, k, N, H, W, Kp, (__nv_bfloat16*)col);
  else im2col_dz_kernel<0><<<grid, block, 0, stream>>>((const __nv_bfloat16*)dz, dzC, Cout, k, N, H, W, Kp, (__nv_bfloat16*)col);
  SG_CHECK_LAUNCH("sg_im2col_dz");
  return SG_OK;
}

extern "C" int sg_wgrad_small_cout(const void* dz, int dzC, const void* xop, int Cout, int k, int Cin, int N, int H, int W,
                                   float* dw, float* ws, long long ws_floats, sg_stream_t stream) {
  SG_CHECK_ARG(dz && xop && dw, "wgrad_small_cout: null pointer");
  SG_CHECK_ARG(Cout >= 1 && Cout <= 3 && dzC >= Cout && Cin == 64 && (k == 7 || k == 3), "wgrad_small_cout: needs Cout <= 3, Cin == 64, k in {3, 7}");
  SG_CHECK_ARG(N > 0 && H > 0 && W > 0, "wgrad_small_cout: empty problem");
  const int tiles = N * sg_cdiv(H, WT_H) * sg_cdiv(W, WT_W);
  const int grid = tiles < SG_WGRAD_SMALL_BLOCKS ? tiles : SG_WGRAD_SMALL_BLOCKS;
  const int elems = Cout * k * k * Cin;
  SG_CHECK_ARG(ws != nullptr && ws_floats >= (long long)grid * elems, "wgrad_small_cout: workspace of %lld floats needed",
               (long long)grid * elems);
  const int PH = WT_H + k - 1, PW = WT_W + k - 1;
  size_t smem = sizeof(__nv_bfloat16) * (size_t)PH * PW * 64 + sizeof(float) * WT_H * WT_W * 3;
  if (k == 7) {
    cudaFuncSetAttribute(wgrad_small_cout_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wgrad_small_cout_kernel<7><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)dz, dzC, (const __nv_bfloat16*)xop, Cout, N, H, W, ws);
  } else {
    cudaFuncSetAttribute(wgrad_small_cout_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wgrad_small_cout_kernel<3><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)dz, dzC, (const __nv_bfloat16*)xop, Cout, N, H, W, ws);
  }
  SG_CHECK_LAUNCH("sg_wgrad_small_cout");
  wgrad_small_reduce_kernel<<<sg_cdiv(elems, 256), 256, 0, stream>>>(ws, grid, elems, dw);
  SG_CHECK_LAUNCH("sg_wgrad_small_cout(reduce)");
  return SG_OK;
}

extern "C" int sg_dgrad_small_cout(const void* dz, int dzC, const float* w, int Cout, int k, int Cin, int N, int H, int W,
                                   void* dx, sg_stream_t stream) {
  SG_CHECK_ARG(dz && w && dx, "dgrad_small_cout: null pointer");
  SG_CHECK_ARG(Cout >= 1 && Cout <= MAXCO && k >= 1 && k <= MAXK && dzC >= Cout, "dgrad_small_cout: Cout must be <= 4, k <= 7");
  SG_CHECK_ARG(Cin == 64 || Cin == 32, "dgrad_small_cout: Cin must be 32 or 64 (got %d)", Cin);
  SG_CHECK_ARG(N > 0 && H > 0 && W > 0, "dgrad_small_cout: empty problem");
  const int P = TILE + k - 1;
  dim3 grid(sg_cdiv(W + k - 1, TILE), sg_cdiv(H + k - 1, TILE), N);
  size_t smem = sizeof(float) * ((size_t)Cout * k * k * Cin + (size_t)Cout * P * P);
  if (Cin == 64) {
    cudaFuncSetAttribute(dgrad_small_cout_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dgrad_small_cout_kernel<64><<<grid, TILE * TILE, smem, stream>>>((const __nv_bfloat16*)dz, dzC, w, Cout, k, N, H, W,
                                                                      (__nv_bfloat16*)dx);
  } else {
    cudaFuncSetAttribute(dgrad_small_cout_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dgrad_small_cout_kernel<32><<<grid, TILE * TILE, smem, stream>>>((const __nv_bfloat16*)dz, dzC, w, Cout, k, N, H, W,
                                                                      (__nv_bfloat16*)dx);
  }
  SG_CHECK_LAUNCH("sg_dgrad_small_cout");
  return SG_OK;
}
