// Channel-compacted layouts.
//
// The scene layout the generator / image discriminator consume (model.py:165-168, trainer.py:246-250) is
// cat(one_hot(class) (x) mask, appearance (x) mask): of its ~180 class channels only the classes present in
// the image (<= max_objects_per_image + 1) are non-zero.  The hot path therefore keeps a layout as
// (N, H, W, Cc = 64) with a per-image channel map  cmap[n][j] -> dense channel (or -1)  and runs the first
// convolution of the consumer with per-image weights  Wc[n][co][tap][j] = W[co][tap][cmap[n][j]]
// (sg_conv_desc_t.w_img_rows): identical sums, minus the terms that multiply a zero.  This file holds
// the weight gather and the adjoint scatter of the per-image weight gradients.
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

typedef __nv_bfloat16 bf16;

__global__ void pack_cmap_kernel(const float* __restrict__ w, int Cout, int taps, int Cin, const int* __restrict__ cmap,
                                 long total, int Cc, bf16* __restrict__ wk) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % Cc);
  long r = idx / Cc;
  const int tap = (int)(r % taps);
  r /= taps;
  const int co = (int)(r % Cout);
  const int n = (int)(r / Cout);
  const int c = cmap[n * Cc + j];
  float v = (c >= 0 && c < Cin) ? w[((long)co * taps + tap) * Cin + c] : 0.f;
  wk[idx] = __float2bfloat16(v);
}

// transposed operand of the input-gradient GEMM: wt[n][j][tap][co_p]
__global__ void pack_cmap_t_kernel(const float* __restrict__ w, int Cout, int taps, int Cin, const int* __restrict__ cmap,
                                   long total, int Cc, int Cout_p, bf16* __restrict__ wt) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % Cout_p);
  long r = idx / Cout_p;
  const int tap = (int)(r % taps);
  r /= taps;
  const int j = (int)(r % Cc);
  const int n = (int)(r / Cc);
  const int c = cmap[n * Cc + j];
  float v = (co < Cout && c >= 0 && c < Cin) ? w[((long)co * taps + tap) * Cin + c] : 0.f;
  wt[idx] = __float2bfloat16(v);
}

// dw[co][tap][c] = sum over (n, j) with cmap[n][j] == c of dwc[n][co][tap][j] — as a GATHER in image order (no
// atomics, no zero-fill): the CTA first inverts the channel map into shared memory (inv[n][c] = j or -1; a dense
// channel occurs at most once per image), then one thread per (co*taps + tap, c) adds its images in ascending n.
__global__ void scatter_cmap_kernel(const float* __restrict__ dwc, const int* __restrict__ cmap, int N, int Cout, int taps,
                                    int Cc, int Cin, int rows_per_block, float* __restrict__ dw) {
  extern __shared__ signed char inv[];       // [N][Cin]
  for (int i = threadIdx.x; i < N * Cin; i += blockDim.x) inv[i] = -1;
  __syncthreads();
  for (int i = threadIdx.x; i < N * Cc; i += blockDim.x) {
    const int c = cmap[i];
    if (c >= 0 && c < Cin) inv[(i / Cc) * Cin + c] = (signed char)(i % Cc);
  }
  __syncthreads();
  const long slab = (long)Cout * taps * Cc;
  const long ct0 = (long)blockIdx.x * rows_per_block;
  for (long i = threadIdx.x; i < (long)rows_per_block * Cin; i += blockDim.x) {
    const long ct = ct0 + i / Cin;
    const int c = (int)(i % Cin);
    if (ct >= (long)Cout * taps) break;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
      const int j = inv[n * Cin + c];
      if (j >= 0) acc += dwc[(long)n * slab + ct * Cc + j];
    }
    dw[ct * Cin + c] = acc;
  }
}

}  // namespace

extern "C" int sg_pack_weight_cmap(const float* w, int Cout, int taps, int Cin, const int* cmap, int N, int Cc, int Cout_p,
                                   void* wk, void* wt, sg_stream_t stream) {
  SG_CHECK_ARG(w && cmap && (wk || wt), "pack_weight_cmap: null pointer");
  SG_CHECK_ARG(Cout > 0 && taps > 0 && Cin > 0 && N > 0 && Cc > 0 && Cc % 8 == 0, "pack_weight_cmap: bad sizes");
  SG_CHECK_ARG(wt == nullptr || (Cout_p >= Cout && Cout_p % 8 == 0), "pack_weight_cmap: bad Cout_p");
  if (wk) {
    long total = (long)N * Cout * taps * Cc;
    pack_cmap_kernel<<<(unsigned)sg_cdiv(total, 256), 256, 0, stream>>>(w, Cout, taps, Cin, cmap, total, Cc, (bf16*)wk);
    SG_CHECK_LAUNCH("sg_pack_weight_cmap");
  }
  if (wt) {
    long total = (long)N * Cc * taps * Cout_p;
    pack_cmap_t_kernel<<<(unsigned)sg_cdiv(total, 256), 256, 0, stream>>>(w, Cout, taps, Cin, cmap, total, Cc, Cout_p, (bf16*)wt);
    SG_CHECK_LAUNCH("sg_pack_weight_cmap(transposed)");
  }
  return SG_OK;
}

extern "C" int sg_wgrad_cmap_scatter(const float* dwc, const int* cmap, int N, int Cout, int taps, int Cc, int Cin, float* dw,
                                     sg_stream_t stream) {
  SG_CHECK_ARG(dwc && cmap && dw, "wgrad_cmap_scatter: null pointer");
  SG_CHECK_ARG(Cout > 0 && taps > 0 && Cin > 0 && N > 0 && Cc > 0, "wgrad_cmap_scatter: bad sizes");
  SG_CHECK_ARG(Cc <= 127 && (size_t)N * Cin <= 160 * 1024, "wgrad_cmap_scatter: channel map too large (N * Cin = %ld, Cc = %d)",
               (long)N * Cin, Cc);
  const long rows = (long)Cout * taps;
  const int rpb = (int)((rows + 591) / 592) > 0 ? (int)((rows + 591) / 592) : 1;      // ~4 CTAs per SM
  const size_t smem = (size_t)N * Cin;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaFuncSetAttribute(scatter_cmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  scatter_cmap_kernel<<<(unsigned)sg_cdiv(rows, rpb), 256, smem, stream>>>(dwc, cmap, N, Cout, taps, Cc, Cin, rpb, dw);
  SG_CHECK_LAUNCH("sg_wgrad_cmap_scatter");
  return SG_OK;
}
