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

// dw[co][tap][c] = sum over (n, j) with cmap[n][j] == c of dwc[n][co][tap][j].  One thread per (co, tap, j)
// walks the images; runs of equal targets (the appearance / image channels map to the same dense channel
// in every image) are summed in a register and flushed with one atomic.
__global__ void scatter_cmap_kernel(const float* __restrict__ dwc, const int* __restrict__ cmap, int N, int Cout, int taps,
                                    int Cc, int Cin, float* __restrict__ dw) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)Cout * taps * Cc;
  if (idx >= total) return;
  const int j = (int)(idx % Cc);
  const long ct = idx / Cc;            // co * taps + tap
  const long slab = (long)Cout * taps * Cc;
  int cur = -1;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    int c = cmap[n * Cc + j];
    if (c < 0 || c >= Cin) continue;
    float v = dwc[(long)n * slab + idx];
    if (c != cur) {
      if (cur >= 0) atomicAdd(dw + ct * Cin + cur, acc);
      cur = c;
      acc = v;
    } else {
      acc += v;
    }
  }
  if (cur >= 0) atomicAdd(dw + ct * Cin + cur, acc);
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
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * taps * Cin, stream);
  long total = (long)Cout * taps * Cc;
  scatter_cmap_kernel<<<(unsigned)sg_cdiv(total, 256), 256, 0, stream>>>(dwc, cmap, N, Cout, taps, Cc, Cin, dw);
  SG_CHECK_LAUNCH("sg_wgrad_cmap_scatter");
  return SG_OK;
}
