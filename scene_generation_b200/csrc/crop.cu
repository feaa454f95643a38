// Differentiable bilinear box crop (reference: scene_generation/bilinear.py:26-130, 246-275).
//   crops[b, c, i, j] = bilinear_zero_pad(feats[map[b], c]; X_j, Y_i)
//   X_j = lin10(j) * (2*x0-1) + lin01(j) * (2*x1-1)   (tensor_linspace, bilinear.py:263-274)
// The reference replicates every image once per box (bilinear.py:80) and fixes the order up with
// an inverse permutation (bilinear.py:94-98); here each output element gathers its 4 taps directly,
// so crops come out in box order by construction.  The adjoint is a gather too: no atomics anywhere.
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

struct CropArgs {
  const float* feats;        // (N, C, H, W) fp32
  const float* boxes;        // (B, 4)
  const long long* map;      // (B,)
  int N, C, H, W, B, HH, WW, align_corners;
  int Cp;                    // NHWC bf16 physical channels
};

__device__ __forceinline__ void crop_axes(const CropArgs& a, int b, int i, int j, SgBilin& ax, SgBilin& ay) {
  const float* bx = a.boxes + 4 * b;
  float x0 = __fsub_rn(__fmul_rn(2.f, bx[0]), 1.f), y0 = __fsub_rn(__fmul_rn(2.f, bx[1]), 1.f);
  float x1 = __fsub_rn(__fmul_rn(2.f, bx[2]), 1.f), y1 = __fsub_rn(__fmul_rn(2.f, bx[3]), 1.f);
  float X = __fadd_rn(__fmul_rn(sg_linspace10(j, a.WW), x0), __fmul_rn(sg_linspace01(j, a.WW), x1));
  float Y = __fadd_rn(__fmul_rn(sg_linspace10(i, a.HH), y0), __fmul_rn(sg_linspace01(i, a.HH), y1));
  ax = sg_axis(X, a.W, a.align_corners);
  ay = sg_axis(Y, a.H, a.align_corners);
}

template <bool NHWC_BF16>
__global__ void crop_fwd_kernel(CropArgs a, void* out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)a.B * a.HH * a.WW;
  if (idx >= total) return;
  int j = idx % a.WW, i = (idx / a.WW) % a.HH, b = idx / ((long)a.WW * a.HH);
  SgBilin ax, ay;
  crop_axes(a, b, i, j, ax, ay);
  long n = a.map[b];
  float wnw = ax.w0 * ay.w0, wne = ax.w1 * ay.w0, wsw = ax.w0 * ay.w1, wse = ax.w1 * ay.w1;
  for (int c = 0; c < (NHWC_BF16 ? a.Cp : a.C); ++c) {
    float v = 0.f;
    if (c < a.C) {
      const float* f = a.feats + ((long)n * a.C + c) * a.H * a.W;
      if (ay.ok0 && ax.ok0) v = __fmaf_rn(wnw, f[ay.i0 * a.W + ax.i0], v);
      if (ay.ok0 && ax.ok1) v = __fmaf_rn(wne, f[ay.i0 * a.W + ax.i0 + 1], v);
      if (ay.ok1 && ax.ok0) v = __fmaf_rn(wsw, f[(ay.i0 + 1) * a.W + ax.i0], v);
      if (ay.ok1 && ax.ok1) v = __fmaf_rn(wse, f[(ay.i0 + 1) * a.W + ax.i0 + 1], v);
    }
    if (NHWC_BF16) ((__nv_bfloat16*)out)[idx * a.Cp + c] = __float2bfloat16(v);
    else ((float*)out)[(((long)b * a.C + c) * a.HH + i) * a.WW + j] = v;
  }
}

// Adjoint as a GATHER (no atomics, fixed summation order): thread = one pixel (y, x) of image n; the boxes of that
// image are visited in box order, their crop rows i and columns j in ascending order, and crop pixel (i, j) adds
// wy * wx * g to the thread's accumulators when one of its four bilinear taps is (y, x).  The axis taps of a box
// (HH + WW entries, the forward's own sg_axis values with the validity folded into the weights) are tabulated in
// shared memory once per CTA and box.  The sample coordinate is linear in the crop index, so the crop rows / columns
// that can touch pixel row y / column x form a short interval: it is estimated from the two end taps of the table
// (with a margin) and every candidate is then checked against the exact table entry; degenerate boxes scan everything.
constexpr int CROP_BWD_THREADS = 256;
constexpr int CROP_MAX_C = 4;

struct CropTap { int i0; float w0, w1; };       // weight 0 where the forward's tap was out of range

// candidate interval [lo, hi) of crop indices whose tap rows can be `y`, from the table's end points (f ~ linear)
__device__ __forceinline__ void crop_candidates(float f_first, float f_last, int n, int y, int& lo, int& hi) {
  lo = 0; hi = n;
  const float span = f_last - f_first;
  if (n < 2 || !isfinite(span) || !(fabsf(span) > 1e-3f)) return;
  const float step = span / (float)(n - 1);
  float a = ((float)(y - 1) - f_first) / step, b = ((float)(y + 1) - f_first) / step;
  if (a > b) { const float t = a; a = b; b = t; }
  if (!isfinite(a) || !isfinite(b)) return;
  lo = max(0, (int)floorf(fminf(a, (float)n)) - 2);
  hi = min(n, (int)ceilf(fmaxf(b, -1.f)) + 3);
  if (hi < lo) hi = lo;
}

template <bool NHWC_BF16>
__global__ void __launch_bounds__(CROP_BWD_THREADS) crop_bwd_kernel(CropArgs a, const void* grad, float* dfeats) {
  extern __shared__ unsigned char crop_smem[];
  CropTap* sAx = reinterpret_cast<CropTap*>(crop_smem);      // [WW]
  CropTap* sAy = sAx + a.WW;                                  // [HH]
  float* sF = reinterpret_cast<float*>(sAy + a.HH);           // [4]: un-normalised coordinate of the first / last tap per axis
  unsigned char* sMine = reinterpret_cast<unsigned char*>(sF + 4);   // [B]: box b crops this CTA's image
  const int n = blockIdx.y;
  for (int b = threadIdx.x; b < a.B; b += CROP_BWD_THREADS) sMine[b] = a.map[b] == n;
  __syncthreads();
  const int p = blockIdx.x * CROP_BWD_THREADS + threadIdx.x;
  const bool live = p < a.H * a.W;
  const int y = live ? p / a.W : -4, x = live ? p % a.W : -4;
  float acc[CROP_MAX_C];
#pragma unroll
  for (int c = 0; c < CROP_MAX_C; ++c) acc[c] = 0.f;
  for (int b = 0; b < a.B; ++b) {
    if (!sMine[b]) continue;                                  // block-uniform
    __syncthreads();
    for (int t = threadIdx.x; t < a.WW + a.HH; t += CROP_BWD_THREADS) {
      SgBilin ax, ay;
      crop_axes(a, b, t < a.WW ? 0 : t - a.WW, t < a.WW ? t : 0, ax, ay);
      const SgBilin& s = t < a.WW ? ax : ay;
      CropTap tap;
      tap.i0 = s.i0;
      tap.w0 = s.ok0 ? s.w0 : 0.f;
      tap.w1 = s.ok1 ? s.w1 : 0.f;
      if (t < a.WW) sAx[t] = tap; else sAy[t - a.WW] = tap;
      // i0 + w1 is the un-normalised sample coordinate (w1 = frac) wherever the tap is finite and sg_axis did not clamp
      // the index (coordinates far outside the image): otherwise no estimate -> every crop index is a candidate
      const int size = t < a.WW ? a.W : a.H;
      const bool fin = (s.w0 != 0.f || s.w1 != 0.f) && s.i0 > -2 && s.i0 < size + 1;
      const float f = fin ? (float)s.i0 + s.w1 : __int_as_float(0x7fc00000);
      if (t == 0) sF[0] = f;
      if (t == a.WW - 1) sF[1] = f;
      if (t == a.WW) sF[2] = f;
      if (t == a.WW + a.HH - 1) sF[3] = f;
    }
    __syncthreads();
    int ilo, ihi, jlo, jhi;
    crop_candidates(sF[2], sF[3], a.HH, y, ilo, ihi);
    crop_candidates(sF[0], sF[1], a.WW, x, jlo, jhi);
    for (int i = ilo; i < ihi; ++i) {
      const CropTap ay = sAy[i];
      float wy;
      if (ay.i0 == y) wy = ay.w0;
      else if (ay.i0 + 1 == y) wy = ay.w1;
      else continue;
      if (wy == 0.f) continue;
      for (int j = jlo; j < jhi; ++j) {
        const CropTap ax = sAx[j];
        float wx;
        if (ax.i0 == x) wx = ax.w0;
        else if (ax.i0 + 1 == x) wx = ax.w1;
        else continue;
        if (wx == 0.f) continue;
        const float wgt = wx * wy;
        if (NHWC_BF16) {
          const __nv_bfloat16* g = (const __nv_bfloat16*)grad + (((long)b * a.HH + i) * a.WW + j) * a.Cp;
          for (int c = 0; c < a.C; ++c) acc[c] = fmaf(wgt, __bfloat162float(g[c]), acc[c]);
        } else {
          const float* g = (const float*)grad + (long)b * a.C * a.HH * a.WW + (long)i * a.WW + j;
          for (int c = 0; c < a.C; ++c) acc[c] = fmaf(wgt, g[(long)c * a.HH * a.WW], acc[c]);
        }
      }
    }
  }
  if (live)
    for (int c = 0; c < a.C; ++c) dfeats[((long)n * a.C + c) * a.H * a.W + p] = acc[c];
}

int check(const CropArgs& a, int fmt) {
  SG_CHECK_ARG(a.N > 0 && a.C > 0 && a.H > 0 && a.W > 0 && a.B >= 0 && a.HH > 0 && a.WW > 0, "crop_bbox: bad sizes");
  SG_CHECK_ARG(fmt == 0 || fmt == 1, "crop_bbox: format must be 0 (NCHW f32) or 1 (NHWC bf16)");
  SG_CHECK_ARG(fmt == 0 || (a.Cp >= a.C && a.Cp % 8 == 0), "crop_bbox: Cp must be a multiple of 8 >= C");
  return SG_OK;
}

}  // namespace

extern "C" int sg_crop_bbox_fwd(const float* feats, const float* boxes, const long long* box_to_feats, int N, int C,
                                int H, int W, int B, int HH, int WW, int align_corners, int out_format, int Cp,
                                void* out, cudaStream_t stream) {
  CropArgs a{feats, boxes, box_to_feats, N, C, H, W, B, HH, WW, align_corners, Cp};
  if (int e = check(a, out_format)) return e;
  if (B == 0) return SG_OK;
  long total = (long)B * HH * WW;
  if (out_format == 1) crop_fwd_kernel<true><<<sg_cdiv(total, 256), 256, 0, stream>>>(a, out);
  else crop_fwd_kernel<false><<<sg_cdiv(total, 256), 256, 0, stream>>>(a, out);
  SG_CHECK_LAUNCH("sg_crop_bbox_fwd");
  return SG_OK;
}

extern "C" int sg_crop_bbox_bwd(const float* boxes, const long long* box_to_feats, int N, int C, int H, int W, int B,
                                int HH, int WW, int align_corners, int grad_format, int Cp, const void* grad_out,
                                float* dfeats, cudaStream_t stream) {
  CropArgs a{nullptr, boxes, box_to_feats, N, C, H, W, B, HH, WW, align_corners, Cp};
  if (int e = check(a, grad_format)) return e;
  SG_CHECK_ARG(dfeats != nullptr, "crop_bbox_bwd: dfeats is null");
  SG_CHECK_ARG(C <= CROP_MAX_C, "crop_bbox_bwd: at most %d feature channels (images)", CROP_MAX_C);
  // every pixel of every image is written by exactly one thread (zeros where no crop touches it): B == 0 included
  dim3 grid(sg_cdiv((long)H * W, CROP_BWD_THREADS), N);
  const size_t smem = sizeof(CropTap) * (size_t)(HH + WW) + 16 + (size_t)B + 16;
  SG_CHECK_ARG(smem <= 48 * 1024, "crop_bbox_bwd: too many boxes / too large crops for the shared-memory tables");
  if (grad_format == 1) crop_bwd_kernel<true><<<grid, CROP_BWD_THREADS, smem, stream>>>(a, grad_out, dfeats);
  else crop_bwd_kernel<false><<<grid, CROP_BWD_THREADS, smem, stream>>>(a, grad_out, dfeats);
  SG_CHECK_LAUNCH("sg_crop_bbox_bwd");
  return SG_OK;
}
