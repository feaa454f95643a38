// Differentiable bilinear box crop (reference: scene_generation/bilinear.py:26-130, 246-275).
//   crops[b, c, i, j] = bilinear_zero_pad(feats[map[b], c]; X_j, Y_i)
//   X_j = lin10(j) * (2*x0-1) + lin01(j) * (2*x1-1)   (tensor_linspace, bilinear.py:263-274)
// The reference replicates every image once per box (bilinear.py:80) and fixes the order up with
// an inverse permutation (bilinear.py:94-98); here each output element gathers its 4 taps directly,
// so crops come out in box order by construction.
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

template <bool NHWC_BF16>
__global__ void crop_bwd_kernel(CropArgs a, const void* grad, float* dfeats) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)a.B * a.HH * a.WW;
  if (idx >= total) return;
  int j = idx % a.WW, i = (idx / a.WW) % a.HH, b = idx / ((long)a.WW * a.HH);
  SgBilin ax, ay;
  crop_axes(a, b, i, j, ax, ay);
  long n = a.map[b];
  float wnw = ax.w0 * ay.w0, wne = ax.w1 * ay.w0, wsw = ax.w0 * ay.w1, wse = ax.w1 * ay.w1;
  for (int c = 0; c < a.C; ++c) {
    float g = NHWC_BF16 ? __bfloat162float(((const __nv_bfloat16*)grad)[idx * a.Cp + c])
                        : ((const float*)grad)[(((long)b * a.C + c) * a.HH + i) * a.WW + j];
    if (g == 0.f) continue;
    float* f = dfeats + ((long)n * a.C + c) * a.H * a.W;
    if (ay.ok0 && ax.ok0) atomicAdd(f + ay.i0 * a.W + ax.i0, wnw * g);
    if (ay.ok0 && ax.ok1) atomicAdd(f + ay.i0 * a.W + ax.i0 + 1, wne * g);
    if (ay.ok1 && ax.ok0) atomicAdd(f + (ay.i0 + 1) * a.W + ax.i0, wsw * g);
    if (ay.ok1 && ax.ok1) atomicAdd(f + (ay.i0 + 1) * a.W + ax.i0 + 1, wse * g);
  }
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
  cudaMemsetAsync(dfeats, 0, sizeof(float) * (size_t)N * C * H * W, stream);
  if (B == 0) return SG_OK;
  long total = (long)B * HH * WW;
  if (grad_format == 1) crop_bwd_kernel<true><<<sg_cdiv(total, 256), 256, 0, stream>>>(a, grad_out, dfeats);
  else crop_bwd_kernel<false><<<sg_cdiv(total, 256), 256, 0, stream>>>(a, grad_out, dfeats);
  SG_CHECK_LAUNCH("sg_crop_bbox_bwd");
  return SG_OK;
}
