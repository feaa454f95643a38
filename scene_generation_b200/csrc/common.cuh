// Shared helpers for libsg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SG_OK 0
#define SG_ERR_ARG 1
#define SG_ERR_CUDA 2

extern thread_local char sg_err_buf[512];
int sg_fail(int code, const char* fmt, ...);

#define SG_CHECK_ARG(cond, ...)                              \
  do {                                                       \
    if (!(cond)) return sg_fail(SG_ERR_ARG, __VA_ARGS__);    \
  } while (0)

#define SG_CHECK_LAUNCH(name)                                                              \
  do {                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess) return sg_fail(SG_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_)); \
    sg_count_launch();                                                                     \
  } while (0)

void sg_count_launch();

static inline int sg_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// dst[i] = sum over p < parts of src[p * part_stride + i], added in part order (runtime.cu)
int sg_sum_parts(const float* src, long n, int parts, long part_stride, float* dst, cudaStream_t stream, const char* what);

// torch.linspace(0,1,steps)[i] in fp32, evaluated from both ends like ATen
// (reference: layout.py:114-115, bilinear.py:263-265).
__device__ __forceinline__ float sg_linspace01(int i, int steps) {
  if (steps <= 1) return 0.f;
  float step = __fdiv_rn(1.0f, (float)(steps - 1));
  return (i < steps / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(steps - i - 1)));
}
// torch.linspace(1,0,steps)[i]
__device__ __forceinline__ float sg_linspace10(int i, int steps) {
  if (steps <= 1) return 1.f;
  float step = __fdiv_rn(-1.0f, (float)(steps - 1));
  return (i < steps / 2) ? __fadd_rn(1.0f, __fmul_rn(step, (float)i)) : __fsub_rn(0.0f, __fmul_rn(step, (float)(steps - i - 1)));
}

// F.grid_sample un-normalisation (zeros padding, bilinear).
__device__ __forceinline__ float sg_unnormalize(float c, int size, int align_corners) {
  return align_corners ? __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1))
                       : __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.f), (float)size), 1.f), 0.5f);
}

struct SgBilin {   // one axis of a bilinear tap: indices i0, i0+1 with weights w0 (for i0), w1
  int i0;
  float w0, w1;
  int ok0, ok1;
};
__device__ __forceinline__ SgBilin sg_axis(float coord, int size, int align_corners) {
  SgBilin a;
  float f = sg_unnormalize(coord, size, align_corners);
  if (!isfinite(f)) { a.i0 = 0; a.w0 = a.w1 = 0.f; a.ok0 = a.ok1 = 0; return a; }
  float fl = floorf(f);
  // clamp before the int conversion so far-away coordinates cannot overflow
  float flc = fminf(fmaxf(fl, -2.f), (float)size + 1.f);
  a.i0 = (int)flc;
  a.w1 = __fsub_rn(f, fl);
  a.w0 = __fsub_rn(__fadd_rn(fl, 1.f), f);
  a.ok0 = (fl >= 0.f) && (fl <= (float)(size - 1));
  a.ok1 = (fl + 1.f >= 0.f) && (fl + 1.f <= (float)(size - 1));
  return a;
}
