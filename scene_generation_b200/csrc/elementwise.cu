// Element-wise / normalisation / re-layout kernels around the tensor-core convolutions.
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

__global__ void cast_pad_kernel(const float* __restrict__ src, long rows, int cols, long ld_src, int ld_dst,
                                __nv_bfloat16* __restrict__ dst) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = rows * ld_dst;
  if (idx >= total) return;
  long r = idx / ld_dst;
  int c = (int)(idx - r * ld_dst);
  dst[idx] = __float2bfloat16(c < cols ? src[r * ld_src + c] : 0.f);
}

}  // namespace

// f32 (rows, cols) with row pitch ld_src -> bf16 (rows, ld_dst), columns >= cols zero filled
extern "C" int sg_cast_pad_bf16(const float* src, long long rows, int cols, long long ld_src, int ld_dst, void* dst,
                                sg_stream_t stream) {
  SG_CHECK_ARG(rows >= 0 && cols > 0 && ld_dst >= cols && ld_src >= cols, "cast_pad_bf16: bad sizes");
  if (rows == 0) return SG_OK;
  long total = rows * ld_dst;
  cast_pad_kernel<<<sg_cdiv(total, 256), 256, 0, stream>>>(src, rows, cols, ld_src, ld_dst, (__nv_bfloat16*)dst);
  SG_CHECK_LAUNCH("sg_cast_pad_bf16");
  return SG_OK;
}
