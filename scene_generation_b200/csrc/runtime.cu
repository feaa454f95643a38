// Runtime entry points of libsg_b200: error reporting, version, launch accounting.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"
#include "../../include/sg_b200.h"

thread_local char sg_err_buf[512] = {0};
static std::atomic<unsigned long long> g_launches{0};

int sg_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(sg_err_buf, sizeof(sg_err_buf), fmt, ap);
  va_end(ap);
  return code;
}

void sg_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" const char* sg_last_error(void) { return sg_err_buf; }
extern "C" const char* sg_version(void) { return "sg_b200 0.1.0 (sm_100a)"; }
extern "C" int sg_arch(void) { return 100; }
extern "C" unsigned long long sg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void sg_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
// launches replayed from a captured CUDA graph (the host enqueued them once, at capture): the caller accounts them
extern "C" void sg_add_launch_count(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// dst[i] = src[0][i] + src[1][i] + ... + src[parts-1][i], added in that order (the second half of every fixed-order
// reduction of the library: producers leave one partial row per CTA / split, this kernel adds the rows).
namespace {
__global__ void sum_parts_kernel(const float* __restrict__ src, long n, int parts, long part_stride, float* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 4 >= n) return;
  if (i * 4 + 4 <= n && (part_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < parts; ++p) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long)p * part_stride) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dst)[i] = acc;
  } else {
    for (long e = i * 4; e < n && e < i * 4 + 4; ++e) {
      float acc = 0.f;
      for (int p = 0; p < parts; ++p) acc += __ldcg(src + (long)p * part_stride + e);
      dst[e] = acc;
    }
  }
}

// Many parts, few elements (bias gradients: up to 296 partial rows of C floats): a serial loop per element would be one
// long latency chain.  G groups of threads add the parts g, g + G, g + 2G, ... of an element each; the G group sums
// are then added in group order.  The grouping depends only on (parts, G): still a fixed order.
template <int G>
__global__ void __launch_bounds__(256) sum_parts_grouped_kernel(const float* __restrict__ src, long n, int parts, long part_stride,
                                                                float* __restrict__ dst) {
  constexpr int COLS = 256 / G;
  __shared__ float red[G][COLS];
  const int col = threadIdx.x % COLS, g = threadIdx.x / COLS;
  const long e = (long)blockIdx.x * COLS + col;
  float acc = 0.f;
  if (e < n)
    for (int p = g; p < parts; p += G) acc += __ldcg(src + (long)p * part_stride + e);
  red[g][col] = acc;
  __syncthreads();
  if (g == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < G; ++k) t += red[k][col];
    dst[e] = t;
  }
}
}  // namespace

int sg_sum_parts(const float* src, long n, int parts, long part_stride, float* dst, cudaStream_t stream, const char* what) {
  if (n <= 0) return SG_OK;
  if (parts > 64 && n <= 65536)
    sum_parts_grouped_kernel<32><<<sg_cdiv(n, 8), 256, 0, stream>>>(src, n, parts, part_stride, dst);
  else if (parts > 8 && n <= 65536)
    sum_parts_grouped_kernel<8><<<sg_cdiv(n, 32), 256, 0, stream>>>(src, n, parts, part_stride, dst);
  else
    sum_parts_kernel<<<sg_cdiv(sg_cdiv(n, 4), 256), 256, 0, stream>>>(src, n, parts, part_stride, dst);
  SG_CHECK_LAUNCH(what);
  return SG_OK;
}
