// GraphTripleConv gather / pooled scatter (reference: scene_generation/graph.py:74-116).
//
//  * gather:  cur_t[t] = [obj[s_t] | pred[t] | obj[o_t]]        (graph.py:79-84)  -- bit-exact copy
//  * pool:    pooled[o] = (sum_{t: s_t=o} new_s[t] + sum_{t: o_t=o} new_o[t]) / max(count_o, 1)
//             (graph.py:94-116).  The reference uses atomic scatter_add; here the incidences of each
//             object are a CSR segment (built once per batch on the host), reduced by one warp per
//             (object, 128-column slab) in the reference's CPU order (all subject uses in triple
//             order, then all object uses) -> deterministic, no atomics.
//  * the two adjoints (pool_bwd is a gather, gather_bwd is a segmented reduce over the same CSR).
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

template <typename OutT>
__device__ __forceinline__ void store_val(OutT* p, float v);
template <>
__device__ __forceinline__ void store_val<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_val<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename OutT>
__global__ void gather_concat_kernel(const float* __restrict__ obj, const float* __restrict__ pred,
                                     const long long* __restrict__ edges, int T, int Do, int Dp, int ld_out,
                                     OutT* __restrict__ out) {
  int t = blockIdx.x;
  if (t >= T) return;
  long long s = edges[2 * t], o = edges[2 * t + 1];
  OutT* row = out + (long)t * ld_out;
  for (int c = threadIdx.x; c < ld_out; c += blockDim.x) {
    float v = 0.f;
    if (c < Do) v = obj[s * Do + c];
    else if (c < Do + Dp) v = pred[(long)t * Dp + (c - Do)];
    else if (c < 2 * Do + Dp) v = obj[o * Do + (c - Do - Dp)];
    store_val(row + c, v);
  }
}

// seg_src[i] = 2*t + role ; role 0 -> columns [0,H) of new_t, role 1 -> columns [col_o, col_o+H)
template <typename OutT>
__global__ void pool_kernel(const float* __restrict__ new_t, int ldt, int col_o, const int* __restrict__ seg_ptr,
                            const int* __restrict__ seg_src, int O, int H, int ld_out, int avg,
                            OutT* __restrict__ out) {
  int o = blockIdx.x;
  int b = seg_ptr[o], e = seg_ptr[o + 1];
  float inv_needed = (float)max(e - b, 1);
  for (int c = threadIdx.x; c < ld_out; c += blockDim.x) {
    float acc = 0.f;
    if (c < H) {
      for (int i = b; i < e; ++i) {
        int src = seg_src[i];
        int t = src >> 1, role = src & 1;
        acc = __fadd_rn(acc, new_t[(long)t * ldt + (role ? col_o : 0) + c]);
      }
      if (avg) acc = __fdiv_rn(acc, inv_needed);
    }
    store_val(out + (long)o * ld_out + c, acc);
  }
}

// d new_t[t] = [ dpooled[s_t]/cnt_s | dnew_p[t] | dpooled[o_t]/cnt_o ]
__global__ void pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ dnew_p,
                                const long long* __restrict__ edges, const int* __restrict__ seg_ptr, int T, int H,
                                int Dout, int avg, int ld_out, float* __restrict__ dnew_t) {
  int t = blockIdx.x;
  long long s = edges[2 * t], o = edges[2 * t + 1];
  float cs = avg ? (float)max(seg_ptr[s + 1] - seg_ptr[s], 1) : 1.f;
  float co = avg ? (float)max(seg_ptr[o + 1] - seg_ptr[o], 1) : 1.f;
  for (int c = threadIdx.x; c < ld_out; c += blockDim.x) {
    float v = 0.f;
    if (c < H) v = __fdiv_rn(dpooled[s * H + c], cs);
    else if (c < H + Dout) v = dnew_p ? dnew_p[(long)t * Dout + (c - H)] : 0.f;
    else if (c < 2 * H + Dout) v = __fdiv_rn(dpooled[o * H + (c - H - Dout)], co);
    dnew_t[(long)t * ld_out + c] = v;
  }
}

// d obj[o] = sum over incidences of the s- or o-slice of d cur_t ; d pred[t] = middle slice
__global__ void gather_bwd_kernel(const float* __restrict__ dcur, int ldc, const int* __restrict__ seg_ptr,
                                  const int* __restrict__ seg_src, int O, int T, int Do, int Dp,
                                  float* __restrict__ dobj, float* __restrict__ dpred) {
  int r = blockIdx.x;
  if (r < O) {
    int b = seg_ptr[r], e = seg_ptr[r + 1];
    for (int c = threadIdx.x; c < Do; c += blockDim.x) {
      float acc = 0.f;
      for (int i = b; i < e; ++i) {
        int src = seg_src[i];
        int t = src >> 1, role = src & 1;
        acc = __fadd_rn(acc, dcur[(long)t * ldc + (role ? Do + Dp : 0) + c]);
      }
      dobj[(long)r * Do + c] = acc;
    }
  } else {
    int t = r - O;
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) dpred[(long)t * Dp + c] = dcur[(long)t * ldc + Do + c];
  }
}

}  // namespace

extern "C" int sg_gconv_gather_fwd(const float* obj_vecs, const float* pred_vecs, const long long* edges, int O, int T,
                                   int Do, int Dp, int out_dtype, int ld_out, void* out, cudaStream_t stream) {
  SG_CHECK_ARG(T >= 0 && Do > 0 && Dp > 0 && ld_out >= 2 * Do + Dp, "gconv_gather_fwd: bad sizes");
  SG_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "gconv_gather_fwd: out_dtype must be 0 (f32) or 1 (bf16)");
  if (T == 0) return SG_OK;
  if (out_dtype == 0) gather_concat_kernel<float><<<T, 128, 0, stream>>>(obj_vecs, pred_vecs, edges, T, Do, Dp, ld_out, (float*)out);
  else gather_concat_kernel<__nv_bfloat16><<<T, 128, 0, stream>>>(obj_vecs, pred_vecs, edges, T, Do, Dp, ld_out, (__nv_bfloat16*)out);
  SG_CHECK_LAUNCH("sg_gconv_gather_fwd");
  return SG_OK;
}

extern "C" int sg_gconv_pool_fwd(const float* new_t, int ldt, int col_o, const int* seg_ptr, const int* seg_src, int O,
                                 int H, int avg, int out_dtype, int ld_out, void* out, cudaStream_t stream) {
  SG_CHECK_ARG(O >= 0 && H > 0 && ld_out >= H, "gconv_pool_fwd: bad sizes");
  SG_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "gconv_pool_fwd: out_dtype must be 0 (f32) or 1 (bf16)");
  if (O == 0) return SG_OK;
  if (out_dtype == 0) pool_kernel<float><<<O, 128, 0, stream>>>(new_t, ldt, col_o, seg_ptr, seg_src, O, H, ld_out, avg, (float*)out);
  else pool_kernel<__nv_bfloat16><<<O, 128, 0, stream>>>(new_t, ldt, col_o, seg_ptr, seg_src, O, H, ld_out, avg, (__nv_bfloat16*)out);
  SG_CHECK_LAUNCH("sg_gconv_pool_fwd");
  return SG_OK;
}

extern "C" int sg_gconv_pool_bwd(const float* dpooled, const float* dnew_p, const long long* edges, const int* seg_ptr,
                                 int T, int H, int Dout, int avg, int ld_out, float* dnew_t, cudaStream_t stream) {
  SG_CHECK_ARG(T >= 0 && H > 0 && Dout > 0 && ld_out >= 2 * H + Dout, "gconv_pool_bwd: bad sizes");
  if (T == 0) return SG_OK;
  pool_bwd_kernel<<<T, 128, 0, stream>>>(dpooled, dnew_p, edges, seg_ptr, T, H, Dout, avg, ld_out, dnew_t);
  SG_CHECK_LAUNCH("sg_gconv_pool_bwd");
  return SG_OK;
}

extern "C" int sg_gconv_gather_bwd(const float* dcur, int ldc, const int* seg_ptr, const int* seg_src, int O, int T,
                                   int Do, int Dp, float* dobj, float* dpred, cudaStream_t stream) {
  SG_CHECK_ARG(O >= 0 && T >= 0 && Do > 0 && Dp > 0 && ldc >= 2 * Do + Dp, "gconv_gather_bwd: bad sizes");
  if (O + T == 0) return SG_OK;
  gather_bwd_kernel<<<O + T, 128, 0, stream>>>(dcur, ldc, seg_ptr, seg_src, O, T, Do, Dp, dobj, dpred);
  SG_CHECK_LAUNCH("sg_gconv_gather_bwd");
  return SG_OK;
}
