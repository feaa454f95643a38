// GraphTripleConv gather / pooled scatter (reference: scene_generation/graph.py:74-116).
//
//  * gather:  cur_t[t] = [obj[s_t] | pred[t] | obj[o_t]]        (graph.py:79-84)  -- bit-exact copy
//  * pool:    pooled[o] = (sum_{t: s_t=o} new_s[t] + sum_{t: o_t=o} new_o[t]) / max(count_o, 1)
//             (graph.py:94-116).  The reference uses atomic scatter_add; here the incidences of each
//             object are a CSR segment (built once per batch on the host), reduced by one warp per
//             (object, 128-column slab): 16-byte loads, incidence indices loaded coalesced and broadcast
//             by warp shuffles, the adds in the reference's CPU order (all subject uses in triple order,
//             then all object uses) -> bit-identical to a sequential scatter_add, no atomics.
//  * the two adjoints (pool_bwd is a gather, gather_bwd is a segmented reduce over the same CSR).
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

constexpr int WARPS = 4;          // warps per CTA; one warp owns one (row, 128-column slab) work item

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// one 128-column slab of an output row: lane l holds columns [4l, 4l+4) of the slab as a float4
template <typename OutT>
__device__ __forceinline__ void store_slab(OutT* row, int c, float4 v, int ncols);
template <>
__device__ __forceinline__ void store_slab<float>(float* row, int c, float4 v, int ncols) {
  if (c + 4 <= ncols && ((reinterpret_cast<uintptr_t>(row + c) & 15) == 0)) {
    *reinterpret_cast<float4*>(row + c) = v;
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 4; ++j)
      if (c + j < ncols) row[c + j] = e[j];
  }
}
template <>
__device__ __forceinline__ void store_slab<__nv_bfloat16>(__nv_bfloat16* row, int c, float4 v, int ncols) {
  if (c + 4 <= ncols && ((reinterpret_cast<uintptr_t>(row + c) & 7) == 0)) {
    __align__(8) __nv_bfloat162 pk[2] = {__floats2bfloat162_rn(v.x, v.y), __floats2bfloat162_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(row + c) = *reinterpret_cast<const uint2*>(pk);
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 4; ++j)
      if (c + j < ncols) row[c + j] = __float2bfloat16(e[j]);
  }
}

// 4 consecutive elements of a row of `n` floats starting at column c (zeros beyond n); 16-byte load when aligned
__device__ __forceinline__ float4 load4(const float* row, int c, int n) {
  if (c + 4 <= n && ((reinterpret_cast<uintptr_t>(row + c) & 15) == 0)) return ld4(row + c);
  float e[4];
  for (int j = 0; j < 4; ++j) e[j] = (c + j < n) ? row[c + j] : 0.f;
  return make_float4(e[0], e[1], e[2], e[3]);
}

// gather (graph.py:79-84): out[t] = [obj[s_t] | pred[t] | obj[o_t]] (+ zero padding up to ld_out).  Warp = (triple,
// 128-column slab); a slab straddling two source segments falls back to per-element selection for those 4 columns.
template <typename OutT>
__global__ void __launch_bounds__(WARPS * 32) gather_concat_kernel(const float* __restrict__ obj, const float* __restrict__ pred,
                                                                   const long long* __restrict__ edges, int T, int Do, int Dp,
                                                                   int ld_out, int slabs, OutT* __restrict__ out) {
  const int item = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= T * slabs) return;
  const int t = item / slabs, c = (item - t * slabs) * 128 + lane * 4;
  if (c >= ld_out) return;
  const long long s = edges[2 * t], o = edges[2 * t + 1];
  const float* so = obj + s * Do;
  const float* sp = pred + (long)t * Dp;
  const float* oo = obj + o * Do;
  float4 v;
  if (c + 4 <= Do) v = load4(so, c, Do);
  else if (c >= Do && c + 4 <= Do + Dp) v = load4(sp, c - Do, Dp);
  else if (c >= Do + Dp && c + 4 <= 2 * Do + Dp) v = load4(oo, c - Do - Dp, Do);
  else {
    float e[4];
    for (int j = 0; j < 4; ++j) {
      const int cc = c + j;
      e[j] = cc < Do ? so[cc] : (cc < Do + Dp ? sp[cc - Do] : (cc < 2 * Do + Dp ? oo[cc - Do - Dp] : 0.f));
    }
    v = make_float4(e[0], e[1], e[2], e[3]);
  }
  store_slab(out + (long)t * ld_out, c, v, ld_out);
}

// Segmented sum over the incidences of one object, in the reference's order (graph.py:100-101,108-109: all subject
// uses in triple order, then all object uses — the CSR order): the warp loads up to 32 incidence indices with one
// coalesced load, broadcasts them lane to lane with shuffles, and every lane adds its 4 columns of each source row
// SEQUENTIALLY (__fadd_rn): bit-identical to a sequential scatter_add, no atomics.
// seg_src[i] = 2*t + role ; role 0 -> columns [0, W) of the source row, role 1 -> columns [col_1, col_1 + W).
__device__ __forceinline__ float4 segment_sum(const float* __restrict__ src, long ld, int col_1, const int* __restrict__ seg_src,
                                              int b, int e, int c, int W, int lane) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = b; i0 < e; i0 += 32) {
    const int mine = (i0 + lane < e) ? seg_src[i0 + lane] : 0;
    const int n = min(32, e - i0);
    for (int j = 0; j < n; ++j) {
      const int sidx = __shfl_sync(0xffffffffu, mine, j);
      const float4 v = load4(src + (long)(sidx >> 1) * ld + ((sidx & 1) ? col_1 : 0), c, W);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
      acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
  }
  return acc;
}

// pool (graph.py:94-116): pooled[o] = (sum of the object's subject / object messages) / max(count, 1)
template <typename OutT>
__global__ void __launch_bounds__(WARPS * 32) pool_kernel(const float* __restrict__ new_t, int ldt, int col_o,
                                                          const int* __restrict__ seg_ptr, const int* __restrict__ seg_src, int O,
                                                          int H, int ld_out, int slabs, int avg, OutT* __restrict__ out) {
  const int item = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= O * slabs) return;                       // warp-uniform
  const int o = item / slabs, c = (item - o * slabs) * 128 + lane * 4;
  const int b = seg_ptr[o], e = seg_ptr[o + 1];
  float4 acc = segment_sum(new_t, ldt, col_o, seg_src, b, e, c, H, lane);
  if (c >= ld_out) return;
  if (avg) {
    const float cnt = (float)max(e - b, 1);
    acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt); acc.z = __fdiv_rn(acc.z, cnt); acc.w = __fdiv_rn(acc.w, cnt);
  }
  store_slab(out + (long)o * ld_out, c, acc, ld_out);   // columns >= H come out as the zero padding
}

// d new_t[t] = [ dpooled[s_t]/cnt_s | dnew_p[t] | dpooled[o_t]/cnt_o ]   (adjoint of pool: a gather)
__global__ void __launch_bounds__(WARPS * 32) pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ dnew_p,
                                                              const long long* __restrict__ edges, const int* __restrict__ seg_ptr,
                                                              int T, int H, int Dout, int avg, int ld_out, int slabs,
                                                              float* __restrict__ dnew_t) {
  const int item = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= T * slabs) return;
  const int t = item / slabs, c = (item - t * slabs) * 128 + lane * 4;
  if (c >= ld_out) return;
  const long long s = edges[2 * t], o = edges[2 * t + 1];
  const float cs = avg ? (float)max(seg_ptr[s + 1] - seg_ptr[s], 1) : 1.f;
  const float co = avg ? (float)max(seg_ptr[o + 1] - seg_ptr[o], 1) : 1.f;
  float e[4];
  if (c + 4 <= H) {
    const float4 v = load4(dpooled + s * H, c, H);
    e[0] = __fdiv_rn(v.x, cs); e[1] = __fdiv_rn(v.y, cs); e[2] = __fdiv_rn(v.z, cs); e[3] = __fdiv_rn(v.w, cs);
  } else if (c >= H && c + 4 <= H + Dout) {
    const float4 v = dnew_p ? load4(dnew_p + (long)t * Dout, c - H, Dout) : make_float4(0.f, 0.f, 0.f, 0.f);
    e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
  } else if (c >= H + Dout && c + 4 <= 2 * H + Dout) {
    const float4 v = load4(dpooled + o * H, c - H - Dout, H);
    e[0] = __fdiv_rn(v.x, co); e[1] = __fdiv_rn(v.y, co); e[2] = __fdiv_rn(v.z, co); e[3] = __fdiv_rn(v.w, co);
  } else {
    for (int j = 0; j < 4; ++j) {
      const int cc = c + j;
      float v = 0.f;
      if (cc < H) v = __fdiv_rn(dpooled[s * H + cc], cs);
      else if (cc < H + Dout) v = dnew_p ? dnew_p[(long)t * Dout + (cc - H)] : 0.f;
      else if (cc < 2 * H + Dout) v = __fdiv_rn(dpooled[o * H + (cc - H - Dout)], co);
      e[j] = v;
    }
  }
  store_slab(dnew_t + (long)t * ld_out, c, make_float4(e[0], e[1], e[2], e[3]), ld_out);
}

// adjoint of the gather: d obj[o] = segmented sum of the s- / o-slices of d cur_t (same CSR, same order);
// d pred[t] = the middle slice (a copy).  Work items: O * slabs_o segment sums, then T * slabs_p copies.
__global__ void __launch_bounds__(WARPS * 32) gather_bwd_kernel(const float* __restrict__ dcur, int ldc,
                                                                const int* __restrict__ seg_ptr, const int* __restrict__ seg_src,
                                                                int O, int T, int Do, int Dp, int slabs_o, int slabs_p,
                                                                float* __restrict__ dobj, float* __restrict__ dpred) {
  const int item = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item < O * slabs_o) {
    const int o = item / slabs_o, c = (item - o * slabs_o) * 128 + lane * 4;
    const float4 acc = segment_sum(dcur, ldc, Do + Dp, seg_src, seg_ptr[o], seg_ptr[o + 1], c, Do, lane);
    if (c < Do) store_slab(dobj + (long)o * Do, c, acc, Do);
  } else if (item < O * slabs_o + T * slabs_p) {
    const int it = item - O * slabs_o;
    const int t = it / slabs_p, c = (it - t * slabs_p) * 128 + lane * 4;
    if (c < Dp) store_slab(dpred + (long)t * Dp, c, load4(dcur + (long)t * ldc + Do, c, Dp), Dp);
  }
}

}  // namespace

extern "C" int sg_gconv_gather_fwd(const float* obj_vecs, const float* pred_vecs, const long long* edges, int O, int T,
                                   int Do, int Dp, int out_dtype, int ld_out, void* out, cudaStream_t stream) {
  SG_CHECK_ARG(T >= 0 && Do > 0 && Dp > 0 && ld_out >= 2 * Do + Dp, "gconv_gather_fwd: bad sizes");
  SG_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "gconv_gather_fwd: out_dtype must be 0 (f32) or 1 (bf16)");
  if (T == 0) return SG_OK;
  const int slabs = sg_cdiv(ld_out, 128);
  const int blocks = sg_cdiv((long)T * slabs, WARPS);
  if (out_dtype == 0)
    gather_concat_kernel<float><<<blocks, WARPS * 32, 0, stream>>>(obj_vecs, pred_vecs, edges, T, Do, Dp, ld_out, slabs, (float*)out);
  else
    gather_concat_kernel<__nv_bfloat16><<<blocks, WARPS * 32, 0, stream>>>(obj_vecs, pred_vecs, edges, T, Do, Dp, ld_out, slabs,
                                                                        (__nv_bfloat16*)out);
  SG_CHECK_LAUNCH("sg_gconv_gather_fwd");
  return SG_OK;
}

extern "C" int sg_gconv_pool_fwd(const float* new_t, int ldt, int col_o, const int* seg_ptr, const int* seg_src, int O,
                                 int H, int avg, int out_dtype, int ld_out, void* out, cudaStream_t stream) {
  SG_CHECK_ARG(O >= 0 && H > 0 && ld_out >= H, "gconv_pool_fwd: bad sizes");
  SG_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "gconv_pool_fwd: out_dtype must be 0 (f32) or 1 (bf16)");
  if (O == 0) return SG_OK;
  const int slabs = sg_cdiv(ld_out, 128);
  const int blocks = sg_cdiv((long)O * slabs, WARPS);
  if (out_dtype == 0)
    pool_kernel<float><<<blocks, WARPS * 32, 0, stream>>>(new_t, ldt, col_o, seg_ptr, seg_src, O, H, ld_out, slabs, avg, (float*)out);
  else
    pool_kernel<__nv_bfloat16><<<blocks, WARPS * 32, 0, stream>>>(new_t, ldt, col_o, seg_ptr, seg_src, O, H, ld_out, slabs, avg,
                                                               (__nv_bfloat16*)out);
  SG_CHECK_LAUNCH("sg_gconv_pool_fwd");
  return SG_OK;
}

extern "C" int sg_gconv_pool_bwd(const float* dpooled, const float* dnew_p, const long long* edges, const int* seg_ptr,
                                 int T, int H, int Dout, int avg, int ld_out, float* dnew_t, cudaStream_t stream) {
  SG_CHECK_ARG(T >= 0 && H > 0 && Dout > 0 && ld_out >= 2 * H + Dout, "gconv_pool_bwd: bad sizes");
  if (T == 0) return SG_OK;
  const int slabs = sg_cdiv(ld_out, 128);
  pool_bwd_kernel<<<sg_cdiv((long)T * slabs, WARPS), WARPS * 32, 0, stream>>>(dpooled, dnew_p, edges, seg_ptr, T, H, Dout, avg, ld_out,
                                                                            slabs, dnew_t);
  SG_CHECK_LAUNCH("sg_gconv_pool_bwd");
  return SG_OK;
}

extern "C" int sg_gconv_gather_bwd(const float* dcur, int ldc, const int* seg_ptr, const int* seg_src, int O, int T,
                                   int Do, int Dp, float* dobj, float* dpred, cudaStream_t stream) {
  SG_CHECK_ARG(O >= 0 && T >= 0 && Do > 0 && Dp > 0 && ldc >= 2 * Do + Dp, "gconv_gather_bwd: bad sizes");
  if (O + T == 0) return SG_OK;
  const int slabs_o = sg_cdiv(Do, 128), slabs_p = sg_cdiv(Dp, 128);
  const long items = (long)O * slabs_o + (long)T * slabs_p;
  gather_bwd_kernel<<<sg_cdiv(items, WARPS), WARPS * 32, 0, stream>>>(dcur, ldc, seg_ptr, seg_src, O, T, Do, Dp, slabs_o, slabs_p, dobj,
                                                                    dpred);
  SG_CHECK_LAUNCH("sg_gconv_gather_bwd");
  return SG_OK;
}
