// tcgen05 / TMA implicit-GEMM convolution family for sm_100a.
//
// Replaces the cuDNN fprop/dgrad/wgrad + cuBLAS GEMM dispatch behind nn.Conv2d / nn.ConvTranspose2d /
// nn.Linear on the reference hot path (generators.py:62-91, layers.py:234-273,
// discriminators.py:87-245, graph.py:85,120).
//
//   conv_tc_kernel : D[128 pixels, BN couts] = sum_taps sum_kblocks A_tap[128, 64] * B_tap[BN, 64]^T
//     A tile = one 5-D TMA box (64 ch, BW, BH, 1 plane, BI imgs) of the NHWC bf16 activation at a
//     tap-shifted coordinate (out-of-range -> zero = padding);  B tile = 3-D TMA box (64 ch, 1 tap, BN)
//     of the [Cout][taps][Cin] bf16 weights.  Both K-major, SWIZZLE_128B; fp32 accumulators in TMEM.
//     Warp roles: w0 TMA producer, w1 TMEM alloc + single-thread tcgen05.mma issuer, w2-5 epilogue
//     (tcgen05.ld -> bias -> InstanceNorm/BatchNorm partial sums -> activation -> vector stores).
//   wgrad_tc_kernel : D[128 couts, BN cins] = sum over 64-pixel boxes dy[pix, co]^T x[pix(+tap), ci]
//     both operands MN-major (pixel = K is the strided smem dimension), split-K reduced in split order.
// Every reduction in this file has a fixed order (no floating-point atomics): results are bit-reproducible.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/sg_b200.h"

namespace {

using namespace ptx;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 tensor [d4][d3][d2][d1][d0] (d0 innermost, contiguous) -> tiled map with the given box
int make_tmap(CUtensorMap* m, const void* base, int rank, const long long* dims, const int* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sg_fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5], estr[5];
  unsigned long long stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    estr[i] = 1;
    stride *= (unsigned long long)dims[i];
    if (i < rank - 1) gstride[i] = stride;
  }
  if (((uintptr_t)base & 15) != 0) return sg_fail(SG_ERR_ARG, "tensor base %p not 16-byte aligned", base);
  if ((dims[0] * 2) % 16 != 0) return sg_fail(SG_ERR_ARG, "innermost extent %lld not a multiple of 8 elements", dims[0]);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sg_fail(SG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SG_OK;
}

// lane l ends up with the sum over the warp of v[l]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      float send = up ? v[i] : v[i + s];
      float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------
struct ConvKParams {
  int Hout, Wout, tiles_w, tiles_h, BW, BH, BI, n_img;
  int kblocks, Cout, stages;
  int w_img_rows;   // > 0: per-image weights, image i uses rows [i * w_img_rows + w_row0, ...) of the weight tensor
  int w_row0;
  int w_col0;       // BMN kernels: first weight column (= output channel offset inside the fprop weight tensor)
  int a_bytes;      // bytes of one A box = 128 * BW*BH*BI (rows beyond the box keep stale smem and are masked)
  int in_h0, in_w0;
  long long os_img, os_h, os_w, os_c;
  int oh_mul, ow_mul;
  void* y;
  int y_dtype;
  const float* bias;
  int act;
  float slope;
  float* stats;
  int stat_slots;       // partial-sum slots per (image, channel): nphases * slots_per_phase
  int slots_per_phase;  // M tiles per image (BI == 1) or 1 (whole images inside a tile)
  int vec_ok;
  sg_phase_t phases[4];
  sg_tap_t taps[SG_MAX_TAPS];
};

constexpr int A_BYTES = 128 * 128;   // 128 rows x 64 bf16

template <int BN>
struct ConvCfg {
  // Pipeline depth is a launch parameter.  Default ("shallow"): two CTAs share an SM (<= 113 KB each), so the
  // prologue / TMEM drain / store epilogue of one tile overlaps the MMA main loop of the other; "deep"
  // (SG_CONV_DEEP=1): one CTA per SM with the whole shared memory as its ring.
  static constexpr int STAGES_DEEP = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int STAGES_SHALLOW = (BN >= 192) ? 2 : (BN == 128 ? 3 : (BN == 64 ? 4 : 6));
  static constexpr int B_BYTES = BN * 128;
  static constexpr int TM_COLS = BN < 32 ? 32 : (BN == 192 ? 256 : BN);   // TMEM allocations are powers of two
  static constexpr int B_STRIDE = (B_BYTES < 1024 ? 1024 : B_BYTES);
  static constexpr int smem_bytes(int stages) { return stages * (A_BYTES + B_STRIDE) + 1024 /*align slack*/ + 256 /*barriers*/; }
};

// The 128 epilogue threads (warps 2..5) synchronise among themselves on named barrier 1.
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

constexpr int EPI_SCRATCH_FLOATS = 128 * 33;   // shared-memory scratch of the epilogue (aliases the drained operand ring)

// Epilogue of one 128 x BN accumulator tile: tcgen05.ld -> bias -> InstanceNorm/BatchNorm partial sums -> activation
// -> vector stores.  Executed by the four epilogue warps (q = TMEM lane quarter).
//
// Statistics are DETERMINISTIC: no atomics.  Every (image, channel) of the output gets `stat_slots` partial
// (sum, sum of squares) pairs — one per (phase, M tile of that image) — each written by exactly one thread after a
// fixed-order reduction (warp shuffle tree over 32 rows, then the warps of the image in order);
// sg_norm_finalize adds the slots in slot order.  Layout: stats[img][slot][Cout][2].
template <int BN>
__device__ __forceinline__ void conv_epilogue(const ConvKParams& p, const sg_phase_t& ph, const uint32_t tmem, const int q,
                                              const int lane, const int w0, const int h0, const int img0, const int n0,
                                              const int slot, float* scratch) {
  const int r = q * 32 + lane;
  const int ww = r % p.BW, hh = (r / p.BW) % p.BH, ii = r / (p.BW * p.BH);
  const int img = img0 + ii, h = h0 + hh, w = w0 + ww;
  const bool valid = (ii < p.BI) && (img < p.n_img) && (h < p.Hout) && (w < p.Wout);
  const long long off = (long long)img * p.os_img + (long long)(h * p.oh_mul + ph.oh_off) * p.os_h +
                        (long long)(w * p.ow_mul + ph.ow_off) * p.os_w;
  // seg_full: the 32 rows of a warp belong to one image.  Otherwise several (tiny) images share a warp: BI > 1, which
  // choose_tile only picks when a whole image fits the tile, i.e. BW == Wout and BH == Hout.
  const bool seg_full = (p.BI == 1) || ((p.BW * p.BH) % 32 == 0);
  constexpr int CH = (BN >= 32) ? 32 : 16;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += CH) {
    if (n0 + c0 >= p.Cout) break;                       // CTA-uniform
    uint32_t raw[32];
    if (CH == 32) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, raw);
    else tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, raw);
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {                      // the bias loads overlap the TMEM read
      const int c = n0 + c0 + j;
      f[j] = (j < CH && p.bias != nullptr && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < CH; ++j) f[j] += __uint_as_float(raw[j]);
    if (p.stats != nullptr) {
      if (seg_full) {
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = valid ? f[j] : 0.f;
          s1[j] = v;
          s2[j] = v * v;
        }
        float t1 = warp_transpose_reduce(s1, lane);
        float t2 = warp_transpose_reduce(s2, lane);
        if (lane < CH) reinterpret_cast<float2*>(scratch)[q * BN + c0 + lane] = make_float2(t1, t2);
      } else {
        epi_bar();                                       // the previous chunk's readers are done with the scratch
#pragma unroll
        for (int j = 0; j < CH; ++j) scratch[r * 33 + j] = f[j];
        epi_bar();
        const int px = p.BW * p.BH;
        for (int idx = r; idx < p.BI * CH; idx += 128) {
          const int si = idx / CH, j = idx - si * CH;
          const int c = n0 + c0 + j, simg = img0 + si;
          if (c < p.Cout && simg < p.n_img) {
            float a = 0.f, b = 0.f;
            for (int rr = si * px; rr < (si + 1) * px; ++rr) {      // rows of this image, in order
              const float v = scratch[rr * 33 + j];
              a += v;
              b = fmaf(v, v, b);
            }
            *reinterpret_cast<float2*>(p.stats + (((long long)simg * p.stat_slots + slot) * p.Cout + c) * 2) = make_float2(a, b);
          }
        }
      }
    }
    if (valid) {
      // one CTA-uniform branch per chunk, not one jump table per element (the epilogue warps run one per scheduler:
      // every taken indirect branch is an exposed refetch)
      if (p.act == SG_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (p.act == SG_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * p.slope;
      } else if (p.act == SG_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = tanhf(f[j]);
      } else if (p.act == SG_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = 1.f / (1.f + __expf(-f[j]));
      }
      const bool full_chunk = (n0 + c0 + CH <= p.Cout) && p.vec_ok;
      if (p.y_dtype == 1) {
        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + off + (long long)(n0 + c0) * p.os_c;
        if (full_chunk) {
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) pk[t] = __floats2bfloat162_rn(f[j + 2 * t], f[j + 2 * t + 1]);
            *reinterpret_cast<uint4*>(yp + j) = *reinterpret_cast<uint4*>(pk);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n0 + c0 + j < p.Cout) yp[(long long)j * p.os_c] = __float2bfloat16(f[j]);
        }
      } else {
        float* yp = reinterpret_cast<float*>(p.y) + off + (long long)(n0 + c0) * p.os_c;
        if (full_chunk) {
#pragma unroll
          for (int j = 0; j < CH; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n0 + c0 + j < p.Cout) yp[(long long)j * p.os_c] = f[j];
        }
      }
    }
  }
  if (p.stats != nullptr && seg_full) {
    // combine the warps of each image of the tile in warp order: gq warps per image, 4 / gq images per tile
    epi_bar();
    const int gq = (p.BI == 1) ? 4 : (p.BW * p.BH) / 32;
    const int groups = 4 / gq;
    for (int idx = r; idx < groups * BN; idx += 128) {
      const int g = idx / BN, cc = idx - g * BN;
      const int c = n0 + cc, simg = img0 + g;
      if (c < p.Cout && simg < p.n_img && g < p.BI) {
        float a = 0.f, b = 0.f;
        for (int qq = g * gq; qq < (g + 1) * gq; ++qq) {
          const float2 v = reinterpret_cast<const float2*>(scratch)[qq * BN + cc];
          a += v.x;
          b += v.y;
        }
        *reinterpret_cast<float2*>(p.stats + (((long long)simg * p.stat_slots + slot) * p.Cout + c) * 2) = make_float2(a, b);
      }
    }
  }
}

// BMN = true: the weight operand is read "transposed" from the fprop tensor [K rows][taps][N cols] (dgrad of a
// convolution uses the SAME bf16 copy of the weights as its fprop): 64 x 64 TMA boxes with the output channel as
// the contiguous dimension, i.e. an MN-major B operand (the layout wgrad_tc_kernel uses for both operands).
template <int BN, bool BMN>
__global__ void __launch_bounds__(192, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BN>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_STRIDE);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.BW, h0 = th * p.BH, img0 = ti * p.BI;
  const int n0 = blockIdx.y * BN;
  const int wrow0 = n0 + img0 * p.w_img_rows + p.w_row0;        // weight-tensor row of this tile's first output channel
  const sg_phase_t ph = p.phases[blockIdx.z];
  const int iters = ph.ntaps * p.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&empty[s], par ^ 1);
        const int tap_i = it / p.kblocks, kb = it - tap_i * p.kblocks;
        const sg_tap_t tp = p.taps[ph.tap_begin + tap_i];
        mbar_expect_tx(&full[s], p.a_bytes + Cfg::B_BYTES);
        tma_load_5d(sA + s * A_BYTES, &tmA, &full[s], kb * 64, w0 + tp.dw + p.in_w0, h0 + tp.dh + p.in_h0, tp.plane, img0);
        if (BMN) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_3d(sB + s * Cfg::B_STRIDE + j * 8192, &tmB, &full[s], p.w_col0 + n0 + 64 * j, tp.wtap, kb * 64);
        } else {
          tma_load_3d(sB + s * Cfg::B_STRIDE, &tmB, &full[s], kb * 64, tp.wtap, wrow0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, BMN ? 1 : 0);
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&full[s], par);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
        // K-major B: 16-element K steps are 32 B apart inside the 128 B rows.  MN-major B: 64 N-elements per 128 B
        // row, 8 K-rows per 1024 B group (SBO), next 64-column block one 8 KB box further (LBO); a K step of 16 rows
        // is 2048 B.
        const uint64_t bd = BMN ? umma_desc_sw128(smem_u32(sB + s * Cfg::B_STRIDE), 8192, 1024)
                                : umma_desc_sw128(smem_u32(sB + s * Cfg::B_STRIDE), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16(tmem, ad + 2 * k, bd + (BMN ? 128 : 2) * k, idesc, (it | k) != 0 ? 1u : 0u);
        mma_commit(&empty[s]);   // frees the smem slot once these MMAs have read it
      }
      mma_commit(accum_full);
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) -------------------
    mbar_wait(accum_full, 0);
    tc_fence_after();
    // every TMA load has landed and every MMA has read its operands: the operand ring is free and serves as scratch
    const int slot = blockIdx.z * p.slots_per_phase + (p.BI == 1 ? th * p.tiles_w + tw : 0);
    conv_epilogue<BN>(p, ph, tmem, warp & 3, lane, w0, h0, img0, n0, slot, reinterpret_cast<float*>(sA));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant for long-K, wide-N GEMMs (the 1024-channel ResnetBlock convolutions and their input gradients):
// two CTAs of a cluster own two neighbouring M tiles and ONE N tile of BN = 256; the leader issues
// tcgen05.mma.cta_group::2 with M = 256: A = 128 rows from each CTA's shared memory, B = 128 output channels from each
// CTA.  Per 64-channel k-block an SM streams 16 KB of A + 16 KB of B for 128 x 256 outputs — half the L2 -> SM operand
// bytes per MMA cycle of the single-CTA 128 x 128 tile, which is what limits that kernel on these shapes (ncu: 604 MB
// through the L2 -> SM path per launch at its ~6.3 kB/clk ceiling).  Used where the pairs alone fill most SMs (the
// input gradients: 64 pairs); the forward GEMM (M = 2048: 32 pairs) stays on the single-CTA kernel — a two-way K split
// of the pairs ran its MMA phase in 38 us at 66 % tensor pipe (ncu, profiles/), but exchanging the fp32 accumulators
// between the two CTAs of a tile cost what the halved operand stream had gained (66 us either way).
// Barrier protocol (CUTLASS PipelineTmaUmmaAsync, 2x1 atom): both producers wait on their own `empty`, only the leader
// arms `full` with the bytes of BOTH CTAs, both issue their TMA loads against the leader's `full`; the leader's MMA
// thread commits to `empty` / `accum_full` of both CTAs (multicast).
template <bool BMN>
__global__ void __launch_bounds__(192, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ ConvKParams p) {
  constexpr int BN = 256;
  constexpr int B_HALF = (BN / 2) * 128;        // bytes of this CTA's half of the weight tile per stage
  constexpr int TM_COLS = BN;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_HALF);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int mt = blockIdx.x;                    // cluster dim x = 2: the pair is (2i, 2i + 1)
  const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.BW, h0 = th * p.BH, img0 = ti * p.BI;
  const int n0 = blockIdx.y * BN;
  const sg_phase_t ph = p.phases[blockIdx.z];
  const int it_begin = 0, it_end = ph.ntaps * p.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) tmem_alloc2<TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                               // peer barriers / TMEM exist before any remote signal
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t par = 0;
      for (int it = it_begin; it < it_end; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&empty[s], par ^ 1);
        const int tap_i = it / p.kblocks, kb = it - tap_i * p.kblocks;
        const sg_tap_t tp = p.taps[ph.tap_begin + tap_i];
        if (crank == 0) mbar_expect_tx(&full[s], 2 * (p.a_bytes + B_HALF));
        const uint32_t bar = leader_bar_addr(&full[s]);
        tma_load_5d_2cta(sA + s * A_BYTES, &tmA, bar, kb * 64, w0 + tp.dw + p.in_w0, h0 + tp.dh + p.in_h0, tp.plane, img0);
        if (BMN) {
#pragma unroll
          for (int j = 0; j < BN / 128; ++j)
            tma_load_3d_2cta(sB + s * B_HALF + j * 8192, &tmB, bar, p.w_col0 + n0 + (int)crank * (BN / 2) + 64 * j, tp.wtap,
                             kb * 64);
        } else {
          tma_load_3d_2cta(sB + s * B_HALF, &tmB, bar, kb * 64, tp.wtap, n0 + (int)crank * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, BMN ? 1 : 0);
      int s = 0;
      uint32_t par = 0;
      for (int it = it_begin; it < it_end; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&full[s], par);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
        const uint64_t bd = BMN ? umma_desc_sw128(smem_u32(sB + s * B_HALF), 8192, 1024)
                                : umma_desc_sw128(smem_u32(sB + s * B_HALF), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_2cta(tmem, ad + 2 * k, bd + (BMN ? 128 : 2) * k, idesc, ((it - it_begin) | k) != 0 ? 1u : 0u);
        mma_commit_2cta(&empty[s], (uint16_t)3);
      }
      mma_commit_2cta(accum_full, (uint16_t)3);
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) -------------------
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int slot = blockIdx.z * p.slots_per_phase + (p.BI == 1 ? th * p.tiles_w + tw : 0);
    conv_epilogue<BN>(p, ph, tmem, warp & 3, lane, w0, h0, img0, n0, slot, reinterpret_cast<float*>(sA));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                               // the leader's MMAs read this CTA's shared memory until the last commit
  if (warp == 1) tmem_dealloc2<TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------
struct WgradKParams {
  int tiles_w, tiles_h, BW, BH, BI;
  int ktiles_total, ktiles_per_split;
  int serial;       // 1: split-K through per-split workspace slabs, added in split order by sg_sum_parts
  int Cout, Cin, w_taps, dw_C, n_ci_tiles, stages;
  int tap_group;    // > 1 (Cin <= 64 only): blockIdx.y owns tap_group taps, tap j in columns [64 j, 64 j + 64) of the N tile
  int ntaps;
  long long dw_split_stride;   // per_image: floats between the dw slabs of consecutive k-splits (= images)
  float* dw;
  float* ws;        // serial: [ksplit][Cout * w_taps * dw_C] partial sums
  sg_wtap_t taps[SG_MAX_TAPS];
};

constexpr int WG_BOX_BYTES = 64 * 128;   // 64 pixels x 64 channels bf16

template <int BN>
struct WgradCfg {
  static constexpr int NB = BN / 64;
  static constexpr int STAGE_BYTES = (2 + NB) * WG_BOX_BYTES;
  static constexpr int STAGES_DEEP = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int STAGES_SHALLOW = (BN >= 192) ? 2 : (BN == 128 ? 3 : 4);      // <= 113 KB: two CTAs per SM
  static constexpr int TM_COLS = BN == 192 ? 256 : BN;
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 + 256; }
};

// Split-K is DETERMINISTIC: every k-split (blockIdx.z) of a dw tile stores its partial sums to its own slab of the
// workspace; sg_sum_parts then adds the slabs in split order (no atomics, nothing waits).
// Tap stacking (tap_group > 1, layers with at most 64 input channels): the N tile holds the SAME 64 input channels of
// tap_group different taps — the dy box is loaded once per k-tile for all of them and the MMAs are N = 64 * tap_group
// wide (an N = 64 MMA occupies the tensor pipe for 54 clocks, an N = 256 one for 128: sg_probe_mma_rate).
template <int BN>
__global__ void __launch_bounds__(192, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ WgradKParams p) {
  using Cfg = WgradCfg<BN>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci_tile = blockIdx.x % p.n_ci_tiles, co_tile = blockIdx.x / p.n_ci_tiles;
  const int TG = p.tap_group;
  const int co0 = co_tile * 128, ci0 = TG > 1 ? 0 : ci_tile * BN;
  const int tap0 = blockIdx.y * TG;                          // first tap of this CTA
  const int ntap = min(TG, p.ntaps - tap0);                  // taps of this CTA that exist
  const sg_wtap_t tp = p.taps[tap0];
  const int kt_begin = blockIdx.z * p.ktiles_per_split;
  const int kt_end = min(kt_begin + p.ktiles_per_split, p.ktiles_total);
  const int iters = max(kt_end - kt_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&empty[s], par ^ 1);
        const int kt = kt_begin + it;
        const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h, ti = kt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, img0 = ti * p.BI;
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        if (TG > 1) {
          // columns of taps that do not exist keep stale shared memory: their accumulator columns are never stored
          mbar_expect_tx(&full[s], (2 + ntap) * WG_BOX_BYTES);
          tma_load_5d(st, &tmA, &full[s], co0, w0 + tp.dwa, h0 + tp.dha, tp.pa, img0);
          tma_load_5d(st + WG_BOX_BYTES, &tmA, &full[s], co0 + 64, w0 + tp.dwa, h0 + tp.dha, tp.pa, img0);
#pragma unroll
          for (int j = 0; j < Cfg::NB; ++j) {
            if (j < ntap) {
              const sg_wtap_t tj = p.taps[tap0 + j];
              tma_load_5d(st + (2 + j) * WG_BOX_BYTES, &tmB, &full[s], 0, w0 + tj.dwb, h0 + tj.dhb, tj.pb, img0);
            }
          }
          continue;
        }
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_5d(st, &tmA, &full[s], co0, w0 + tp.dwa, h0 + tp.dha, tp.pa, img0);
        tma_load_5d(st + WG_BOX_BYTES, &tmA, &full[s], co0 + 64, w0 + tp.dwa, h0 + tp.dha, tp.pa, img0);
#pragma unroll
        for (int j = 0; j < Cfg::NB; ++j)
          tma_load_5d(st + (2 + j) * WG_BOX_BYTES, &tmB, &full[s], ci0 + 64 * j, w0 + tp.dwb, h0 + tp.dhb, tp.pb, img0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      int s = 0;
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it, ++s) {
        if (s == STAGES) { s = 0; par ^= 1; }
        mbar_wait(&full[s], par);
        tc_fence_after();
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        // MN-major SW128: 64 MN-elements contiguous (128 B), 8 K-rows per 1024 B group (SBO),
        // next 64-element MN block one TMA box further (LBO)
        const uint64_t ad = umma_desc_sw128(smem_u32(st), WG_BOX_BYTES, 1024);
        const uint64_t bd = umma_desc_sw128(smem_u32(st + 2 * WG_BOX_BYTES), WG_BOX_BYTES, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(tmem, ad + 128 * k, bd + 128 * k, idesc, (it | k) != 0 ? 1u : 0u);
        mma_commit(&empty[s]);
      }
      mma_commit(accum_full);
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    float* base = p.serial ? p.ws + (long long)blockIdx.z * p.dw_split_stride
                           : p.dw + (long long)blockIdx.z * p.dw_split_stride;       // per_image: one slab per image
    if (iters > 0) {
      mbar_wait(accum_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        // tap stacking: column block j = c0 / 64 is tap tap0 + j, its columns are input channels c0 % 64 ...
        const int tj = TG > 1 ? c0 >> 6 : 0;
        const int cc = TG > 1 ? (c0 & 63) : ci0 + c0;        // first input channel of this chunk
        if (TG > 1 ? tj >= ntap : cc >= p.Cin) break;
        if (cc >= p.Cin) continue;
        const int wtap = p.taps[tap0 + tj].wtap;
        uint32_t raw[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, raw);
        tmem_ld_wait();
        if (co < p.Cout) {
          float* dst = base + ((long long)co * p.w_taps + wtap) * p.dw_C + cc;
          const bool vec = (cc + 32 <= p.Cin) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              __stcg(reinterpret_cast<float4*>(dst + j), make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]),
                                                                     __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cc + j < p.Cin) __stcg(dst + j, __uint_as_float(raw[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TM_COLS>(tmem);
}

// Choose the TMA box (BW, BH, BI) of one M tile.  exact=false (conv): any box with BW*BH*BI <= rows —
// smem rows beyond the box stay stale and their accumulator rows are masked in the epilogue, so
// odd extents (10x10, 65x65, 134x134) do not round up to powers of two.  exact=true (wgrad): the box
// is the reduction dimension and must fill the stage exactly.  Minimises the number of tiles.
void choose_tile(int rows, bool exact, int H, int W, int N, int* BW, int* BH, int* BI) {
  long best = -1;
  int best_rows = 0;
  for (int bw = 1; bw <= rows && bw <= 256; ++bw) {
    if (!exact && bw > W) break;
    for (int bh = 1; bw * bh <= rows && bh <= 256; ++bh) {
      if (!exact && bh > H) break;
      int bi = 1;
      if (bw >= W && bh >= H) bi = rows / (bw * bh);       // several images per tile only if one image fits
      if (bi > 256) bi = 256;
      if (!exact && bi > N) bi = N;
      if (exact && bw * bh * bi != rows) continue;
      long cost = (long)sg_cdiv(W, bw) * sg_cdiv(H, bh) * sg_cdiv(N, bi);
      int used = bw * bh * bi;
      if (best < 0 || cost < best || (cost == best && (bw > *BW || (bw == *BW && used < best_rows)))) {
        best = cost;
        best_rows = used;
        *BW = bw; *BH = bh; *BI = bi;
      }
    }
  }
}

// SG_CONV_DEEP=1: one CTA per SM with a deep ring (the round-1 baseline configuration, kept for A/B runs)
bool deep_pipeline() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SG_CONV_DEEP");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int BN, bool BMN>
int launch_conv_t(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvKParams& kp, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ConvCfg<BN>::smem_bytes(ConvCfg<BN>::STAGES_DEEP));
    if (e != cudaSuccess) return sg_fail(SG_ERR_CUDA, "conv_tc smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  // a grid that does not even fill the SMs once gains nothing from co-residency: give each CTA the deep ring
  const bool one_wave = (long)grid.x * grid.y * grid.z <= 148;
  kp.stages = (deep_pipeline() || one_wave) ? ConvCfg<BN>::STAGES_DEEP : ConvCfg<BN>::STAGES_SHALLOW;
  static_assert(ConvCfg<BN>::STAGES_SHALLOW * A_BYTES >= EPI_SCRATCH_FLOATS * 4, "epilogue scratch must fit the A ring");
  conv_tc_kernel<BN, BMN><<<grid, 192, ConvCfg<BN>::smem_bytes(kp.stages), stream>>>(tmA, tmB, kp);
  SG_CHECK_LAUNCH("sg_conv_tc");
  return SG_OK;
}

template <int BN>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvKParams& kp, dim3 grid, bool bmn, cudaStream_t stream) {
  if (bmn) {
    if constexpr (BN >= 64) return launch_conv_t<BN, true>(tmA, tmB, kp, grid, stream);
  }
  return launch_conv_t<BN, false>(tmA, tmB, kp, grid, stream);
}

// SG_CONV_2CTA=0 disables the CTA-pair kernel (A/B switch)
bool two_cta_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SG_CONV_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <bool BMN>
int launch_conv2(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvKParams& kp, dim3 grid, cudaStream_t stream) {
  constexpr int STAGES = 6;                                          // 32 KB per stage and CTA
  constexpr int SMEM = STAGES * (A_BYTES + 128 * 128) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return sg_fail(SG_ERR_CUDA, "conv_tc2 smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  kp.stages = STAGES;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BMN>, tmA, tmB, kp);
  if (e != cudaSuccess) return sg_fail(SG_ERR_CUDA, "sg_conv_tc (CTA-pair launch): %s", cudaGetErrorString(e));
  SG_CHECK_LAUNCH("sg_conv_tc");
  return SG_OK;
}

// SG_CONV_NO192=1 keeps the N tiles at powers of two (A/B switch for the 192-wide tile)
bool no192() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SG_CONV_NO192");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// SG_WGRAD_TAPSTACK=0 disables tap stacking in sg_wgrad_tc (A/B switch)
bool tap_stack_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SG_WGRAD_TAPSTACK");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BN>
int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, WgradKParams& kp, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WgradCfg<BN>::smem_bytes(WgradCfg<BN>::STAGES_DEEP));
    if (e != cudaSuccess) return sg_fail(SG_ERR_CUDA, "wgrad_tc smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const bool one_wave = (long)grid.x * grid.y * grid.z <= 148;
  kp.stages = (deep_pipeline() || one_wave) ? WgradCfg<BN>::STAGES_DEEP : WgradCfg<BN>::STAGES_SHALLOW;
  wgrad_tc_kernel<BN><<<grid, 192, WgradCfg<BN>::smem_bytes(kp.stages), stream>>>(tmA, tmB, kp);
  SG_CHECK_LAUNCH("sg_wgrad_tc");
  return SG_OK;
}

// tile geometry of a convolution launch (shared by sg_conv_tc and sg_conv_stats_slots)
int conv_geometry(const sg_conv_desc_t* d, ConvKParams& kp) {
  SG_CHECK_ARG(d != nullptr, "sg_conv_tc: null descriptor");
  SG_CHECK_ARG(d->nphases >= 1 && d->nphases <= 4, "sg_conv_tc: nphases %d out of range", d->nphases);
  SG_CHECK_ARG(d->Hout > 0 && d->Wout > 0 && d->x_N > 0 && d->w_Cout > 0, "sg_conv_tc: empty problem");
  memset(&kp, 0, sizeof(kp));
  choose_tile(128, false, d->Hout, d->Wout, d->x_N, &kp.BW, &kp.BH, &kp.BI);
  kp.a_bytes = 128 * kp.BW * kp.BH * kp.BI;
  kp.Hout = d->Hout; kp.Wout = d->Wout;
  kp.tiles_w = sg_cdiv(d->Wout, kp.BW); kp.tiles_h = sg_cdiv(d->Hout, kp.BH);
  kp.n_img = d->x_N;
  kp.slots_per_phase = kp.BI == 1 ? kp.tiles_w * kp.tiles_h : 1;
  kp.stat_slots = d->nphases * kp.slots_per_phase;
  return SG_OK;
}

}  // namespace

extern "C" int sg_conv_stats_slots(const sg_conv_desc_t* d, int* slots) {
  ConvKParams kp;
  SG_CHECK_ARG(slots != nullptr, "sg_conv_stats_slots: null output");
  if (int e = conv_geometry(d, kp)) return e;
  *slots = kp.stat_slots;
  return SG_OK;
}

extern "C" int sg_conv_tc(const sg_conv_desc_t* d, sg_stream_t stream) {
  ConvKParams kp;
  if (int e = conv_geometry(d, kp)) return e;
  SG_CHECK_ARG(d->x && d->w && d->y, "sg_conv_tc: null tensor pointer");
  SG_CHECK_ARG(d->x_C % 8 == 0 && d->w_C % 8 == 0, "sg_conv_tc: channel counts must be multiples of 8 (x_C=%d w_C=%d)", d->x_C, d->w_C);
  SG_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= SG_MAX_TAPS, "sg_conv_tc: ntaps %d out of range", d->ntaps);
  SG_CHECK_ARG(d->y_dtype == 0 || d->y_dtype == 1, "sg_conv_tc: y_dtype must be 0 (f32) or 1 (bf16)");
  SG_CHECK_ARG(d->y_os_c >= 1, "sg_conv_tc: y_os_c (channel stride) must be >= 1");
  for (int i = 0; i < d->nphases; ++i)
    SG_CHECK_ARG(d->phases[i].ntaps >= 1 && d->phases[i].tap_begin >= 0 && d->phases[i].tap_begin + d->phases[i].ntaps <= d->ntaps,
                 "sg_conv_tc: phase %d tap range invalid", i);
  const int img_tiles = sg_cdiv(d->x_N, kp.BI);
  const bool bmn = d->w_mn != 0;
  SG_CHECK_ARG(!bmn || (d->w_img_rows == 0 && d->w_rows > 0 && d->w_col0 >= 0 && d->w_col0 % 8 == 0 &&
                        d->w_col0 + d->w_Cout <= d->w_C),
               "sg_conv_tc: w_mn needs one weight tensor, w_rows > 0 and columns [w_col0, w_col0 + w_Cout) inside w_C");
  const int k_extent = bmn ? d->w_rows : d->w_C;      // extent of the contraction inside the weight tensor
  kp.kblocks = sg_cdiv(d->x_C < k_extent ? d->x_C : k_extent, 64);
  kp.w_col0 = bmn ? d->w_col0 : 0;
  kp.Cout = d->w_Cout;
  SG_CHECK_ARG(d->w_img_rows >= 0, "sg_conv_tc: w_img_rows must be >= 0");
  SG_CHECK_ARG(d->w_img_rows == 0 || kp.BI == 1, "sg_conv_tc: per-image weights need tiles within one image (H*W >= 128)");
  kp.w_img_rows = d->w_img_rows;
  SG_CHECK_ARG(d->w_row0 >= 0 && (d->w_img_rows == 0 ? d->w_row0 == 0 : d->w_row0 + d->w_Cout <= d->w_img_rows),
               "sg_conv_tc: w_row0 / w_Cout outside the per-image weight rows");
  kp.w_row0 = d->w_row0;
  kp.in_h0 = d->in_h0; kp.in_w0 = d->in_w0;
  kp.os_img = d->y_os_img; kp.os_h = d->y_os_h; kp.os_w = d->y_os_w; kp.os_c = d->y_os_c;
  kp.oh_mul = d->oh_mul; kp.ow_mul = d->ow_mul;
  kp.y = d->y; kp.y_dtype = d->y_dtype; kp.bias = d->bias; kp.act = d->act; kp.slope = d->slope; kp.stats = d->stats;
  SG_CHECK_ARG(d->stats == nullptr || d->stats_slots == kp.stat_slots,
               "sg_conv_tc: stats_slots = %d, this launch writes %d partial-sum slots per (image, channel) "
               "(ask sg_conv_stats_slots)", d->stats_slots, kp.stat_slots);
  const int va = d->y_dtype == 1 ? 8 : 4;
  kp.vec_ok = (d->y_os_c == 1) && (d->y_os_img % va == 0) && (d->y_os_h % va == 0) && (d->y_os_w % va == 0) &&
              (((uintptr_t)d->y) % 16 == 0);
  for (int i = 0; i < 4; ++i) kp.phases[i] = d->phases[i < d->nphases ? i : 0];
  for (int i = 0; i < d->ntaps; ++i) kp.taps[i] = d->taps[i];

  CUtensorMap tmA, tmB;
  long long adims[5] = {d->x_C, d->x_W, d->x_H, d->x_P, d->x_N};
  int abox[5] = {64, kp.BW, kp.BH, 1, kp.BI};
  if (int e = make_tmap(&tmA, d->x, 5, adims, abox)) return e;
  // N tile: the MMA time of a CTA grows with BN, the number of waves over the 148 SMs shrinks with it
  int BN = bmn ? 64 : 16;
  if (d->w_Cout > 16) {
    const long m_ctas = (long)kp.tiles_w * kp.tiles_h * img_tiles * d->nphases;
    double best_t = 0;
    const int cands[4] = {256, 192, 128, 64};
    for (int ci = 0; ci < 4; ++ci) {
      const int cand = cands[ci];
      if (cand > 64 && cand / 2 >= d->w_Cout) continue;             // do not tile far beyond Cout
      // 192 only where it wastes less of the last tile than the powers of two (192-channel mask_net, 3 x 64 ...)
      if (cand == 192 && (sg_cdiv(d->w_Cout, 192) * 192 - d->w_Cout >= 64 || no192())) continue;
      long ctas = m_ctas * sg_cdiv(d->w_Cout, cand);
      double t = (double)((ctas + 147) / 148) * (cand + 48);        // +48: per-CTA prologue/epilogue in units of N columns
      if (best_t == 0 || t < best_t) { best_t = t; BN = cand; }
    }
  }
  long long bdims[3] = {d->w_C, d->w_taps,
                        bmn ? (long long)d->w_rows
                            : (d->w_img_rows > 0 ? (long long)d->w_img_rows * d->x_N : (long long)d->w_Cout)};
  int m_tiles = kp.tiles_w * kp.tiles_h * img_tiles;
  // CTA pairs (M = 256 x N = 256 per pair) for long-K, wide-N launches whose pairs fill most of the SMs (the input
  // gradients of the 1024-channel ResnetBlock convolutions: 64 pairs) but whose single-CTA grid is at most two waves
  const long iters = (long)d->ntaps * kp.kblocks;
  const int n_tiles256 = sg_cdiv(d->w_Cout, 256);
  const int pairs = ((m_tiles + 1) / 2) * n_tiles256 * d->nphases;
  if (two_cta_enabled() && BN == 256 && d->w_Cout >= 256 && d->w_img_rows == 0 && m_tiles >= 2 && iters >= 64 &&
      2 * pairs >= 96 && 2 * pairs <= 2 * 148) {
    m_tiles = (m_tiles + 1) & ~1;       // an odd tail tile gets a fully masked partner
    int bbox2[3] = {64, 1, bmn ? 64 : 128};
    if (int e = make_tmap(&tmB, d->w, 3, bdims, bbox2)) return e;
    dim3 grid2(m_tiles, n_tiles256, d->nphases);
    return bmn ? launch_conv2<true>(tmA, tmB, kp, grid2, stream) : launch_conv2<false>(tmA, tmB, kp, grid2, stream);
  }
  int bbox[3] = {64, 1, bmn ? 64 : BN};
  if (int e = make_tmap(&tmB, d->w, 3, bdims, bbox)) return e;
  dim3 grid(m_tiles, sg_cdiv(d->w_Cout, BN), d->nphases);
  switch (BN) {
    case 256: return launch_conv<256>(tmA, tmB, kp, grid, bmn, stream);
    case 192: return launch_conv<192>(tmA, tmB, kp, grid, bmn, stream);
    case 128: return launch_conv<128>(tmA, tmB, kp, grid, bmn, stream);
    case 64: return launch_conv<64>(tmA, tmB, kp, grid, bmn, stream);
    default: return launch_conv<16>(tmA, tmB, kp, grid, false, stream);
  }
}

extern "C" int sg_wgrad_tc(const sg_wgrad_desc_t* d, sg_stream_t stream) {
  SG_CHECK_ARG(d != nullptr && d->dy && d->x && d->dw, "sg_wgrad_tc: null pointer");
  SG_CHECK_ARG(d->dy_C % 8 == 0 && d->x_C % 8 == 0, "sg_wgrad_tc: channel counts must be multiples of 8");
  SG_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= SG_MAX_TAPS, "sg_wgrad_tc: ntaps %d out of range", d->ntaps);
  SG_CHECK_ARG(d->Hred > 0 && d->Wred > 0 && d->N > 0 && d->Cout > 0 && d->Cin > 0, "sg_wgrad_tc: empty problem");
  WgradKParams kp;
  memset(&kp, 0, sizeof(kp));
  choose_tile(64, true, d->Hred, d->Wred, d->N, &kp.BW, &kp.BH, &kp.BI);
  kp.tiles_w = sg_cdiv(d->Wred, kp.BW); kp.tiles_h = sg_cdiv(d->Hred, kp.BH);
  kp.ktiles_total = kp.tiles_w * kp.tiles_h * sg_cdiv(d->N, kp.BI);
  kp.Cout = d->Cout; kp.Cin = d->Cin; kp.w_taps = d->w_taps; kp.dw_C = d->dw_C; kp.dw = d->dw;
  for (int i = 0; i < d->ntaps; ++i) kp.taps[i] = d->taps[i];
  int BN = (d->Cin > 128 && d->Cin <= 192 && !no192()) ? 192 : (d->Cin > 128 ? 256 : (d->Cin > 64 ? 128 : 64));
  // tap stacking for layers with at most 64 input channels: taps whose dy box is the same share an N tile
  kp.tap_group = 1;
  kp.ntaps = d->ntaps;
  if (d->Cin <= 64 && d->ntaps >= 3 && tap_stack_enabled()) {
    bool same_a = true;
    for (int i = 1; i < d->ntaps; ++i)
      same_a = same_a && d->taps[i].dha == d->taps[0].dha && d->taps[i].dwa == d->taps[0].dwa && d->taps[i].pa == d->taps[0].pa;
    if (same_a) {
      kp.tap_group = (d->ntaps % 4 != 0 && d->ntaps % 3 == 0 && !no192()) ? 3 : 4;
      BN = 64 * kp.tap_group;
    }
  }
  const int tap_groups = sg_cdiv(d->ntaps, kp.tap_group);
  kp.n_ci_tiles = kp.tap_group > 1 ? 1 : sg_cdiv(d->Cin, BN);
  const int co_tiles = sg_cdiv(d->Cout, 128);
  int ksplit = d->ksplit;
  const long base_ctas = (long)co_tiles * kp.n_ci_tiles * tap_groups;
  if (d->per_image) {
    // one k-split per image: dw is [N][Cout][w_taps][dw_C], every slab is owned by the CTAs of one image
    SG_CHECK_ARG(kp.BI == 1, "sg_wgrad_tc: per_image needs reduction tiles within one image (Hred*Wred >= 64)");
    ksplit = d->N;
    kp.dw_split_stride = (long long)d->Cout * d->w_taps * d->dw_C;
  } else if (ksplit <= 0) {
    // fill the 2 x 148 co-resident CTA slots once
    ksplit = (int)((2 * 148 + base_ctas - 1) / base_ctas);
    if (ksplit > kp.ktiles_total) ksplit = kp.ktiles_total;
    if (ksplit < 1) ksplit = 1;
    // keep at least 8 k-tiles per split so the pipeline prologue amortises
    while (ksplit > 1 && kp.ktiles_total / ksplit < 8) --ksplit;
  }
  const long long dw_floats = (long long)d->Cout * d->w_taps * d->dw_C;
  if (!d->per_image && ksplit > 1) {
    // the splits leave their partial sums in the caller's workspace: as many splits as it (and the tickets) allow
    const long long cap = d->ws != nullptr ? d->ws_floats / dw_floats : 1;
    if (ksplit > cap) ksplit = cap < 1 ? 1 : (int)cap;
  }
  kp.ktiles_per_split = sg_cdiv(kp.ktiles_total, ksplit);
  ksplit = sg_cdiv(kp.ktiles_total, kp.ktiles_per_split);
  kp.serial = ksplit > 1 && !d->per_image;
  if (kp.serial) {
    kp.ws = d->ws;
    kp.dw_split_stride = dw_floats;
  }
  CUtensorMap tmA, tmB;
  long long adims[5] = {d->dy_C, d->dy_W, d->dy_H, d->dy_P, d->N};
  long long bdims[5] = {d->x_C, d->x_W, d->x_H, d->x_P, d->N};
  int box[5] = {64, kp.BW, kp.BH, 1, kp.BI};
  if (int e = make_tmap(&tmA, d->dy, 5, adims, box)) return e;
  if (int e = make_tmap(&tmB, d->x, 5, bdims, box)) return e;
  dim3 grid(co_tiles * kp.n_ci_tiles, tap_groups, ksplit);
  int e;
  switch (BN) {
    case 256: e = launch_wgrad<256>(tmA, tmB, kp, grid, stream); break;
    case 192: e = launch_wgrad<192>(tmA, tmB, kp, grid, stream); break;
    case 128: e = launch_wgrad<128>(tmA, tmB, kp, grid, stream); break;
    default: e = launch_wgrad<64>(tmA, tmB, kp, grid, stream); break;
  }
  if (e != SG_OK || !kp.serial) return e;
  return sg_sum_parts(d->ws, dw_floats, ksplit, dw_floats, d->dw, stream, "sg_wgrad_tc(split sum)");
}


// ------------------------------------------------------------------------------------------------
// PROBE (round-2 groundwork, not on any product path): can one shared-memory HALO tile feed all taps of a 3x3
// convolution through tap-shifted UMMA descriptors?  The halo of a 16x8-pixel tile (18 x 10 pixels x 64 channels,
// one TMA box, SWIZZLE_128B) is loaded ONCE; tap (dh, dw) then reads the 128 output pixels' operand rows as 16
// groups of 8 consecutive pixels starting at row (dh * 10 + dw), group stride (SBO) = 10 rows = 1280 B — i.e. a
// start address that is 128-byte but not 1024-byte aligned and an SBO that is not a multiple of the swizzle period.
// mode 0: descriptor base_offset = 0; mode 1: base_offset = (start >> 7) & 7 (the field PTX defines for start
// addresses that are not aligned to the swizzle repeat).  If either mode reproduces the reference convolution, the
// 7x7 / 3x3 convolutions can stop re-reading their input tile once per tap (49x / 9x less L2 -> SM traffic for A).
namespace {

__global__ void __launch_bounds__(128) probe_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          float* y, int mode) {
  constexpr int HALO_ROWS = 18 * 10, A_HALO = HALO_ROWS * 128, B_TAP = 64 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((A_HALO + 1023) / 1024) * 1024;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + 9 * B_TAP);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, A_HALO + 9 * B_TAP);
    tma_load_5d(sA, &tmA, full, 0, 0, 0, 0, 0);
    for (int tap = 0; tap < 9; ++tap) tma_load_3d(sB + tap * B_TAP, &tmB, full, 0, tap, 0);
    mbar_wait(full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    for (int tap = 0; tap < 9; ++tap) {
      const int dh = tap / 3, dw = tap % 3;
      const uint32_t start = smem_u32(sA + (dh * 10 + dw) * 128);
      uint64_t ad = umma_desc_sw128(start, 16, 10 * 128);
      if (mode == 1) ad |= (uint64_t)((start >> 7) & 7u) << 49;
      const uint64_t bd = umma_desc_sw128(smem_u32(sB + tap * B_TAP), 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
    }
    mma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t raw[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, raw);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) y[row * 64 + c0 + j] = __uint_as_float(raw[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace

// PROBE: issue rate of tcgen05.mma (M = 128, K = 16, N = BN) from one thread over fixed shared-memory operands —
// the floor under every kernel of this file.  mode 0: the four K steps of one 64-wide k-block, one accumulator;
// mode 1: the same K step every time; mode 2: four accumulators round robin.  out[0] = clocks per MMA.
namespace {
template <int BN>
__global__ void __launch_bounds__(128) probe_mma_rate_kernel(int iters, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * A_BYTES;              // A region: 32 KB (modes 3 / 4 read 16 groups 1792 B apart)
  uint64_t* done = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  for (int i = threadIdx.x; i < (2 * A_BYTES + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(tmem_slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    // modes 3 / 4: the A operand as a tap-shifted view of a halo tile (start 3 rows into a swizzle atom / aligned
    // start, 8-row groups 14 rows apart): groups that straddle two 1024-byte swizzle atoms
    const uint64_t ad = mode == 3 ? umma_desc_sw128(smem_u32(sA) + 3 * 128, 16, 14 * 128)
                      : mode == 4 ? umma_desc_sw128(smem_u32(sA), 16, 14 * 128)
                                  : umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t bd = umma_desc_sw128(smem_u32(sB), 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kk = mode == 1 ? 0 : k;   // (modes 3 / 4 step through K like mode 0)
        const uint32_t acc = mode == 2 ? (uint32_t)(((it * 4 + k) & (BN > 128 ? 1 : 3)) * BN) : 0u;
        mma_bf16(tmem + acc, ad + 2 * kk, bd + 2 * kk, idesc, 1u);
      }
    }
    mma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    out[0] = (float)(t1 - t0) / (4.f * iters);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
}  // namespace

extern "C" int sg_probe_mma_rate(int BN, int iters, int mode, float* out, sg_stream_t stream) {
  SG_CHECK_ARG(out && iters > 0 && mode >= 0 && mode <= 4, "sg_probe_mma_rate: bad arguments");
  const int smem = 2 * A_BYTES + 256 * 128 + 1024 + 256;
#define SG_PROBE_RATE(N)                                                                                           \
  do {                                                                                                             \
    cudaFuncSetAttribute(probe_mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);             \
    probe_mma_rate_kernel<N><<<1, 128, smem, stream>>>(iters, mode, out);                                          \
  } while (0)
  switch (BN) {
    case 16: SG_PROBE_RATE(16); break;
    case 64: SG_PROBE_RATE(64); break;
    case 128: SG_PROBE_RATE(128); break;
    case 256: SG_PROBE_RATE(256); break;
    default: return sg_fail(SG_ERR_ARG, "sg_probe_mma_rate: BN must be 16, 64, 128 or 256");
  }
#undef SG_PROBE_RATE
  SG_CHECK_LAUNCH("sg_probe_mma_rate");
  return SG_OK;
}

/* x: bf16 [1][1][18][10][64] (reflection / zero halo already materialised), w: bf16 [64][9][64], y: f32 [128][64]
 * (pixel = h * 8 + w of the 16 x 8 tile, then output channel). */
extern "C" int sg_probe_shifted_desc(const void* x, const void* w, float* y, int mode, sg_stream_t stream) {
  SG_CHECK_ARG(x && w && y && (mode == 0 || mode == 1), "sg_probe_shifted_desc: bad arguments");
  CUtensorMap tmA, tmB;
  long long adims[5] = {64, 10, 18, 1, 1};
  int abox[5] = {64, 10, 18, 1, 1};
  if (int e = make_tmap(&tmA, x, 5, adims, abox)) return e;
  long long bdims[3] = {64, 9, 64};
  int bbox[3] = {64, 1, 64};
  if (int e = make_tmap(&tmB, w, 3, bdims, bbox)) return e;
  const int smem = 24 * 1024 + 9 * 8192 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(probe_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return sg_fail(SG_ERR_CUDA, "probe smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  probe_shift_kernel<<<1, 128, smem, stream>>>(tmA, tmB, y, mode);
  SG_CHECK_LAUNCH("sg_probe_shifted_desc");
  return SG_OK;
}
