// Multi-tensor Adam step that also refreshes the bf16 tensor-core operand of every weight it updates.
//
// Replaces torch.optim.Adam.step (trainer.py:60,80,106,133 of the reference; fused multi-tensor kernel in
// PyTorch) PLUS the per-step re-packing of the f32 masters into bf16 [Cout][taps][Cin_p] operands: one pass
// reads p, g, m, v (16 B / parameter) and writes p, m, v and the bf16 copy (14 B / parameter).
// HBM-bound: 30 B per parameter, 197.6 M parameters per iteration at cfg-2.
//
// Arithmetic (Adam, no weight decay, no amsgrad — the reference's configuration):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with t read from the parameter's device-side step counter (already incremented by the caller), so the
// launch can sit inside a captured CUDA graph.
#include "common.cuh"
#include "../../include/sg_b200.h"

namespace {

typedef __nv_bfloat16 bf16;

// One launch covers up to 16 blocks per SM (148 x 16 = 2368 chunks of 16 k elements = 38.8 M parameters): enough
// resident warps to keep ~100 KB of loads in flight per SM.  The tensor / chunk tables travel as kernel parameters
// (19 KB; CUDA >= 12.1 allows 32 KB on sm_70+), so a launch needs no host-to-device copy and is graph-capturable.
constexpr int ADAM_MAX_TENSORS = 96;
constexpr int ADAM_MAX_BLOCKS = 2368;
constexpr int ADAM_CHUNK = 16384;      // elements per block
constexpr int ADAM_THREADS = 256;

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* wk;            // bf16 operand [rows][Cp] of the master viewed as [rows][C], or nullptr
  const float* step;   // device-side step counter (float, as torch's capturable Adam keeps it)
  long long n;
  int C, Cp;
  int vec;             // every pointer 16-byte aligned (wk: 8), n % 4 == 0, C == Cp  -> float4 path
  int pad_;
};

struct AdamArgs {
  AdamTensor t[ADAM_MAX_TENSORS];
  int blk_chunk[ADAM_MAX_BLOCKS];
  unsigned char blk_tensor[ADAM_MAX_BLOCKS];
  double lr, b1d, b2d;        // bias corrections are evaluated in double (as torch's fused kernel does)
  float b1, b2, omb1, omb2, eps;   // omb = 1 - beta, rounded from the double difference
};

struct AdamC {
  float b1, b2, omb1, omb2, step_size, inv_bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamC& c) {
  m = c.b1 * m + c.omb1 * g;
  v = c.b2 * v + c.omb2 * g * g;
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  p -= c.step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_pack_kernel(const __grid_constant__ AdamArgs a) {
  const AdamTensor& t = a.t[a.blk_tensor[blockIdx.x]];
  const long long base = (long long)a.blk_chunk[blockIdx.x] * ADAM_CHUNK;
  const long long left = t.n - base;
  const int n = (int)(left < ADAM_CHUNK ? left : ADAM_CHUNK);
  const double step = (double)__ldg(t.step);
  AdamC c;
  c.b1 = a.b1; c.b2 = a.b2; c.omb1 = a.omb1; c.omb2 = a.omb2; c.eps = a.eps;
  c.step_size = (float)(a.lr / (1.0 - pow(a.b1d, step)));
  c.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow(a.b2d, step)));
  float* p = t.p + base;
  const float* g = t.g + base;
  float* m = t.m + base;
  float* v = t.v + base;
  if (t.vec) {
    bf16* wk = t.wk ? t.wk + base : nullptr;
    // two float4 groups per trip: all eight loads are issued before the first dependent store
    constexpr int STRIDE = ADAM_THREADS * 4;
    for (int i = threadIdx.x * 4; i < n; i += 2 * STRIDE) {
      const bool two = i + STRIDE < n;
      const int j = two ? i + STRIDE : i;
      float4 P0 = *reinterpret_cast<const float4*>(p + i), P1 = *reinterpret_cast<const float4*>(p + j);
      const float4 G0 = __ldcs(reinterpret_cast<const float4*>(g + i)), G1 = __ldcs(reinterpret_cast<const float4*>(g + j));
      float4 M0 = *reinterpret_cast<const float4*>(m + i), M1 = *reinterpret_cast<const float4*>(m + j);
      float4 V0 = *reinterpret_cast<const float4*>(v + i), V1 = *reinterpret_cast<const float4*>(v + j);
      adam_one(P0.x, G0.x, M0.x, V0.x, c);
      adam_one(P0.y, G0.y, M0.y, V0.y, c);
      adam_one(P0.z, G0.z, M0.z, V0.z, c);
      adam_one(P0.w, G0.w, M0.w, V0.w, c);
      *reinterpret_cast<float4*>(p + i) = P0;
      *reinterpret_cast<float4*>(m + i) = M0;
      *reinterpret_cast<float4*>(v + i) = V0;
      if (wk) {
        __align__(8) __nv_bfloat162 pk[2] = {__floats2bfloat162_rn(P0.x, P0.y), __floats2bfloat162_rn(P0.z, P0.w)};
        *reinterpret_cast<uint2*>(wk + i) = *reinterpret_cast<uint2*>(pk);
      }
      if (two) {
        adam_one(P1.x, G1.x, M1.x, V1.x, c);
        adam_one(P1.y, G1.y, M1.y, V1.y, c);
        adam_one(P1.z, G1.z, M1.z, V1.z, c);
        adam_one(P1.w, G1.w, M1.w, V1.w, c);
        *reinterpret_cast<float4*>(p + j) = P1;
        *reinterpret_cast<float4*>(m + j) = M1;
        *reinterpret_cast<float4*>(v + j) = V1;
        if (wk) {
          __align__(8) __nv_bfloat162 pk[2] = {__floats2bfloat162_rn(P1.x, P1.y), __floats2bfloat162_rn(P1.z, P1.w)};
          *reinterpret_cast<uint2*>(wk + j) = *reinterpret_cast<uint2*>(pk);
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += ADAM_THREADS) {
      float P = p[i], M = m[i], V = v[i];
      adam_one(P, g[i], M, V, c);
      p[i] = P;
      m[i] = M;
      v[i] = V;
      if (t.wk) {
        const long long e = base + i;
        const long long row = e / t.C;
        t.wk[row * t.Cp + (e - row * t.C)] = __float2bfloat16(P);     // pad columns [C, Cp) stay zero from the first pack
      }
    }
  }
}

}  // namespace

extern "C" int sg_adam_pack(int n_tensors, void* const* p, void* const* g, void* const* m, void* const* v, void* const* wk,
                            void* const* step, const long long* numel, const int* C, const int* Cp, double lr, double beta1,
                            double beta2, double eps, sg_stream_t stream) {
  SG_CHECK_ARG(n_tensors >= 0 && p && g && m && v && wk && step && numel && C && Cp, "sg_adam_pack: null argument array");
  SG_CHECK_ARG(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "sg_adam_pack: bad hyper-parameters");
  static thread_local AdamArgs a;     // 19 KB: kept off the stack, one per calling thread (the entry points stay
                                      // re-entrant across threads and streams)
  a.lr = lr; a.b1d = beta1; a.b2d = beta2;
  a.b1 = (float)beta1; a.b2 = (float)beta2; a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2); a.eps = (float)eps;
  int nt = 0, nb = 0;
  auto flush = [&]() -> int {
    if (nb == 0) { nt = 0; return SG_OK; }
    adam_pack_kernel<<<nb, ADAM_THREADS, 0, stream>>>(a);
    SG_CHECK_LAUNCH("sg_adam_pack");
    nt = nb = 0;
    return SG_OK;
  };
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] == 0) continue;
    SG_CHECK_ARG(p[i] && g[i] && m[i] && v[i] && step[i] && numel[i] > 0, "sg_adam_pack: tensor %d has a null pointer", i);
    SG_CHECK_ARG(wk[i] == nullptr || (C[i] > 0 && Cp[i] >= C[i] && numel[i] % C[i] == 0),
                 "sg_adam_pack: tensor %d: operand row length %d / pitch %d does not fit %lld elements", i, C[i], Cp[i], numel[i]);
    const long long chunks = (numel[i] + ADAM_CHUNK - 1) / ADAM_CHUNK;
    long long c = 0;
    while (c < chunks) {
      if (nt == ADAM_MAX_TENSORS || nb == ADAM_MAX_BLOCKS)
        if (int e = flush()) return e;
      AdamTensor& t = a.t[nt];
      t.p = (float*)p[i]; t.g = (const float*)g[i]; t.m = (float*)m[i]; t.v = (float*)v[i];
      t.wk = (bf16*)wk[i]; t.step = (const float*)step[i];
      t.n = numel[i]; t.C = wk[i] ? C[i] : 1; t.Cp = wk[i] ? Cp[i] : 1;
      const uintptr_t al = (uintptr_t)p[i] | (uintptr_t)g[i] | (uintptr_t)m[i] | (uintptr_t)v[i];
      t.vec = (al % 16 == 0) && (numel[i] % 4 == 0) && (wk[i] == nullptr || (C[i] == Cp[i] && (uintptr_t)wk[i] % 8 == 0));
      t.pad_ = 0;
      while (c < chunks && nb < ADAM_MAX_BLOCKS) {
        a.blk_tensor[nb] = (unsigned char)nt;
        a.blk_chunk[nb] = (int)c;
        ++nb; ++c;
      }
      ++nt;
    }
  }
  return flush();
}
