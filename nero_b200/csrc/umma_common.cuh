// Device helpers shared by the tcgen05 GEMM kernels (k_umma_linear.cu, k_umma_chain.cu).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#ifndef NERO_PF_MODE
#define NERO_PF_MODE 2   // aux-operand L2 prefetch: 0 off, 1 per-sector prefetch.global.L2, 2 bulk prefetch per row
#endif

namespace nero {

// compile-time epilogue kinds
enum EpiKind : int { EK_BIAS_SOFTPLUS = 0, EK_BIAS_RELU = 1, EK_BIAS_GENERIC = 2, EK_DACT_SOFTPLUS = 3, EK_DACT_RELU = 4,
                     EK_DACT_NONE = 5, EK_TANGENT = 6 };

constexpr int kStagePitch = 20;                  // floats per row of the per-warp transpose buffer (16 cols + pad)
constexpr uint32_t kStageWarpBytes = 32 * kStagePitch * 4;

// fast softplus(beta=100): max error ~1e-9 absolute (ex2/lg2 approximations, 1+t rounding)
__device__ __forceinline__ float softplus100_fast(float a) {
  const float z = 100.0f * a;
  const float t = exp2f(fminf(z, 20.0f) * 1.4426950408889634f);      // MUFU.EX2
  const float h = __log2f(1.0f + t) * (0.6931471805599453f * 0.01f);  // MUFU.LG2
  return z > 20.0f ? a : h;
}
__device__ __forceinline__ float dsoftplus100_from_h_fast(float h) {
  const float z = 100.0f * h;
  return z > 20.0f ? 1.0f : 1.0f - exp2f(-z * 1.4426950408889634f);
}

// batched versions over a 16-element register block: the three stages are written as separate loops so that the
// sixteen MUFU chains are independent and issue back to back (ptxas otherwise serialises them through one register)
__device__ __forceinline__ float ex2_ftz(float x) {   // MUFU.EX2 without the denormal-range fix-up of exp2f()
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softplus100_fast16(float* x) {
  float t[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = ex2_ftz(fminf(x[j], 0.2f) * (100.0f * 1.4426950408889634f));
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = lg2_ftz(1.0f + t[j]);
  // softplus(x) >= x everywhere and the clamped branch saturates at softplus(0.2) = 0.2 + 2e-11, so the select
  // "x > 0.2 ? x : ..." is a plain max (one FMNMX instead of FSETP + FSEL)
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], t[j] * (0.6931471805599453f * 0.01f));
}
// sigma(100 a) from h = softplus(a) (possibly stored scaled by 1/hscale): 1 - 2^(-100*log2e*hscale*h); for
// 100*hscale*h > 20 the exponential is < 2.1e-9, below fp32 resolution of 1 - t, so no branch is needed
__device__ __forceinline__ void dsoftplus100_from_h_fast16(float* h, float hscale) {
  const float k = -(100.0f * 1.4426950408889634f) * hscale;
#pragma unroll
  for (int j = 0; j < 16; ++j) h[j] = 1.0f - ex2_ftz(k * h[j]);
}

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const uint32_t hb = *reinterpret_cast<uint32_t*>(&h);
  const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  hi = hb;
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// 16 columns of one TMEM lane group
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// coalesced load of a [32 rows x 16 cols] block (row-major global, leading dim ld) into registers of the
// "lane = row" layout, through the warp's transpose buffer.  Rows >= rows_valid / cols >= cols_valid read as 0.
// Full, 16-byte aligned blocks take a branch-free fast path (4 LDG.128 + 4 STS.128 + 4 LDS.128).
__device__ __forceinline__ void load_block16(const float* __restrict__ g, int ld, int rows_valid, int cols_valid, bool vec,
                                             float* stage, int lane, float* r) {
  float* sdst = stage + (lane >> 2) * kStagePitch + (lane & 3) * 4;
  if (vec && rows_valid >= 32 && cols_valid >= 16) {
    const float* src = g + size_t(lane >> 2) * ld + (lane & 3) * 4;
    const size_t ld8 = size_t(ld) * 8;
    float4 x0 = *reinterpret_cast<const float4*>(src);
    float4 x1 = *reinterpret_cast<const float4*>(src + ld8);
    float4 x2 = *reinterpret_cast<const float4*>(src + 2 * ld8);
    float4 x3 = *reinterpret_cast<const float4*>(src + 3 * ld8);
    *reinterpret_cast<float4*>(sdst) = x0;
    *reinterpret_cast<float4*>(sdst + 8 * kStagePitch) = x1;
    *reinterpret_cast<float4*>(sdst + 16 * kStagePitch) = x2;
    *reinterpret_cast<float4*>(sdst + 24 * kStagePitch) = x3;
  } else {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), q = (lane & 3) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows_valid) {
        const float* p = g + size_t(row) * ld + q;
        if (vec && q + 3 < cols_valid) x = *reinterpret_cast<const float4*>(p);
        else {
          if (q < cols_valid) x.x = p[0];
          if (q + 1 < cols_valid) x.y = p[1];
          if (q + 2 < cols_valid) x.z = p[2];
          if (q + 3 < cols_valid) x.w = p[3];
        }
      }
      *reinterpret_cast<float4*>(stage + row * kStagePitch + q) = x;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 x = *reinterpret_cast<const float4*>(stage + lane * kStagePitch + j * 4);
    r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w;
  }
  __syncwarp();
}
// split version of load_block16 for software pipelining: issue the coalesced global loads now ...
__device__ __forceinline__ void issue_block16(const float* __restrict__ g, int ld, int rows_valid, int cols_valid, bool vec, int lane,
                                              float4 (&x)[4]) {
  if (vec && rows_valid >= 32 && cols_valid >= 16) {
    const float* src = g + size_t(lane >> 2) * ld + (lane & 3) * 4;
    const size_t ld8 = size_t(ld) * 8;
#pragma unroll
    for (int it = 0; it < 4; ++it) x[it] = __ldg(reinterpret_cast<const float4*>(src + it * ld8));
  } else {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), q = (lane & 3) * 4;
      x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows_valid) {
        const float* p = g + size_t(row) * ld + q;
        if (vec && q + 3 < cols_valid) x[it] = *reinterpret_cast<const float4*>(p);
        else {
          if (q < cols_valid) x[it].x = p[0];
          if (q + 1 < cols_valid) x[it].y = p[1];
          if (q + 2 < cols_valid) x[it].z = p[2];
          if (q + 3 < cols_valid) x[it].w = p[3];
        }
      }
    }
  }
}
// ... and transpose them into the "lane = row" register layout later
__device__ __forceinline__ void finish_block16(const float4 (&x)[4], float* stage, int lane, float* r) {
  float* sdst = stage + (lane >> 2) * kStagePitch + (lane & 3) * 4;
#pragma unroll
  for (int it = 0; it < 4; ++it) *reinterpret_cast<float4*>(sdst + it * 8 * kStagePitch) = x[it];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 y = *reinterpret_cast<const float4*>(stage + lane * kStagePitch + j * 4);
    r[4 * j] = y.x; r[4 * j + 1] = y.y; r[4 * j + 2] = y.z; r[4 * j + 3] = y.w;
  }
  __syncwarp();
}
// the reverse: registers ("lane = row") -> coalesced global store of columns [0, cols_valid)
__device__ __forceinline__ void store_block16(float* __restrict__ g, int ld, int rows_valid, int cols_valid, bool vec,
                                              float* stage, int lane, const float* r) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(stage + lane * kStagePitch + j * 4) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  __syncwarp();
  if (vec && rows_valid >= 32 && cols_valid >= 16) {
    const float* ssrc = stage + (lane >> 2) * kStagePitch + (lane & 3) * 4;
    float* dst = g + size_t(lane >> 2) * ld + (lane & 3) * 4;
    const size_t ld8 = size_t(ld) * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(ssrc);
    const float4 x1 = *reinterpret_cast<const float4*>(ssrc + 8 * kStagePitch);
    const float4 x2 = *reinterpret_cast<const float4*>(ssrc + 16 * kStagePitch);
    const float4 x3 = *reinterpret_cast<const float4*>(ssrc + 24 * kStagePitch);
    *reinterpret_cast<float4*>(dst) = x0;
    *reinterpret_cast<float4*>(dst + ld8) = x1;
    *reinterpret_cast<float4*>(dst + 2 * ld8) = x2;
    *reinterpret_cast<float4*>(dst + 3 * ld8) = x3;
  } else {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), q = (lane & 3) * 4;
      if (row < rows_valid && q < cols_valid) {
        const float4 x = *reinterpret_cast<const float4*>(stage + row * kStagePitch + q);
        float* p = g + size_t(row) * ld + q;
        if (vec && q + 3 < cols_valid) *reinterpret_cast<float4*>(p) = x;
        else {
          p[0] = x.x;
          if (q + 1 < cols_valid) p[1] = x.y;
          if (q + 2 < cols_valid) p[2] = x.z;
          if (q + 3 < cols_valid) p[3] = x.w;
        }
      }
    }
  }
  __syncwarp();
}
// prefetch the [32 rows x ncols] window of an aux matrix into L2: one bulk prefetch (UBLKPF) per row segment
__device__ __forceinline__ void prefetch_rows_l2(const float* __restrict__ g, int ld, int rows_valid, int ncols, int lane) {
  if (!g || rows_valid <= 0 || ncols <= 0) return;
#if NERO_PF_MODE == 2
  const uint32_t bytes = uint32_t(((ncols * 4) + 15) & ~15);
  if (lane < rows_valid) {
    const float* p = g + size_t(lane) * ld;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
  }
#elif NERO_PF_MODE == 1
  const int sectors = (ncols * 4 + 31) / 32;            // one prefetch per 32-byte sector
  for (int i = lane; i < rows_valid * sectors; i += 32) {
    const int r = i / sectors, l = i % sectors;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(g + size_t(r) * ld + l * 8));
  }
#endif
}
__device__ __forceinline__ bool vec_ok(const float* p, int ld) {
  return ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}


}  // namespace nero
