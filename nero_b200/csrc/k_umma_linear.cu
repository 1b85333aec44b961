// tcgen05 (UMMA) linear-layer kernel: OUT[M, N] = epilogue( A[M, K] * W[N, K]^T ).
//
//  * A is fp32 row-major in HBM.  Producer warps load it (coalesced float4), split every value into
//    bf16 hi + bf16 lo (error-compensated "split-bf16": x ~= hi + lo, ~2^-17 relative) and store both planes
//    into the K-major SWIZZLE_128B shared-memory operand layout the tensor core reads.
//  * W is pre-split and pre-swizzled once per step by nero_prep_weight (k_weights.cu) into the exact
//    shared-memory image, so a K-chunk of it is ONE cp.async.bulk (UBLKCP) from L2, completing on an mbarrier.
//  * One elected thread issues tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 in TMEM), three MMAs per k-step:
//    A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (the lo*lo term, ~2^-18 relative, is dropped).
//  * The fp32 accumulator tile (128 x N) lives in TMEM, double buffered (2 x 256 columns), so the epilogue
//    warps (tcgen05.ld -> bias/activation/derivative products -> HBM) overlap the next tile's MMAs.
//  * Persistent CTAs (one per SM), tiles of 128 rows, row count read from device memory (no host sync for the
//    data-dependent number of live samples).
//
// Reference op sites replaced: every nn.Linear (+weight_norm, +Softplus/ReLU/Sigmoid/exp) of
// network/field.py:130-147 (SDFNetwork), :258-283 (NeRFNetwork), :310-346 (make_predictor), and the
// autograd-generated input-gradient GEMMs of their backward / double-backward.
#include "common.cuh"
#include "ptx.cuh"

namespace nero {

struct LinearParams {
  const float* A; int lda; int k_valid;
  const uint8_t* wimg; int n_pad; int k_chunks;
  const float* bias; int n_bias;
  float* out; int ldo; int ncol_out; float oscale;
  int mode; int act; float act_param;
  const float* H; int ldh; float hscale; int dact;
  const float* V; int ldv;
  float* out2; int ldo2;
  const float* addend; int ldadd;
  int ncol_main; float* tail; int ldt;
  const int* m_ptr; int m_cap;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 4, kProdWarps = 4;
constexpr int kLinearThreads = (kEpiWarps + kProdWarps + 2) * 32;  // + MMA warp + W-loader warp
constexpr uint32_t kABytes = BM * 128;                             // one bf16 plane of an A stage (16 KB)

template <int NPAD> struct LinCfg {
  static constexpr uint32_t b_plane = NPAD * 128;                  // one bf16 plane of a W stage
  static constexpr uint32_t stage_bytes = 2 * kABytes + 2 * b_plane;
  static constexpr int stages = (stage_bytes * 4 <= 200 * 1024) ? 4 : (stage_bytes * 3 <= 200 * 1024) ? 3 : 2;
  static constexpr uint32_t smem_bytes = stages * stage_bytes + 1024 /*align slack*/ + 1024 /*bias*/ + 256 /*barriers*/;
};

template <int NPAD>
__global__ void __launch_bounds__(kLinearThreads, 1) umma_linear_kernel(const LinearParams p) {
  using Cfg = LinCfg<NPAD>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes + 1024);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;      // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = p.m_ptr ? *p.m_ptr : p.m_cap;
  if (M > p.m_cap) M = p.m_cap;
  const int num_tiles = (M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kProdWarps + 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kEpiWarps + kProdWarps + 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < NPAD; i += blockDim.x) s_bias[i] = (p.bias && i < p.n_bias) ? p.bias[i] : 0.0f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ============================== epilogue warps: TMEM -> registers -> HBM
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const int row = tile * BM + warp * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(acc * 256);
#pragma unroll 1
      for (int cc = 0; cc < (NPAD + 31) / 32; ++cc) {
        float v[32];
        tmem_ld32(taddr + cc * 32, v);
        tmem_ld_wait();
        const int c0 = cc * 32;
        if (row_ok && c0 < p.ncol_out) {
          float o1[32];
          float o2[32];
          const float* hrow = p.H ? p.H + size_t(row) * p.ldh + c0 : nullptr;
          const float* vrow = p.V ? p.V + size_t(row) * p.ldv + c0 : nullptr;
          const float* arow = p.addend ? p.addend + size_t(row) * p.ldadd + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c0 + j;
            const float a = v[j];
            float r = 0.0f, r2 = 0.0f;
            if (col < p.ncol_out) {
              if (p.mode == EPI_BIAS_ACT) {
                r = p.oscale * apply_act(a + s_bias[col], p.act, p.act_param);
              } else if (col < p.ncol_main) {
                const float s = hrow ? dact_from_h(hrow[j] * p.hscale, p.dact) : 1.0f;
                r = p.oscale * s * a;
                if (arow) r += arow[j];
                if (p.mode == EPI_TANGENT) r2 = 100.0f * (1.0f - s) * vrow[j] * a;
              } else {
                r = p.oscale * a;  // tail columns (skip-connection branch)
              }
            }
            o1[j] = r; o2[j] = r2;
          }
          // ---- stores
          const int nmain = (p.mode == EPI_BIAS_ACT) ? p.ncol_out : min(p.ncol_out, p.ncol_main);
          float* orow = p.out + size_t(row) * p.ldo + c0;
          const bool vec_ok = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (c0 + 32 <= nmain);
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + j) = make_float4(o1[j], o1[j + 1], o1[j + 2], o1[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c0 + j < nmain) orow[j] = o1[j];
          }
          if (p.mode == EPI_TANGENT) {
            float* o2row = p.out2 + size_t(row) * p.ldo2 + c0;
            const bool vec2 = ((p.ldo2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out2) & 15) == 0) && (c0 + 32 <= nmain);
            if (vec2) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o2row + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (c0 + j < nmain) o2row[j] = o2[j];
            }
          }
          if (p.mode != EPI_BIAS_ACT && p.tail && c0 + 32 > p.ncol_main) {
            float* trow = p.tail + size_t(row) * p.ldt;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = c0 + j;
              if (col >= p.ncol_main && col < p.ncol_out) trow[col - p.ncol_main] = o1[j];
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  } else if (warp < kEpiWarps + kProdWarps) {
    // ============================== A producers: fp32 HBM -> split bf16 -> swizzled smem
    const int pw = warp - kEpiWarps;
    const int rsub = lane >> 4;          // 2 rows per load instruction
    const int col4 = (lane & 15) * 4;    // 16 lanes x float4 = 64 columns
    int g = 0;                           // global chunk counter (ring position)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int c = 0; c < p.k_chunks; ++c, ++g) {
        const int s = g % STAGES;
        float4 x[16];
        const int kcol = c * BK + col4;
        const bool col_ok = kcol < p.k_valid;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int r = tile * BM + pw * 32 + i * 2 + rsub;
          if (col_ok && r < M) x[i] = __ldg(reinterpret_cast<const float4*>(p.A + size_t(r) * p.lda + kcol));
          else x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
        uint8_t* a_hi = stage_base + s * Cfg::stage_bytes;
        uint8_t* a_lo = a_hi + kABytes;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t r = pw * 32 + i * 2 + rsub;
          const uint32_t off = sw128_offset(r, col4);
          __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
          split_bf16(x[i].x, h0, l0); split_bf16(x[i].y, h1, l1);
          split_bf16(x[i].z, h2, l2); split_bf16(x[i].w, h3, l3);
          __nv_bfloat162 hh0 = __halves2bfloat162(h0, h1), hh1 = __halves2bfloat162(h2, h3);
          __nv_bfloat162 ll0 = __halves2bfloat162(l0, l1), ll1 = __halves2bfloat162(l2, l3);
          uint2 hv, lv;
          hv.x = *reinterpret_cast<uint32_t*>(&hh0); hv.y = *reinterpret_cast<uint32_t*>(&hh1);
          lv.x = *reinterpret_cast<uint32_t*>(&ll0); lv.y = *reinterpret_cast<uint32_t*>(&ll1);
          *reinterpret_cast<uint2*>(a_hi + off) = hv;
          *reinterpret_cast<uint2*>(a_lo + off) = lv;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else if (warp == kEpiWarps + kProdWarps) {
    // ============================== MMA issuer (one elected lane)
    constexpr uint32_t idesc = make_idesc_bf16(BM, NPAD);
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(acc * 256);
      for (int c = 0; c < p.k_chunks; ++c, ++g) {
        const int s = g % STAGES;
        mbar_wait(&full[s], (g / STAGES) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = smem_u32(stage_base + s * Cfg::stage_bytes);
          const uint32_t a_lo = a_hi + kABytes;
          const uint32_t b_hi = a_hi + 2 * kABytes;
          const uint32_t b_lo = b_hi + Cfg::b_plane;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t dah = make_desc_k_sw128(a_hi + k * 32), dal = make_desc_k_sw128(a_lo + k * 32);
            const uint64_t dbh = make_desc_k_sw128(b_hi + k * 32), dbl = make_desc_k_sw128(b_lo + k * 32);
            umma_bf16(d_tmem, dal, dbh, idesc, (c | k) != 0);   // small terms first
            umma_bf16(d_tmem, dah, dbl, idesc, 1);
            umma_bf16(d_tmem, dah, dbh, idesc, 1);
          }
          umma_commit(&empty[s]);
          if (c == p.k_chunks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== W loader (one elected lane): bulk copy of the pre-swizzled image
    int g = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int c = 0; c < p.k_chunks; ++c, ++g) {
        const int s = g % STAGES;
        mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* b_hi = stage_base + s * Cfg::stage_bytes + 2 * kABytes;
          mbar_arrive_expect_tx(&full[s], 2 * Cfg::b_plane);
          bulk_copy_g2s(b_hi, p.wimg + size_t(c) * 2 * Cfg::b_plane, 2 * Cfg::b_plane, &full[s]);
        }
        __syncwarp();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kProdWarps + 1) tmem_dealloc<512>(tmem_base);
}

template <int NPAD>
static int launch_linear(const LinearParams& p, cudaStream_t stream) {
  using Cfg = LinCfg<NPAD>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_linear_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  int tiles_cap = (p.m_cap + BM - 1) / BM;
  if (tiles_cap <= 0) return NERO_OK;
  int grid = tiles_cap < kNumSMs ? tiles_cap : kNumSMs;
  umma_linear_kernel<NPAD><<<grid, kLinearThreads, Cfg::smem_bytes, stream>>>(p);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

int linear_dispatch(const LinearParams& p, cudaStream_t stream) {
  if (p.k_chunks <= 0 || (p.k_valid & 3) || (p.lda & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15)) return NERO_ERR_ARG;
  if (p.mode == EPI_TANGENT && (!p.H || !p.V || !p.out2)) return NERO_ERR_ARG;
  switch (p.n_pad) {
    case 16: return launch_linear<16>(p, stream);
    case 64: return launch_linear<64>(p, stream);
    case 128: return launch_linear<128>(p, stream);
    case 224: return launch_linear<224>(p, stream);
    case 256: return launch_linear<256>(p, stream);
    default: return NERO_ERR_ARG;
  }
}

}  // namespace nero
