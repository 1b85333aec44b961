// tcgen05 (UMMA) linear-layer kernel: OUT[M, N] = epilogue( A[M, K] * W[N, K]^T ).
//
//  * A is fp32 row-major in HBM.  EIGHT producer warps (two groups alternating K-chunks, so one group's global
//    loads are in flight while the other converts) load it with coalesced float4, split every value into
//    bf16 hi + bf16 lo (error-compensated "split-bf16": x ~= hi + lo, ~2^-17 relative; packed cvt.rn.bf16x2) and
//    store both planes into the K-major SWIZZLE_128B shared-memory operand layout the tensor core reads.
//  * W is pre-split and pre-swizzled once per step by nero_prep_weight (k_weights.cu) into the exact
//    shared-memory image, so a K-chunk of it is ONE cp.async.bulk (UBLKCP) from L2, completing on an mbarrier.
//  * One elected thread issues tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 in TMEM), three MMAs per k-step:
//    A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  (the lo*lo term, ~2^-18 relative, is dropped).
//  * The fp32 accumulator tile (128 x N) lives in TMEM, double buffered (2 x 256 columns), so the EIGHT epilogue
//    warps (tcgen05.ld -> bias/activation/derivative products -> HBM) overlap the next tile's MMAs.  Every global
//    access of the epilogue is staged through a per-warp shared-memory transpose buffer so that warps read and
//    write 64-byte row segments (8 rows per instruction) instead of one row per lane.
//  * Epilogue math is specialised at compile time (template EPI) and uses ex2/lg2 approximations for
//    softplus(beta=100) and its derivative (absolute error ~1e-9 on activations of O(1)).
//  * Persistent CTAs (one per SM), tiles of 128 rows, row count read from device memory (no host sync for the
//    data-dependent number of live samples).
//
// Reference op sites replaced: every nn.Linear (+weight_norm, +Softplus/ReLU/Sigmoid/exp) of
// network/field.py:130-147 (SDFNetwork), :258-283 (NeRFNetwork), :310-346 (make_predictor), and the
// autograd-generated input-gradient GEMMs of their backward / double-backward.
#include "umma_common.cuh"

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
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kMmaWarp = kEpiWarps + kProdWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kLinearThreads = (kEpiWarps + kProdWarps + 2) * 32;
constexpr uint32_t kABytes = BM * 128;           // one bf16 plane of an A stage (16 KB)

template <int NPAD> struct LinCfg {
  static constexpr uint32_t b_plane = NPAD * 128;
  static constexpr uint32_t stage_bytes = 2 * kABytes + 2 * b_plane;
  static constexpr int stages = (stage_bytes * 4 <= 200 * 1024) ? 4 : (stage_bytes * 3 <= 200 * 1024) ? 3 : 2;
  static constexpr uint32_t epi_bytes = kEpiWarps * kStageWarpBytes;   // 20 KB
  static constexpr uint32_t smem_bytes = stages * stage_bytes + epi_bytes + 1024 /*align*/ + 1024 /*bias*/ + 256 /*barriers*/;
};

template <int NPAD, int EPI>
__global__ void __launch_bounds__(kLinearThreads, 1) umma_linear_kernel(const LinearParams p) {
  using Cfg = LinCfg<NPAD>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* s_epi = reinterpret_cast<float*>(smem + STAGES * Cfg::stage_bytes);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::stage_bytes + Cfg::epi_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes + Cfg::epi_bytes + 1024);
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
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kProdWarps / 2 + 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kLoadWarp) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = (p.bias && i < p.n_bias) ? p.bias[i] : 0.0f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ============================== epilogue warps: TMEM -> registers -> (smem transpose) -> HBM
    const int lg = warp & 3;            // TMEM lane group (rows lg*32 .. +31 of the tile)
    const int half = warp >> 2;         // which 16-column blocks (interleaved) this warp handles
    float* stg = s_epi + warp * (32 * kStagePitch);
    constexpr int NBLK = NPAD / 16;
    constexpr bool kBias = (EPI == EK_BIAS_SOFTPLUS || EPI == EK_BIAS_RELU || EPI == EK_BIAS_GENERIC);
    const int nmain = kBias ? p.ncol_out : min(p.ncol_out, p.ncol_main);
    const bool v_out = vec_ok(p.out, p.ldo);
    int it = 0;
    // aux operands (H / addend / V) of the DACT / TANGENT epilogues are pulled into L2 one tile ahead, so the
    // per-block loads below see L2 latency instead of DRAM latency (each half-warp-group prefetches its own rows once)
    auto prefetch_tile = [&](int tile) {
      if constexpr (!kBias) {
        if (half == 0 && tile < num_tiles) {
          const int r0 = tile * BM + lg * 32;
          const int rv = min(32, M - r0);
          if constexpr (EPI != EK_DACT_NONE) prefetch_rows_l2(p.H ? p.H + size_t(r0) * p.ldh : nullptr, p.ldh, rv, nmain, lane);
          prefetch_rows_l2(p.addend ? p.addend + size_t(r0) * p.ldadd : nullptr, p.ldadd, rv, nmain, lane);
          if constexpr (EPI == EK_TANGENT) prefetch_rows_l2(p.V + size_t(r0) * p.ldv, p.ldv, rv, nmain, lane);
        }
      }
    };
    prefetch_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      prefetch_tile(tile + gridDim.x);
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const int row0 = tile * BM + lg * 32;
      const int rows_valid = min(32, M - row0);       // may be <= 0
      const uint32_t taddr = tmem_base + (uint32_t(lg * 32) << 16) + uint32_t(acc * 256);
#pragma unroll 1
      for (int b = half; b < NBLK; b += 2) {
        const int c0 = b * 16;
        if (c0 >= p.ncol_out) break;
        float v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (rows_valid <= 0) continue;
        float r[16];
        if constexpr (kBias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += s_bias[c0 + j];
          if constexpr (EPI == EK_BIAS_SOFTPLUS) softplus100_fast16(v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float y = v[j];
            if constexpr (EPI == EK_BIAS_RELU) y = fmaxf(y, 0.0f);
            else if constexpr (EPI == EK_BIAS_GENERIC) y = apply_act(y, p.act, p.act_param);
            r[j] = p.oscale * y;
          }
          store_block16(p.out + size_t(row0) * p.ldo + c0, p.ldo, rows_valid, nmain - c0, v_out, stg, lane, r);
        } else {
          const int cm = nmain - c0;   // main columns in this block (may be <= 0)
          float s[16];
          if constexpr (EPI == EK_DACT_NONE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = 1.0f;
          } else {
            load_block16(p.H + size_t(row0) * p.ldh + c0, p.ldh, rows_valid, cm, vec_ok(p.H, p.ldh), stg, lane, s);
            if constexpr (EPI == EK_DACT_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) s[j] = s[j] > 0.0f ? 1.0f : 0.0f;
            } else {
              dsoftplus100_from_h_fast16(s, p.hscale);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = p.oscale * s[j] * v[j];
          if (p.addend) {
            float a[16];
            load_block16(p.addend + size_t(row0) * p.ldadd + c0, p.ldadd, rows_valid, cm, vec_ok(p.addend, p.ldadd), stg, lane, a);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] += a[j];
          }
          if (p.tail && c0 + 16 > p.ncol_main) {
            // skip-branch columns (>= ncol_main) carry oscale*acc and go to the tail buffer
            float t[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) t[j] = p.oscale * v[j];
            const int shift = max(p.ncol_main - c0, 0);       // first tail column inside this block
            // write columns [shift, min(16, ncol_out - c0)) of t to tail[:, c0 + shift - ncol_main ...]
            const int ncols = min(16, p.ncol_out - c0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j >= shift && j < ncols && lane < rows_valid) p.tail[size_t(row0 + lane) * p.ldt + (c0 + j - p.ncol_main)] = t[j];
          }
          if (cm > 0) store_block16(p.out + size_t(row0) * p.ldo + c0, p.ldo, rows_valid, cm, v_out, stg, lane, r);
          if constexpr (EPI == EK_TANGENT) {
            if (cm > 0) {
              float vv[16];
              load_block16(p.V + size_t(row0) * p.ldv + c0, p.ldv, rows_valid, cm, vec_ok(p.V, p.ldv), stg, lane, vv);
#pragma unroll
              for (int j = 0; j < 16; ++j) vv[j] = 100.0f * (1.0f - s[j]) * vv[j] * v[j];
              store_block16(p.out2 + size_t(row0) * p.ldo2 + c0, p.ldo2, rows_valid, cm, vec_ok(p.out2, p.ldo2), stg, lane, vv);
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
    const int pw = (warp - kEpiWarps) & 3;       // 32-row slab
    const int grp = (warp - kEpiWarps) >> 2;     // chunk parity handled by this group
    const int rsub = lane >> 4;                  // 2 rows per load instruction
    const int col4 = (lane & 15) * 4;            // 16 lanes x float4 = 64 columns
    int g = 0;                                   // global chunk counter (ring position)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int c = 0; c < p.k_chunks; ++c, ++g) {
        if ((g & 1) != grp) continue;
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
          uint2 hv, lv;
          split2(x[i].x, x[i].y, hv.x, lv.x);
          split2(x[i].z, x[i].w, hv.y, lv.y);
          *reinterpret_cast<uint2*>(a_hi + off) = hv;
          *reinterpret_cast<uint2*>(a_lo + off) = lv;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else if (warp == kMmaWarp) {
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
  if (warp == kLoadWarp) tmem_dealloc<512>(tmem_base);
}

template <int NPAD, int EPI>
static int launch_linear(const LinearParams& p, cudaStream_t stream) {
  using Cfg = LinCfg<NPAD>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_linear_kernel<NPAD, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  int tiles_cap = (p.m_cap + BM - 1) / BM;
  if (tiles_cap <= 0) return NERO_OK;
  int grid = tiles_cap < kNumSMs ? tiles_cap : kNumSMs;
  umma_linear_kernel<NPAD, EPI><<<grid, kLinearThreads, Cfg::smem_bytes, stream>>>(p);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

template <int NPAD>
static int dispatch_epi(const LinearParams& p, cudaStream_t stream) {
  if (p.mode == EPI_BIAS_ACT) {
    if (p.act == ACT_SOFTPLUS100) return launch_linear<NPAD, EK_BIAS_SOFTPLUS>(p, stream);
    if (p.act == ACT_RELU) return launch_linear<NPAD, EK_BIAS_RELU>(p, stream);
    return launch_linear<NPAD, EK_BIAS_GENERIC>(p, stream);
  }
  if (p.mode == EPI_TANGENT) return launch_linear<NPAD, EK_TANGENT>(p, stream);
  if (!p.H || p.dact == ACT_NONE) return launch_linear<NPAD, EK_DACT_NONE>(p, stream);
  if (p.dact == ACT_RELU) return launch_linear<NPAD, EK_DACT_RELU>(p, stream);
  return launch_linear<NPAD, EK_DACT_SOFTPLUS>(p, stream);
}

int linear_dispatch(const LinearParams& p, cudaStream_t stream) {
  if (p.k_chunks <= 0 || (p.k_valid & 3) || (p.lda & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15)) return NERO_ERR_ARG;
  if (p.mode == EPI_TANGENT && (!p.H || !p.V || !p.out2 || p.dact != ACT_SOFTPLUS100)) return NERO_ERR_ARG;
  if (p.mode == EPI_MUL_DACT && p.dact != ACT_NONE && !p.H) return NERO_ERR_ARG;
  switch (p.n_pad) {
    case 16: return dispatch_epi<16>(p, stream);
    case 64: return dispatch_epi<64>(p, stream);
    case 128: return dispatch_epi<128>(p, stream);
    case 224: return dispatch_epi<224>(p, stream);
    case 256: return dispatch_epi<256>(p, stream);
    default: return NERO_ERR_ARG;
  }
}

}  // namespace nero
