// tcgen05 weight-gradient kernel: dW[n, k] = sum_m dY[m, n] * X[m, k]  (+ sum_m dY2[m, n] * X2[m, k]).
//
// The contraction runs over SAMPLES (rows of both fp32 row-major inputs), so both tensor-core operands are
// transposed on the fly: 16 producer warps read columns of dY / X (coalesced across lanes), split to bf16 hi/lo
// (packed cvt.rn.bf16x2) and write 8-sample k-chunks (one 16-byte store, bank-conflict free) into the same
// K-major SWIZZLE_128B layout the linear kernel uses (tile row = feature index, K = sample index).
// Split-K over CTAs: CTA (p, t) owns a contiguous range of 64-sample chunks of output row tile t, accumulates a
// 128 x NW fp32 tile in TMEM and ADDS it into the one [rows x ld] accumulator of the layer with vectorised L2
// reductions (red.global.add.v4.f32): the accumulator (256 KB) stays L2-resident, there is no [P][rows][ld] partial
// round trip through HBM any more (round 1: 29 MB written + re-read per layer).  nero_wgrad_finish (k_weights.cu)
// then applies the weight-norm chain rule to that single matrix.  The caller zeroes the accumulator.
// The optional second pair implements the double-backward term of the SDF network,
//   dW_k = abar_k^T h_k + v_k^T ubar_k      (SURVEY.md Appendix A.3, K4),
// in ONE accumulator.  Column sums of dY (bias gradient) are produced by the dY producer threads for free.

#include "common.cuh"
#include "ptx.cuh"
#include "umma_common.cuh"

namespace nero {

struct WgradParams {
  const float* dY; int ldy; const float* X; int ldx;
  const float* dY2; int ldy2; const float* X2; int ldx2;
  int n0; int n_valid;   // output rows [n0 + 128*blockIdx.y, +128) = dY columns; columns >= n_valid read as zero
  int k0; int k_valid;   // output cols [k0, k0+NW) = X columns; columns >= k_valid read as zero
  float* partial; int ld_partial; int rows_partial;  // [rows_partial][ld_partial], accumulated into (zeroed by the caller)
  float* bias_partial;                               // [rows_partial] (may be null), accumulated into
  const int* m_ptr; int m_cap;
};

constexpr int WG_BM = 128, WG_BK = 64;
constexpr int kWgProdWarps = 16;
constexpr int kWgThreads = (kWgProdWarps + 1) * 32;  // warps 0-7: dY producers (0-3 also epilogue), 8-15: X producers, 16: MMA
constexpr uint32_t kWgABytes = WG_BM * 128;

template <int NW> struct WgCfg {
  static constexpr uint32_t b_plane = NW * 128;
  static constexpr uint32_t stage_bytes = 2 * kWgABytes + 2 * b_plane;
  static constexpr int stages = (stage_bytes * 3 <= 200 * 1024) ? 3 : 2;
  static constexpr uint32_t smem_bytes = stages * stage_bytes + 1024 + 256 + 512;
};

__device__ __forceinline__ void wg_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const uint32_t hb = *reinterpret_cast<uint32_t*>(&h);
  const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  hi = hb;
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// one thread transposes NOCT x 8 consecutive samples of one column into 16-byte k-chunks (hi and lo planes).
// load_octets issues all NOCT*8 loads (before the smem slot is known to be free); store_octets converts and stores.
template <int NOCT>
__device__ __forceinline__ void load_octets(const float* __restrict__ src, int ld, int col, bool col_ok, int s0, int M, uint32_t oct0,
                                            float (&x)[NOCT][8]) {
#pragma unroll
  for (int o = 0; o < NOCT; ++o)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = s0 + (oct0 + o) * 8 + i;
      x[o][i] = (col_ok && s < M) ? __ldg(src + size_t(s) * ld + col) : 0.0f;
    }
}
template <int NOCT>
__device__ __forceinline__ float store_octets(const float (&x)[NOCT][8], uint8_t* plane_hi, uint8_t* plane_lo, uint32_t row, uint32_t oct0) {
  float sum = 0.0f;
#pragma unroll
  for (int o = 0; o < NOCT; ++o) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wg_split2(x[o][2 * i], x[o][2 * i + 1], hw[i], lw[i]);
      sum += x[o][2 * i] + x[o][2 * i + 1];
    }
    const uint32_t off = row * 128u + (((oct0 + o) ^ (row & 7u)) << 4);
    *reinterpret_cast<uint4*>(plane_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(plane_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
  return sum;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// epilogue of both kernels: the CTA's 128 x NW accumulator tile (TMEM) is added into the layer's accumulator matrix
template <int NW>
__device__ __forceinline__ void wg_epilogue(const WgradParams& p, uint32_t tmem_base, int pw, int lane, int n0, bool any, const float* s_bias) {
  const int orow = n0 + pw * 32 + lane;
  float* prow = p.partial + size_t(orow) * p.ld_partial + p.k0;
  const uint32_t taddr = tmem_base + (uint32_t(pw * 32) << 16);
  if (any) {
#pragma unroll 1
    for (int cc = 0; cc < (NW + 31) / 32; ++cc) {
      float v[32];
      tmem_ld32(taddr + cc * 32, v);
      tmem_ld_wait();
      if (orow < p.rows_partial) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (cc * 32 + j < NW) red_add_v4(prow + cc * 32 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    if (p.bias_partial && p.k0 == 0 && orow < p.rows_partial) atomicAdd(&p.bias_partial[orow], s_bias[pw * 32 + lane]);
  }
}

template <int NW>
__global__ void __launch_bounds__(kWgThreads, 1) umma_wgrad_kernel(const WgradParams p) {
  using Cfg = WgCfg<NW>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::stage_bytes + 256);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = p.m_ptr ? *p.m_ptr : p.m_cap;
  if (M > p.m_cap) M = p.m_cap;
  const int P = gridDim.x;
  const int n0 = p.n0 + int(blockIdx.y) * WG_BM;
  const int total_chunks = (M + WG_BK - 1) / WG_BK;
  const int cpp = (total_chunks + P - 1) / P;
  const int c_begin = min(total_chunks, int(blockIdx.x) * cpp), c_end = min(total_chunks, c_begin + cpp);
  const int npairs = p.dY2 ? 2 : 1;
  const int nc1 = c_end - c_begin;
  const int nchunks = nc1 * npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kWgProdWarps); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 128) s_bias[threadIdx.x] = 0.0f;
  if (warp == kWgProdWarps) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWgProdWarps) {
    const bool is_a = warp < 8;
    const int pw = warp & 3;             // 32-column slab
    const int oh = (warp >> 2) & 1;      // which half of the 8 octets
    float bias_acc = 0.0f;
    for (int g = 0; g < nchunks; ++g) {
      const int s = g % STAGES;
      const int pair = g / nc1;
      const int chunk = c_begin + g % nc1;
      const int s0 = chunk * WG_BK;
      uint8_t* st = smem + s * Cfg::stage_bytes;
      if (is_a) {
        const float* src = pair ? p.dY2 : p.dY;
        const int ld = pair ? p.ldy2 : p.ldy;
        const uint32_t row = pw * 32 + lane;
        const int col = n0 + int(row);
        const bool ok = col < p.n_valid;
        float x[4][8];
        load_octets<4>(src, ld, col, ok, s0, M, oh * 4, x);
        mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
        const float sacc = store_octets<4>(x, st, st + kWgABytes, row, oh * 4);
        if (pair == 0) bias_acc += sacc;
      } else {
        const float* src = pair ? p.X2 : p.X;
        const int ld = pair ? p.ldx2 : p.ldx;
        uint8_t* bh = st + 2 * kWgABytes;
        constexpr int NJ = (NW + 127) / 128;
        float x[NJ][4][8];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const uint32_t row = j * 128 + pw * 32 + lane;
          const int col = p.k0 + int(row);
          load_octets<4>(src, ld, col, row < NW && col < p.k_valid, s0, M, oh * 4, x[j]);
        }
        mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const uint32_t row = j * 128 + pw * 32 + lane;
          if (row < NW) store_octets<4>(x[j], bh, bh + Cfg::b_plane, row, oh * 4);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    if (is_a) atomicAdd(&s_bias[pw * 32 + lane], bias_acc);
    // every producer warp joins the barrier below (named barrier 1) so that s_bias is complete for the epilogue
    asm volatile("bar.sync 1, %0;" ::"n"(kWgProdWarps * 32));
    if (warp < 4) {
      // -------- epilogue: TMEM tile -> added into the layer's accumulator
      mbar_wait(tfull, 0);
      tcgen05_fence_after();
      wg_epilogue<NW>(p, tmem_base, pw, lane, n0, nchunks > 0, s_bias);
    }
  } else {
    // -------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(WG_BM, NW);
    for (int g = 0; g < nchunks; ++g) {
      const int s = g % STAGES;
      mbar_wait(&full[s], (g / STAGES) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * Cfg::stage_bytes);
        const uint32_t a_lo = a_hi + kWgABytes;
        const uint32_t b_hi = a_hi + 2 * kWgABytes;
        const uint32_t b_lo = b_hi + Cfg::b_plane;
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {
          const uint64_t dah = make_desc_k_sw128(a_hi + k * 32), dal = make_desc_k_sw128(a_lo + k * 32);
          const uint64_t dbh = make_desc_k_sw128(b_hi + k * 32), dbl = make_desc_k_sw128(b_lo + k * 32);
          umma_bf16(tmem_base, dal, dbh, idesc, (g | k) != 0);
          umma_bf16(tmem_base, dah, dbl, idesc, 1);
          umma_bf16(tmem_base, dah, dbh, idesc, 1);
        }
        umma_commit(&empty[s]);
        if (g == nchunks - 1) umma_commit(tfull);
      }
      __syncwarp();
    }
    if (nchunks == 0 && elect_one()) mbar_arrive(tfull);
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kWgProdWarps) tmem_dealloc<256>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------------
// MN-major variant.  Both operands are contracted over their ROW index (samples), i.e. in memory the MMA's M / N
// dimension (features) is the contiguous one: that is exactly an MN-major UMMA operand.  Producers therefore copy rows
// as they lie -- two 16-byte global loads (8 features of one sample), split to bf16 hi/lo, one 16-byte shared store per
// plane into the MN-major SWIZZLE_128B atom -- instead of transposing with scalar loads.
// Stage layout per plane: [feature block of 64][sample group of 8][8 rows x 128 B]  (SBO = 1024, LBO = 4096 with 32-sample stages).
// Requires 16-byte aligned rows (ld % 4 == 0, column origin % 4 == 0); wgrad_dispatch checks and otherwise uses the
// transposing kernel above.
// The 16 producer warps work as TWO GROUPS on alternating 32-sample stages: while one group converts and stores its stage,
// the other group's loads for the next stage are in flight.  (Measured neutral against all 16 warps sharing 64-sample stages,
// 70.1 vs 70.7 us per 256x256 layer at 127 k rows: of those ~55 us are the main loop at 4.75 TB/s -- two-pair launches take
// 125 us -- and ~15 us are fill, drain and the accumulate epilogue.)
constexpr int MN_BK = 32, kMnGroupWarps = kWgProdWarps / 2;
constexpr uint32_t kMnSbo = 1024, kMnLbo = (MN_BK / 8) * 1024;
template <int NW> struct WgMnCfg {
  static constexpr uint32_t a_plane = WG_BM * MN_BK * 2, b_plane = NW * MN_BK * 2;
  static constexpr uint32_t stage_bytes = 2 * a_plane + 2 * b_plane;
  static constexpr int stages = 4;
  static constexpr uint32_t smem_bytes = stages * stage_bytes + 1024 + 256 + 512;
};

__device__ __forceinline__ uint32_t mn_chunk_offset(uint32_t chunk, uint32_t s) {   // chunk = feature/8, s = sample in stage
  const uint32_t r = s & 7u;
  return (chunk >> 3) * kMnLbo + (s >> 3) * kMnSbo + r * 128u + (((chunk & 7u) ^ r) << 4);
}

// Epilogue of the MN-major kernel: all 16 producer warps move the 128 x NW accumulator tile TMEM -> shared (thread = row, as
// TMEM hands it out) and then ADD it into the layer's accumulator row by row, 32 lanes x 16 bytes = 512 contiguous bytes per
// reduction instruction (the thread-per-row epilogue of the transposing kernel issues every 16-byte reduction to a different
// row: 32 separate L2 transactions per instruction).
// CTAs start at different rows so that they do not all hit the same L2 lines at the same time.
template <int NW>
__device__ __forceinline__ void wg_epilogue_rows(const WgradParams& p, uint32_t tmem_base, float* tile, int warp, int lane, int n0,
                                                 const float* s_bias) {
  constexpr int LD = NW + 4;                              // row stride of the shared tile (floats): 16-byte aligned, 4 banks off
  constexpr int SLICE = NW / 4;                           // columns per warp: 64 / 32 / 16
  const int quarter = warp & 3, slice = warp >> 2;        // TMEM lane quarter this warp may read, column slice it takes
  {
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(slice * SLICE);
    float* dst = tile + row * LD + slice * SLICE;
#pragma unroll
    for (int c = 0; c < SLICE; c += 16) {
      float v[16];
      tmem_ld16(taddr + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kWgProdWarps * 32));
  const int r_first = (int(blockIdx.x) * 5) & 127;
  for (int i = warp; i < WG_BM; i += kWgProdWarps) {
    const int r = (i + r_first) & 127;
    const int orow = n0 + r;
    if (orow >= p.rows_partial) continue;
    float* prow = p.partial + size_t(orow) * p.ld_partial + p.k0;
#pragma unroll
    for (int c = lane * 4; c < NW; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(tile + r * LD + c);
      red_add_v4(prow + c, v.x, v.y, v.z, v.w);
    }
    if (lane == 0 && p.bias_partial && p.k0 == 0) atomicAdd(&p.bias_partial[orow], s_bias[r]);
  }
}

template <int NW>
__global__ void __launch_bounds__(kWgThreads, 1) umma_wgrad_mn_kernel(const WgradParams p) {
  using Cfg = WgMnCfg<NW>;
  constexpr int STAGES = Cfg::stages;
  constexpr int CB = NW / 8;                 // 16-byte chunks per X row
  constexpr int SPW = 32 / CB;               // X sample rows per warp instruction
  constexpr int ATASKS = MN_BK / 2;          // dY tasks (2 sample rows each)
  constexpr int TASKS = ATASKS + MN_BK / SPW;   // + X tasks: 48 / 32 / 24 per stage
  constexpr int TPW = TASKS / kMnGroupWarps;    // tasks per warp of the stage's group: 6 / 4 / 3
  static_assert(TASKS % kMnGroupWarps == 0 && TPW >= 2 && STAGES % 2 == 0, "task split");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::stage_bytes + 256);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = p.m_ptr ? *p.m_ptr : p.m_cap;
  if (M > p.m_cap) M = p.m_cap;
  const int P = gridDim.x;
  const int n0 = p.n0 + int(blockIdx.y) * WG_BM;
  const int total_chunks = (M + MN_BK - 1) / MN_BK;
  const int cpp = (total_chunks + P - 1) / P;
  const int c_begin = min(total_chunks, int(blockIdx.x) * cpp), c_end = min(total_chunks, c_begin + cpp);
  const int npairs = p.dY2 ? 2 : 1;
  const int nc1 = c_end - c_begin;
  const int nchunks = nc1 * npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kMnGroupWarps); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 128) s_bias[threadIdx.x] = 0.0f;
  if (warp == kWgProdWarps) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWgProdWarps) {
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // this lane's role in each of its tasks (constant over stages)
    const uint32_t a_chunk = lane & 15, a_sub = lane >> 4;
    const uint32_t b_chunk = lane % CB, b_sub = lane / CB;
    const int a_col = n0 + int(a_chunk) * 8, b_col = p.k0 + int(b_chunk) * 8;
    const int group = warp / kMnGroupWarps, wi = warp % kMnGroupWarps;
    for (int g = group; g < nchunks; g += 2) {        // this group's stages
      const int s = g % STAGES;
      const int pair = g / nc1;
      const int chunk = c_begin + g % nc1;
      const int s0 = chunk * MN_BK;
      uint8_t* st = smem + s * Cfg::stage_bytes;
      const float* srcA = pair ? p.dY2 : p.dY;
      const int ldA = pair ? p.ldy2 : p.ldy;
      const float* srcB = pair ? p.X2 : p.X;
      const int ldB = pair ? p.ldx2 : p.ldx;
      float4 v[TPW][2];
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int t = wi + kMnGroupWarps * i;
        const bool isA = t < ATASKS;                             // warp-uniform
        const int sl = isA ? 2 * t + int(a_sub) : (t - ATASKS) * SPW + int(b_sub);
        const int srow = s0 + sl;
        const int col = isA ? a_col : b_col;
        const int lim = isA ? p.n_valid : p.k_valid;
        const float* rp = (isA ? srcA : srcB) + size_t(srow) * (isA ? ldA : ldB) + col;
        const bool rok = srow < M;
        v[i][0] = (rok && col < lim) ? __ldg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i][1] = (rok && col + 4 < lim) ? __ldg(reinterpret_cast<const float4*>(rp + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int t = wi + kMnGroupWarps * i;
        const bool isA = t < ATASKS;
        const int sl = isA ? 2 * t + int(a_sub) : (t - ATASKS) * SPW + int(b_sub);
        const int col = isA ? a_col : b_col;
        const int lim = isA ? p.n_valid : p.k_valid;
        float x[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) if (col + j >= lim) x[j] = 0.f;      // partial float4 at the valid-column boundary
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wg_split2(x[2 * j], x[2 * j + 1], hw[j], lw[j]);
        uint8_t* hi = isA ? st : st + 2 * Cfg::a_plane;
        uint8_t* lo = isA ? st + Cfg::a_plane : st + 2 * Cfg::a_plane + Cfg::b_plane;
        const uint32_t off = mn_chunk_offset(isA ? a_chunk : b_chunk, uint32_t(sl));
        *reinterpret_cast<uint4*>(hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        if (isA && pair == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bsum[j] += x[j];
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], 16);
      if (lane < 16) atomicAdd(&s_bias[a_chunk * 8 + j], bsum[j]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kWgProdWarps * 32));
    // -------- epilogue: TMEM tile -> shared (the stage buffers are free once the last MMA has completed) -> accumulator rows
    mbar_wait(tfull, 0);
    tcgen05_fence_after();
    if (nchunks > 0) wg_epilogue_rows<NW>(p, tmem_base, reinterpret_cast<float*>(smem), warp, lane, n0, s_bias);
  } else {
    // -------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16_mn(WG_BM, NW);
    for (int g = 0; g < nchunks; ++g) {
      const int s = g % STAGES;
      mbar_wait(&full[s], (g / STAGES) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * Cfg::stage_bytes);
        const uint32_t a_lo = a_hi + Cfg::a_plane;
        const uint32_t b_hi = a_hi + 2 * Cfg::a_plane;
        const uint32_t b_lo = b_hi + Cfg::b_plane;
#pragma unroll
        for (int k = 0; k < MN_BK / 16; ++k) {     // 16 samples = two 8-row groups per MMA
          const uint32_t ko = k * 2 * kMnSbo;
          const uint64_t dah = make_desc_mn_sw128(a_hi + ko, kMnLbo, kMnSbo), dal = make_desc_mn_sw128(a_lo + ko, kMnLbo, kMnSbo);
          const uint64_t dbh = make_desc_mn_sw128(b_hi + ko, kMnLbo, kMnSbo), dbl = make_desc_mn_sw128(b_lo + ko, kMnLbo, kMnSbo);
          umma_bf16(tmem_base, dal, dbh, idesc, (g | k) != 0);
          umma_bf16(tmem_base, dah, dbl, idesc, 1);
          umma_bf16(tmem_base, dah, dbh, idesc, 1);
        }
        umma_commit(&empty[s]);
        if (g == nchunks - 1) umma_commit(tfull);
      }
      __syncwarp();
    }
    if (nchunks == 0 && elect_one()) mbar_arrive(tfull);
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kWgProdWarps) tmem_dealloc<256>(tmem_base);
}

template <int NW>
static int launch_wgrad_mn(const WgradParams& p, int P, int n_tiles, cudaStream_t stream) {
  using Cfg = WgMnCfg<NW>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_wgrad_mn_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  umma_wgrad_mn_kernel<NW><<<dim3(P, n_tiles), kWgThreads, Cfg::smem_bytes, stream>>>(p);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

template <int NW>
static int launch_wgrad(const WgradParams& p, int P, int n_tiles, cudaStream_t stream) {
  using Cfg = WgCfg<NW>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_wgrad_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  umma_wgrad_kernel<NW><<<dim3(P, n_tiles), kWgThreads, Cfg::smem_bytes, stream>>>(p);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

// Computes partial[P][rows_partial][ld_partial] for output rows [0, n_rows_pad) and columns [0, k_pad)
// (k_pad multiple of 64), decomposed into 128-row tiles (grid.y) x {256,128,64}-column tiles (separate launches).
int wgrad_dispatch(WgradParams p, int n_rows_pad, int k_pad, int P, cudaStream_t stream) {
  if ((p.ld_partial & 3) || k_pad % 64 || n_rows_pad % 16 || P <= 0) return NERO_ERR_ARG;
  const int n_tiles = (n_rows_pad + 127) / 128;
  p.n0 = 0;
  // the MN-major kernel copies rows in 16-byte pieces: it needs 16-byte aligned rows and column origins; otherwise the
  // transposing kernel runs.
  // (Tried in round 2 and removed: feeding the MN-major kernel from a TMA ring of raw fp32 row blocks [2 x 48 KB] with the 16
  //  warps converting shared -> shared -- 81.6 us vs 71.3 us per 256x256 layer at 127 k rows, profiles/r02r_*: with only two
  //  refillable stages the bytes in flight are no more than the register-staged loads already keep in flight;
  //  and computing both 128-row tiles of a 256-row layer in one CTA [32-sample stages, X converted once]: 85.0 us,
  //  profiles/r02t_*.  The loop is bound by load latency + conversion time per stage, i.e. by the bytes one stage keeps in
  //  flight, not by the amount of conversion work.)
  auto al = [](const float* q, int ld) { return q == nullptr || ((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld & 3) == 0); };
  const bool mn = al(p.dY, p.ldy) && al(p.X, p.ldx) && al(p.dY2, p.ldy2) && al(p.X2, p.ldx2);
  int k0 = 0;
  while (k0 < k_pad) {
    const int rem = k_pad - k0;
    p.k0 = k0;
    int rc;
    if (rem >= 256) { rc = mn ? launch_wgrad_mn<256>(p, P, n_tiles, stream) : launch_wgrad<256>(p, P, n_tiles, stream); k0 += 256; }
    else if (rem >= 128) { rc = mn ? launch_wgrad_mn<128>(p, P, n_tiles, stream) : launch_wgrad<128>(p, P, n_tiles, stream); k0 += 128; }
    else { rc = mn ? launch_wgrad_mn<64>(p, P, n_tiles, stream) : launch_wgrad<64>(p, P, n_tiles, stream); k0 += 64; }
    if (rc != NERO_OK) return rc;
  }
  return NERO_OK;
}

}  // namespace nero
