// Fused MLP-CHAIN kernel on tcgen05: a 128-row tile of samples flows through a whole sequence of linear layers
// without its activations ever leaving the SM between layers.
//
//   TMEM (512 columns x 128 lanes x 32 bit, all of it):
//     [  0,256)  fp32 accumulator of the current layer (D operand)
//     [256,384)  A operand, bf16 "hi" plane, two K elements per 32-bit column  (A read by tcgen05.mma FROM TMEM)
//     [384,512)  A operand, bf16 "lo" plane                                    (split-bf16: x ~= hi + lo)
//   Per layer: the MMA warp issues, per 16-wide k-step, A_lo*W_hi + A_hi*W_lo + A_hi*W_hi into the accumulator; the
//   epilogue warps then read the accumulator (tcgen05.ld, one matrix row per thread), apply bias/activation or the
//   derivative products of the gradient sweeps, and write the NEXT layer's A operand straight back into TMEM
//   (tcgen05.st) as packed split-bf16.
//
//   Round 2 (v3): every HBM operand of the epilogue moves through TMA.  The epilogue warps form TEAMS of four warps (one
//   per TMEM lane quarter); a team works on [128 rows x 16 columns] units.  Per team a FIFO of 8 KB shared-memory slots
//   (SWIZZLE_64B boxes, bank-conflict free for one-row-per-thread 16-byte accesses) is filled by cp.async.bulk.tensor
//   loads that one lane of the aux-loader warp issues up to kSlotsIn entries AHEAD -- across units, layers and tiles, i.e.
//   while the tensor core is still busy with the MMAs whose epilogue will consume them -- and results leave through
//   cp.async.bulk.tensor stores that one lane of the store warp issues from kSlotsOut staging slots.  Shipped
//   configuration: 4 teams (16 epilogue warps + MMA, W-loader, aux-loader and store warp = 640 threads).  No thread ever waits on a global load; the former
//   long-scoreboard stalls (ncu, round 1: 6.6 of 12.3 cycles per issued instruction) are gone.
//   Weights stream from L2 as pre-swizzled images: one stage = the hi and lo planes of one 64-wide K chunk for one
//   half of the N range (32 KB, cp.async.bulk, kWStages-deep mbarrier ring).
//
// This realises SURVEY.md K2/K3/K4/K10/K13 "activations stay on-chip across layers".
//
// Skip connection of the SDF network (field.py:139-140): a layer with `csrc` set takes columns >= ncol_out of the
// next A operand from that buffer, where ray_fill / pe_tangent pre-stored PE/sqrt2 resp. its tangent.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "umma_common.cuh"
#include "tma.cuh"

namespace nero {

constexpr int kMaxChainLayers = 10;

// ---------------------------------------------------------------------------------------------- C ABI (host) structs
struct ChainLayer {
  const uint8_t* wimg; const float* bias;
  float* save; const float* H; const float* addend; const float* V; float* out2; float* tail;
  int n_pad, k_chunks, n_bias, ncol_out, ncol_main, kind, act;
  int ld_save, ldh, ldadd, ldv, ldo2, ldt;
  float oscale, hscale, act_param;
  int write_a, a_blocks;      // write_a: the output becomes the next A operand; a_blocks: 16-col blocks of it to define
  const float* csrc; int ld_csrc;   // skip-concat source: columns >= ncol_out of the next A operand come from csrc[row, col]
  int pad_;
};
struct ChainParams {
  const float* A0; int lda0; int k_valid0;
  int n_layers; const int* m_ptr; int m_cap;
  ChainLayer L[kMaxChainLayers];
};

// ---------------------------------------------------------------------------------------------- device-side records
enum ChainOp : int { OP_H = 0, OP_AUX2 = 1, OP_CSRC = 2, OP_SAVE = 3, OP_OUT2 = 4, OP_COUNT = 5 };

struct ChainLayerDev {
  const uint8_t* wimg; const float* bias;
  float* save; const float* H; const float* aux2; float* out2; float* tail; const float* csrc;
  int n_pad, k_chunks, n_bias, ncol_out, ncol_main, kind, act;
  int ld_save, ldh, ld2, ldo2, ldt, ld_csrc;
  float oscale, hscale, act_param;
  int write_a, a_blocks;
  int aux2_is_addend;          // aux2 = addend (derivative kinds) or V (tangent kind)
  int tma;                     // bit o set: operand o moves through TMA (tensor map maps[map[o]])
  int map[OP_COUNT];
  int n_units;                 // 16-column units the epilogue walks (accumulator columns + next-A columns to define)
  int n_full;                  // units completely inside the valid columns: the ones with TMA traffic
  int in_ops, out_ops;         // operand bits loaded / stored through TMA by those units
};
constexpr int kMaxMaps = 1 + kMaxChainLayers * OP_COUNT;
struct ChainParamsDev {
  const float* A0; int lda0; int k_valid0; int a0_tma; int a0_units;
  int n_layers; const int* m_ptr; int m_cap;
  ChainLayerDev L[kMaxChainLayers];
  CUtensorMap maps[kMaxMaps];        // maps[0]: A0
};

// ---------------------------------------------------------------------------------------------- configuration
#ifndef NERO_TEAMS
#define NERO_TEAMS 4
#endif
#ifndef NERO_SLOTS_IN
#define NERO_SLOTS_IN 2
#endif
#ifndef NERO_SLOTS_OUT
#define NERO_SLOTS_OUT 2
#endif
constexpr int CH_BM = 128, CH_BK = 64;
constexpr int kTeams = NERO_TEAMS;                 // epilogue teams (4 warps each, one per TMEM lane quarter)
constexpr int kChEpiWarps = 4 * kTeams;
constexpr int kChMmaWarp = kChEpiWarps, kChLoadWarp = kChEpiWarps + 1, kChAuxWarp = kChEpiWarps + 2, kChStoreWarp = kChEpiWarps + 3;
constexpr int kChThreads = (kChEpiWarps + 4) * 32;
#ifndef NERO_WSTAGES
#define NERO_WSTAGES 3
#endif
constexpr int kWStages = NERO_WSTAGES;
// Timing experiments only (tools/bench_chain.py A/B builds; results are WRONG with any bit set): 1|2 = the MMA issuer and the
// epilogue ignore each other's barriers (each free-runs at its own speed), 4 = no MMA instructions are issued,
// 8 = no epilogue math / aux traffic, 16 = only the hi plane of W is loaded (half the bytes).
#ifndef NERO_CHAIN_EXP
#define NERO_CHAIN_EXP 0
#endif
constexpr int kExp = NERO_CHAIN_EXP;
constexpr uint32_t kWStageBytes = 2 * 128 * 128;   // hi + lo planes of up to 128 weight rows x 64 K (one N half of a K chunk)
constexpr int kSlotsIn = NERO_SLOTS_IN, kSlotsOut = NERO_SLOTS_OUT;         // per team
constexpr uint32_t kSlotBytes = CH_BM * 16 * 4;    // [128 rows x 16 fp32] = 8 KB
constexpr uint32_t kSlotsBytes = kTeams * (kSlotsIn + kSlotsOut) * kSlotBytes;
constexpr uint32_t kChBiasBytes = 2 * 256 * 4;
constexpr uint32_t kChBarBytes = 512;
constexpr uint32_t kChSmemBytes = kWStages * kWStageBytes + kSlotsBytes + kChBiasBytes + kChBarBytes;
static_assert(kChSmemBytes <= 232448, "shared memory budget");
static_assert((2 * kWStages + 5 + 2 * kTeams * (kSlotsIn + kSlotsOut)) * 8 + 8 <= kChBarBytes, "barrier area");
constexpr uint32_t kAccCol = 0, kAHiCol = 256, kALoCol = 384;

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// write 16 fp32 values of this lane's row (columns c0..c0+15 of the next A operand) as packed split-bf16 into TMEM
__device__ __forceinline__ void write_a16(uint32_t tmem_lane_base, int c0, const float* y) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split2(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
  tmem_st8(tmem_lane_base + kAHiCol + (c0 >> 1), hi);
  tmem_st8(tmem_lane_base + kALoCol + (c0 >> 1), lo);
}

// one row (this thread's) of a [128 x 16] fp32 slot in the SWIZZLE_64B layout: 64-byte rows, the 16-byte chunk index is
// XORed with bits 1..2 of the row (the pattern cuTensorMapEncodeTiled(..., SWIZZLE_64B) uses; slots are 1024-B aligned).
// `rowoff` = row*64, `sw` = ((row >> 1) & 3) << 4 are per-thread constants.
__device__ __forceinline__ void slot_read16(const uint8_t* slot, uint32_t rowoff, uint32_t sw, float* x) {
  const uint8_t* base = slot + rowoff;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(base + ((j << 4) ^ sw));
    x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
  }
}
__device__ __forceinline__ void slot_write16(uint8_t* slot, uint32_t rowoff, uint32_t sw, const float* x) {
  uint8_t* base = slot + rowoff;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(base + ((j << 4) ^ sw)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
}
// direct (non-TMA) access of this thread's row: narrow / unaligned operands and the ragged last tile
__device__ __noinline__ void row_load16(const float* __restrict__ g, int ld, long row, int c0, int ncols, bool row_ok, float* x) {
  const float* p = g + row * ld + c0;
  const bool vec = row_ok && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (vec && 4 * j + 3 < ncols) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + 4 * j));
      x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[4 * j + i] = (row_ok && 4 * j + i < ncols) ? __ldg(p + 4 * j + i) : 0.0f;
    }
  }
}
__device__ __noinline__ void row_store16(float* __restrict__ g, int ld, long row, int c0, int ncols, bool row_ok, const float* x) {
  if (!row_ok) return;
  float* p = g + row * ld + c0;
  const bool vec = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (vec && 4 * j + 3 < ncols) {
      *reinterpret_cast<float4*>(p + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (4 * j + i < ncols) p[4 * j + i] = x[4 * j + i];
    }
  }
}

// ---------------------------------------------------------------------------------------------- unit bookkeeping
// TMA traffic follows ONE regular pattern per layer, so that the service lanes (aux loader / store warp) and the epilogue
// threads agree on the FIFO order without exchanging anything: the units that lie completely inside the valid columns
// (u < n_full = nmain / 16) load the operands in `in_ops` (H, then aux2) and store the results in `out_ops` (save, then
// out2) through TMA, team t owning units t, t + kTeams, ...; everything else (the unit that straddles the last valid
// column -- TMA bounds the innermost dimension in 16-byte granules: a box clipped at column 217 still wrote 217..219 --,
// the skip-concat source, narrow or unaligned operands, the rows of a ragged last tile) is accessed row by row.
__device__ __forceinline__ uint32_t unit_in_mask(const ChainLayerDev& L, int u) {
  return u < L.n_full ? uint32_t(L.in_ops) & ~(1u << OP_CSRC) : uint32_t(L.in_ops) & (1u << OP_CSRC);
}
__device__ __forceinline__ uint32_t unit_out_mask(const ChainLayerDev& L, int u, bool tile_tma_ok) {
  return (tile_tma_ok && u < L.n_full) ? uint32_t(L.out_ops) : 0u;
}
__device__ __forceinline__ bool tile_tma_store_ok(const ChainParamsDev& p, int tile, int M) {
  // TMA stores write whole boxes clipped at the tensor extent (m_cap rows): on a ragged tile with M < m_cap the rows
  // beyond M lie inside the extent and must not be written, so that tile stores row by row
  return (M - tile * CH_BM >= CH_BM) || (M >= p.m_cap);
}

// per-thread view of its team's FIFOs
struct TeamCtx {
  uint8_t* slots_in;        // kSlotsIn x 8 KB
  uint8_t* slots_out;       // kSlotsOut x 8 KB
  uint64_t* in_full;        // [kSlotsIn]   TMA load landed                     (aux loader -> team)
  uint64_t* in_empty;       // [kSlotsIn]   all four warps have read the slot   (team -> aux loader)
  uint64_t* out_full;       // [kSlotsOut]  all four warps have written         (team -> store warp)
  uint64_t* out_empty;      // [kSlotsOut]  the TMA store has read the slot     (store warp -> team)
  int q, lane;
  uint32_t rowoff, sw;      // this thread's row offset / swizzle term inside a slot
  uint32_t in_seq, out_seq; // entries consumed / produced so far
};
__device__ __forceinline__ const uint8_t* in_wait(TeamCtx& t) {
  const uint32_t s = t.in_seq % kSlotsIn;
  mbar_wait(&t.in_full[s], (t.in_seq / kSlotsIn) & 1);
  return t.slots_in + s * kSlotBytes;
}
// release the oldest input slot: call after the values read from it have been USED (so the reads have completed)
__device__ __forceinline__ void in_release(TeamCtx& t) {
  __syncwarp();
  if (t.lane == 0) mbar_arrive(&t.in_empty[t.in_seq % kSlotsIn]);
  ++t.in_seq;
}
__device__ __forceinline__ uint8_t* out_acquire(TeamCtx& t) {
  const uint32_t s = t.out_seq % kSlotsOut;
  mbar_wait(&t.out_empty[s], ((t.out_seq / kSlotsOut) & 1) ^ 1);
  return t.slots_out + s * kSlotBytes;
}
__device__ __forceinline__ void out_publish(TeamCtx& t) {
  fence_proxy_async_smem();
  __syncwarp();
  if (t.lane == 0) mbar_arrive(&t.out_full[t.out_seq % kSlotsOut]);
  ++t.out_seq;
}

// ---------------------------------------------------------------------------------------------- epilogue of one unit
// FAST: the unit lies completely inside the valid columns, every operand it touches moves through TMA and the tile may be
// stored by TMA -- the common case; all run-time conditionals of the general unit fold away.
template <int KIND, bool FAST, typename PreWrite>
__device__ __forceinline__ void epi_unit(const ChainParamsDev& p, int l, TeamCtx& t, uint32_t tl, int u, int tile, int rows_valid,
                                         bool tile_tma_ok, const float* s_bias, PreWrite&& before_a_write) {
  constexpr bool kBias = (KIND == EK_BIAS_SOFTPLUS || KIND == EK_BIAS_RELU || KIND == EK_BIAS_GENERIC);
  const ChainLayerDev& L = p.L[l];
  const int c0 = u * 16;
  const int nblk = L.n_pad >> 4;
  const int nmain = kBias ? L.ncol_out : min(L.ncol_out, L.ncol_main);
  const int nblk_a = L.write_a ? max(L.a_blocks, nblk) : 0;
  const int rl = t.q * 32 + t.lane;                 // row of this thread inside the tile
  const long row = long(tile) * CH_BM + rl;
  const bool row_ok = rl < rows_valid;
  const uint32_t in_mask = FAST ? 0u : unit_in_mask(L, u);
  const uint32_t out_mask = FAST ? 0u : unit_out_mask(L, u, tile_tma_ok);
  const bool computed = FAST || (u < nblk && c0 < L.ncol_out);
  const bool has_main = FAST || (computed && c0 < nmain);
  const bool st_save = has_main && L.save != nullptr;
  const bool st_out2 = (KIND == EK_TANGENT) && has_main;
  const bool tma_save = FAST ? st_save : (out_mask >> OP_SAVE) & 1;
  const bool tma_out2 = FAST ? st_out2 : (out_mask >> OP_OUT2) & 1;
  float r[16];
  float q2[16];
  if (computed) {
    float v[16];
    tmem_ld16(tl + kAccCol + c0, v);
    if constexpr (kBias) {
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += s_bias[c0 + j];
      if constexpr (KIND == EK_BIAS_SOFTPLUS) softplus100_fast16(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float y = v[j];
        if constexpr (KIND == EK_BIAS_RELU) y = fmaxf(y, 0.0f);
        else if constexpr (KIND == EK_BIAS_GENERIC) y = apply_act(y, L.act, L.act_param);
        r[j] = y;
      }
      if (L.oscale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] *= L.oscale;
      }
    } else {
      float s[16];
      bool rel_h = false;
      if constexpr (KIND == EK_DACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s[j] = 1.0f;
      } else {
        if (FAST || (in_mask & (1u << OP_H))) { slot_read16(in_wait(t), t.rowoff, t.sw, s); rel_h = true; }
        else if (has_main) row_load16(L.H, L.ldh, row, c0, nmain - c0, row_ok, s);
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = 0.0f;
        }
        if constexpr (KIND == EK_DACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = s[j] > 0.0f ? 1.0f : 0.0f;
        } else {
          dsoftplus100_from_h_fast16(s, L.hscale);
        }
        if (!has_main) {
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = 0.0f;
        }
      }
      if (rel_h) in_release(t);
      float a2[16];
      const bool has2 = has_main && L.aux2 != nullptr;
      bool rel_2 = false;
      if (FAST ? has2 : bool(in_mask & (1u << OP_AUX2))) { slot_read16(in_wait(t), t.rowoff, t.sw, a2); rel_2 = true; }
      else if (has2) row_load16(L.aux2, L.ld2, row, c0, nmain - c0, row_ok, a2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = s[j] * v[j];
      if (L.oscale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] *= L.oscale;
      }
      if constexpr (KIND == EK_TANGENT) {
        if (has_main) {
#pragma unroll
          for (int j = 0; j < 16; ++j) q2[j] = 100.0f * (1.0f - s[j]) * a2[j] * v[j];
        }
      } else {
        if (has2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] += a2[j];
        }
      }
      if (rel_2) in_release(t);
      if (!FAST && L.tail && c0 + 16 > L.ncol_main && row_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + j;
          if (col >= L.ncol_main && col < L.ncol_out) L.tail[row * L.ldt + (col - L.ncol_main)] = L.oscale * v[j];
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = 0.0f;
  }
  // ---- next A operand
  if (FAST ? (L.write_a != 0) : (u < nblk_a)) {
    before_a_write(u);
    if (!FAST && c0 + 16 > nmain) {     // columns >= nmain: the skip-concat source or zero
      float cc[16];
      const bool tma_c = (in_mask >> OP_CSRC) & 1;
      if (tma_c) slot_read16(in_wait(t), t.rowoff, t.sw, cc);
      else if (L.csrc) row_load16(L.csrc, L.ld_csrc, row, c0, 256 - c0, row_ok, cc);
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = (c0 + j >= nmain) ? (L.csrc ? cc[j] : 0.0f) : r[j];
      if (tma_c) in_release(t);
      write_a16(tl, c0, a);
    } else {
      write_a16(tl, c0, r);
    }
  }
  // ---- results to HBM
  if (tma_save) { slot_write16(out_acquire(t), t.rowoff, t.sw, r); out_publish(t); }
  else if (st_save) row_store16(L.save, L.ld_save, row, c0, nmain - c0, row_ok, r);
  if constexpr (KIND == EK_TANGENT) {
    if (tma_out2) { slot_write16(out_acquire(t), t.rowoff, t.sw, q2); out_publish(t); }
    else if (st_out2) row_store16(L.out2, L.ldo2, row, c0, nmain - c0, row_ok, q2);
  }
}

// all units of one layer for this thread's team
template <int KIND, typename Hook, typename PreWrite>
__device__ __forceinline__ void epi_layer(const ChainParamsDev& p, int l, TeamCtx& t, int u0, uint32_t tl, int tile, int rows_valid,
                                          bool tile_tma_ok, const float* s_bias, Hook&& before_unit, PreWrite&& before_a_write) {
  constexpr bool kBias = (KIND == EK_BIAS_SOFTPLUS || KIND == EK_BIAS_RELU || KIND == EK_BIAS_GENERIC);
  const ChainLayerDev& L = p.L[l];
  const int nmain = kBias ? L.ncol_out : min(L.ncol_out, L.ncol_main);
  // units [0, n_fast) take the fast path: completely inside the valid columns, every operand they touch through TMA
  bool fast_ok = tile_tma_ok && rows_valid == CH_BM;
  if (L.save && !(L.out_ops & (1 << OP_SAVE))) fast_ok = false;
  if (!kBias) {
    if (KIND != EK_DACT_NONE && !(L.in_ops & (1 << OP_H))) fast_ok = false;
    if (L.aux2 && !(L.in_ops & (1 << OP_AUX2))) fast_ok = false;
  }
  if (KIND == EK_TANGENT && !(L.out_ops & (1 << OP_OUT2))) fast_ok = false;
  const int n_fast = fast_ok ? L.n_full : 0;
  int u = u0;
  for (; u < n_fast; u += kTeams) { before_unit(u); epi_unit<KIND, true>(p, l, t, tl, u, tile, rows_valid, tile_tma_ok, s_bias, before_a_write); }
  for (; u < L.n_units; u += kTeams) { before_unit(u); epi_unit<KIND, false>(p, l, t, tl, u, tile, rows_valid, tile_tma_ok, s_bias, before_a_write); }
}

// FAM: 0 = bias/activation epilogues (forward chains), 1 = derivative-product epilogues (gradient sweeps), 2 = tangent
// sweep.  One instantiation per family keeps the register pressure of each kernel low.
template <int FAM>
__global__ void __launch_bounds__(kChThreads, 1) umma_chain_kernel(const __grid_constant__ ChainParamsDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + kWStages * kWStageBytes;
  float* s_bias2 = reinterpret_cast<float*>(s_slots + kSlotsBytes);   // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_slots + kSlotsBytes + kChBiasBytes);
  uint64_t* full = bars;                               // [kWStages]  W stage landed
  uint64_t* empty = bars + kWStages;                   // [kWStages]  W stage consumed
  uint64_t* a_ready = bars + 2 * kWStages;             // [2] next A operand written for K < 128 (+ accumulator half 0 drained) / all of it
  uint64_t* acc_ready = bars + 2 * kWStages + 2;       // [2] accumulator N half 0 / half 1 of the current layer complete
  uint64_t* a_free = bars + 2 * kWStages + 4;          // the MMAs of this layer no longer read A columns K < 128
  uint64_t* b_in_full = bars + 2 * kWStages + 5;       // [kTeams][kSlotsIn]
  uint64_t* b_in_empty = b_in_full + kTeams * kSlotsIn;
  uint64_t* b_out_full = b_in_empty + kTeams * kSlotsIn;   // [kTeams][kSlotsOut]
  uint64_t* b_out_empty = b_out_full + kTeams * kSlotsOut;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_out_empty + kTeams * kSlotsOut);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = p.m_ptr ? *p.m_ptr : p.m_cap;
  if (M > p.m_cap) M = p.m_cap;
  const int num_tiles = (M + CH_BM - 1) / CH_BM;
  const int tile_step = gridDim.x;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("nero: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_ready, kChEpiWarps);
    mbar_init(a_ready + 1, kChEpiWarps);
    mbar_init(acc_ready, 1);
    mbar_init(acc_ready + 1, 1);
    mbar_init(a_free, 1);
    for (int i = 0; i < kTeams * kSlotsIn; ++i) { mbar_init(&b_in_full[i], 1); mbar_init(&b_in_empty[i], 4); }
    for (int i = 0; i < kTeams * kSlotsOut; ++i) { mbar_init(&b_out_full[i], 4); mbar_init(&b_out_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == kChLoadWarp) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kChEpiWarps) {
    // ============================== epilogue teams
    const int team = warp >> 2;
    TeamCtx t;
    t.q = warp & 3; t.lane = lane;
    t.slots_in = s_slots + team * (kSlotsIn + kSlotsOut) * kSlotBytes;
    t.slots_out = t.slots_in + kSlotsIn * kSlotBytes;
    t.in_full = b_in_full + team * kSlotsIn;
    t.in_empty = b_in_empty + team * kSlotsIn;
    t.out_full = b_out_full + team * kSlotsOut;
    t.out_empty = b_out_empty + team * kSlotsOut;
    t.in_seq = 0; t.out_seq = 0;
    const int rl = t.q * 32 + lane;
    t.rowoff = uint32_t(rl) * 64u;
    t.sw = uint32_t((rl >> 1) & 3) << 4;
    const uint32_t tl = tmem_base + (uint32_t(t.q * 32) << 16);
    uint32_t acc_phase = 0, layer_seq = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
      const int rows_valid = min(CH_BM, M - tile * CH_BM);
      const long row = long(tile) * CH_BM + rl;
      const bool row_ok = rl < rows_valid;
      const bool tile_tma_ok = tile_tma_store_ok(p, tile, M);
      // ---- first A operand: fp32 rows from HBM -> split-bf16 in TMEM
      for (int u = team; u < ((kExp & 8) ? 0 : p.a0_units); u += kTeams) {
        float x[16];
        if (p.a0_tma) {
          slot_read16(in_wait(t), t.rowoff, t.sw, x);
          write_a16(tl, u * 16, x);
          in_release(t);
        } else {
          row_load16(p.A0, p.lda0, row, u * 16, p.k_valid0 - u * 16, row_ok, x);
          write_a16(tl, u * 16, x);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a_ready); mbar_arrive(a_ready + 1); }
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayerDev& L = p.L[l];
        // bias of this layer -> one of two smem buffers: written while the MMAs run; the named barrier below orders it
        // against the reads (two buffers + one barrier per layer make the reuse race-free).  The buffers alternate with a
        // counter that runs ACROSS tiles: indexed by (l & 1), a chain with an odd number of layers used the same buffer for
        // the last layer of one tile and the first layer of the next, and a warp that was a whole A0 conversion ahead
        // overwrote the bias the slower warps were still adding (seen as rare wrong SDF values of the 9-layer sampling
        // chain when the weight images were cold in L2; tools/stress_chain.py)
        float* sb = s_bias2 + (layer_seq & 1u) * 256;
        ++layer_seq;
        {
          const int i = warp * 32 + lane;
          if (i < 256) sb[i] = (L.bias && i < L.n_bias) ? __ldg(L.bias + i) : 0.0f;
        }
        if (!(kExp & 2)) mbar_wait(acc_ready, acc_phase);
        tcgen05_fence_after();
        asm volatile("bar.sync 1, %0;" ::"n"(kChEpiWarps * 32));
        // The layer runs as two N halves (k_umma: half 0 = accumulator columns [0, h_units*16)): the units of half 0 are
        // processed while the tensor core still works on half 1, and once this warp has written its part of the next A
        // operand's columns K < 128 the next layer's first MMAs may start (a_ready[0]) under the rest of this epilogue.
        const int h_units = (L.n_pad > 128 ? L.n_pad / 2 : L.n_pad) >> 4;
        const bool hand_over = l + 1 < p.n_layers;
        bool got_acc1 = false, got_free = false, lo_sent = false;
        auto before_unit = [&](int u) {
          if (u >= h_units && !got_acc1) { if (!(kExp & 2)) mbar_wait(acc_ready + 1, acc_phase); tcgen05_fence_after(); got_acc1 = true; }
          if (u >= 8 && !lo_sent) {      // this warp's share of A columns K < 128 is complete (units ascend)
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0 && hand_over) mbar_arrive(a_ready);
            lo_sent = true;
          }
        };
        // the next A operand (columns K < 128) is written over the one the half-1 MMAs of THIS layer still read: wait for
        // them right before the first such write (by then the unit's own math has covered most of that time)
        auto before_a_write = [&](int u) {
          if (u < 8 && !got_acc1 && !got_free) { if (!(kExp & 2)) mbar_wait(a_free, acc_phase); tcgen05_fence_after(); got_free = true; }
        };
        const int u0 = (team + l) % kTeams;          // rotate the unit -> team assignment (evens out an uneven split of the 16 units)
#define NERO_EPI_CALL(K) epi_layer<K>(p, l, t, u0, tl, tile, rows_valid, tile_tma_ok, sb, before_unit, before_a_write)
        if constexpr ((kExp & 8) != 0) {
        } else if constexpr (FAM == 0) {
          if (L.kind == EK_BIAS_SOFTPLUS) NERO_EPI_CALL(EK_BIAS_SOFTPLUS);
          else if (L.kind == EK_BIAS_RELU) NERO_EPI_CALL(EK_BIAS_RELU);
          else NERO_EPI_CALL(EK_BIAS_GENERIC);
        } else if constexpr (FAM == 1) {
          if (L.kind == EK_DACT_SOFTPLUS) NERO_EPI_CALL(EK_DACT_SOFTPLUS);
          else if (L.kind == EK_DACT_RELU) NERO_EPI_CALL(EK_DACT_RELU);
          else NERO_EPI_CALL(EK_DACT_NONE);
        } else {
          NERO_EPI_CALL(EK_TANGENT);
        }
#undef NERO_EPI_CALL
        // every barrier of the layer is passed exactly once per thread (phase bookkeeping), all MMAs of the layer are done
        if (!got_acc1 && !(kExp & 2)) { mbar_wait(acc_ready + 1, acc_phase); tcgen05_fence_after(); }
        if (!got_free && !(kExp & 2)) mbar_wait(a_free, acc_phase);
        acc_phase ^= 1;
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        // the last layer of a tile hands over nothing: the next tile's A operand is announced after its conversion
        if (lane == 0 && hand_over) {
          if (!lo_sent) mbar_arrive(a_ready);
          mbar_arrive(a_ready + 1);
        }
      }
    }
  } else if (warp == kChMmaWarp) {
    // ============================== MMA issuer
    int g = 0;
    uint32_t a_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayerDev& L = p.L[l];
        const int nh = L.n_pad > 128 ? 2 : 1;
        const uint32_t half_rows = uint32_t(L.n_pad) / nh;
        const uint32_t idesc = make_idesc_bf16(CH_BM, half_rows);
        const uint32_t b_plane = half_rows * 128u;
        const int kc = L.k_chunks;
        bool hi_ok = false;
        for (int hf = 0; hf < nh; ++hf) {
          const uint32_t d_col = tmem_base + kAccCol + uint32_t(hf) * half_rows;
          for (int c = 0; c < kc; ++c, ++g) {
            if (hf == 0 && c == 0 && !(kExp & 1)) { mbar_wait(a_ready, a_phase); tcgen05_fence_after(); }           // A columns K < 128, half 0 drained
            // all of A written, half 1 drained.  Also required before this layer's a_free / acc_ready[1] commits are
            // issued: every epilogue thread passes the previous layer's phase of those barriers before it arrives on
            // a_ready[1], so waiting here keeps a barrier from completing two phases ahead of a waiter (single-half
            // layers would otherwise commit them after a_ready[0] alone).
            const bool commits_late = (hf == nh - 1) && (c == (kc < 2 ? kc : 2) - 1 || c == kc - 1);
            if (!hi_ok && (c == 2 || hf == 1 || commits_late)) { if (!(kExp & 1)) mbar_wait(a_ready + 1, a_phase); tcgen05_fence_after(); hi_ok = true; }
            const int s = g % kWStages;
            mbar_wait(&full[s], (g / kWStages) & 1);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t b_hi = smem_u32(s_w + s * kWStageBytes);
              const uint32_t b_lo = b_hi + b_plane;
#pragma unroll
              for (int k = 0; k < ((kExp & 4) ? 0 : CH_BK / 16); ++k) {
                const uint32_t a_col = uint32_t(c * 32 + k * 8);
                const uint64_t dbh = make_desc_k_sw128(b_hi + k * 32), dbl = make_desc_k_sw128(b_lo + k * 32);
                umma_bf16_ts(d_col, tmem_base + kALoCol + a_col, dbh, idesc, (c | k) != 0);
                umma_bf16_ts(d_col, tmem_base + kAHiCol + a_col, dbl, idesc, 1);
                umma_bf16_ts(d_col, tmem_base + kAHiCol + a_col, dbh, idesc, 1);
              }
              umma_commit(&empty[s]);
              const bool last_c = c == kc - 1;
              if (hf == 0 && last_c) umma_commit(acc_ready);
              // A columns K < 128 are dead once the last half has passed its K chunks 0 and 1
              if (hf == nh - 1 && c == (kc < 2 ? kc : 2) - 1) umma_commit(a_free);
              if (hf == nh - 1 && last_c) umma_commit(acc_ready + 1);
            }
            __syncwarp();
          }
        }
        if (!hi_ok && !(kExp & 1)) { mbar_wait(a_ready + 1, a_phase); tcgen05_fence_after(); }      // keep the phases in step
        a_phase ^= 1;
      }
    }
  } else if (warp == kChLoadWarp) {
    // ============================== W loader
    int g = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayerDev& L = p.L[l];
        const int nh = L.n_pad > 128 ? 2 : 1;
        const uint32_t half_rows = uint32_t(L.n_pad) / nh;
        const uint32_t plane = uint32_t(L.n_pad) * 128u, hp = half_rows * 128u;
        for (int hf = 0; hf < nh; ++hf) {
          for (int c = 0; c < L.k_chunks; ++c, ++g) {
            const int s = g % kWStages;
            mbar_wait(&empty[s], ((g / kWStages) & 1) ^ 1);
            if (elect_one()) {
              // rows [hf*half_rows, +half_rows) of the chunk's hi plane and of its lo plane
              const uint8_t* src = L.wimg + size_t(c) * 2u * plane + size_t(hf) * hp;
              mbar_arrive_expect_tx(&full[s], (kExp & 16) ? hp : 2u * hp);
              bulk_copy_g2s(s_w + s * kWStageBytes, src, hp, &full[s]);
              if (!(kExp & 16)) bulk_copy_g2s(s_w + s * kWStageBytes + hp, src + plane, hp, &full[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kChAuxWarp) {
    // ============================== aux loader: lane t fills the input FIFO of team t, as far ahead as slots are free
    if (lane < kTeams && !(kExp & 8)) {
      const int tm = lane;
      uint64_t* fb = b_in_full + tm * kSlotsIn;
      uint64_t* eb = b_in_empty + tm * kSlotsIn;
      uint8_t* slots = s_slots + tm * (kSlotsIn + kSlotsOut) * kSlotBytes;
      uint32_t seq = 0;
      auto issue = [&](const CUtensorMap* map, int c0, int r0) {
        const uint32_t s_ = seq % kSlotsIn;
        mbar_wait(&eb[s_], ((seq / kSlotsIn) & 1) ^ 1);          // the team has read the previous contents
        mbar_arrive_expect_tx(&fb[s_], kSlotBytes);
        tma_load_2d(slots + s_ * kSlotBytes, map, c0, r0, &fb[s_]);
        ++seq;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
        const int r0 = tile * CH_BM;
        if (p.a0_tma)
          for (int u = tm; u < p.a0_units; u += kTeams) issue(&p.maps[0], u * 16, r0);
        for (int l = 0; l < p.n_layers; ++l) {
          const ChainLayerDev& L = p.L[l];
          if (!L.in_ops) continue;
          const CUtensorMap* mh = (L.in_ops & (1 << OP_H)) ? &p.maps[L.map[OP_H]] : nullptr;
          const CUtensorMap* m2 = (L.in_ops & (1 << OP_AUX2)) ? &p.maps[L.map[OP_AUX2]] : nullptr;
          const CUtensorMap* mc = (L.in_ops & (1 << OP_CSRC)) ? &p.maps[L.map[OP_CSRC]] : nullptr;
          int u = (tm + l) % kTeams;
          for (; u < L.n_full; u += kTeams) {
            if (mh) issue(mh, u * 16, r0);
            if (m2) issue(m2, u * 16, r0);
          }
          if (mc)      // the skip-concat source of the units at and beyond the last valid column
            for (; u < L.n_units; u += kTeams) issue(mc, u * 16, r0);
        }
      }
    }
  } else {
    // ============================== store warp: lane t drains the output FIFO of team t with TMA stores
    if (lane < kTeams && !(kExp & 8)) {
      const int tm = lane;
      uint64_t* fb = b_out_full + tm * kSlotsOut;
      uint64_t* eb = b_out_empty + tm * kSlotsOut;
      uint8_t* slots = s_slots + (tm * (kSlotsIn + kSlotsOut) + kSlotsIn) * kSlotBytes;
      uint32_t seq = 0;
      auto issue = [&](const CUtensorMap* map, int c0, int r0) {
        const uint32_t s_ = seq % kSlotsOut;
        mbar_wait(&fb[s_], (seq / kSlotsOut) & 1);                 // all four warps of the team have written their rows
        tma_store_2d(map, slots + s_ * kSlotBytes, c0, r0);
        tma_commit();
        if (seq > 0) {                                             // the previous store has read its slot: hand it back
          tma_wait_read<1>();
          mbar_arrive(&eb[(seq - 1) % kSlotsOut]);
        }
        ++seq;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
        if (!tile_tma_store_ok(p, tile, M)) continue;
        const int r0 = tile * CH_BM;
        for (int l = 0; l < p.n_layers; ++l) {
          const ChainLayerDev& L = p.L[l];
          if (!L.out_ops) continue;
          const CUtensorMap* ms = (L.out_ops & (1 << OP_SAVE)) ? &p.maps[L.map[OP_SAVE]] : nullptr;
          const CUtensorMap* m2 = (L.out_ops & (1 << OP_OUT2)) ? &p.maps[L.map[OP_OUT2]] : nullptr;
          for (int u = (tm + l) % kTeams; u < L.n_full; u += kTeams) {
            if (ms) issue(ms, u * 16, r0);
            if (m2) issue(m2, u * 16, r0);
          }
        }
      }
      tma_wait_all();      // outstanding TMA stores must complete before the CTA exits
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kChLoadWarp) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- host side
// a [rows x cols] fp32 window (leading dimension ld) as a 2-D tensor map with [128 x 16] SWIZZLE_64B boxes
static bool tma_eligible(const float* ptr, int ld, int cols) { return tma_eligible_f32(ptr, ld, cols); }
static int make_map(CUtensorMap* out, const float* ptr, int ld, int rows, int cols) {
  return tma_make_map_f32(out, ptr, ld, rows, cols, 16, CH_BM, 64);
}

int chain_dispatch(const ChainParams& p, cudaStream_t stream) {
  if (p.n_layers <= 0 || p.n_layers > kMaxChainLayers || (p.lda0 & 3) || p.k_valid0 > 256 || !p.A0) return NERO_ERR_ARG;
  for (int l = 0; l < p.n_layers; ++l) {
    const ChainLayer& L = p.L[l];
    if (L.n_pad % 16 || L.n_pad < 16 || L.n_pad > 256 || L.k_chunks < 1 || L.k_chunks > 4 || L.ncol_out > L.n_pad) return NERO_ERR_ARG;
    if (L.n_pad > 128 && (L.n_pad / 2) % 16) return NERO_ERR_ARG;
    if (L.kind == EK_TANGENT && (!L.H || !L.V || !L.out2)) return NERO_ERR_ARG;
    if ((L.kind == EK_DACT_SOFTPLUS || L.kind == EK_DACT_RELU) && !L.H) return NERO_ERR_ARG;
  }
  int fam = -1;
  for (int l = 0; l < p.n_layers; ++l) {
    const int k = p.L[l].kind;
    const int f = k <= EK_BIAS_GENERIC ? 0 : (k == EK_TANGENT ? 2 : 1);
    if (fam >= 0 && f != fam) return NERO_ERR_ARG;   // a chain is homogeneous: forward, gradient sweep or tangent sweep
    fam = f;
  }
  const int tiles_cap = (p.m_cap + CH_BM - 1) / CH_BM;
  if (tiles_cap <= 0) return NERO_OK;

  static ChainParamsDev d;       // (host calls are serialised by the Python GIL; the struct is copied at launch)
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  d.A0 = p.A0; d.lda0 = p.lda0; d.k_valid0 = p.k_valid0; d.n_layers = p.n_layers; d.m_ptr = p.m_ptr; d.m_cap = p.m_cap;
  d.a0_units = p.L[0].k_chunks * 4;           // every column the first layer's MMAs read (zeros beyond k_valid0)
  int nmaps = 1;
  d.a0_tma = tma_eligible(p.A0, p.lda0, p.k_valid0) ? 1 : 0;
  if (d.a0_tma && make_map(&d.maps[0], p.A0, p.lda0, p.m_cap, p.k_valid0) != NERO_OK) return NERO_ERR_CUDA;
  for (int l = 0; l < p.n_layers; ++l) {
    const ChainLayer& s = p.L[l];
    ChainLayerDev& o = d.L[l];
    o.wimg = s.wimg; o.bias = s.bias; o.save = s.save; o.H = s.H; o.out2 = s.out2; o.tail = s.tail; o.csrc = s.csrc;
    o.n_pad = s.n_pad; o.k_chunks = s.k_chunks; o.n_bias = s.n_bias; o.ncol_out = s.ncol_out; o.ncol_main = s.ncol_main; o.kind = s.kind; o.act = s.act;
    o.ld_save = s.ld_save; o.ldh = s.ldh; o.ldo2 = s.ldo2; o.ldt = s.ldt; o.ld_csrc = s.ld_csrc;
    o.oscale = s.oscale; o.hscale = s.hscale; o.act_param = s.act_param; o.write_a = s.write_a; o.a_blocks = s.a_blocks;
    const bool tangent = s.kind == EK_TANGENT;
    o.aux2 = tangent ? s.V : (s.kind >= EK_DACT_SOFTPLUS ? s.addend : nullptr);
    o.ld2 = tangent ? s.ldv : s.ldadd;
    o.aux2_is_addend = tangent ? 0 : 1;
    const int nblk = s.n_pad / 16;
    const int nblk_a = s.write_a ? (s.a_blocks > nblk ? s.a_blocks : nblk) : 0;
    o.n_units = nblk > nblk_a ? nblk : nblk_a;
    const bool bias_kind = s.kind <= EK_BIAS_GENERIC;
    const int nmain = bias_kind ? s.ncol_out : (s.ncol_out < s.ncol_main ? s.ncol_out : s.ncol_main);
    o.tma = 0;
    struct { int op; const float* ptr; int ld; int cols; bool use; } ops[OP_COUNT] = {
        {OP_H, s.H, s.ldh, nmain, !bias_kind && s.kind != EK_DACT_NONE},
        {OP_AUX2, o.aux2, o.ld2, nmain, !bias_kind},
        {OP_CSRC, s.csrc, s.ld_csrc, 256, s.write_a != 0},
        {OP_SAVE, s.save, s.ld_save, nmain, true},
        {OP_OUT2, s.out2, s.ldo2, nmain, tangent}};
    for (auto& q : ops) {
      o.map[q.op] = 0;
      if (!q.use || !tma_eligible(q.ptr, q.ld, q.cols)) continue;
      if (make_map(&d.maps[nmaps], q.ptr, q.ld, p.m_cap, q.cols) != NERO_OK) return NERO_ERR_CUDA;
      o.map[q.op] = nmaps++;
      o.tma |= 1 << q.op;
    }
    const int ncomp = nmain < s.ncol_out ? nmain : s.ncol_out;
    o.n_full = ncomp / 16;
    o.in_ops = (o.n_full ? (o.tma & ((1 << OP_H) | (1 << OP_AUX2))) : 0) | (o.tma & (1 << OP_CSRC));
    o.out_ops = o.n_full ? (o.tma & ((1 << OP_SAVE) | (1 << OP_OUT2))) : 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(umma_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(umma_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  const int grid = tiles_cap < kNumSMs ? tiles_cap : kNumSMs;
  if (fam == 0) umma_chain_kernel<0><<<grid, kChThreads, kChSmemBytes, stream>>>(d);
  else if (fam == 1) umma_chain_kernel<1><<<grid, kChThreads, kChSmemBytes, stream>>>(d);
  else umma_chain_kernel<2><<<grid, kChThreads, kChSmemBytes, stream>>>(d);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

}  // namespace nero
