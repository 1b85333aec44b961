// Fused MLP-CHAIN kernel on tcgen05: a 128-row tile of samples flows through a whole sequence of linear layers
// without its activations ever leaving the SM between layers.
//
//   TMEM (512 columns x 128 lanes x 32 bit, all of it):
//     [  0,256)  fp32 accumulator of the current layer (D operand)
//     [256,384)  A operand, bf16 "hi" plane, two K elements per 32-bit column  (A read by tcgen05.mma FROM TMEM)
//     [384,512)  A operand, bf16 "lo" plane                                    (split-bf16: x ~= hi + lo)
//   Per layer: the MMA warp issues, per 16-wide k-step, A_lo*W_hi + A_hi*W_lo + A_hi*W_hi into the accumulator;
//   twelve epilogue warps then read the accumulator (tcgen05.ld), apply bias/activation or the derivative products
//   of the gradient sweeps, store what the backward pass needs to HBM (coalesced through a per-warp transpose
//   buffer) and write the NEXT layer's A operand straight back into TMEM (tcgen05.st) as packed split-bf16.
//   Weights stream from L2 as pre-swizzled images (one cp.async.bulk per 64-wide K chunk, 3-stage mbarrier ring).
//   Aux operands of the derivative epilogues (H, addend, V) are prefetched into L2 one layer ahead.
//
// This realises SURVEY.md K2/K3/K4/K10/K13 "activations stay on-chip across layers": compared with one
// nero_linear launch per layer it removes the A-operand HBM round trip (1 KB/sample/layer) and the producer
// load->convert->store latency chain, which the ncu profiles of round 1 (profiles/r01b_*) show to dominate.
//
// Skip connection of the SDF network (field.py:139-140): a layer with `concat` set takes columns >= ncol_out of the
// next A operand from its own `save` buffer, where ray_fill / pe_tangent pre-stored PE/sqrt2 resp. its tangent.
#include "umma_common.cuh"

#ifndef NERO_EARLY_ISSUE
#define NERO_EARLY_ISSUE 0   // issue all aux-operand loads of a block before reading the accumulator
#endif

namespace nero {

constexpr int kMaxChainLayers = 10;

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

constexpr int CH_BM = 128, CH_BK = 64;
constexpr int kChEpiWarps = 12;
constexpr int kChMmaWarp = kChEpiWarps, kChLoadWarp = kChEpiWarps + 1;
constexpr int kChThreads = (kChEpiWarps + 2) * 32;
#ifndef NERO_EPI_V2
#define NERO_EPI_V2 1
#endif
// NERO_NSPLIT=1 (needs the v2 epilogue): layers wider than 128 columns run as two N halves so that the epilogue of half 0
// overlaps the MMAs of half 1.  Measured on B200 (round 1): 15.6 -> 16.0-16.5 ms/step, i.e. a LOSS -- the epilogue of half
// 0 may not overwrite the A operand in TMEM before the half-1 MMAs have read it (a_free barrier), N=128 MMAs have less
// reuse of the A operand, and the epilogue, not the MMA, is the longer phase.  Kept for experiments, off by default.
#ifndef NERO_NSPLIT
#define NERO_NSPLIT 0
#endif
#if NERO_NSPLIT
constexpr int kChStages = 6;
constexpr uint32_t kChStageBytes = 2 * 128 * 128;                       // half a W chunk: hi + lo planes of up to 128 rows
#else
constexpr int kChStages = 3;
constexpr uint32_t kChStageBytes = 2 * 256 * 128;                       // W chunk: hi + lo planes of up to 256 rows
#endif
constexpr uint32_t kChEpiBytes = kChEpiWarps * kStageWarpBytes;         // 30 KB
constexpr uint32_t kChBiasBytes = 2 * 256 * 4;
constexpr uint32_t kChSmemBytes = kChStages * kChStageBytes + kChEpiBytes + kChBiasBytes + 1024 + 256;
constexpr uint32_t kAccCol = 0, kAHiCol = 256, kALoCol = 384;

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

template <int KIND>
__device__ __forceinline__ void chain_epilogue_layer(const ChainLayer& L, uint32_t tmem_lane_base, int row0, int rows_valid,
                                                     int third, int lane, float* stg, const float* s_bias) {
  constexpr bool kBias = (KIND == EK_BIAS_SOFTPLUS || KIND == EK_BIAS_RELU || KIND == EK_BIAS_GENERIC);
  const int nblk = L.n_pad >> 4;
  const int nmain = kBias ? L.ncol_out : min(L.ncol_out, L.ncol_main);
  const bool v_save = L.save && vec_ok(L.save, L.ld_save);
  // when this layer feeds the next one, every 16-column block the next layer's MMAs read must be defined
  const int nblk_a = L.write_a ? max(L.a_blocks, nblk) : 0;
  const int nb_loop = max(nblk, nblk_a);
#pragma unroll 1
  for (int b = third; b < nb_loop; b += 3) {
    const int c0 = b * 16;
    float r[16];
    if (b < nblk && c0 < L.ncol_out) {
      float v[16];
      const int cm = nmain - c0;   // main columns in this block (may be <= 0)
#if NERO_EARLY_ISSUE
      float4 hraw[4], araw[4], vraw[4];
      if constexpr (!kBias) {      // all aux loads of this block are in flight before the accumulator is read
        if constexpr (KIND != EK_DACT_NONE) issue_block16(L.H + size_t(row0) * L.ldh + c0, L.ldh, rows_valid, cm, vec_ok(L.H, L.ldh), lane, hraw);
        if (L.addend) issue_block16(L.addend + size_t(row0) * L.ldadd + c0, L.ldadd, rows_valid, cm, vec_ok(L.addend, L.ldadd), lane, araw);
        if constexpr (KIND == EK_TANGENT) issue_block16(L.V + size_t(row0) * L.ldv + c0, L.ldv, rows_valid, cm, vec_ok(L.V, L.ldv), lane, vraw);
      }
#endif
      tmem_ld16(tmem_lane_base + kAccCol + c0, v);
      tmem_ld_wait();
      if constexpr (kBias) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += s_bias[c0 + j];
        if constexpr (KIND == EK_BIAS_SOFTPLUS) softplus100_fast16(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float y = v[j];
          if constexpr (KIND == EK_BIAS_RELU) y = fmaxf(y, 0.0f);
          else if constexpr (KIND == EK_BIAS_GENERIC) y = apply_act(y, L.act, L.act_param);
          r[j] = L.oscale * y;
        }
      } else {
        float s[16];
        if constexpr (KIND == EK_DACT_NONE) {
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = 1.0f;
        } else {
#if NERO_EARLY_ISSUE
          finish_block16(hraw, stg, lane, s);
#else
          load_block16(L.H + size_t(row0) * L.ldh + c0, L.ldh, rows_valid, cm, vec_ok(L.H, L.ldh), stg, lane, s);
#endif
          if constexpr (KIND == EK_DACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = s[j] > 0.0f ? 1.0f : 0.0f;
          } else {
            dsoftplus100_from_h_fast16(s, L.hscale);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = L.oscale * s[j] * v[j];
        if (L.addend) {
          float a[16];
#if NERO_EARLY_ISSUE
          finish_block16(araw, stg, lane, a);
#else
          load_block16(L.addend + size_t(row0) * L.ldadd + c0, L.ldadd, rows_valid, cm, vec_ok(L.addend, L.ldadd), stg, lane, a);
#endif
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] += a[j];
        }
        if (L.tail && c0 + 16 > L.ncol_main) {
          const int shift = max(L.ncol_main - c0, 0);
          const int ncols = min(16, L.ncol_out - c0);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j >= shift && j < ncols && lane < rows_valid)
              L.tail[size_t(row0 + lane) * L.ldt + (c0 + j - L.ncol_main)] = L.oscale * v[j];
        }
        if constexpr (KIND == EK_TANGENT) {
          float vv[16];
#if NERO_EARLY_ISSUE
          finish_block16(vraw, stg, lane, vv);
#else
          load_block16(L.V + size_t(row0) * L.ldv + c0, L.ldv, rows_valid, cm, vec_ok(L.V, L.ldv), stg, lane, vv);
#endif
          if (cm > 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) vv[j] = 100.0f * (1.0f - s[j]) * vv[j] * v[j];
            store_block16(L.out2 + size_t(row0) * L.ldo2 + c0, L.ldo2, rows_valid, cm, vec_ok(L.out2, L.ldo2), stg, lane, vv);
          }
        }
      }
      if (L.save && nmain - c0 > 0) store_block16(L.save + size_t(row0) * L.ld_save + c0, L.ld_save, rows_valid, nmain - c0, v_save, stg, lane, r);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = 0.0f;
    }
    if (b < nblk_a) {
      // columns >= nmain of the next A operand: the skip-concat source (from the save buffer) or zero
      if (c0 + 16 > nmain) {
        float cc[16];
        if (L.csrc) load_block16(L.csrc + size_t(row0) * L.ld_csrc + c0, L.ld_csrc, rows_valid, 256 - c0, vec_ok(L.csrc, L.ld_csrc), stg, lane, cc);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j >= nmain) r[j] = L.csrc ? cc[j] : 0.0f;
      }
      write_a16(tmem_lane_base, c0, r);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue v2: accumulator fragments in the 16x256b TMEM load shape (thread t of a warp holds, per 8-column repeat,
// columns 2(t%4), 2(t%4)+1 of rows t/4 and t/4+8 -- the m16n8 fragment).  A quad then covers 32 contiguous bytes of one
// matrix row, so aux operands (H, addend, V) are read and results written DIRECTLY from/to global memory with full
// 32-byte sectors: no shared-memory transposes, no warp syncs, and the loads are in flight before the accumulator
// read completes.  The packed split-bf16 A operand of the next layer is one 32-bit word per (row, column pair) = the
// 16x128b store shape with the same thread <-> (row, column) map.
#ifndef NERO_EPI_V2
#define NERO_EPI_V2 1
#endif

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_16x128b_x4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// element e = 4j + 2hh + q of a [16 rows x 32 cols] unit  <->  row rlo + 8hh, column c0 + 8j + 2a + q
struct FragPos { int rlo; int a; bool plo, phi; };

__device__ __forceinline__ bool vec2_ok(const float* p, int ld) {
  return ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(p) & 7) == 0);
}
// g: matrix origin (row 0, first column of the layer window); columns >= ncols and masked rows read as zero
__device__ __forceinline__ void frag_load(const float* __restrict__ g, int ld, const FragPos& f, int c0, int ncols, bool v2, float* x) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = c0 + 8 * j + 2 * f.a;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float2 val = make_float2(0.f, 0.f);
      if ((hh ? f.phi : f.plo) && col < ncols) {
        const float* p = g + size_t(f.rlo + 8 * hh) * ld + col;
        if (v2 && col + 1 < ncols) val = *reinterpret_cast<const float2*>(p);
        else { val.x = p[0]; if (col + 1 < ncols) val.y = p[1]; }
      }
      x[4 * j + 2 * hh] = val.x; x[4 * j + 2 * hh + 1] = val.y;
    }
  }
}
__device__ __forceinline__ void frag_store(float* __restrict__ g, int ld, const FragPos& f, int c0, int ncols, bool v2, const float* x) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = c0 + 8 * j + 2 * f.a;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if ((hh ? f.phi : f.plo) && col < ncols) {
        float* p = g + size_t(f.rlo + 8 * hh) * ld + col;
        if (v2 && col + 1 < ncols) *reinterpret_cast<float2*>(p) = make_float2(x[4 * j + 2 * hh], x[4 * j + 2 * hh + 1]);
        else { p[0] = x[4 * j + 2 * hh]; if (col + 1 < ncols) p[1] = x[4 * j + 2 * hh + 1]; }
      }
    }
  }
}
// next layer's A operand: 32 columns of this unit as packed split-bf16 (16 packed columns per plane)
__device__ __forceinline__ void frag_write_a(uint32_t tl, int c0, const float* r) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    split2(r[4 * j], r[4 * j + 1], hi[2 * j], lo[2 * j]);              // row t/4
    split2(r[4 * j + 2], r[4 * j + 3], hi[2 * j + 1], lo[2 * j + 1]);  // row t/4 + 8
  }
  tmem_st_16x128b_x4(tl + kAHiCol + (c0 >> 1), hi);
  tmem_st_16x128b_x4(tl + kALoCol + (c0 >> 1), lo);
}

template <int KIND>
__device__ __forceinline__ void chain_epilogue_layer2(const ChainLayer& L, uint32_t tmem_base, int lg, int grp_row0, int rows_valid,
                                                      int third, int lane, const float* s_bias, uint64_t* acc_ready1, uint64_t* a_free, uint32_t phase1) {
  constexpr bool kBias = (KIND == EK_BIAS_SOFTPLUS || KIND == EK_BIAS_RELU || KIND == EK_BIAS_GENERIC);
  const int nmain = kBias ? L.ncol_out : min(L.ncol_out, L.ncol_main);
  const int ncols_a = L.write_a ? max(L.a_blocks * 16, L.n_pad) : 0;     // columns of the next A operand to define
  const int n_units = ((max(L.n_pad, ncols_a) + 31) >> 5) * 2;
  const bool v2_save = L.save && vec2_ok(L.save, L.ld_save);
#pragma unroll 1
  for (int u = third; u < n_units; u += 3) {
    const int h = u & 1, c0 = (u >> 1) * 32;
    if (acc_ready1 && c0 >= 128) {      // second N half of the accumulator: its MMAs ran while half 0 was processed
      mbar_wait(acc_ready1, phase1);
      tcgen05_fence_after();
      acc_ready1 = nullptr;
    }
    const int rl = h * 16 + (lane >> 2);
    FragPos f;
    f.rlo = grp_row0 + rl; f.a = lane & 3; f.plo = rl < rows_valid; f.phi = rl + 8 < rows_valid;
    const uint32_t tl = tmem_base + (uint32_t(lg * 32 + h * 16) << 16);
    float r[16];
    if (c0 < L.ncol_out) {
      float v[16];
      if constexpr (kBias) {
        tmem_ld_16x256b_x4(tl + kAccCol + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 b = *reinterpret_cast<const float2*>(s_bias + ((c0 + 8 * j + 2 * f.a) & 255));
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.x; v[4 * j + 3] += b.y;
        }
        if constexpr (KIND == EK_BIAS_SOFTPLUS) softplus100_fast16(v);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float y = v[e];
          if constexpr (KIND == EK_BIAS_RELU) y = fmaxf(y, 0.0f);
          else if constexpr (KIND == EK_BIAS_GENERIC) y = apply_act(y, L.act, L.act_param);
          r[e] = y;
        }
        if (L.oscale != 1.0f) {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] *= L.oscale;
        }
      } else {
        float s[16];
        if constexpr (KIND != EK_DACT_NONE) frag_load(L.H, L.ldh, f, c0, nmain, vec2_ok(L.H, L.ldh), s);   // in flight during the TMEM read
        tmem_ld_16x256b_x4(tl + kAccCol + c0, v);
        tmem_ld_wait();
        if constexpr (KIND == EK_DACT_NONE) {
#pragma unroll
          for (int e = 0; e < 16; ++e) s[e] = 1.0f;
        } else if constexpr (KIND == EK_DACT_RELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) s[e] = s[e] > 0.0f ? 1.0f : 0.0f;
        } else {
          dsoftplus100_from_h_fast16(s, L.hscale);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = s[e] * v[e];
        if (L.oscale != 1.0f) {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] *= L.oscale;
        }
        if (L.addend) {
          float ad[16];
          frag_load(L.addend, L.ldadd, f, c0, nmain, vec2_ok(L.addend, L.ldadd), ad);
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] += ad[e];
        }
        if (L.tail && c0 + 32 > L.ncol_main) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int col = c0 + 8 * (e >> 2) + 2 * f.a + (e & 1);
            const bool pr = (e & 2) ? f.phi : f.plo;
            if (pr && col >= L.ncol_main && col < L.ncol_out)
              L.tail[size_t(f.rlo + ((e & 2) ? 8 : 0)) * L.ldt + (col - L.ncol_main)] = L.oscale * v[e];
          }
        }
        if constexpr (KIND == EK_TANGENT) {
          float vv[16];
          frag_load(L.V, L.ldv, f, c0, nmain, vec2_ok(L.V, L.ldv), vv);
#pragma unroll
          for (int e = 0; e < 16; ++e) vv[e] = 100.0f * (1.0f - s[e]) * vv[e] * v[e];
          frag_store(L.out2, L.ldo2, f, c0, nmain, vec2_ok(L.out2, L.ldo2), vv);
        }
      }
      if (L.save) frag_store(L.save, L.ld_save, f, c0, nmain, v2_save, r);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = 0.0f;
    }
    if (c0 < ncols_a) {
      if (c0 + 32 > nmain) {   // columns >= nmain: the skip-concat source (from the save buffer) or zero
        float cc[16];
        if (L.csrc) frag_load(L.csrc, L.ld_csrc, f, c0, 256, vec2_ok(L.csrc, L.ld_csrc), cc);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c0 + 8 * (e >> 2) + 2 * f.a + (e & 1);
          if (col >= nmain) r[e] = L.csrc ? cc[e] : 0.0f;
        }
      }
      // two-halves layers: the half-1 MMAs of THIS layer still read K < 128 of the current A operand while half 0 is
      // processed; its first k-chunks retire long before the first unit gets here, but the order must not rest on timing
      if (a_free && c0 < 128) {
        mbar_wait(a_free, phase1);
        tcgen05_fence_after();
        a_free = nullptr;
      }
      frag_write_a(tl, c0, r);
    }
  }
}

// FAM: 0 = bias/activation epilogues (forward chains), 1 = derivative-product epilogues (gradient sweeps), 2 = tangent
// sweep.  One instantiation per family keeps the register pressure of each kernel low enough for ptxas to overlap
// the sixteen independent MUFU chains of a block instead of serialising them through one register.
template <int FAM>
__global__ void __launch_bounds__(kChThreads, 1) umma_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_epi = reinterpret_cast<float*>(smem + kChStages * kChStageBytes);
  float* s_bias2 = reinterpret_cast<float*>(smem + kChStages * kChStageBytes + kChEpiBytes);   // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChStages * kChStageBytes + kChEpiBytes + kChBiasBytes);
  uint64_t* full = bars;                        // [stages]  W chunk landed
  uint64_t* empty = bars + kChStages;           // [stages]  W chunk consumed
  uint64_t* a_ready = bars + 2 * kChStages;     // A operand written + accumulator drained (12 epilogue warps)
  uint64_t* acc_ready = bars + 2 * kChStages + 1;   // [2]: accumulator columns of N half 0 / half 1 complete
  uint64_t* a_free = bars + 2 * kChStages + 3;      // half-1 MMAs are done reading K < 128 of the current A operand
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kChStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = p.m_ptr ? *p.m_ptr : p.m_cap;
  if (M > p.m_cap) M = p.m_cap;
  const int num_tiles = (M + CH_BM - 1) / CH_BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kChStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_ready, kChEpiWarps);
    mbar_init(acc_ready, 1);
    mbar_init(acc_ready + 1, 1);
    mbar_init(a_free, 1);
    fence_mbar_init();
  }
  if (warp == kChLoadWarp) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kChEpiWarps) {
    // ============================== epilogue warps
    const int lg = warp & 3, third = warp >> 2;
    float* stg = s_epi + warp * (32 * kStagePitch);
    const uint32_t tl = tmem_base + (uint32_t(lg * 32) << 16);
    uint32_t acc_phase = 0, acc_phase1 = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row0 = tile * CH_BM + lg * 32;
      const int rows_valid = min(32, M - row0);
      // ---- first A operand: fp32 rows from HBM -> split-bf16 in TMEM
      {
#if NERO_EPI_V2
        const bool v0 = vec2_ok(p.A0, p.lda0);
        const int nu0 = p.L[0].k_chunks * 4;      // [16 x 32] units: every column the first layer's MMAs read (zeros beyond k_valid0)
        for (int u = third; u < nu0; u += 3) {
          const int h = u & 1, c0 = (u >> 1) * 32;
          const int rl = h * 16 + (lane >> 2);
          FragPos f;
          f.rlo = row0 + rl; f.a = lane & 3; f.plo = rl < rows_valid; f.phi = rl + 8 < rows_valid;
          float x[16];
          frag_load(p.A0, p.lda0, f, c0, p.k_valid0, v0, x);
          frag_write_a(tmem_base + (uint32_t(lg * 32 + h * 16) << 16), c0, x);
        }
#else
        const bool v0 = vec_ok(p.A0, p.lda0);
        const int nb0 = p.L[0].k_chunks * 4;      // every column the first layer's MMAs read (zeros beyond k_valid0)
        for (int b = third; b < nb0; b += 3) {
          float x[16];
          load_block16(p.A0 + size_t(row0) * p.lda0 + b * 16, p.lda0, rows_valid, p.k_valid0 - b * 16, v0, stg, lane, x);
          write_a16(tl, b * 16, x);
        }
#endif
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayer& L = p.L[l];
        // prefetch the aux operands of THIS layer's epilogue into L2 while the MMAs run
        if (third == 0 && rows_valid > 0 && L.kind >= EK_DACT_SOFTPLUS) {
          const int nm = min(L.ncol_out, L.ncol_main);
          if (L.kind != EK_DACT_NONE) prefetch_rows_l2(L.H + size_t(row0) * L.ldh, L.ldh, rows_valid, nm, lane);
          prefetch_rows_l2(L.addend ? L.addend + size_t(row0) * L.ldadd : nullptr, L.ldadd, rows_valid, nm, lane);
          if (L.kind == EK_TANGENT) prefetch_rows_l2(L.V + size_t(row0) * L.ldv, L.ldv, rows_valid, nm, lane);
        }
        // bias of this layer -> smem buffer (l & 1): written while the MMAs run; the named barrier below orders it
        // against the reads (two buffers + one barrier per layer make the reuse race-free)
        float* sb = s_bias2 + (l & 1) * 256;
        if (warp < 8) {
          const int i = warp * 32 + lane;
          sb[i] = (L.bias && i < L.n_bias) ? __ldg(L.bias + i) : 0.0f;
        }
        const bool two_halves = NERO_NSPLIT && L.n_pad > 128;
        mbar_wait(acc_ready, acc_phase);
        acc_phase ^= 1;
        tcgen05_fence_after();
        asm volatile("bar.sync 1, %0;" ::"n"(kChEpiWarps * 32));
#if NERO_EPI_V2
#define NERO_EPI_CALL(K) chain_epilogue_layer2<K>(L, tmem_base, lg, row0, rows_valid, third, lane, sb, two_halves ? acc_ready + 1 : nullptr, two_halves ? a_free : nullptr, acc_phase1)
#else
#define NERO_EPI_CALL(K) chain_epilogue_layer<K>(L, tl, row0, rows_valid, third, lane, stg, sb)
#endif
        if constexpr (FAM == 0) {
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
        if (two_halves) acc_phase1 ^= 1;
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        // the last layer of the last tile has no consumer, every other (tile, layer) hands over to the MMA warp
        if (lane == 0 && l + 1 < p.n_layers) mbar_arrive(a_ready);
      }
    }
  } else if (warp == kChMmaWarp) {
    // ============================== MMA issuer
    int g = 0;
    uint32_t a_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayer& L = p.L[l];
        const int nh = (NERO_NSPLIT && L.n_pad > 128) ? 2 : 1;
        const uint32_t half_rows = uint32_t(L.n_pad) / nh;
        const uint32_t idesc = make_idesc_bf16(CH_BM, half_rows);
        const uint32_t b_plane = half_rows * 128u;
        mbar_wait(a_ready, a_phase);
        a_phase ^= 1;
        tcgen05_fence_after();
        for (int hf = 0; hf < nh; ++hf) {
          const uint32_t d_col = tmem_base + kAccCol + uint32_t(hf) * half_rows;
          for (int c = 0; c < L.k_chunks; ++c, ++g) {
            const int s = g % kChStages;
            mbar_wait(&full[s], (g / kChStages) & 1);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t b_hi = smem_u32(smem + s * kChStageBytes);
              const uint32_t b_lo = b_hi + b_plane;
#pragma unroll
              for (int k = 0; k < CH_BK / 16; ++k) {
                const uint32_t a_col = uint32_t(c * 32 + k * 8);
                const uint64_t dbh = make_desc_k_sw128(b_hi + k * 32), dbl = make_desc_k_sw128(b_lo + k * 32);
                umma_bf16_ts(d_col, tmem_base + kALoCol + a_col, dbh, idesc, (c | k) != 0);
                umma_bf16_ts(d_col, tmem_base + kAHiCol + a_col, dbl, idesc, 1);
                umma_bf16_ts(d_col, tmem_base + kAHiCol + a_col, dbh, idesc, 1);
              }
              umma_commit(&empty[s]);
              if (nh == 2 && hf == 1 && c == min(L.k_chunks, 2) - 1) umma_commit(a_free);   // K < 128 of A no longer needed
              if (c == L.k_chunks - 1) umma_commit(acc_ready + hf);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ============================== W loader
    int g = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayer& L = p.L[l];
        const int nh = (NERO_NSPLIT && L.n_pad > 128) ? 2 : 1;
        const uint32_t half_rows = uint32_t(L.n_pad) / nh;
        const uint32_t plane = uint32_t(L.n_pad) * 128u, hp = half_rows * 128u;
        for (int hf = 0; hf < nh; ++hf) {
          for (int c = 0; c < L.k_chunks; ++c, ++g) {
            const int s = g % kChStages;
            mbar_wait(&empty[s], ((g / kChStages) & 1) ^ 1);
            if (elect_one()) {
              // rows [hf*half_rows, +half_rows) of the chunk's hi plane and of its lo plane
              const uint8_t* src = L.wimg + size_t(c) * 2u * plane + size_t(hf) * hp;
              mbar_arrive_expect_tx(&full[s], 2u * hp);
              bulk_copy_g2s(smem + s * kChStageBytes, src, hp, &full[s]);
              bulk_copy_g2s(smem + s * kChStageBytes + hp, src + plane, hp, &full[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kChLoadWarp) tmem_dealloc<512>(tmem_base);
}

int chain_dispatch(const ChainParams& p, cudaStream_t stream) {
  if (p.n_layers <= 0 || p.n_layers > kMaxChainLayers || (p.lda0 & 3) || p.k_valid0 > 256) return NERO_ERR_ARG;
  for (int l = 0; l < p.n_layers; ++l) {
    const ChainLayer& L = p.L[l];
    if (L.n_pad % 16 || L.n_pad < 16 || L.n_pad > 256 || L.k_chunks < 1 || L.k_chunks > 4 || L.ncol_out > L.n_pad) return NERO_ERR_ARG;
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
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(umma_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(umma_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(umma_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes) != cudaSuccess)
      return NERO_ERR_CUDA;
    attr_set = true;
  }
  const int tiles_cap = (p.m_cap + CH_BM - 1) / CH_BM;
  if (tiles_cap <= 0) return NERO_OK;
  const int grid = tiles_cap < kNumSMs ? tiles_cap : kNumSMs;
  if (fam == 0) umma_chain_kernel<0><<<grid, kChThreads, kChSmemBytes, stream>>>(p);
  else if (fam == 1) umma_chain_kernel<1><<<grid, kChThreads, kChSmemBytes, stream>>>(p);
  else umma_chain_kernel<2><<<grid, kChThreads, kChSmemBytes, stream>>>(p);
  NERO_LAUNCH_CHECK();
  return NERO_OK;
}

}  // namespace nero
