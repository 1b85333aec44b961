// Row-wise global I/O for "one thread = one sample" kernels.
// A thread that reads or writes its own row of a [M x ld] matrix touches one 4-byte piece of 32 different rows per warp
// instruction: 32 sectors move for 128 useful bytes.  These helpers route the rows through a per-warp shared-memory tile
// (lane l owns tile row l while computing) so that global memory only sees whole rows accessed by the 32 lanes together.
// LD must be odd (>= the widest row): (LD * l + c) % 32 is then conflict free when every lane walks its own row.
#pragma once
#include "common.cuh"

namespace nero {

// dst[r][c] = scale * tile[r][c]   for r < rows, c < ncols      (dst points at row 0 / column 0 of the warp's block)
template <int LD>
__device__ __forceinline__ void warp_rows_store(const float (*tile)[LD], float* dst, size_t ld, int ncols, int rows, int lane,
                                                float scale = 1.0f) {
  __syncwarp();
  for (int r = 0; r < rows; ++r)
    for (int c = lane; c < ncols; c += 32) dst[size_t(r) * ld + c] = tile[r][c] * scale;
  __syncwarp();
}
// The same for a compile-time row width of 4*NC4 columns: the (row, float4) pairs are dealt to the lanes round-robin, one
// 16-byte store each (4x fewer instructions than the scalar walk, no idle lanes in the last partial pass).  Falls back to the
// scalar walk when the destination is not 16-byte aligned.
template <int LD, int NC4>
__device__ __forceinline__ void warp_rows_store4(const float (*tile)[LD], float* dst, size_t ld, int rows, int lane, float scale = 1.0f) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) || (ld & 3)) { warp_rows_store<LD>(tile, dst, ld, 4 * NC4, rows, lane, scale); return; }
  __syncwarp();
  const int total = rows * NC4;
  for (int idx = lane; idx < total; idx += 32) {
    const int r = idx / NC4, q = idx - r * NC4;
    const float* t = tile[r] + 4 * q;
    *reinterpret_cast<float4*>(dst + size_t(r) * ld + 4 * q) = make_float4(t[0] * scale, t[1] * scale, t[2] * scale, t[3] * scale);
  }
  __syncwarp();
}
// tile[r][c] = a[r][c] (+ b[r][c])
template <int LD>
__device__ __forceinline__ void warp_rows_load(float (*tile)[LD], const float* a, size_t lda, const float* b, size_t ldb, int ncols,
                                               int rows, int lane) {
  __syncwarp();
  for (int r = 0; r < rows; ++r)
    for (int c = lane; c < ncols; c += 32) tile[r][c] = a[size_t(r) * lda + c] + (b ? b[size_t(r) * ldb + c] : 0.0f);
  __syncwarp();
}

}  // namespace nero
