// TMA helpers shared by the tcgen05 kernels: device-side 2-D tiled loads / stores and the host-side tensor-map encoder
// (cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint, so the library does not link libcuda).
#pragma once
#include <cuda.h>

#include <cstdio>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace nero {

// 2-D tiled load global -> shared (arrives on an mbarrier with the box byte count), store shared -> global
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int r0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(r0)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int r0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(r0)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct TmaMapKey {
  uintptr_t ptr; int ld, rows, cols, box_cols, box_rows, swizzle;
  bool operator==(const TmaMapKey& o) const {
    return ptr == o.ptr && ld == o.ld && rows == o.rows && cols == o.cols && box_cols == o.box_cols && box_rows == o.box_rows && swizzle == o.swizzle;
  }
};
struct TmaMapKeyHash {
  size_t operator()(const TmaMapKey& k) const {
    size_t h = std::hash<uintptr_t>()(k.ptr);
    for (int v : {k.ld, k.rows, k.cols, k.box_cols, k.box_rows, k.swizzle}) h = (h ^ size_t(v)) * 0x9E3779B97F4A7C15ull;
    return h;
  }
};

inline bool tma_eligible_f32(const float* ptr, int ld, int cols) {
  return ptr && cols >= 8 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0;
}

// A [rows x cols] fp32 window (leading dimension ld) as a 2-D tensor map with [box_rows x box_cols] boxes; columns / rows
// outside the window read as zero.  swizzle: 0 = none, 64 = SWIZZLE_64B.  Maps are cached by every argument.
inline int tma_make_map_f32(CUtensorMap* out, const float* ptr, int ld, int rows, int cols, int box_cols, int box_rows, int swizzle) {
  static std::unordered_map<TmaMapKey, CUtensorMap, TmaMapKeyHash> cache;
  static std::mutex mu;
  const TmaMapKey key{reinterpret_cast<uintptr_t>(ptr), ld, rows, cols, box_cols, box_rows, swizzle};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return NERO_OK; }
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) return NERO_ERR_CUDA;
  const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
  const cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    fprintf(stderr, "nero_b200: cuTensorMapEncodeTiled failed (%d) for ptr %p ld %d rows %d cols %d box %dx%d\n", int(rc), (const void*)ptr, ld,
            rows, cols, box_rows, box_cols);
    return NERO_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return NERO_OK;
}

}  // namespace nero
