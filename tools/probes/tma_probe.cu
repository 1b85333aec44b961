// Developer probe (run on the GPU box): per-SM throughput of cp.async.bulk.tensor 2-D loads / stores for several box shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int PEND>
__global__ void store_kernel(const __grid_constant__ CUtensorMap map, int box_cols, int box_rows, int ncol_boxes, int nrow_boxes, int slots, int slot_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x != 0) return;
  int n = 0;
  for (int rb = blockIdx.x; rb < nrow_boxes; rb += gridDim.x)
    for (int cb = 0; cb < ncol_boxes; ++cb, ++n) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map)),
                   "r"(smem_u32(smem + (n % slots) * slot_bytes)), "r"(cb * box_cols), "r"(rb * box_rows) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PEND) : "memory");
    }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void load_kernel(const __grid_constant__ CUtensorMap map, int box_cols, int box_rows, int ncol_boxes, int nrow_boxes, int slots, int slot_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x != 0) return;
  for (int i = 0; i < slots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  int n = 0;
  for (int rb = blockIdx.x; rb < nrow_boxes; rb += gridDim.x)
    for (int cb = 0; cb < ncol_boxes; ++cb, ++n) {
      const int s = n % slots;
      if (n >= slots) {      // wait for the previous load into this slot
        uint32_t ok = 0;
        const uint32_t par = ((n / slots) - 1) & 1;
        while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(par) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(slot_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem + s * slot_bytes)),
                   "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bars[s])), "r"(cb * box_cols), "r"(rb * box_rows) : "memory");
    }
  for (int i = 0; i < slots && i < n; ++i) {
    const int m = n - 1 - i;       // last loads
    const int s = m % slots;
    uint32_t ok = 0;
    const uint32_t par = (m / slots) & 1;
    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(par) : "memory");
  }
}

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  const int rows = 127232, cols = 256;
  float* buf; cudaMalloc(&buf, size_t(rows) * cols * 4 * 8);    // 8 buffers (like 8 layers) = 1 GB
  cudaMemset(buf, 0, size_t(rows) * cols * 4 * 8);
  struct Shape { int bc, br; CUtensorMapSwizzle sw; const char* name; };
  Shape shapes[] = {{16, 128, CU_TENSOR_MAP_SWIZZLE_64B, "128x16 sw64"}, {32, 128, CU_TENSOR_MAP_SWIZZLE_128B, "128x32 sw128"},
                    {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, "64x32 sw128"}, {16, 32, CU_TENSOR_MAP_SWIZZLE_64B, "32x16 sw64"},
                    {64, 128, CU_TENSOR_MAP_SWIZZLE_NONE, "128x64 none"}, {256, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "32x256 none"}};
  cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(store_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (auto& sh : shapes) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows) * 8};
    const cuuint64_t strides[1] = {cuuint64_t(cols) * 4};
    const cuuint32_t box[2] = {cuuint32_t(sh.bc), cuuint32_t(sh.br)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sh.sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("%s: encode failed %d\n", sh.name, int(rc)); continue; }
    const int slot_bytes = sh.bc * sh.br * 4;
    const int ncb = cols / sh.bc, nrb = rows * 8 / sh.br;
    const double bytes = double(rows) * 8 * cols * 4;
    for (int mode = 0; mode < 5; ++mode) {
      const int slots = mode == 0 ? 1 : (mode == 1 ? 4 : (mode == 2 ? 8 : (mode == 3 ? 4 : 8)));
      if (slots * slot_bytes > 190 * 1024) continue;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) store_kernel<0><<<148, 32, slots * slot_bytes>>>(m, sh.bc, sh.br, ncb, nrb, slots, slot_bytes);
        else if (mode == 1) store_kernel<3><<<148, 32, slots * slot_bytes>>>(m, sh.bc, sh.br, ncb, nrb, slots, slot_bytes);
        else if (mode == 2) store_kernel<7><<<148, 32, slots * slot_bytes>>>(m, sh.bc, sh.br, ncb, nrb, slots, slot_bytes);
        else load_kernel<<<148, 32, slots * slot_bytes>>>(m, sh.bc, sh.br, ncb, nrb, slots, slot_bytes);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      const double us_per_box = ms * 1e3 / (double(ncb) * nrb / 148.0);
      printf("%-14s %s slots/pending %d: %7.1f us, %6.1f GB/s chip, %5.1f B/cyc/SM @1.9GHz, %.2f us per box per SM %s\n", sh.name, mode < 3 ? "store" : "load ", slots, ms * 1e3,
             bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9, us_per_box, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
