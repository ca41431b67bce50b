// How fast can one SM pull an L2-resident activation into shared memory with TMA, and what happens when many SMs pull the
// SAME rows at the same instant (the access pattern of a "feature-owner" GEMM phase, where every CTA needs the whole
// [64 rows x K] operand)?  box {64 k, 64 rows} = 8 KB tiles out of a [64, K] bf16 matrix, `stages` loads in flight.
//   same = 1: every CTA reads the same matrix; same = 0: CTA c reads its own copy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_ingest_bench l2_ingest_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(64, 1) ingest_kernel(const __grid_constant__ CUtensorMap map, int kb, int stages, int reps, int same,
                                                       unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sb + stages * 8192;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (stages + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int row0 = same ? 0 : blockIdx.x * 64;
  if (threadIdx.x == 0) {
    uint32_t it = 0;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < kb; ++k, ++it) {
        const uint32_t s = it % stages, ph = (it / stages) & 1;
        mbar_wait(bars + 8 * (stages + s), ph ^ 1);
        mbar_expect(bars + 8 * s, 8192);
        tma2d(sb + s * 8192, &map, bars + 8 * s, k * 64, row0);
      }
  } else if (threadIdx.x == 32) {
    uint32_t it = 0;
    unsigned long long acc = 0;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < kb; ++k, ++it) {
        const uint32_t s = it % stages, ph = (it / stages) & 1;
        mbar_wait(bars + 8 * s, ph);
        acc += *reinterpret_cast<volatile unsigned*>(smem_raw + (sb - smem_u32(smem_raw)) + s * 8192);
        mbar_arrive(bars + 8 * (stages + s));
      }
    if (acc == 0x1234567) sink[0] = acc;
  }
}

int main() {
  const int K = 1600, kb = K / 64, reps = 400;
  char* buf;
  cudaMalloc(&buf, (size_t)148 * 64 * K * 2);
  cudaMemset(buf, 1, (size_t)148 * 64 * K * 2);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Enc enc = (Enc)fn;
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)148 * 64}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
  enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  for (int same : {1, 0})
    for (int grid : {1, 8, 37, 100, 148})
      for (int stages : {2, 4, 8, 16}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const size_t smem = stages * 8192 + 1024 + 512;
        ingest_kernel<<<grid, 64, smem>>>(m, kb, stages, 20, same, sink);   // warm L2
        cudaEventRecord(e0);
        ingest_kernel<<<grid, 64, smem>>>(m, kb, stages, reps, same, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)reps * kb * 8192;
        printf("%s matrix, %3d CTAs, %2d tiles in flight: %.1f GB/s per SM, %.2f TB/s total, %.0f ns per 8 KB tile (%s)\n", same ? "same" : "own ",
               grid, stages, bytes / ms / 1e6, bytes * grid / ms / 1e9, ms * 1e6 / (reps * kb), cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
