// Microbenchmark: how fast can 148 persistent CTAs stream a weight matrix from HBM into shared memory?
//   mode 0: TMA tensor loads, box {64 k, 128 rows} out of a K-major [rows, K] bf16 matrix (row pitch 2K bytes)
//   mode 1: cp.async.bulk of contiguous 16 KB tiles (pre-tiled layout)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_bench stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

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
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "l"(0x12F0000000000000ull) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(0x12F0000000000000ull) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const char* base, int tiles_r, int kb,
                                                        int stages, int reps, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sb + stages * 16384;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (stages + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long units = (long long)tiles_r * kb;
  const long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    uint32_t it = 0;
    for (int r = 0; r < reps; ++r)
      for (long long u = u0; u < u1; ++u, ++it) {
        const uint32_t s = it % stages, ph = (it / stages) & 1;
        mbar_wait(bars + 8 * (stages + s), ph ^ 1);
        mbar_expect(bars + 8 * s, 16384);
        if (MODE == 0) tma2d(sb + s * 16384, &map, bars + 8 * s, (int)(u % kb) * 64, (int)(u / kb) * 128);
        else bulk1d(sb + s * 16384, base + u * 16384, 16384, bars + 8 * s);
      }
  } else if (threadIdx.x == 32) {
    uint32_t it = 0;
    unsigned long long acc = 0;
    for (int r = 0; r < reps; ++r)
      for (long long u = u0; u < u1; ++u, ++it) {
        const uint32_t s = it % stages, ph = (it / stages) & 1;
        mbar_wait(bars + 8 * s, ph);
        acc += *reinterpret_cast<volatile unsigned*>(smem_raw + (sb - smem_u32(smem_raw)) + s * 16384);
        mbar_arrive(bars + 8 * (stages + s));
      }
    if (acc == 0x1234567) sink[0] = acc;
  }
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 6400, K = argc > 2 ? atoi(argv[2]) : 1600;
  const int layers = 48;  // cycle over distinct matrices so nothing is served from L2
  const int tiles_r = (rows + 127) / 128, kb = K / 64;
  const size_t mat_bytes = (size_t)tiles_r * 128 * K * 2;
  char* buf;
  cudaMalloc(&buf, mat_bytes * layers);
  cudaMemset(buf, 1, mat_bytes * layers);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Enc enc = (Enc)fn;
  cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int stages : {4, 8, 12}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        for (int l = 0; l < layers; ++l) {
          CUtensorMap m;
          cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)tiles_r * 128}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
          cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
          enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf + l * mat_bytes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          const size_t smem = stages * 16384 + 1024 + 256;
          if (mode == 0) stream_kernel<0><<<148, 128, smem>>>(m, buf + l * mat_bytes, tiles_r, kb, stages, 1, sink);
          else stream_kernel<1><<<148, 128, smem>>>(m, buf + l * mat_bytes, tiles_r, kb, stages, 1, sink);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("mode %d (%s) rows %d K %d stages %2d: %.3f ms for %.1f MB -> %.0f GB/s (%s)\n", mode, mode ? "bulk 16KB contiguous" : "TMA 2D box", rows, K,
             stages, best, mat_bytes * layers / 1e6, mat_bytes * layers / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  // one persistent launch over all matrices (no launch gaps): bulk mode with reps over a 48x larger "matrix"
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)tiles_r * 128 * layers}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      const size_t smem = 12 * 16384 + 1024 + 256;
      if (mode == 0) stream_kernel<0><<<148, 128, smem>>>(m, buf, tiles_r * layers, kb, 12, 1, sink);
      else stream_kernel<1><<<148, 128, smem>>>(m, buf, tiles_r * layers, kb, 12, 1, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("single launch, mode %d, 12 stages: %.3f ms -> %.0f GB/s (%s)\n", mode, best, mat_bytes * layers / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
