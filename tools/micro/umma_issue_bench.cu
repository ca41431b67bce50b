// Microbenchmark: what slows the tcgen05.mma issue loop of a real kernel down?  One CTA (or 148), warp 0 issues groups
// of 4 MMAs (128 x N x 16) the way the kernels do (converged warp, elect_one, descriptors from loop-carried registers,
// one commit per group); optionally the other warps of the CTA spin on an mbarrier that never completes, like idle
// producer / epilogue warps do.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../clip-image-captioning_b200/csrc -o umma_issue_bench umma_issue_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace ccb;

__global__ void __launch_bounds__(384, 1) issue_kernel(int N, int iters, int spin_warps, int mode, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bars = base + 4 * 49152;   // 8 stages of 48 KB
  const uint32_t done_bar = bars, never_bar = bars + 8, commit_bar = bars + 16, slot = bars + 64;   // commit_bar .. +24: four barriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u + (i * 2654435761u & 0x007f007fu);
  if (threadIdx.x == 0) {
    ptx::mbar_init(done_bar, 1);
    ptx::mbar_init(never_bar, 1);
    ptx::mbar_init(commit_bar, 1);
    ptx::mbar_init(commit_bar + 8, 1);
    ptx::mbar_init(commit_bar + 16, 1);
    ptx::mbar_init(commit_bar + 24, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) ptx::tmem_alloc<512>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 4 * 49152 + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(gen + 4 * 49152 + 128);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (warp == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint64_t desc0 = ptx::umma_desc_k_sw128(base);
    const uint32_t stage_bytes = 49152u;
    uint32_t s = 0, ph = 0;
    const int S = 4 + (iters < 0);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      ptx::tc_fence_after();
      if (mode == 0) {            // converged warp, elected issue (the kernels' pattern)
        if (ptx::elect_one()) {
          const uint64_t adesc = desc0 + static_cast<uint64_t>((s * stage_bytes) >> 4);
          const uint64_t bdesc = adesc + static_cast<uint64_t>(16384u >> 4);
          ptx::umma_bf16(tmem + (ph << 8), adesc, bdesc, idesc, it > 0 ? 1u : 0u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 2u, bdesc + 2u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 4u, bdesc + 4u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 6u, bdesc + 6u, idesc, 1u);
          ptx::umma_commit(commit_bar);
        }
        __syncwarp();
      } else if (mode == 2) {     // the decode kernel's pair loop: two (passing) mbarrier waits, 8 MMAs, three commits
        ptx::mbar_wait(never_bar, 1);
        ptx::mbar_wait(commit_bar + 8, 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t adesc = desc0 + static_cast<uint64_t>((s * stage_bytes) >> 4);
          const uint64_t bdesc = adesc + static_cast<uint64_t>(16384u >> 4);
          ptx::umma_bf16(tmem + (ph << 8), adesc, bdesc, idesc, it > 0 ? 1u : 0u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 2u, bdesc + 2u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 4u, bdesc + 4u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 6u, bdesc + 6u, idesc, 1u);
          if ((it & 3) == 3) ptx::umma_commit(commit_bar + 16);
          if (it & 1) {
            ptx::umma_commit(commit_bar);
            ptx::umma_commit(commit_bar + 24);
          }
        }
        __syncwarp();
      } else {                    // lane 0 only
        if (lane == 0) {
          const uint64_t adesc = desc0 + static_cast<uint64_t>((s * stage_bytes) >> 4);
          const uint64_t bdesc = adesc + static_cast<uint64_t>(16384u >> 4);
          ptx::umma_bf16(tmem + (ph << 8), adesc, bdesc, idesc, it > 0 ? 1u : 0u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 2u, bdesc + 2u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 4u, bdesc + 4u, idesc, 1u);
          ptx::umma_bf16(tmem + (ph << 8), adesc + 6u, bdesc + 6u, idesc, 1u);
          ptx::umma_commit(commit_bar);
        }
        __syncwarp();
      }
      if (++s == static_cast<uint32_t>(S)) { s = 0; ph ^= 1; }
    }
    if (lane == 0) {
      ptx::umma_commit(done_bar);
      ptx::mbar_wait(done_bar, 0);
      cycles[blockIdx.x] = clock64() - t0;
      *stop = 1;
      ptx::mbar_arrive(never_bar);   // releases the spinners
    }
    __syncwarp();
  } else if (warp <= spin_warps) {
    // idle role: waits on a barrier that completes only at the end (mbarrier.try_wait spin, as in the kernels)
    while (!ptx::mbar_try_wait(never_bar, 0)) {}
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

int main() {
  long long* dc;
  cudaMalloc(&dc, 148 * 8);
  const int smem = 4 * 49152 / 2 + 49152 + 2048;   // 4 stages used
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int iters = 20000;
  for (int nblk : {1, 148})
    for (int N : {64, 256})
      for (int mode : {0, 2, 1})
        for (int spin : {0, 11}) {
          issue_kernel<<<nblk, 384, 227 * 1024 - 1024>>>(N, iters, spin, mode, dc);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
          printf("%3d CTAs N=%3d %s spinning warps %2d: %.1f cycles per tcgen05.mma (%s)\n", nblk, N, mode == 0 ? "elect " : mode == 2 ? "decode" : "lane0 ", spin,
                 double(c) / (iters * 4.0), cudaGetErrorString(e));
        }
  (void)smem;
  return 0;
}
