// Microbenchmark: cost of one tcgen05.mma (kind::f16, bf16 operands from shared memory) as a function of the shape
// M x N (K = 16), and the TMEM lane layout of the M = 64 accumulator.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../clip-image-captioning_b200/csrc -o umma_shape_bench umma_shape_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ptx.cuh"
using namespace ccb;

// smem: A tile [128 rows x 64 k] SW128 (16 KB) at 0, B tile [256 rows x 64 k] SW128 (32 KB) at 16 KB
__global__ void __launch_bounds__(128, 1) shape_kernel(int M, int N, int iters, int mode, long long* cycles, float* dump, const __nv_bfloat16* a_src,
                                                       const __nv_bfloat16* b_src) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar = base + 49152, slot = bar + 8;
  // fill A / B with the 128B swizzle applied by hand: element (r, k) of a tile lives at r*128 + ((k/8 ^ (r%8))*16) + (k%8)*2
  for (int idx = threadIdx.x; idx < 128 * 64; idx += blockDim.x) {
    const int r = idx / 64, k = idx % 64;
    *reinterpret_cast<__nv_bfloat16*>(gen + r * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2) = a_src[idx];
  }
  for (int idx = threadIdx.x; idx < 256 * 64; idx += blockDim.x) {
    const int r = idx / 64, k = idx % 64;
    *reinterpret_cast<__nv_bfloat16*>(gen + 16384 + r * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2) = b_src[idx];
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(bar + 16, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 49152 + 8);
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, N);
    const uint64_t ad = ptx::umma_desc_k_sw128(base), bd = ptx::umma_desc_k_sw128(base + 16384);
    // one k-block (4 MMAs) for the layout dump
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, ad + 2u * k, bd + 2u * k, idesc, k > 0);
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    const long long t0 = clock64();
    const unsigned long long g0 = ptx::globaltimer_ns();
    // mode 0: back-to-back MMAs, one commit at the end; 1: a commit (to a barrier nobody waits on) after every 4 MMAs;
    // 2: commit + wait after every 4 MMAs (issue -> complete -> mbarrier -> issue latency); 3: like 1 with rotating
    // operand addresses (4 x 4 KB apart)
    const uint32_t bar2 = bar + 16;
    if (mode == 0) {
      for (int it = 0; it < iters; ++it)
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, ad + 2u * k, bd + 2u * k, idesc, 1u);
    } else if (mode == 1 || mode == 3) {
      for (int it = 0; it < iters; ++it) {
        const uint64_t o = mode == 3 ? static_cast<uint64_t>(((it & 3) * 4096) >> 4) : 0;
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, ad + o + 2u * k, bd + o + 2u * k, idesc, 1u);
        ptx::umma_commit(bar2);
      }
    } else if (mode == 2) {
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, ad + 2u * k, bd + 2u * k, idesc, 1u);
        ptx::umma_commit(bar2);
        ptx::mbar_wait(bar2, it & 1);
      }
    } else if (mode == 4) {   // tcgen05.fence::after_thread_sync before every group of 4
      for (int it = 0; it < iters; ++it) {
        ptx::tc_fence_after();
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, ad + 2u * k, bd + 2u * k, idesc, 1u);
      }
    } else if (mode == 6) {   // in-situ style: stage index / phase / descriptors recomputed per group from loop-carried registers
      uint32_t st = 0, ph = 0;
      const int S = iters > 0 ? 3 : 1;       // opaque to the compiler
      const uint32_t stage_bytes = static_cast<uint32_t>(N > 0 ? 4096 : 1);
      const uint64_t desc0 = ptx::umma_desc_k_sw128(base);
      for (int it = 0; it < iters; ++it) {
        ptx::tc_fence_after();
        {
          const uint64_t adesc = desc0 + static_cast<uint64_t>((st * stage_bytes) >> 4);
          const uint64_t bdesc = adesc + static_cast<uint64_t>(16384u >> 4);
          ptx::umma_bf16(tmem + 256 + (ph << 4), adesc, bdesc, idesc, it > 0 ? 1u : 0u);
          ptx::umma_bf16(tmem + 256 + (ph << 4), adesc + 2u, bdesc + 2u, idesc, 1u);
          ptx::umma_bf16(tmem + 256 + (ph << 4), adesc + 4u, bdesc + 4u, idesc, 1u);
          ptx::umma_bf16(tmem + 256 + (ph << 4), adesc + 6u, bdesc + 6u, idesc, 1u);
        }
        if (++st == static_cast<uint32_t>(S)) { st = 0; ph ^= 1; }
      }
    } else {                  // 5: the accumulator is re-initialised (accumulate = 0) at the first MMA of every 25th group
      for (int it = 0; it < iters; ++it)
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, ad + 2u * k, bd + 2u * k, idesc, (it % 25 == 0 && k == 0) ? 0u : 1u);
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 1);
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
    cycles[1] = static_cast<long long>(ptx::globaltimer_ns() - g0);
  }
  __syncthreads();
  ptx::tc_fence_after();
  // dump TMEM lanes 0..127, columns 0..N-1 of the first accumulator
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    ptx::tmem_ld8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int v = 0; v < 8; ++v) if (blockIdx.x == 0) dump[(warp * 32 + lane) * 256 + c0 + v] = __uint_as_float(r[v]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  std::vector<__nv_bfloat16> ha(128 * 64), hb(256 * 64);
  // A[r, k] = (r + 1) if k == 0 else 0;  B[n, k] = (n + 1) if k == 0 else 0  ->  D[r, n] = (r + 1) * (n + 1) (exact in bf16 for small values)
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) ha[r * 64 + k] = __float2bfloat16(k == 0 ? float(r + 1) : 0.f);
  for (int n = 0; n < 256; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(k == 0 ? (n < 8 ? float(n + 1) : 1.f) : 0.f);
  if (getenv("RANDOM_OPERANDS")) {
    srand(1);
    for (auto& v : ha) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-3f);
    for (auto& v : hb) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-3f);
  }
  __nv_bfloat16 *da, *db; long long* dc; float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dc, 16); cudaMalloc(&dd, 128 * 256 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(shape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int shapes[][2] = {{128, 64}, {128, 128}, {128, 256}};
  for (int nblk : {1, 148})
  for (auto& s : shapes) {
    printf("[%d CTAs] ", nblk);
    const int iters = 100000;
    cudaMemset(dd, 0, 128 * 256 * 4);
    double res[7];
    cudaError_t e = cudaSuccess;
    for (int mode = 0; mode < 1; ++mode) {
      shape_kernel<<<nblk, 128, 60000>>>(s[0], s[1], iters, mode, dc, dd, da, db);
      e = cudaDeviceSynchronize();
      long long cyc[2] = {0, 0};
      cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
      res[mode] = double(cyc[0]) / (iters * 4.0);
      if (mode == 0) printf("(back-to-back: %.2f ns per mma -> %.0f MHz) ", double(cyc[1]) / (iters * 4.0), double(cyc[0]) / double(cyc[1]) * 1e3);
    }
    printf("M=%3d N=%3d: cycles per tcgen05.mma: back-to-back %.1f | commit per 4: %.1f | commit+wait per 4: %.1f | rotating operands: %.1f | fence per 4: %.1f | acc reset per 100: %.1f | in-situ style loop: %.1f (%s)\n",
           s[0], s[1], res[0], res[1], res[2], res[3], res[4], res[5], res[6], cudaGetErrorString(e));
    if (s[0] == 64 && s[1] == 64) {
      std::vector<float> hd(128 * 256);
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      // D[r, 0] = r + 1: which lane holds row r?
      printf("  M=64 layout, column 0 per lane (value = row + 1, 0 = unused):\n  ");
      for (int l = 0; l < 128; ++l) printf("%g%s", hd[l * 256], (l % 32 == 31) ? "\n  " : " ");
      printf("column 1 of lanes 0..3: %g %g %g %g (expect 2x column 0)\n", hd[1], hd[256 + 1], hd[512 + 1], hd[768 + 1]);
    }
  }
  return 0;
}
