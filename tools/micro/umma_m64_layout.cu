// Which TMEM lanes does a tcgen05.mma cta_group::1 M = 64 accumulator occupy?  D[i][j] = i + 1 (A[i][0] = i + 1, B[j][0] = 1),
// then every warp reads its 32 lanes x 64 columns and prints lane -> value.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_m64_layout umma_m64_layout.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(float* out, int M) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);            // 128 rows x 64 k, 128B swizzle
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gen + 16384);    // 64 rows x 64 k
  uint64_t* bar = reinterpret_cast<uint64_t*>(gen + 16384 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(gen + 16384 + 8192 + 64);
  for (int i = threadIdx.x; i < (16384 + 8192) / 2; i += 128) A[i] = __float2bfloat16(0.f);
  __syncthreads();
  if (threadIdx.x < 128) {
    const int r = threadIdx.x;   // k = 0: chunk 0 ^ (r & 7)
    A[(r * 128 + ((0 ^ (r & 7)) * 16)) / 2] = __float2bfloat16((float)(r + 1));
    if (r < 64) B[(r * 128 + ((0 ^ (r & 7)) * 16)) / 2] = __float2bfloat16(1.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // zero the accumulator region first with an M = 128 MMA of zeros?  Simpler: the M = 64 MMA writes with accumulate = 0; lanes it
  // does not touch keep whatever they held, so fill all 128 lanes with -1 through a first M = 128 MMA against a B of... skip:
  // unwritten lanes are reported as they are (garbage is recognisable: not an integer in 1..64).
  if (threadIdx.x == 0) {
    auto desc = [](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFF) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(1024 >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(desc(base)), "l"(desc(base + 16384)), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[8];
  for (int c0 = 0; c0 < 64; c0 += 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int v = 0; v < 8; ++v) out[(warp * 32 + lane) * 64 + c0 + v] = __uint_as_float(r[v]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int M : {128, 64}) {
    cudaMemset(d, 0, 128 * 64 * 4);
    k<<<1, 128, 16384 + 8192 + 256 + 1024>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M = %d: %s\n", M, cudaGetErrorString(e));
    static float h[128 * 64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int l = 0; l < 128; ++l) {
      bool same = true;
      for (int c = 1; c < 64; ++c) same = same && h[l * 64 + c] == h[l * 64];
      printf("lane %3d: col0 %g col63 %g %s\n", l, h[l * 64], h[l * 64 + 63], same ? "" : "(columns differ)");
    }
  }
  return 0;
}
