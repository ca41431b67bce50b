// bf16 x bf16 -> fp32 GEMM for sm_100a: TMA (128B-swizzled K-major tiles) -> tcgen05.mma with the accumulator
// in TMEM -> tcgen05.ld epilogue with fused bias / activation / residual and an optional transposed store.
//
//   acc[i, j] = sum_k A[i, k] * B[j, k]          A: [Ra, K] bf16 row-major, B: [Rb, K] bf16 row-major
//
// "normal" orientation   : A = activations (tokens x K), B = weights (features x K) -> out[token, feature]
// "swapped" orientation  : A = weights (features x K),  B = activations (tokens x K) -> out[token, feature]
//                          written through the transposed store.  Used when tokens <= 256 (decode), so the
//                          128-row MMA tile is filled by weight rows and every weight byte is streamed once.
//
// One CTA computes one 128 x BN tile (optionally one K-split of it).  Warp roles: warp 0 = TMA producer,
// warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (one TMEM lane quadrant
// each).  Split-K is deterministic: every split stores its raw fp32 tile to a workspace, the CTA that
// arrives last on the tile's semaphore sums the splits in fixed order 0..S-1 and runs the epilogue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace ccb {

struct GemmParams {
  int Ra, Rb;        // valid rows of A / B (tile rows beyond are zero-filled by TMA and masked at the store)
  int k_blocks;      // K / 64
  int split_k;       // >= 1
  // epilogue
  void* out;         // f32 or bf16
  int out_bf16;
  int transposed;    // 0: out[i*ldo + j]   1: out[j*ldo + i]
  long long ldo;
  const float* bias;       // per feature: index j (normal) or i (transposed); may be null
  const float* residual;   // f32, same indexing as out with leading dim ldr; may be null (may alias out)
  long long ldr;
  int act;
  // normal-mode row remap: out_row = (i / rg_in) * rg_out + rg_off + (i % rg_in); rg_in == 0 -> identity
  int rg_in, rg_out, rg_off;
  // split-K workspace and per-tile semaphores
  float* ws;
  int* sem;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreads = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                           const __grid_constant__ CUtensorMap tma_b,
                                                           const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout: full[kStages], empty[kStages], tmem_full, then the TMEM address slot and a flag word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 1);
  const uint32_t flag_slot = tmem_slot + 4;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * Cfg::kStageBytes + 8 * (2 * kStages + 1));
  volatile uint32_t* flag_ptr = tmem_slot_ptr + 1;
  (void)flag_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // K range of this split
  const int kb_per = (p.k_blocks + p.split_k - 1) / p.split_k;
  const int kb_begin = blockIdx.z * kb_per;
  int kb_end = kb_begin + kb_per;
  if (kb_end > p.k_blocks) kb_end = p.k_blocks;
  const int nkb = kb_end - kb_begin;  // host guarantees >= 1 for every split

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int row_a0 = blockIdx.x * Cfg::BM;
  const int row_b0 = blockIdx.y * BN;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // weights are streamed once (evict-first); activations are re-read by many CTAs (evict-last)
      const uint64_t hint_a = p.transposed ? ptx::kEvictFirst : ptx::kEvictLast;
      const uint64_t hint_b = p.transposed ? ptx::kEvictLast : ptx::kEvictNormal;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        ptx::mbar_arrive_expect_tx(full_bar(s), Cfg::kStageBytes);
        const int kcoord = (kb_begin + it) * Cfg::BK;
        ptx::tma_load_2d(sa, &tma_a, full_bar(s), kcoord, row_a0, hint_a);
        ptx::tma_load_2d(sb, &tma_b, full_bar(s), kcoord, row_b0, hint_b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(Cfg::BM, BN);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(full_bar(s), ph);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        const uint64_t adesc = ptx::umma_desc_k_sw128(sa);
        const uint64_t bdesc = ptx::umma_desc_k_sw128(sb);
#pragma unroll
        for (int k = 0; k < Cfg::BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
          ptx::umma_bf16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
      }
      ptx::umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int i = row_a0 + quad * 32 + lane;
    const bool i_ok = i < p.Ra;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    bool do_epilogue = true;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    float* ws_tile = nullptr;
    if (p.split_k > 1) {
      ws_tile = p.ws + static_cast<size_t>(tile_id) * p.split_k * (BN * 128);
      float* mine = ws_tile + static_cast<size_t>(blockIdx.z) * (BN * 128) + quad * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 32; ++v) mine[(c0 + v) * 128] = __uint_as_float(r[v]);
      }
      __threadfence();
      ptx::named_bar_sync(1, 128);
      if (warp == 2 && lane == 0) {
        const int prev = atomicAdd(p.sem + tile_id, 1);
        const int last = (prev == p.split_k - 1);
        if (last) p.sem[tile_id] = 0;  // self-reset for the next launch
        *flag_ptr = last;
      }
      ptx::named_bar_sync(1, 128);
      do_epilogue = (*flag_ptr != 0);
      if (do_epilogue) __threadfence();
    }

    if (do_epilogue) {
      long long out_row = i;
      if (!p.transposed && p.rg_in > 0) out_row = static_cast<long long>(i / p.rg_in) * p.rg_out + p.rg_off + (i % p.rg_in);
      const float bias_i = (p.transposed && p.bias != nullptr && i_ok) ? p.bias[i] : 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float acc[32];
        if (p.split_k > 1) {
#pragma unroll
          for (int v = 0; v < 32; ++v) acc[v] = 0.f;
          // fixed summation order 0..S-1 (deterministic); loads of 4 splits are in flight together
          for (int s0 = 0; s0 < p.split_k; s0 += 4) {
            float t[4][32];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (s0 + u < p.split_k) {
                const float* src = ws_tile + static_cast<size_t>(s0 + u) * (BN * 128) + quad * 32 + lane;
#pragma unroll
                for (int v = 0; v < 32; ++v) t[u][v] = ptx::ldcg_f32(src + (c0 + v) * 128);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (s0 + u < p.split_k) {
#pragma unroll
                for (int v = 0; v < 32; ++v) acc[v] += t[u][v];
              }
            }
          }
        } else {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 32; ++v) acc[v] = __uint_as_float(r[v]);
        }
        const int j0 = row_b0 + c0;
        if (j0 >= p.Rb) break;
        if (p.transposed) {
          // lanes hold consecutive features i -> coalesced stores for every token j
#pragma unroll
          for (int v = 0; v < 32; ++v) {
            const int j = j0 + v;
            if (i_ok && j < p.Rb) {
              float x = acc[v] + bias_i;
              x = apply_act(x, p.act);
              if (p.residual) x += p.residual[static_cast<long long>(j) * p.ldr + i];
              if (p.out_bf16)
                reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(j) * p.ldo + i] = __float2bfloat16_rn(x);
              else
                reinterpret_cast<float*>(p.out)[static_cast<long long>(j) * p.ldo + i] = x;
            }
          }
        } else if (i_ok) {
          const bool full = (j0 + 32 <= p.Rb);
          const bool vec_ok = full && ((p.ldo & 7) == 0) && (p.residual == nullptr || (p.ldr & 3) == 0);
          if (vec_ok) {
#pragma unroll
            for (int v = 0; v < 32; v += 4) {
              float x[4] = {acc[v], acc[v + 1], acc[v + 2], acc[v + 3]};
              if (p.bias) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + j0 + v));
                x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) x[q] = apply_act(x[q], p.act);
              if (p.residual) {
                const float4 r4 = *reinterpret_cast<const float4*>(p.residual + out_row * p.ldr + j0 + v);
                x[0] += r4.x; x[1] += r4.y; x[2] += r4.z; x[3] += r4.w;
              }
              acc[v] = x[0]; acc[v + 1] = x[1]; acc[v + 2] = x[2]; acc[v + 3] = x[3];
            }
            if (p.out_bf16) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + j0;
#pragma unroll
              for (int v = 0; v < 32; v += 8) {
                uint4 pk;
                pk.x = pack_bf16x2(acc[v], acc[v + 1]);
                pk.y = pack_bf16x2(acc[v + 2], acc[v + 3]);
                pk.z = pack_bf16x2(acc[v + 4], acc[v + 5]);
                pk.w = pack_bf16x2(acc[v + 6], acc[v + 7]);
                *reinterpret_cast<uint4*>(o + v) = pk;
              }
            } else {
              float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + j0;
#pragma unroll
              for (int v = 0; v < 32; v += 4)
                *reinterpret_cast<float4*>(o + v) = make_float4(acc[v], acc[v + 1], acc[v + 2], acc[v + 3]);
            }
          } else {
#pragma unroll
            for (int v = 0; v < 32; ++v) {
              const int j = j0 + v;
              if (j < p.Rb) {
                float x = acc[v] + (p.bias ? p.bias[j] : 0.f);
                x = apply_act(x, p.act);
                if (p.residual) x += p.residual[out_row * p.ldr + j];
                if (p.out_bf16)
                  reinterpret_cast<__nv_bfloat16*>(p.out)[out_row * p.ldo + j] = __float2bfloat16_rn(x);
                else
                  reinterpret_cast<float*>(p.out)[out_row * p.ldo + j] = x;
              }
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace ccb
