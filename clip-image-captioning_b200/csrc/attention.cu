// Attention kernels.
//  * attention_prefill: full attention over a short sequence (S <= 256) for one (batch, head) per CTA.  Serves
//    the ViT (S=50, hd=64, non-causal), the prefix mapper (S=clip_len+P, hd=d/8, non-causal; reference
//    layers/MultiHeadAttention.py:24-41) and the LM prefill (causal; HF GPT2Attention / GPTJAttention), where
//    it also writes K/V into the paged cache.
//  * attention_decode: one query token per row against the paged KV cache; memory-bound, 16-byte loads,
//    online softmax, 4 warps per (row, head) each striding over the context, merged through shared memory.
#include "common.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace ccb {

namespace {

__device__ __forceinline__ size_t kv_index(const KvCache& c, int layer, int kv, int page, int h, int t) {
  return ((((static_cast<size_t>(layer) * 2 + kv) * c.num_pages + page) * c.H + h) * c.page_tokens + t) * c.hd;
}

// GPT-J rotary ("rotate_every_two", HF modeling_gptj.py:47-67): pairs (2i, 2i+1) of the first rotary_dim dims.
__device__ __forceinline__ void rotary_pair(float& a, float& b, int pair_idx, int pos, int rotary_dim) {
  const float inv_freq = powf(10000.f, -2.f * pair_idx / rotary_dim);
  float sn, cs;
  sincosf(pos * inv_freq, &sn, &cs);
  const float a2 = a * cs - b * sn;
  const float b2 = b * cs + a * sn;
  a = a2;
  b = b2;
}

// ------------------------------------------------------------------------------------------------ prefill
// smem: Ks [S][hd/2+1] words (bf16x2, padded -> conflict-free column walks), Vs [S][hd/2] words,
//       per-warp q row (hd floats) and probability row (S floats).
constexpr int kPrefillWarps = 8;

__global__ void __launch_bounds__(kPrefillWarps * 32) attention_prefill_kernel(
    const bf16* __restrict__ qkv, bf16* __restrict__ out, int S, int H, int hd, float scale, int causal,
    KvCache cache, int layer, const int* __restrict__ block_table, int pos0, int rotary_dim, int write_cache,
    const uint8_t* __restrict__ key_mask) {
  extern __shared__ uint32_t smem_u[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int d = H * hd;
  const int hw = hd / 2;        // words per row
  const int kstride = hw + 1;
  uint32_t* Ks = smem_u;
  uint32_t* Vs = Ks + S * kstride;
  // q rows are read back as float2: keep the fp32 region 8-byte aligned (S * (2 * hw + 1) words can be odd)
  float* qbuf = reinterpret_cast<float*>(smem_u + ((S * (kstride + hw) + 1) & ~1));  // [warps][hd]
  float* pbuf = qbuf + kPrefillWarps * hd;                        // [warps][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + static_cast<size_t>(b) * S * 3 * d;

  // stage K and V of this head (bf16 pairs); apply rotary to K first when requested
  for (int idx = threadIdx.x; idx < S * hw; idx += blockDim.x) {
    const int j = idx / hw, w = idx % hw;
    const bf16* rowp = base + static_cast<size_t>(j) * 3 * d;
    uint32_t kw = *reinterpret_cast<const uint32_t*>(rowp + d + h * hd + 2 * w);
    const uint32_t vw = *reinterpret_cast<const uint32_t*>(rowp + 2 * d + h * hd + 2 * w);
    if (rotary_dim > 0 && 2 * w < rotary_dim) {
      float2 kf = unpack_bf16x2(kw);
      rotary_pair(kf.x, kf.y, w, pos0 + j, rotary_dim);
      kw = pack_bf16x2(kf.x, kf.y);
    }
    Ks[j * kstride + w] = kw;
    Vs[j * hw + w] = vw;
    if (write_cache) {
      const int pos = pos0 + j;
      const int page = block_table[static_cast<size_t>(b) * cache.max_pages_per_row + pos / cache.page_tokens];
      const int t = pos % cache.page_tokens;
      *reinterpret_cast<uint32_t*>(cache.base + kv_index(cache, layer, 0, page, h, t) + 2 * w) = kw;
      *reinterpret_cast<uint32_t*>(cache.base + kv_index(cache, layer, 1, page, h, t) + 2 * w) = vw;
    }
  }
  __syncthreads();

  float* myq = qbuf + warp * hd;
  float* myp = pbuf + warp * S;
  for (int i = warp; i < S; i += kPrefillWarps) {
    // q row -> fp32 in shared (rotary applied)
    const bf16* qrow = base + static_cast<size_t>(i) * 3 * d + h * hd;
    for (int w = lane; w < hw; w += 32) {
      float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(qrow + 2 * w));
      if (rotary_dim > 0 && 2 * w < rotary_dim) rotary_pair(qf.x, qf.y, w, pos0 + i, rotary_dim);
      myq[2 * w] = qf.x;
      myq[2 * w + 1] = qf.y;
    }
    __syncwarp();
    const int jmax = causal ? i + 1 : S;  // keys [0, jmax)
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = c * 32 + lane;
      float acc = -INFINITY;
      if (c * 32 < jmax && j < jmax && (key_mask == nullptr || key_mask[static_cast<size_t>(b) * S + j])) {
        acc = 0.f;
        const uint32_t* kr = Ks + j * kstride;
        for (int w = 0; w < hw; ++w) {
          const float2 kf = unpack_bf16x2(kr[w]);
          const float2 qf = *reinterpret_cast<const float2*>(myq + 2 * w);
          acc = fmaf(qf.x, kf.x, acc);
          acc = fmaf(qf.y, kf.y, acc);
        }
        acc *= scale;
      }
      sc[c] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float e = (sc[c] == -INFINITY) ? 0.f : expf(sc[c] - mx);
      sc[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = c * 32 + lane;
      if (j < jmax) myp[j] = sc[c] * inv;
    }
    __syncwarp();
    bf16* orow = out + (static_cast<size_t>(b) * S + i) * d + h * hd;
    for (int w = lane; w < hw; w += 32) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < jmax; ++j) {
        const float p = myp[j];
        const float2 vf = unpack_bf16x2(Vs[j * hw + w]);
        a0 = fmaf(p, vf.x, a0);
        a1 = fmaf(p, vf.y, a1);
      }
      *reinterpret_cast<uint32_t*>(orow + 2 * w) = pack_bf16x2(a0, a1);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ decode
// One WARP per (row, head); 4 warps per CTA.  The step is bound by the K/V bytes it has to read, so the kernel is
// organised for memory-level parallelism: every lane owns whole tokens (t = lane, lane + 32, ...) and fetches their
// 128..512-byte K (then V) rows with independent 16-byte loads (HD/8 in flight per token slot), instead of a few
// lanes sharing one token behind a block-table lookup chain.
//   pass 1: s_t = q . K_t for the lane's tokens (q pre-scaled, fp32 in shared memory, read as broadcasts)
//   softmax statistics across the warp (+ the new token, whose k/v come from the qkv buffer, not the cache)
//   pass 2: per 64-dim chunk, acc[64] += p_t * V_t over the lane's tokens, then a 62-shuffle transpose-reduce
//           leaves dims (2*lane, 2*lane+1) of the chunk in each lane -> one coalesced 128-byte store per warp.
constexpr int kDecodeWarps = 4;

template <int HD>
__global__ void __launch_bounds__(kDecodeWarps * 32) attention_decode_kernel(
    const bf16* __restrict__ qkv, bf16* __restrict__ out, int rows, int H, float scale, KvCache cache, int layer,
    const int* __restrict__ block_table, const int* __restrict__ ctx_len, int rotary_dim, int sc_cap) {
  constexpr int E = HD / 32;  // dims per lane in the "own dims" layout (2, 4, 8): pairs stay inside a lane
  extern __shared__ float dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int unit = blockIdx.x * kDecodeWarps + warp;  // (row, head)
  if (unit >= rows * H) return;
  const int b = unit / H, h = unit % H;
  const int d = H * HD;
  float* qs = dsm + warp * (HD + sc_cap);  // [HD] scaled query
  float* sc = qs + HD;                      // [ctx] scores
  const int ctx = ctx_len[b];               // tokens already cached; the new token sits at position ctx
  const bf16* row = qkv + static_cast<size_t>(b) * 3 * d;
  const int* bt = block_table + static_cast<size_t>(b) * cache.max_pages_per_row;

  // ---- q / k_new / v_new for this lane's E dims; rotary on q and k_new; append k_new / v_new to the cache
  float s_new = 0.f;
  {
    float q[E], kn[E], vn[E];
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      const int dim = lane * E + e;
      float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + h * HD + dim));
      float2 kf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + d + h * HD + dim));
      const float2 vf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + 2 * d + h * HD + dim));
      if (rotary_dim > 0 && dim < rotary_dim) {
        rotary_pair(qf.x, qf.y, dim / 2, ctx, rotary_dim);
        rotary_pair(kf.x, kf.y, dim / 2, ctx, rotary_dim);
        kf = unpack_bf16x2(pack_bf16x2(kf.x, kf.y));  // exactly what later steps read back from the cache
      }
      q[e] = qf.x * scale; q[e + 1] = qf.y * scale;
      kn[e] = kf.x; kn[e + 1] = kf.y;
      vn[e] = vf.x; vn[e + 1] = vf.y;
    }
    const int page = bt[ctx / cache.page_tokens];
    const int tin = ctx % cache.page_tokens;
    bf16* kdst = cache.base + kv_index(cache, layer, 0, page, h, tin) + lane * E;
    bf16* vdst = cache.base + kv_index(cache, layer, 1, page, h, tin) + lane * E;
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      *reinterpret_cast<uint32_t*>(kdst + e) = pack_bf16x2(kn[e], kn[e + 1]);
      *reinterpret_cast<uint32_t*>(vdst + e) = pack_bf16x2(vn[e], vn[e + 1]);
      qs[lane * E + e] = q[e];
      qs[lane * E + e + 1] = q[e + 1];
      s_new = fmaf(q[e], kn[e], s_new);
      s_new = fmaf(q[e + 1], kn[e + 1], s_new);
    }
    s_new = warp_sum(s_new);
  }
  __syncwarp();

  // ---- pass 1: scores of the cached tokens
  float mx = s_new;
  for (int t = lane; t < ctx; t += 32) {
    const int page = bt[t / cache.page_tokens];
    const uint4* kp = reinterpret_cast<const uint4*>(cache.base + kv_index(cache, layer, 0, page, h, t % cache.page_tokens));
    uint4 kk[HD / 8];
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) kk[c] = ptx::ld_nc_u4(kp + c);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const float4 q0 = *reinterpret_cast<const float4*>(qs + c * 8);
      const float4 q1 = *reinterpret_cast<const float4*>(qs + c * 8 + 4);
      const float2 k0 = unpack_bf16x2(kk[c].x), k1 = unpack_bf16x2(kk[c].y), k2 = unpack_bf16x2(kk[c].z), k3 = unpack_bf16x2(kk[c].w);
      s = fmaf(q0.x, k0.x, s); s = fmaf(q0.y, k0.y, s); s = fmaf(q0.z, k1.x, s); s = fmaf(q0.w, k1.y, s);
      s = fmaf(q1.x, k2.x, s); s = fmaf(q1.y, k2.y, s); s = fmaf(q1.z, k3.x, s); s = fmaf(q1.w, k3.y, s);
    }
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  const float p_new = expf(s_new - mx);
  float lsum = 0.f;
  for (int t = lane; t < ctx; t += 32) {
    const float p = expf(sc[t] - mx);
    sc[t] = p;  // own slots only: no cross-lane hazard
    lsum += p;
  }
  lsum = warp_sum(lsum) + p_new;
  const float inv = 1.f / lsum;

  // ---- pass 2: output, 64 dims at a time
#pragma unroll 1
  for (int c0 = 0; c0 < HD; c0 += 64) {
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int t = lane; t < ctx; t += 32) {
      const int page = bt[t / cache.page_tokens];
      const uint4* vp = reinterpret_cast<const uint4*>(cache.base + kv_index(cache, layer, 1, page, h, t % cache.page_tokens) + c0);
      uint4 vv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vv[c] = ptx::ld_nc_u4(vp + c);
      const float p = sc[t];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 v0 = unpack_bf16x2(vv[c].x), v1 = unpack_bf16x2(vv[c].y), v2 = unpack_bf16x2(vv[c].z), v3 = unpack_bf16x2(vv[c].w);
        acc[c * 8 + 0] = fmaf(p, v0.x, acc[c * 8 + 0]); acc[c * 8 + 1] = fmaf(p, v0.y, acc[c * 8 + 1]);
        acc[c * 8 + 2] = fmaf(p, v1.x, acc[c * 8 + 2]); acc[c * 8 + 3] = fmaf(p, v1.y, acc[c * 8 + 3]);
        acc[c * 8 + 4] = fmaf(p, v2.x, acc[c * 8 + 4]); acc[c * 8 + 5] = fmaf(p, v2.y, acc[c * 8 + 5]);
        acc[c * 8 + 6] = fmaf(p, v3.x, acc[c * 8 + 6]); acc[c * 8 + 7] = fmaf(p, v3.y, acc[c * 8 + 7]);
      }
    }
    // transpose-reduce over the 32 lanes: after the round with step s a lane keeps the half of the remaining
    // dims selected by its bit s, so the survivors of lane l are dims 2l and 2l+1 of this chunk
#pragma unroll
    for (int step = 16, half = 32; step >= 1; step >>= 1, half >>= 1) {
      const bool up = (lane & step) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float keep = up ? acc[j + half] : acc[j];
        const float send = up ? acc[j] : acc[j + half];
        acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
      }
    }
    const int dim = c0 + 2 * lane;
    const float2 vn = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + 2 * d + h * HD + dim));
    const float o0 = (acc[0] + p_new * vn.x) * inv, o1 = (acc[1] + p_new * vn.y) * inv;
    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(b) * d + h * HD + dim) = pack_bf16x2(o0, o1);
  }
}

}  // namespace

int attention_prefill(const bf16* qkv, bf16* out, int B, int S, int H, int hd, float scale, int causal,
                      const KvCache* cache, int layer, const int* block_table, int pos0, int rotary_dim,
                      const uint8_t* key_mask, cudaStream_t s) {
  if (B <= 0 || S <= 0) return 0;
  if (S > 256 || hd % 2) return (int)cudaErrorInvalidValue;
  const size_t smem = (static_cast<size_t>(S) * (hd / 2 + 1) + static_cast<size_t>(S) * (hd / 2) + 1) * 4 +
                      static_cast<size_t>(kPrefillWarps) * (hd + S) * 4;
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = 200 * 1024;
  }
  KvCache c;
  if (cache) c = *cache;
  attention_prefill_kernel<<<dim3(H, B), kPrefillWarps * 32, smem, s>>>(qkv, out, S, H, hd, scale, causal, c, layer,
                                                                        block_table, pos0, rotary_dim,
                                                                        cache != nullptr ? 1 : 0, key_mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int attention_decode(const bf16* qkv, bf16* out, int B, int H, int hd, float scale, const KvCache* cache, int layer,
                     const int* block_table, const int* ctx_len, int rotary_dim, cudaStream_t s) {
  if (B <= 0) return 0;
  if (!cache) return (int)cudaErrorInvalidValue;
  // per-warp shared memory: hd floats of q + one score per cached token
  const int sc_cap = (cache->max_pages_per_row + 3) & ~3;  // max_pages_per_row == max_ctx >= any context length
  const size_t smem = static_cast<size_t>(kDecodeWarps) * (hd + sc_cap) * sizeof(float);
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  const int grid = (B * H + kDecodeWarps - 1) / kDecodeWarps;
#define CCB_LAUNCH_DECODE(HDV)                                                                                      \
  do {                                                                                                              \
    static size_t configured = 0;                                                                                   \
    if (smem > 48 * 1024 && smem > configured) {                                                                    \
      cudaError_t e = cudaFuncSetAttribute(attention_decode_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           200 * 1024);                                                             \
      if (e != cudaSuccess) return (int)e;                                                                          \
      configured = 200 * 1024;                                                                                      \
    }                                                                                                               \
    cudaError_t le = launch_kernel(attention_decode_kernel<HDV>, dim3(grid), dim3(kDecodeWarps * 32), smem, s, true, \
                                   qkv, out, B, H, scale, *cache, layer, block_table, ctx_len, rotary_dim, sc_cap);   \
    if (le != cudaSuccess) return (int)le;                                                                          \
  } while (0)
  if (hd == 64)
    CCB_LAUNCH_DECODE(64);
  else if (hd == 128)
    CCB_LAUNCH_DECODE(128);
  else if (hd == 256)
    CCB_LAUNCH_DECODE(256);
  else
    return (int)cudaErrorInvalidValue;
#undef CCB_LAUNCH_DECODE
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ccb
