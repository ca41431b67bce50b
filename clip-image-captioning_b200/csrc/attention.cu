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
// CTA = 4 warps = one (row, head).  Lanes are grouped LPT = HD/8 per token (8 bf16 = 16 bytes per lane);
// warp w visits token groups w, w+4, ...  Online softmax per lane-group; merged across groups and warps.
template <int HD>
__global__ void __launch_bounds__(128) attention_decode_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                               int H, float scale, KvCache cache, int layer,
                                                               const int* __restrict__ block_table,
                                                               const int* __restrict__ ctx_len, int rotary_dim) {
  constexpr int LPT = HD / 8;     // lanes per token
  constexpr int TPW = 32 / LPT;   // tokens per warp iteration
  const int h = blockIdx.x, b = blockIdx.y;
  const int d = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / LPT, sub = lane % LPT;
  const int ctx = ctx_len[b];  // tokens already cached; the new token sits at position ctx
  const bf16* row = qkv + static_cast<size_t>(b) * 3 * d;
  const int* bt = block_table + static_cast<size_t>(b) * cache.max_pages_per_row;

  // this lane's 8 q values and the new token's k/v chunk
  float q[8], kn[8], vn[8];
  {
    const uint4 qu = *reinterpret_cast<const uint4*>(row + h * HD + sub * 8);
    const uint4 ku = *reinterpret_cast<const uint4*>(row + d + h * HD + sub * 8);
    const uint4 vu = *reinterpret_cast<const uint4*>(row + 2 * d + h * HD + sub * 8);
    const uint32_t qa[4] = {qu.x, qu.y, qu.z, qu.w}, ka[4] = {ku.x, ku.y, ku.z, ku.w}, va[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 a = unpack_bf16x2(qa[t]), k2 = unpack_bf16x2(ka[t]), v2 = unpack_bf16x2(va[t]);
      if (rotary_dim > 0 && sub * 8 + 2 * t < rotary_dim) {
        rotary_pair(a.x, a.y, sub * 4 + t, ctx, rotary_dim);
        rotary_pair(k2.x, k2.y, sub * 4 + t, ctx, rotary_dim);
        // keep the cached key bf16-rounded exactly as later steps will read it
        k2 = unpack_bf16x2(pack_bf16x2(k2.x, k2.y));
      }
      q[2 * t] = a.x * scale; q[2 * t + 1] = a.y * scale;
      kn[2 * t] = k2.x; kn[2 * t + 1] = k2.y;
      vn[2 * t] = v2.x; vn[2 * t + 1] = v2.y;
    }
  }
  // append the new token's K/V to the cache (warp 0, first lane group)
  if (warp == 0 && grp == 0) {
    const int page = bt[ctx / cache.page_tokens];
    const int t = ctx % cache.page_tokens;
    uint4 kp, vp;
    kp.x = pack_bf16x2(kn[0], kn[1]); kp.y = pack_bf16x2(kn[2], kn[3]);
    kp.z = pack_bf16x2(kn[4], kn[5]); kp.w = pack_bf16x2(kn[6], kn[7]);
    vp.x = pack_bf16x2(vn[0], vn[1]); vp.y = pack_bf16x2(vn[2], vn[3]);
    vp.z = pack_bf16x2(vn[4], vn[5]); vp.w = pack_bf16x2(vn[6], vn[7]);
    *reinterpret_cast<uint4*>(cache.base + kv_index(cache, layer, 0, page, h, t) + sub * 8) = kp;
    *reinterpret_cast<uint4*>(cache.base + kv_index(cache, layer, 1, page, h, t) + sub * 8) = vp;
  }

  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;

  // cached tokens [0, ctx); token ctx (the new one) is folded in by warp 0 / group 0 from registers
  for (int t0 = warp * TPW; t0 < ctx; t0 += 4 * TPW) {
    const int tok = t0 + grp;
    const bool valid = tok < ctx;
    uint4 ku = make_uint4(0, 0, 0, 0), vu = make_uint4(0, 0, 0, 0);
    if (valid) {
      const int page = bt[tok / cache.page_tokens];
      const int t = tok % cache.page_tokens;
      ku = ptx::ld_nc_u4(cache.base + kv_index(cache, layer, 0, page, h, t) + sub * 8);
      vu = ptx::ld_nc_u4(cache.base + kv_index(cache, layer, 1, page, h, t) + sub * 8);
    }
    const uint32_t ka[4] = {ku.x, ku.y, ku.z, ku.w}, va[4] = {vu.x, vu.y, vu.z, vu.w};
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 k2 = unpack_bf16x2(ka[t]);
      s = fmaf(q[2 * t], k2.x, s);
      s = fmaf(q[2 * t + 1], k2.y, s);
    }
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (valid) {
      const float mn = fmaxf(m, s);
      const float corr = expf(m - mn);  // m = -inf on first use -> 0
      const float p = expf(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 v2 = unpack_bf16x2(va[t]);
        acc[2 * t] = acc[2 * t] * corr + p * v2.x;
        acc[2 * t + 1] = acc[2 * t + 1] * corr + p * v2.y;
      }
      m = mn;
    }
  }
  if (warp == 0 && grp == 0) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(q[e], kn[e], s);
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync((LPT == 32) ? 0xffffffffu : ((1u << LPT) - 1u), s, o);
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn);
    const float p = expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = acc[e] * corr + p * vn[e];
    m = mn;
  }

  // merge the 4*TPW partial states: through shared memory, summed in a fixed order by the first LPT lanes
  __shared__ float sm_m[4 * 32 / 1];
  __shared__ float sm_l[4 * 32];
  __shared__ float sm_acc[4 * 32 * 8];
  const int slot = warp * 32 + lane;
  sm_m[slot] = m;
  sm_l[slot] = l;
#pragma unroll
  for (int e = 0; e < 8; ++e) sm_acc[slot * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < LPT) {
    float M = -INFINITY;
    for (int w = 0; w < 4; ++w)
      for (int g = 0; g < TPW; ++g) M = fmaxf(M, sm_m[w * 32 + g * LPT + threadIdx.x]);
    float L = 0.f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    for (int w = 0; w < 4; ++w)
      for (int g = 0; g < TPW; ++g) {
        const int sl = w * 32 + g * LPT + threadIdx.x;
        const float mm = sm_m[sl];
        if (mm == -INFINITY) continue;
        const float f = expf(mm - M);
        L += sm_l[sl] * f;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += sm_acc[sl * 8 + e] * f;
      }
    const float inv = 1.f / L;
    uint4 pk;
    pk.x = pack_bf16x2(o[0] * inv, o[1] * inv);
    pk.y = pack_bf16x2(o[2] * inv, o[3] * inv);
    pk.z = pack_bf16x2(o[4] * inv, o[5] * inv);
    pk.w = pack_bf16x2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(b) * d + h * HD + threadIdx.x * 8) = pk;
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
  if (hd == 64)
    attention_decode_kernel<64><<<dim3(H, B), 128, 0, s>>>(qkv, out, H, scale, *cache, layer, block_table, ctx_len,
                                                           rotary_dim);
  else if (hd == 128)
    attention_decode_kernel<128><<<dim3(H, B), 128, 0, s>>>(qkv, out, H, scale, *cache, layer, block_table, ctx_len,
                                                            rotary_dim);
  else if (hd == 256)
    attention_decode_kernel<256><<<dim3(H, B), 128, 0, s>>>(qkv, out, H, scale, *cache, layer, block_table, ctx_len,
                                                            rotary_dim);
  else
    return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ccb
