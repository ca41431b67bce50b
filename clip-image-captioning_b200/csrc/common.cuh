// Shared device helpers: activations (bit-for-bit the same formulas the reference's torch ops use), bf16
// packing, warp/block reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ccb {

enum Act : int {
  ACT_NONE = 0,
  ACT_RELU = 1,       // mapper default (reference train.py:78, layers/Transformer.py:120)
  ACT_QUICKGELU = 2,  // CLIP ViT: x * sigmoid(1.702 x)
  ACT_GELU_NEW = 3,   // GPT-2 / GPT-J "gelu_new" (tanh form)
  ACT_GELU_ERF = 4,   // nnf.gelu (layers/Transformer.py:124)
  ACT_ELU = 5,        // nnf.elu  (layers/Transformer.py:122)
  ACT_SELU = 6,       // nnf.selu (layers/Transformer.py:126)
  ACT_TANH = 7,       // upstream ClipCap MLP mapper (nn.Tanh between the two Linear layers)
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_QUICKGELU: return x / (1.f + expf(-1.702f * x));
    case ACT_TANH: return tanhf(x);
    case ACT_GELU_NEW: {
      const float k = 0.7978845608028654f;  // sqrt(2/pi)
      float u = k * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.f + tanhf(u));
    }
    case ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
    case ACT_ELU: return x > 0.f ? x : expm1f(x);
    case ACT_SELU: {
      const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
      return scale * (x > 0.f ? x : alpha * expm1f(x));
    }
    default: return x;
  }
}

// out-of-line copy for cold code (GEMM epilogues): one call instead of an inlined switch per element
static __device__ __noinline__ float apply_act_call(float x, int act) { return apply_act(x, act); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float bf16_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum / max through shared scratch (>= 32 floats). All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

}  // namespace ccb
