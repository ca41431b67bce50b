// Image preprocessing of CLIP's `_transform` on the device (the step in front of `encode_image`: clip.load(...)[1],
// called at inference.py:310, evaluate_model.py:458-463; same pipeline as blip_test.py:22-26):
//   Resize(n_px, BICUBIC) on a PIL image -> CenterCrop(n_px) -> ToTensor -> Normalize(mean, std)
// PIL's resize is reproduced exactly (Pillow src/libImaging/Resample.c, 8 bits per channel): separable bicubic
// (a = -0.5) with the support widened by the down-scaling factor (antialiasing), coefficients normalised in double
// precision and rounded to 22-bit fixed point, horizontal pass rounded to uint8 before the vertical pass.  Only the
// pixels that survive the centre crop are computed.  One image per call; everything is asynchronous on `stream`.
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace ccb {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS
constexpr int kMaxTaps = 512;

__device__ __forceinline__ double bicubic_filter(double x) {   // Resample.c bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output positions [first, first + count) of one axis.
// One thread per output position; coeffs[i * ksize + k], bounds[2 i] = first input index, bounds[2 i + 1] = taps.
__global__ void resample_coeffs_kernel(int in_size, int out_size, int first, int count, int ksize, int* __restrict__ bounds,
                                       int* __restrict__ coeffs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int xx = first + i;
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const double center = (xx + 0.5) * scale;
  const double ss = 1.0 / filterscale;
  int xmin = static_cast<int>(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) ww += bicubic_filter((x + xmin - center + 0.5) * ss);
  int* k = coeffs + static_cast<long long>(i) * ksize;
  for (int x = 0; x < ksize; ++x) {
    double w = 0.0;
    if (x < xmax) {
      w = bicubic_filter((x + xmin - center + 0.5) * ss);
      if (ww != 0.0) w /= ww;
    }
    k[x] = w < 0 ? static_cast<int>(-0.5 + w * (1 << kPrecisionBits)) : static_cast<int>(0.5 + w * (1 << kPrecisionBits));
  }
  bounds[2 * i] = xmin;
  bounds[2 * i + 1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {   // Resample.c clip8: lookup[(in) >> PRECISION_BITS], clamped to 0..255
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: tmp[y, i, c] for every input row y and the `count` kept output columns
__global__ void resample_h_kernel(const uint8_t* __restrict__ img, int H, int W, int count, int ksize,
                                  const int* __restrict__ bounds, const int* __restrict__ coeffs, uint8_t* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (i >= count) return;
  const int xmin = bounds[2 * i], n = bounds[2 * i + 1];
  const int* k = coeffs + static_cast<long long>(i) * ksize;
  const uint8_t* row = img + (static_cast<long long>(y) * W + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int kk = k[x];
    s0 += row[3 * x] * kk;
    s1 += row[3 * x + 1] * kk;
    s2 += row[3 * x + 2] * kk;
  }
  uint8_t* o = tmp + (static_cast<long long>(y) * count + i) * 3;
  o[0] = static_cast<uint8_t>(clip8(s0));
  o[1] = static_cast<uint8_t>(clip8(s1));
  o[2] = static_cast<uint8_t>(clip8(s2));
}

// vertical pass over tmp [rows_in, cols, 3] + ToTensor + Normalize -> out [3, count, cols] f32
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int cols, int count, int ksize,
                                       const int* __restrict__ bounds, const int* __restrict__ coeffs, float m0, float m1, float m2,
                                       float d0, float d1, float d2, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (x >= cols) return;
  const int ymin = bounds[2 * j], n = bounds[2 * j + 1];
  const int* k = coeffs + static_cast<long long>(j) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const uint8_t* p = tmp + (static_cast<long long>(ymin + y) * cols + x) * 3;
    const int kk = k[y];
    s0 += p[0] * kk;
    s1 += p[1] * kk;
    s2 += p[2] * kk;
  }
  const long long plane = static_cast<long long>(count) * cols;
  const long long o = static_cast<long long>(j) * cols + x;
  // ToTensor: uint8 -> float / 255; Normalize: (x - mean) / std, both in fp32 like torchvision
  out[o] = (static_cast<float>(clip8(s0)) / 255.0f - m0) / d0;
  out[plane + o] = (static_cast<float>(clip8(s1)) / 255.0f - m1) / d1;
  out[2 * plane + o] = (static_cast<float>(clip8(s2)) / 255.0f - m2) / d2;
}

// no resize needed along either axis (input already has the target size): crop + normalise only
__global__ void crop_norm_kernel(const uint8_t* __restrict__ img, int W, int top, int left, int n_px, float m0, float m1, float m2,
                                 float d0, float d1, float d2, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (x >= n_px) return;
  const uint8_t* p = img + (static_cast<long long>(top + j) * W + left + x) * 3;
  const long long plane = static_cast<long long>(n_px) * n_px, o = static_cast<long long>(j) * n_px + x;
  out[o] = (static_cast<float>(p[0]) / 255.0f - m0) / d0;
  out[plane + o] = (static_cast<float>(p[1]) / 255.0f - m1) / d1;
  out[2 * plane + o] = (static_cast<float>(p[2]) / 255.0f - m2) / d2;
}

int taps_for(int in_size, int out_size) {   // Resample.c: ksize = (int)ceil(support) * 2 + 1
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  int c = static_cast<int>(support);
  if (c < support) ++c;
  return c * 2 + 1;
}

}  // namespace

size_t preprocess_scratch_bytes(int H, int W, int new_h, int new_w, int n_px) {
  const size_t kh = taps_for(W, new_w), kv = taps_for(H, new_h);
  size_t b = static_cast<size_t>(n_px) * (2 + kh) * 4 + static_cast<size_t>(n_px) * (2 + kv) * 4;   // bounds + coeffs of both axes
  b += static_cast<size_t>(H) * n_px * 3;                                                            // horizontal pass
  return b + 256;
}

int preprocess_image(const uint8_t* rgb, int H, int W, int new_h, int new_w, int top, int left, int n_px, const float* mean,
                     const float* stdv, float* out, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  if (H <= 0 || W <= 0 || new_h < n_px || new_w < n_px || top < 0 || left < 0 || top + n_px > new_h || left + n_px > new_w)
    return (int)cudaErrorInvalidValue;
  if (scratch_bytes < preprocess_scratch_bytes(H, W, new_h, new_w, n_px)) return (int)cudaErrorInvalidValue;
  if (new_h == H && new_w == W) {   // torchvision returns the image unchanged
    crop_norm_kernel<<<dim3((n_px + 127) / 128, n_px), 128, 0, s>>>(rgb, W, top, left, n_px, mean[0], mean[1], mean[2], stdv[0],
                                                                    stdv[1], stdv[2], out);
    return (int)cudaGetLastError();
  }
  const int kh = taps_for(W, new_w), kv = taps_for(H, new_h);
  if (kh > kMaxTaps || kv > kMaxTaps) return (int)cudaErrorInvalidValue;
  int* bounds_h = static_cast<int*>(scratch);
  int* coeffs_h = bounds_h + 2 * n_px;
  int* bounds_v = coeffs_h + static_cast<size_t>(n_px) * kh;
  int* coeffs_v = bounds_v + 2 * n_px;
  uint8_t* tmp = reinterpret_cast<uint8_t*>(coeffs_v + static_cast<size_t>(n_px) * kv);
  // PIL runs the horizontal pass only when the width changes and the vertical pass only when the height changes; a pass
  // over an unchanged axis has the single coefficient 1.0 per pixel (scale 1: taps of the neighbours are exactly 0 for
  // the bicubic kernel at integer offsets), so always running both is the same arithmetic
  resample_coeffs_kernel<<<(n_px + 127) / 128, 128, 0, s>>>(W, new_w, left, n_px, kh, bounds_h, coeffs_h);
  resample_coeffs_kernel<<<(n_px + 127) / 128, 128, 0, s>>>(H, new_h, top, n_px, kv, bounds_v, coeffs_v);
  resample_h_kernel<<<dim3((n_px + 127) / 128, H), 128, 0, s>>>(rgb, H, W, n_px, kh, bounds_h, coeffs_h, tmp);
  resample_v_norm_kernel<<<dim3((n_px + 127) / 128, n_px), 128, 0, s>>>(tmp, n_px, n_px, kv, bounds_v, coeffs_v, mean[0], mean[1],
                                                                        mean[2], stdv[0], stdv[1], stdv[2], out);
  return (int)cudaGetLastError();
}

}  // namespace ccb
