// Host side of the tcgen05 GEMM: tensor-map encoding, orientation / tile / split-K selection, launch.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "gemm_sm100.cuh"
#include "internal.h"
#include "mega.h"

namespace ccb {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
char g_err[512] = {0};
std::once_flag g_attr_once;
int g_attr_status = 0;

int fail(const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return -1;
}

// 2-D bf16 tensor [rows, K] with row pitch `ld` elements; box = {64 (K), box_rows}; 128B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail("GEMM operand not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail("GEMM operand row pitch not a multiple of 16 bytes");
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[64];
    snprintf(b, sizeof(b), "%d", static_cast<int>(r));
    return fail("cuTensorMapEncodeTiled failed (CUresult %s)", b);
  }
  return 0;
}

template <int BN>
int set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GemmCfg<BN>::kSmemBytes);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
  return 0;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GemmCfg<BN>::kThreads);
  cfg.dynamicSmemBytes = GemmCfg<BN>::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.split_k > 1) {
    // the S splits of one output tile form one cluster and reduce over distributed shared memory
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = static_cast<unsigned>(p.split_k);
    ++na;
  }
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<BN>, ta, tb, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("GEMM launch failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace

const char* gemm_last_error() { return g_err; }

int gemm_make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows) {
  if (!g_encode) return fail("gemm_init() was not called");
  return make_tmap(m, ptr, rows, K, ld, box_rows);
}

int gemm_init(int device) {
  (void)device;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
      return fail("cuTensorMapEncodeTiled not available from the driver: %s", cudaGetErrorString(e));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  std::call_once(g_attr_once, [] {
    int r = 0;
    r |= set_attr<32>();
    r |= set_attr<64>();
    r |= set_attr<128>();
    r |= set_attr<256>();
    g_attr_status = r;
  });
  return g_attr_status;
}

int gemm_launch(const GemmArgs& a, const GemmWorkspace& w, cudaStream_t stream) {
  if (!g_encode) return fail("gemm_init() was not called");
  if (a.K <= 0 || a.K % 64 != 0) return fail("GEMM K must be a positive multiple of 64");
  if (a.tokens <= 0 || a.features <= 0) return fail("GEMM with empty extent");

  const bool swapped = a.force_orientation ? (a.force_orientation == 2) : (a.tokens <= 256);
  if (swapped && a.rg_in > 0) return fail("row remap is only supported in the normal orientation");
  if (swapped && a.tokens > 256) return fail("swapped orientation needs tokens <= 256");

  int bn;
  if (a.force_bn) {
    bn = a.force_bn;
  } else if (swapped) {
    bn = a.tokens <= 32 ? 32 : a.tokens <= 64 ? 64 : a.tokens <= 128 ? 128 : 256;
  } else {
    bn = a.features >= 256 ? 256 : a.features > 64 ? 128 : 64;
  }
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return fail("unsupported BN");
  if (swapped && bn < a.tokens) return fail("swapped orientation: BN smaller than token count");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  const bf16* A = swapped ? a.weight : a.act;
  const bf16* B = swapped ? a.act : a.weight;
  p.Ra = swapped ? a.features : a.tokens;
  p.Rb = swapped ? a.tokens : a.features;
  const long long ldA = swapped ? a.K : a.lda;
  const long long ldB = swapped ? a.lda : a.K;
  p.k_blocks = a.K / 64;
  p.out = a.out;
  p.out_bf16 = a.out_bf16;
  p.transposed = swapped ? 1 : 0;
  p.ldo = a.ldo;
  p.bias = a.bias;
  p.residual = a.residual;
  p.ldr = a.ldr;
  p.act = a.act_fn;
  p.rg_in = a.rg_in;
  p.rg_out = a.rg_out;
  p.rg_off = a.rg_off;
  p.trace = w.trace ? w.trace + (w.trace_count++ % w.trace_launches) * w.trace_stride : nullptr;
  p.pdl = (a.allow_pdl && pdl_enabled()) ? 1 : 0;
  // the vector store path assumes 16-byte aligned bases; fall back to scalar stores through odd ldo otherwise
  if (!swapped) {
    if ((reinterpret_cast<uintptr_t>(a.out) & 15) || (a.residual && (reinterpret_cast<uintptr_t>(a.residual) & 15)) ||
        (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)))
      return fail("normal-orientation GEMM needs 16-byte aligned out / residual / bias");
  }

  dim3 grid((p.Ra + 127) / 128, (p.Rb + bn - 1) / bn, 1);
  const int tiles = grid.x * grid.y;
  int split = 1;
  if (a.force_split) {
    split = a.force_split;
  } else if (tiles < w.num_sms) {
    // fill the machine without spilling into a second wave (one CTA per SM): tiles * split <= number of SMs,
    // each split >= 4 k-blocks
    split = w.num_sms / tiles;
    const int max_by_k = p.k_blocks / 4 > 0 ? p.k_blocks / 4 : 1;
    if (split > max_by_k) split = max_by_k;
  }
  if (split > p.k_blocks) split = p.k_blocks;
  if (split > 8) split = 8;  // portable cluster size
  // every split must own at least one k-block: ceil(kb/split)*(split-1) < kb
  while (split > 1 && ((p.k_blocks + split - 1) / split) * (split - 1) >= p.k_blocks) --split;
  p.split_k = split;
  grid.z = split;

  CUtensorMap ta, tb;
  if (make_tmap(&ta, A, p.Ra, a.K, ldA, 128)) return -1;
  if (make_tmap(&tb, B, p.Rb, a.K, ldB, bn)) return -1;

  switch (bn) {
    case 32: return launch<32>(ta, tb, p, grid, stream);
    case 64: return launch<64>(ta, tb, p, grid, stream);
    case 128: return launch<128>(ta, tb, p, grid, stream);
    default: return launch<256>(ta, tb, p, grid, stream);
  }
}

}  // namespace ccb
