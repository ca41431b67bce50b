// Persistent decode-step kernel, second generation (up to 64 rows): the GPT-2 layer stack of one decode step
// (lms/GPT2.py:17-19 -> HF GPT2LMHeadModel, one new token per row against the KV cache) in FIVE grid-wide phases per
// layer instead of the eight of decode_mega.cu:
//
//   A  qkv = LayerNorm_1(h) Wqkv^T + b          D  mlp = gelu_new(LayerNorm_2(h) Wfc^T + b)
//   B  att = attention(qkv, KV cache)           E  h  += mlp Wfc2^T + b
//   C  h  += att Wproj^T + b
//
// STATUS: correct (tests/test_gpu_mega.py, GPT2-XL tokens) and measured SLOWER than decode_mega.cu (4.4 vs 2.4 ms per step
// on B200; DESIGN.md section 3.1b has the per-phase timeline and the reasons), so it is opt-in: CCB_MEGA2=1 or
// ccb_debug_set_mega(ctx, 3).  Kept as the measured starting point of the cluster design.
//
// What went away are the phases that only folded split-K partials out of an L2 workspace (ln_1, ln_2, gelu, and the q/k/v
// fold in front of attention).  The means:
//   * Thread-block clusters of 4 CTAs.  A weight matrix is cut into 64-feature row tiles (tcgen05.mma M = 64: feature
//     16 q + i of a tile lives in TMEM lane 32 q + i); a tile belongs to ONE cluster, whose CTAs split K four ways.  The
//     four fp32 partials [64 features x 64 rows] are reduce-scattered over distributed shared memory (CTA j of the cluster
//     finalises rows 16 j .. 16 j + 15: every CTA pushes the rows of the other three with 128-bit st.shared::cluster and
//     signals a remote mbarrier), summed in rank order (deterministic) and finished in registers: bias, then bf16 q/k/v,
//     or gelu_new -> bf16, or the fp32 residual update of h.
//   * LayerNorm moved into the consumer.  The finaliser of a residual update knows h for 64 features x 16 rows, so it
//     publishes the tile's (mean, M2) per row [tiles][64 rows]; a consumer CTA Chan-combines the pairs of a row (25 for
//     d = 1600), reads its K slice of h (fp32, L2), normalises and writes the bf16 operand tile straight into the
//     128B-swizzled shared-memory layout tcgen05.mma reads.  No x round trip, no LayerNorm phase.
//   * Everything else as in decode_mega.cu: a W-producer warp that streams this CTA's weight tiles of ALL layers in
//     consumption order through a TMA ring (it runs ahead across phases), an X-producer warp (TMA loads of att / mlp
//     tiles), one MMA-issuer warp (tcgen05.mma 64 x 64 x 16, accumulators in TMEM), 8 compute warps, bounded spins; ring
//     slots are handed over in chunks (up to 4 weight tiles / 16 MMAs per barrier wait).
//   Warps 1-3 and two attention-only warps take attention units too (13 warps x 132 CTAs >= 64 rows x 25 heads).
//
// Tile -> cluster: tile t of GEMM kind k belongs to cluster (t + off_k) % ncl (host-planned offsets balance the bytes
// each CTA streams); a cluster owns at most four tiles of a kind and runs them together (one X tile feeds all of them).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "mega.h"
#include "ptx.cuh"

namespace ccb {

namespace {

#include "mega_common.cuh"

constexpr int kThreads2 = 448;            // 14 warps
constexpr int kAttnWarps2 = 13;           // compute warps 0-7, warps 1-3 (X producer, MMA issuer, TMEM owner), warps 12-13
constexpr int kHelpers2 = 5;
constexpr int kS = 4;                     // cluster size == K split
constexpr int kN = 64;                    // rows of a step (MMA N); R <= kN
constexpr int kMT = 64;                   // features per row tile (MMA M)
constexpr int kMaxT = 4;                  // row tiles of one GEMM a cluster may own
constexpr int kRowsPerCta = kN / kS;      // rows a CTA finalises
constexpr int kWTile = kMT * 64 * 2;      // one weight tile: 64 features x 64 k bf16 (8 KB)
constexpr int kWChunk = 4 * kWTile;       // hand-over unit of the weight ring: the tiles of up to 4 (k block, row tile) pairs
constexpr int kNWC = 3;                   // weight chunks in flight
constexpr int kXStage = kN * 128;         // one activation tile: 64 rows x 64 k bf16 (8 KB)
constexpr int kNX = 8;                    // activation tiles in flight
constexpr int kRecvBuf = (kS - 1) * kMT * kRowsPerCta * 4;   // [3 remote ranks][64 features][16 rows] f32 = 12 KB per tile
constexpr int kGBHalf = 512;              // floats of gamma (then beta) of this CTA's K slice (<= 8 k blocks: d <= 2048)
constexpr int kSredFloats = kMaxT * 4 * kRowsPerCta * 2;     // [tile][4 quadrant warps][16 rows](sum, sum of squares)
// ---- shared-memory layout (offsets from the 1024-aligned base)
constexpr int kOffX = kNWC * kWChunk;
constexpr int kOffRecv = kOffX + kNX * kXStage;
constexpr int kURegion = kNX * kXStage + kMaxT * kRecvBuf;   // X ring + receive buffers; during attention: 8 KB of K/V staging per warp
static_assert(kAttnWarps2 * 8192 <= kURegion, "U region");
constexpr int kOffStrips = kOffX + kURegion;
constexpr int kOffRowstat = kOffStrips + kAttnWarps2 * 192 * 4;
constexpr int kOffSred = kOffRowstat + kN * 2 * 4;
constexpr int kOffGB = kOffSred + kSredFloats * 4;
constexpr int kOffRed = kOffGB + 2 * kGBHalf * 4;
constexpr int kOffSched = kOffRed + 64;
constexpr int kOffBars = kOffSched + 64;
constexpr int kBarWFull = kOffBars, kBarWEmpty = kBarWFull + 8 * kNWC, kBarXLFull = kBarWEmpty + 8 * kNWC, kBarXTFull = kBarXLFull + 8 * kNX,
              kBarXEmpty = kBarXTFull + 8 * kNX, kBarTFull = kBarXEmpty + 8 * kNX, kBarRecv = kBarTFull + 8, kBarXgo = kBarRecv + 8 * kMaxT,
              kBarVgo = kBarXgo + 8, kBarVdone = kBarVgo + 8, kOffTmemSlot = kBarVdone + 8;
constexpr int kSmem2 = kOffTmemSlot + 16 + 1024 /*alignment*/;
static_assert(kSmem2 <= 227 * 1024, "shared memory");

enum Flavor : int { FL_QKV = 0, FL_RESID = 1, FL_GELU = 2 };

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const uint64_t t0 = ptx::globaltimer_ns();
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0x3ff) == 0 && ptx::globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_dsmem_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ldcg_f1(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// 16 consecutive fp32 columns of this warp's TMEM quadrant
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// (count, mean, M2) of two disjoint sets -> their union (Chan et al.); b may be empty
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  const float nt = n + nb;
  if (nb > 0.f) {
    const float delta = meanb - mean, f = nb / nt;
    mean += delta * f;
    m2 += m2b + delta * delta * n * f;
    n = nt;
  }
}

// This CTA's share of one GEMM kind (the same in every layer): its cluster's tiles t0, t0 + ncl, ... and its K slice.
struct Slice {
  int nt, t0, kb0, kb1;
};
__device__ __forceinline__ int units_of(const Slice& s) { return s.nt > 0 ? s.kb1 - s.kb0 : 0; }
// k blocks handed over together: (k blocks) x (row tiles) <= 4 weight tiles and <= 16 MMAs per barrier wait
__device__ __forceinline__ int chunk_kb(const Slice& s) { return s.nt == 1 ? 4 : s.nt == 2 ? 2 : 1; }

// Activation ring bookkeeping, kept in step by every role that walks the schedule.  The ring restarts at slot 0 with
// every GEMM (it has drained by then); a chunk of ck k blocks occupies the slots [c * ck, c * ck + ck) and is handed
// over through the barriers of its first slot.  Parity bits per slot: uses as a first slot (empty), completions of the
// LayerNorm-staged full barrier, of the TMA full barrier.
struct XRing {
  uint32_t emask, lmask, tmask;
  __device__ __forceinline__ void skip(const Slice& sl) {   // a GEMM whose tiles another producer stages
    const int n = units_of(sl);
    if (n == 0) return;
    const int ck = chunk_kb(sl), nc = kNX / ck;
    int c = 0;
    for (int i = 0; i < n; i += ck) {
      emask ^= 1u << (c * ck);
      if (++c == nc) c = 0;
    }
  }
};

// gamma / beta of this CTA's K slice -> shared memory (issued before the grid-barrier wait: the vectors do not depend on
// the step, and fetched from HBM inside the staging loop they cost a DRAM round trip per tile)
__device__ __forceinline__ void prefetch_gb(uint32_t base, int ct, const Slice& sl, const float* gamma, const float* beta) {
  const int n4 = units_of(sl) * 16;   // float4 per vector
  for (int i = ct; i < 2 * n4; i += kComputeThreads) {
    const int vec = i >= n4 ? 1 : 0, idx = i - vec * n4;
    cp_async16(base + kOffGB + (vec * kGBHalf + idx * 4) * 4, (vec ? beta : gamma) + sl.kb0 * 64 + idx * 4, 16u);
  }
  cp_async_commit();
}

// LayerNorm-staged operand of this CTA's K slice: row statistics of h from the per-tile (mean, M2) table, then
// x[row, kb*64 .. +64) = bf16((h - mean) * rstd * gamma + beta) written as 128B-swizzled K-major tiles (what a TMA box
// {64, 64 rows} with SWIZZLE_128B would write).  thread = (row = ct / 4, 16 k values per tile).  The phase is a chain of
// L2 round trips, so the loads of the statistics and of the first four tiles are issued together, and the next pair is
// requested before a pair is converted.
__device__ __noinline__ void stage_ln_tiles(const Mega2Params& p, uint32_t base, uint8_t* gen, int ct, XRing& xr_io, const Slice sl) {
  const int row = ct >> 2, c4 = ct & 3, lane = ct & 31;
  const int kb0 = sl.kb0, kb1 = sl.kb1, ck = chunk_kb(sl), nc = kNX / ck;
  const bool row_ok = row < p.R;
  const int d = p.d;
  const float* hrow = p.h + static_cast<size_t>(row) * d + c4 * 16;
  float* rowstat = reinterpret_cast<float*>(gen + kOffRowstat);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 b0[8], b1[8];   // two pairs of tiles in flight
  auto load_pair = [&](float4 (&b)[8], int kb) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) b[j * 4 + i] = (row_ok && kb + j < kb1) ? ldcg_f4(hrow + (kb + j) * 64 + i * 4) : z4;
  };
  // ---- statistics: tiles part, part + 4, ... of this row
  {
    const int tiles = (d + kMT - 1) / kMT;
    const float* st = p.stats + static_cast<size_t>(row) * 2;
    float2 sv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = c4 + 4 * j;
      sv[j] = (row_ok && t < tiles) ? ldcg_f2(st + static_cast<size_t>(t) * (kN * 2)) : make_float2(0.f, 0.f);
    }
    load_pair(b0, kb0);
    load_pair(b1, kb0 + 2);
    float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = c4 + 4 * j;
      if (row_ok && t < tiles) chan_merge(n, mean, m2, static_cast<float>(min(kMT, d - t * kMT)), sv[j].x, sv[j].y);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o), qb = __shfl_xor_sync(0xffffffffu, m2, o);
      chan_merge(n, mean, m2, nb, mb, qb);   // (only the c4 == 0 lane's result is used: the merge is not symmetric in fp)
    }
    if (c4 == 0) {
      rowstat[row * 2] = mean;
      rowstat[row * 2 + 1] = row_ok ? rsqrtf(m2 / static_cast<float>(d) + p.eps) : 0.f;
    }
    cp_async_wait<0>();   // this thread's share of gamma / beta (prefetch_gb)
    ptx::named_bar_sync(2, kComputeThreads);
  }
  const float mean = rowstat[row * 2], rstd = rowstat[row * 2 + 1];
  const uint32_t sw = static_cast<uint32_t>(row & 7);
  const uint32_t gb = base + kOffGB + static_cast<uint32_t>(c4) * 64u;
  uint32_t emask = xr_io.emask;
  int c = 0;   // chunk slot
  auto emit = [&](const float4* v4, int i) {   // tile i of the slice
    uint32_t pk[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 g = lds_f4(gb + static_cast<uint32_t>(i * 64 + q * 4) * 4u), b = lds_f4(gb + static_cast<uint32_t>(kGBHalf + i * 64 + q * 4) * 4u);
      const float4 v = v4[q];
      pk[2 * q] = row_ok ? pack_bf16x2((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y) : 0u;
      pk[2 * q + 1] = row_ok ? pack_bf16x2((v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w) : 0u;
    }
    const int j = i & (ck - 1);
    const uint32_t first = static_cast<uint32_t>(c * ck);
    if (j == 0) {
      ptx::mbar_wait(base + kBarXEmpty + 8u * first, ((emask >> first) & 1u) ^ 1u);
      emask ^= 1u << first;
    }
    const uint32_t dst = base + kOffX + (first + static_cast<uint32_t>(j)) * kXStage + static_cast<uint32_t>(row) * 128u;
    sts_u4(dst + (((2u * c4) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    sts_u4(dst + (((2u * c4 + 1u) ^ sw) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
    if (j == ck - 1 || kb0 + i + 1 == kb1) {
      ptx::fence_proxy_async();   // generic stores -> the MMA's async-proxy reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(base + kBarXLFull + 8u * first);
      if (++c == nc) c = 0;
    }
  };
  const int n = kb1 - kb0;
#pragma unroll 1
  for (int i = 0; i < n; i += 4) {
    emit(b0, i);
    if (i + 1 < n) emit(b0 + 4, i + 1);
    load_pair(b0, kb0 + i + 4);
    if (i + 2 < n) emit(b1, i + 2);
    if (i + 3 < n) emit(b1 + 4, i + 3);
    load_pair(b1, kb0 + i + 6);
  }
  xr_io.emask = emask;
}

struct Epi2 {
  uint32_t tmem_base;
  uint32_t tph;        // parity of t_full
  uint32_t rph;        // parity bits of recv_full[0 .. kMaxT)
};

// Epilogue of one GEMM kind for this cluster's tiles: TMEM partial -> reduce-scatter over DSMEM -> rank-ordered sum ->
// bias + flavour -> global.  An M = 64 accumulator keeps feature 16 q + i in TMEM lane 32 q + i (i < 16): compute warp w
// reads quadrant q = w % 4 (lanes 0-15 carry data) and column half w / 4 (rows 32 half .. 32 half + 31).
// Receive buffer of a tile: [slot][feature 64][row 16] f32, the four 16-byte chunks of a feature XOR-swizzled with
// (feature / 2) % 4 so that the finaliser's 128-bit reads (one feature per lane, 64-byte pitch) are conflict free.
template <int FL>
__device__ __noinline__ void epilogue2(const Mega2Params& p, uint32_t base, uint8_t* gen, int ct, int layer, Epi2& ec_io, const Slice sl, int kind,
                                       uint32_t rank, const float* bias) {
  if (sl.nt == 0) return;
  const int kbt = p.g[kind].kb, rows_out = p.g[kind].rows_out, ncl = p.ncl;
  const bool have = sl.kb1 > sl.kb0;
  const int lane = ct & 31, cw = ct >> 5, quad = cw & 3, half = cw >> 2;
  const int R = p.R, d = p.d;
  const int fl = quad * 16 + (lane & 15);          // feature inside the tile (lanes 16-31 mirror 0-15 and stay idle)
  const bool lane_ok = lane < 16;
  const uint32_t chsw = static_cast<uint32_t>((fl >> 1) & 3);
  const uint32_t tmem_q = ec_io.tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  Epi2 ec = ec_io;
  // biases of the tiles this warp group finalises (ti = half, half + 2): requested before any wait
  float bvs[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int gf = (sl.t0 + (half + 2 * k) * ncl) * kMT + fl;
    bvs[k] = (lane_ok && half + 2 * k < sl.nt && gf < rows_out) ? __ldg(bias + gf) : 0.f;
  }
  if (have) {
    ptx::mbar_wait(base + kBarTFull, ec.tph);
    ec.tph ^= 1u;
    ptx::tc_fence_after();
  }
  if (ct == 0) MEGA_RSTAMP(layer, kind * 8 + 3);
  // ---- push the rows of the other CTAs: columns [0,16) of this half belong to rank 2*half, [16,32) to 2*half + 1.
  // All tiles first, one release per destination barrier afterwards (a single drain of the remote stores).
  if (have) {
#pragma unroll 1
    for (int ti = 0; ti < sl.nt; ++ti) {
      uint32_t r[32];
      ptx::tmem_ld32(tmem_q + static_cast<uint32_t>(ti * kN + half * 32), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t dest = static_cast<uint32_t>(2 * half + hh);
        if (dest == rank || !lane_ok) continue;
        const uint32_t slot = rank < dest ? rank : rank - 1u;   // my slot among the dest's three remote ranks
        const uint32_t local = base + kOffRecv + static_cast<uint32_t>(ti) * kRecvBuf + (slot * kMT + static_cast<uint32_t>(fl)) * 64u;
        const uint32_t remote = ptx::mapa_shared(local, dest);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_dsmem_v4(remote + ((static_cast<uint32_t>(c) ^ chsw) << 4), r[hh * 16 + 4 * c], r[hh * 16 + 4 * c + 1], r[hh * 16 + 4 * c + 2],
                      r[hh * 16 + 4 * c + 3]);
      }
    }
    ptx::tc_fence_before();
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll 1
    for (int ti = 0; ti < sl.nt; ++ti) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t dest = static_cast<uint32_t>(2 * half + hh);
        if (dest != rank) ptx::mbar_arrive_cluster(ptx::mapa_shared(base + kBarRecv + 8u * ti, dest));   // release.cluster
      }
    }
  }
  if (ct == 0) MEGA_RSTAMP(layer, kind * 8 + 4);
  // ---- my 16 rows: warp group `half` finalises the tiles ti = half, half + 2; own partial from TMEM, the others from the
  // receive buffer
  const int row0 = static_cast<int>(rank) * kRowsPerCta;
#pragma unroll 1
  for (int ti = half; ti < sl.nt; ti += 2) {
    const int tile = sl.t0 + ti * ncl;
    const int gf = tile * kMT + fl;
    const bool f_ok = lane_ok && gf < rows_out;
    float hv[16];
    if (FL == FL_RESID) {
#pragma unroll
      for (int v = 0; v < 16; ++v) hv[v] = (f_ok && row0 + v < R) ? ldcg_f1(p.h + static_cast<size_t>(row0 + v) * d + gf) : 0.f;
    }
    const float bv = ti == half ? bvs[0] : bvs[1];
    uint32_t own[16];
    if (have) {
      ptx::tc_fence_after();
      tmem_ld16(tmem_q + static_cast<uint32_t>(ti * kN) + rank * kRowsPerCta, own);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
    }
    mbar_wait_cluster(base + kBarRecv + 8u * ti, (ec.rph >> ti) & 1u);
    ec.rph ^= 1u << ti;
    if (ct == 0 && ti == 0) MEGA_RSTAMP(layer, kind * 8 + 5);
    float acc[16];
#pragma unroll
    for (int v = 0; v < 16; ++v) acc[v] = 0.f;
    const uint32_t rb = base + kOffRecv + static_cast<uint32_t>(ti) * kRecvBuf + static_cast<uint32_t>(fl) * 64u;
#pragma unroll
    for (int s = 0; s < kS; ++s) {                       // rank order: deterministic
      if (s * kbt / kS >= (s + 1) * kbt / kS) continue;  // that rank had no K slice
      if (static_cast<uint32_t>(s) == rank) {
#pragma unroll
        for (int v = 0; v < 16; ++v) acc[v] += __uint_as_float(own[v]);
      } else {
        const uint32_t slot = static_cast<uint32_t>(s) < rank ? s : s - 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 e = lds_f4(rb + slot * (kMT * 64u) + ((static_cast<uint32_t>(c) ^ chsw) << 4));
          acc[4 * c] += e.x; acc[4 * c + 1] += e.y; acc[4 * c + 2] += e.z; acc[4 * c + 3] += e.w;
        }
      }
    }
    if (FL == FL_QKV) {
      bf16* out = p.qkv + gf;
#pragma unroll
      for (int v = 0; v < 16; ++v)
        if (f_ok && row0 + v < R) out[static_cast<size_t>(row0 + v) * rows_out] = __float2bfloat16_rn(acc[v] + bv);
    } else if (FL == FL_GELU) {
      bf16* out = p.mlp + gf;
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        // same form as the GEMM epilogue's gelu_new: 0.5 x (1 + tanh(u)) == x - x / (1 + exp(2u))
        const float xx = acc[v] + bv;
        const float u = 0.7978845608028654f * (xx + 0.044715f * xx * xx * xx);
        const float o = xx - __fdividef(xx, 1.f + __expf(2.f * u));
        if (f_ok && row0 + v < R) out[static_cast<size_t>(row0 + v) * rows_out] = __float2bfloat16_rn(o);
      }
    } else {
      // residual update + this tile's (mean, M2) of every row for the LayerNorm of the consumer.  One pass: 16 sums and
      // 16 sums of squares per lane, reduced over the 16 feature lanes by a transposed butterfly (30 shuffles): lane i
      // ends up with the totals of row i.
      float rs[32];
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const float nv = hv[v] + (acc[v] + bv);
        if (f_ok && row0 + v < R) p.h[static_cast<size_t>(row0 + v) * d + gf] = nv;
        rs[v] = f_ok ? nv : 0.f;
        rs[16 + v] = f_ok ? nv * nv : 0.f;
      }
      // level with lane bit b: keep the half of the rows whose bit matches this lane's, add the partner's
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const bool up = (lane & 8) != 0;
        const float s_keep = up ? rs[8 + v] : rs[v], s_send = up ? rs[v] : rs[8 + v];
        const float q_keep = up ? rs[24 + v] : rs[16 + v], q_send = up ? rs[16 + v] : rs[24 + v];
        rs[v] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 8);
        rs[16 + v] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 8);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const bool up = (lane & 4) != 0;
        const float s_keep = up ? rs[4 + v] : rs[v], s_send = up ? rs[v] : rs[4 + v];
        const float q_keep = up ? rs[20 + v] : rs[16 + v], q_send = up ? rs[16 + v] : rs[20 + v];
        rs[v] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 4);
        rs[16 + v] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 4);
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const bool up = (lane & 2) != 0;
        const float s_keep = up ? rs[2 + v] : rs[v], s_send = up ? rs[v] : rs[2 + v];
        const float q_keep = up ? rs[18 + v] : rs[16 + v], q_send = up ? rs[16 + v] : rs[18 + v];
        rs[v] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 2);
        rs[16 + v] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 2);
      }
      {
        const bool up = (lane & 1) != 0;
        const float s_keep = up ? rs[1] : rs[0], s_send = up ? rs[0] : rs[1];
        const float q_keep = up ? rs[17] : rs[16], q_send = up ? rs[16] : rs[17];
        rs[0] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 1);
        rs[16] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 1);
      }
      // lane i (< 16) now holds row i of this warp's 16 features
      float* sr = reinterpret_cast<float*>(gen + kOffSred) + ti * (4 * kRowsPerCta * 2);
      if (lane_ok) {
        sr[(quad * kRowsPerCta + lane) * 2] = rs[0];
        sr[(quad * kRowsPerCta + lane) * 2 + 1] = rs[16];
      }
      if (half == 0) ptx::named_bar_sync(3, 128); else ptx::named_bar_sync(4, 128);   // the four warps of this finaliser group
      if (quad == 0 && lane_ok) {
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          sum += sr[(q * kRowsPerCta + lane) * 2];
          sq += sr[(q * kRowsPerCta + lane) * 2 + 1];
        }
        const float nt = static_cast<float>(min(kMT, rows_out - tile * kMT));
        const float mean = sum / nt;
        *reinterpret_cast<float2*>(p.stats + (static_cast<size_t>(tile) * kN + row0 + lane) * 2) = make_float2(mean, fmaxf(sq - sum * mean, 0.f));
      }
    }
  }
  ec_io = ec;
}

// h[t] = wte[token] + wpe[position] and the per-tile statistics, rows t = cta, cta + ncta, ...
__device__ __noinline__ void embed_phase2(const Mega2Params& p, ComputeCtx& cc, int cta) {
  const int d = p.d, nq = d >> 2;
  for (int t = cta; t < p.R; t += p.ncta) {
    const int tok = p.tokens[t];
    const uint2* te = reinterpret_cast<const uint2*>(p.wte + static_cast<size_t>(tok) * d);
    const uint2* pe = p.wpe ? reinterpret_cast<const uint2*>(p.wpe + static_cast<size_t>(p.ctx_len[t]) * d) : nullptr;
    float4* hrow = reinterpret_cast<float4*>(p.h + static_cast<size_t>(t) * d);
    // a warp iteration covers exactly two 64-feature tiles (16 lanes x 4 features each)
    for (int q0 = cc.cw * 32; q0 < nq; q0 += kComputeThreads) {
      const int q = q0 + cc.lane;
      const bool ok = q < nq;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const uint2 e = __ldg(te + q);
        const float2 e0 = unpack_bf16x2(e.x), e1 = unpack_bf16x2(e.y);
        a = make_float4(e0.x, e0.y, e1.x, e1.y);
        if (pe != nullptr) {
          const uint2 w = __ldg(pe + q);
          const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y);
          a.x += w0.x; a.y += w0.y; a.z += w1.x; a.w += w1.y;
        }
        hrow[q] = a;
      }
      const int tile = (q0 >> 4) + (cc.lane >> 4);
      const int nv = max(1, min(kMT, d - tile * kMT));
      float sum = (a.x + a.y) + (a.z + a.w);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / static_cast<float>(nv);
      const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
      float m2 = ok ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      if ((cc.lane & 15) == 0 && tile * kMT < d) *reinterpret_cast<float2*>(p.stats + (static_cast<size_t>(tile) * kN + t) * 2) = make_float2(mean, m2);
    }
  }
}

// x[t] = LayerNorm_f(h[t]) for the lm_head GEMM, rows t = cta, cta + ncta, ...
__device__ __noinline__ void lnf_phase2(const Mega2Params& p, ComputeCtx& cc, int cta) {
  const int d = p.d, nq = d >> 2, tiles = (d + kMT - 1) / kMT;
  for (int t = cta; t < p.R; t += p.ncta) {
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int tl = cc.lane; tl < tiles; tl += 32) {
      const float2 st = ldcg_f2(p.stats + (static_cast<size_t>(tl) * kN + t) * 2);
      chan_merge(n, mean, m2, static_cast<float>(min(kMT, d - tl * kMT)), st.x, st.y);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o), qb = __shfl_xor_sync(0xffffffffu, m2, o);
      chan_merge(n, mean, m2, nb, mb, qb);
    }
    mean = __shfl_sync(0xffffffffu, mean, 0);   // (lane 0's merge order: every warp computes the same bits)
    m2 = __shfl_sync(0xffffffffu, m2, 0);
    const float rstd = rsqrtf(m2 / static_cast<float>(d) + p.eps);
    const float* hrow = p.h + static_cast<size_t>(t) * d;
    uint2* xr = reinterpret_cast<uint2*>(p.x + static_cast<size_t>(t) * d);
    for (int q = cc.ct; q < nq; q += kComputeThreads) {
      const float4 v = ldcg_f4(hrow + q * 4);
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.lnf_g) + q), b4 = __ldg(reinterpret_cast<const float4*>(p.lnf_b) + q);
      uint2 pk;
      pk.x = pack_bf16x2((v.x - mean) * rstd * g4.x + b4.x, (v.y - mean) * rstd * g4.y + b4.y);
      pk.y = pack_bf16x2((v.z - mean) * rstd * g4.z + b4.z, (v.w - mean) * rstd * g4.w + b4.w);
      xr[q] = pk;
    }
  }
}

__global__ void __launch_bounds__(kThreads2, 1) decode_mega2_kernel(const __grid_constant__ CUtensorMap xmap_att,
                                                                    const __grid_constant__ CUtensorMap xmap_mlp,
                                                                    const __grid_constant__ Mega2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;   // (the same offset in every CTA of the cluster: same kernel, same layout)
  uint8_t* gen = smem_raw + (base - raw_u32);
  float* strips = reinterpret_cast<float*>(gen + kOffStrips);
  Slice* sched = reinterpret_cast<Slice*>(gen + kOffSched);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + kOffTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = p.ncta;
  const uint32_t rank = ptx::cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNWC; ++s) {
      ptx::mbar_init(base + kBarWFull + 8u * s, 1);
      ptx::mbar_init(base + kBarWEmpty + 8u * s, 1);
    }
    for (int s = 0; s < kNX; ++s) {
      ptx::mbar_init(base + kBarXLFull + 8u * s, 8);   // one arrival per compute warp
      ptx::mbar_init(base + kBarXTFull + 8u * s, 1);
      ptx::mbar_init(base + kBarXEmpty + 8u * s, 1);
    }
    ptx::mbar_init(base + kBarTFull, 1);
    for (int t = 0; t < kMaxT; ++t) ptx::mbar_init(base + kBarRecv + 8u * t, 3 * 4);   // 3 remote CTAs x the 4 warps that hold my rows
    ptx::mbar_init(base + kBarXgo, 1);
    ptx::mbar_init(base + kBarVgo, 1);
    ptx::mbar_init(base + kBarVdone, kHelpers2);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x >= 32 && threadIdx.x < 36) {
    const int kind = threadIdx.x - 32;
    const Mega2Gemm g = p.g[kind];
    Slice s;
    int t0 = cta / kS - g.off;
    if (t0 < 0) t0 += p.ncl;
    s.t0 = t0;
    s.nt = t0 < g.tiles ? (g.tiles - t0 + p.ncl - 1) / p.ncl : 0;   // tiles t0, t0 + ncl, ... (<= kMaxT, host-checked)
    s.kb0 = static_cast<int>(rank) * g.kb / kS;
    s.kb1 = (static_cast<int>(rank) + 1) * g.kb / kS;
    sched[kind] = s;
  }
  if (warp == 3) ptx::tmem_alloc<kMaxT * kN>(base + kOffTmemSlot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // every CTA of the cluster has initialised its barriers before anyone signals a peer
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ W producer: chunks of <= 4 weight tiles
    uint32_t wc = 0, ph = 0;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
#pragma unroll 1
      for (int kind = 0; kind < 4; ++kind) {
        const CUtensorMap* wm = p.wmaps + (l * 4 + kind);
        const Slice sl = sched[kind];
        if (units_of(sl) == 0) continue;
        const int ck = chunk_kb(sl);
#pragma unroll 1
        for (int kb = sl.kb0; kb < sl.kb1; kb += ck) {
          const int nk = min(ck, sl.kb1 - kb);
          ptx::mbar_wait(base + kBarWEmpty + 8u * wc, ph ^ 1);
          if (ptx::elect_one()) {
            const uint32_t full = base + kBarWFull + 8u * wc;
            ptx::mbar_arrive_expect_tx(full, static_cast<uint32_t>(nk * sl.nt) * kWTile);
            uint32_t dst = base + wc * kWChunk;
            for (int j = 0; j < nk; ++j)
              for (int ti = 0; ti < sl.nt; ++ti, dst += kWTile)
                ptx::tma_load_2d(dst, wm, full, (kb + j) * 64, (sl.t0 + ti * p.ncl) * kMT, ptx::kEvictFirst);
          }
          __syncwarp();
          if (++wc == kNWC) { wc = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ X producer (att / mlp tiles) + attention helper
    ComputeCtx hc;
    init_attn_ctx(p, hc, 8, kAttnWarps2, lane, cta, nullptr, gen + kOffX, base + kOffX, strips);
    XRing xr = {0u, 0u, 0u};
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
#pragma unroll 1
      for (int kind = 0; kind < 4; ++kind) {
        if (kind == 1) helper_attention<false>(p, hc, cta, l, base + kBarVgo, base + kBarVdone, false);
        const Slice sl = sched[kind];
        if ((kind & 1) == 0) {
          xr.skip(sl);   // staged by the compute warps
          continue;
        }
        if (units_of(sl) == 0) continue;
        // barriers of a layer: #5l+1 after A, +2 after B, +3 after C, +4 after D, +5 after E (#0 after the embedding)
        const unsigned nb = static_cast<unsigned>(5 * l + (kind == 1 ? 2 : 4)) + 1u;
        ptx::mbar_wait(base + kBarXgo, kind == 3 ? 1u : 0u);   // this CTA's compute warps arrived there (their (2l + kind/2)-th signal)
        if (lane == 0) {
          poll_counter(p.sync, nb * static_cast<unsigned>(ncta));
          fence_proxy_async_all();
          MEGA_RSTAMP(l, kind * 8 + 1);
        }
        __syncwarp();
        const CUtensorMap* xm = kind == 1 ? &xmap_att : &xmap_mlp;
        const int ck = chunk_kb(sl), nc = kNX / ck;
        int c = 0;
#pragma unroll 1
        for (int kb = sl.kb0; kb < sl.kb1; kb += ck) {
          const int nk = min(ck, sl.kb1 - kb);
          const uint32_t first = static_cast<uint32_t>(c * ck);
          ptx::mbar_wait(base + kBarXEmpty + 8u * first, ((xr.emask >> first) & 1u) ^ 1u);
          xr.emask ^= 1u << first;
          if (lane == 0) {
            const uint32_t full = base + kBarXTFull + 8u * first;
            ptx::mbar_arrive_expect_tx(full, static_cast<uint32_t>(nk) * kXStage);
            for (int j = 0; j < nk; ++j)
              ptx::tma_load_2d(base + kOffX + (first + j) * kXStage, xm, full, (kb + j) * 64, 0, ptx::kEvictLast);
          }
          __syncwarp();
          if (++c == nc) c = 0;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer + attention helper
    ComputeCtx hc;
    init_attn_ctx(p, hc, 9, kAttnWarps2, lane, cta, nullptr, gen + kOffX, base + kOffX, strips);
    const uint32_t idesc = ptx::umma_idesc_bf16(kMT, kN);
    const uint64_t wdesc0 = ptx::umma_desc_k_sw128(base), xdesc0 = ptx::umma_desc_k_sw128(base + kOffX);
    constexpr uint32_t kTileStep = kWTile >> 4;   // == kXStage >> 4
    XRing xr = {0u, 0u, 0u};
    uint32_t wc = 0, wph = 0;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
#pragma unroll 1
      for (int kind = 0; kind < 4; ++kind) {
        if (kind == 1) helper_attention<false>(p, hc, cta, l, base + kBarVgo, base + kBarVdone, false);
        const Slice sl = sched[kind];
        if (units_of(sl) == 0) continue;
        const bool tma_kind = (kind & 1) != 0;
        const int ck = chunk_kb(sl), nc = kNX / ck, nt = sl.nt;
        int c = 0;
#pragma unroll 1
        for (int kb = sl.kb0; kb < sl.kb1; kb += ck) {
          const int nk = min(ck, sl.kb1 - kb);
          const uint32_t first = static_cast<uint32_t>(c * ck);
          if (tma_kind) {
            ptx::mbar_wait(base + kBarXTFull + 8u * first, (xr.tmask >> first) & 1u);
            xr.tmask ^= 1u << first;
          } else {
            ptx::mbar_wait(base + kBarXLFull + 8u * first, (xr.lmask >> first) & 1u);
            xr.lmask ^= 1u << first;
          }
          ptx::mbar_wait(base + kBarWFull + 8u * wc, wph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            if (kb == sl.kb0) MEGA_RSTAMP(l, kind * 8 + 7);
            uint64_t wd = wdesc0 + static_cast<uint64_t>(wc * (kWChunk >> 4));
            uint64_t xd = xdesc0 + static_cast<uint64_t>(first * kTileStep);
            for (int j = 0; j < nk; ++j, xd += kTileStep) {
              const uint32_t accf = kb + j > sl.kb0 ? 1u : 0u;
              for (int ti = 0; ti < nt; ++ti, wd += kTileStep) {
                const uint32_t tacc = tmem_base + static_cast<uint32_t>(ti * kN);
                ptx::umma_bf16(tacc, wd, xd, idesc, accf);
                ptx::umma_bf16(tacc, wd + 2u, xd + 2u, idesc, 1u);
                ptx::umma_bf16(tacc, wd + 4u, xd + 4u, idesc, 1u);
                ptx::umma_bf16(tacc, wd + 6u, xd + 6u, idesc, 1u);
              }
            }
            ptx::umma_commit(base + kBarWEmpty + 8u * wc);
            ptx::umma_commit(base + kBarXEmpty + 8u * first);
            if (kb + ck >= sl.kb1) ptx::umma_commit(base + kBarTFull);
          }
          __syncwarp();
          if (++wc == kNWC) { wc = 0; wph ^= 1; }
          if (++c == nc) c = 0;
        }
      }
    }
  } else if (warp == 3 || warp >= kComputeWarp0 + 8) {
    // ------------------------------------------------------------------ TMEM owner / attention-only warps
    ComputeCtx hc;
    init_attn_ctx(p, hc, warp == 3 ? 10 : warp - 1, kAttnWarps2, lane, cta, nullptr, gen + kOffX, base + kOffX, strips);
    for (int l = 0; l < p.L; ++l) helper_attention<false>(p, hc, cta, l, base + kBarVgo, base + kBarVdone, false);
  } else {
    // ------------------------------------------------------------------ compute warps
    ComputeCtx cc;
    cc.ct = threadIdx.x - kComputeWarp0 * 32;
    cc.red = reinterpret_cast<float*>(gen + kOffRed);
    cc.red_it = 0;
    cc.sched = nullptr;
    cc.ctr = p.sync;
    cc.ncta = ncta;
    cc.trace = p.trace ? p.trace + static_cast<size_t>(cta) * (2 * (p.nbar + 2)) : nullptr;
    cc.trace_it = 0;
    init_attn_ctx(p, cc, warp - kComputeWarp0, kAttnWarps2, lane, cta, nullptr, gen + kOffX, base + kOffX, strips);
    const int ct = cc.ct;
    Epi2 ec = {tmem_base, 0u, 0u};
    XRing xr = {0u, 0u, 0u};
    unsigned nb = 0;   // grid barriers this CTA has arrived at

    embed_phase2(p, cc, cta);
    grid_arrive(cc, 0, 0, 0, false);   // #0
    ++nb;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
      const MegaLayer* ly = p.layers + l;
      cc.cur_layer = l;
      // ---- A: qkv = LN_1(h) Wqkv^T + b
      {
        const Slice sl = sched[0];
        const bool work = units_of(sl) > 0;
        if (work) prefetch_gb(base, ct, sl, ly->ln1_g, ly->ln1_b);
        grid_wait(cc, nb);
        if (ct == 0) MEGA_RSTAMP(l, 0);
        if (work) stage_ln_tiles(p, base, gen, ct, xr, sl);
        if (ct == 0) MEGA_RSTAMP(l, 2);
        epilogue2<FL_QKV>(p, base, gen, ct, l, ec, sl, 0, rank, ly->b_qkv);
        if (ct == 0) MEGA_RSTAMP(l, 6);
        grid_arrive(cc, 0, 0, 0, false);
        ++nb;
      }
      // ---- B: attention (the first K/V batches travel while the barrier completes; warp 0 polls it)
      if (cc.cw != 0) attention_phase<false>(p, cc, cta, l, nullptr, true);
      grid_wait(cc, nb);
      if (ct == 0) ptx::mbar_arrive(base + kBarVgo);
      attention_phase<false>(p, cc, cta, l, nullptr, false);
      grid_arrive(cc, base + kBarXgo, base + kBarVdone, static_cast<uint32_t>(l) & 1u, true);
      ++nb;
      // ---- C: h += att Wproj^T + b
      {
        const Slice sl = sched[1];
        grid_wait(cc, nb);
        if (ct == 0) MEGA_RSTAMP(l, 8);
        xr.skip(sl);
        epilogue2<FL_RESID>(p, base, gen, ct, l, ec, sl, 1, rank, ly->b_proj);
        if (ct == 0) MEGA_RSTAMP(l, 14);
        grid_arrive(cc, 0, 0, 0, false);
        ++nb;
      }
      // ---- D: mlp = gelu_new(LN_2(h) Wfc^T + b)
      {
        const Slice sl = sched[2];
        const bool work = units_of(sl) > 0;
        if (work) prefetch_gb(base, ct, sl, ly->ln2_g, ly->ln2_b);
        grid_wait(cc, nb);
        if (ct == 0) MEGA_RSTAMP(l, 16);
        if (work) stage_ln_tiles(p, base, gen, ct, xr, sl);
        if (ct == 0) MEGA_RSTAMP(l, 18);
        epilogue2<FL_GELU>(p, base, gen, ct, l, ec, sl, 2, rank, ly->b_fc);
        if (ct == 0) MEGA_RSTAMP(l, 22);
        grid_arrive(cc, base + kBarXgo, 0, 0, true);
        ++nb;
      }
      // ---- E: h += mlp Wfc2^T + b
      {
        const Slice sl = sched[3];
        grid_wait(cc, nb);
        if (ct == 0) MEGA_RSTAMP(l, 24);
        xr.skip(sl);
        epilogue2<FL_RESID>(p, base, gen, ct, l, ec, sl, 3, rank, ly->b_fc2);
        if (ct == 0) MEGA_RSTAMP(l, 30);
        grid_arrive(cc, 0, 0, 0, false);
        ++nb;
      }
    }
    grid_wait(cc, nb);
    lnf_phase2(p, cc, cta);
    grid_arrive(cc, 0, 0, 0, false);   // exit barrier; CTA 0 re-arms the counter for the next launch
    ++nb;
    if (cta == 0 && ct == 0) {
      poll_counter(p.sync, nb * static_cast<unsigned>(ncta));
      *reinterpret_cast<volatile unsigned*>(p.sync) = 0u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  // no CTA leaves while a peer may still write into its shared memory / signal its barriers
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (warp == 3) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kMaxT * kN>(tmem_base);
  }
}

cudaLaunchConfig_t mega2_config(int ncta, size_t smem, cudaStream_t s, cudaLaunchAttribute* attr) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ncta);
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they wait on one another
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // (profilers replay the launch and refuse the cooperative attribute together with clusters: CCB_MEGA2_NOCOOP=1 drops it;
  //  co-residency then rests on the occupancy query alone)
  static const bool nocoop = [] { const char* e = getenv("CCB_MEGA2_NOCOOP"); return e && e[0] == '1'; }();
  cfg.numAttrs = nocoop ? 1 : 2;
  return cfg;
}


}  // namespace

int mega2_init(int* max_clusters) {
  cudaError_t e = cudaFuncSetAttribute(decode_mega2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = mega2_config(kS, kSmem2, nullptr, attr);
  cfg.numAttrs = 1;   // (the occupancy query takes the cluster shape only)
  int n = 0;
  e = cudaOccupancyMaxActiveClusters(&n, decode_mega2_kernel, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return static_cast<int>(e);
  }
  *max_clusters = n;
  return 0;
}

// Tile -> cluster offsets: kinds in order of bytes per (tile, CTA), each placed where the most loaded cluster stays lowest.
bool mega2_plan(MegaState& m, int d, int ff, int ncl) {
  const int rows[4] = {3 * d, d, ff, d};
  const int ks[4] = {d, d, d, ff};
  std::vector<long long> load(ncl, 0);
  int order[4] = {0, 1, 2, 3};
  for (int i = 0; i < 4; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (ks[order[j]] > ks[order[i]]) std::swap(order[i], order[j]);
  for (int oi = 0; oi < 4; ++oi) {
    const int k = order[oi];
    Mega2Gemm& g = m.g2[k];
    g.rows_out = rows[k];
    g.kb = ks[k] / 64;
    g.tiles = (rows[k] + kMT - 1) / kMT;
    if (ks[k] % 64 || g.tiles > kMaxT * ncl || d > 2048) return false;
    long long best_max = -1, best_sq = -1;
    int best_off = 0;
    for (int off = 0; off < ncl; ++off) {
      long long mx = 0, sq = 0;
      for (int c = 0; c < ncl; ++c) {
        int t0 = c - off;
        if (t0 < 0) t0 += ncl;
        const int nt = t0 < g.tiles ? (g.tiles - t0 + ncl - 1) / ncl : 0;
        const long long v = load[c] + static_cast<long long>(nt) * g.kb;
        mx = std::max(mx, v);
        sq += v * v;
      }
      if (best_max < 0 || mx < best_max || (mx == best_max && sq < best_sq)) {
        best_max = mx;
        best_sq = sq;
        best_off = off;
      }
    }
    g.off = best_off;
    for (int c = 0; c < ncl; ++c) {
      int t0 = c - best_off;
      if (t0 < 0) t0 += ncl;
      load[c] += static_cast<long long>(t0 < g.tiles ? (g.tiles - t0 + ncl - 1) / ncl : 0) * g.kb;
    }
  }
  m.ncl = ncl;
  return true;
}

int mega2_launch(const Mega2Params& p_in, cudaStream_t s) {
  Mega2Params p = p_in;
  if (p.d > 2048 || p.d % 64 || p.ff % 64 || p.R > kN || p.R < 1 || p.ncta != p.ncl * kS) return static_cast<int>(cudaErrorInvalidValue);
  p.nW = kNWC;
  {
    const char* lay = getenv("CCB_MEGA_LAYERS");  // tuning / debugging only: run the first n layers
    if (lay && atoi(lay) > 0 && atoi(lay) < p.L) p.L = atoi(lay);
  }
  p.nbar = 5 * p.L + 1;
  CUtensorMap ma, mm;
  if (gemm_make_tmap(&ma, p.att, p.R, p.d, p.d, kN)) return -1;
  if (gemm_make_tmap(&mm, p.mlp, p.R, p.ff, p.ff, kN)) return -1;
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg = mega2_config(p.ncta, kSmem2, s, attr);
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_mega2_kernel, ma, mm, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace ccb
