// sd_cnn.cuh -- CNN encoder kernels (world_model/networks.py:59-98,192-234) for sm_100a.
//
// One encoder stage is  conv k=5 'SAME' stride 1 (+bias) -> MaxPool2d(2,2) -> RMSNorm over channels (eps 1e-4) -> SiLU.
// The forward kernels compute a stage as an implicit GEMM on tcgen05 with the pooling window folded into the tile:
//
//   a tile is 128 POOLED pixels; the four conv outputs of a pooling window live in four TMEM accumulators
//   acc[q = 2 dy + dx][128 x Cout] (q*64 columns apart), so the epilogue thread that owns TMEM lane r sees all four
//   members of window r in its own registers: max / arg-max, + bias, RMSNorm over its Cout channels, SiLU and the
//   bf16 NHWC store of the next stage's input need no cross-thread traffic at all.
//
//   K loop = the 6 x 6 distinct input shifts (sy, sx) = (dy + ky, dx + kx): "view" (sy, sx) is the 128 x Cin matrix of
//   input pixels (2 py + sy - 2, 2 px + sx - 2); accumulator (dy, dx) takes it with tap (sy - dy, sx - dx), so one
//   staged view feeds up to four MMAs (36 view loads for 100 tap-products).
//
// Operands are K-major in the UMMA canonical NO-swizzle layout, stored chunk-major: [K/8][rows][16 B] -- a core matrix
// (8 rows x 16 B) is 128 contiguous bytes, LBO (next 8 K elements) = rows*16 B, SBO (next 8 rows) = 128 B.  The
// software im2col producer (4 warps, thread = tile row) writes it with 16-byte cp.async (zero-fill = the SAME padding):
// a warp's 32 rows of one chunk are 512 contiguous bytes, so there is nothing to swizzle.
//
// Warp roles (416 threads): 0-3 producers, 4 = TMEM owner + single-thread MMA issuer, 5-8 and 9-12 two epilogue groups
// (TMEM lane quarter = warp & 3).  CTAs are persistent over tiles; two accumulator sets (2 x 256 TMEM columns), one per
// epilogue group, let the epilogues of tiles i and i+1 overlap each other and the MMAs of tile i+2.  Every wait is bounded
// (trap instead of hang).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sd_tc.cuh"

namespace sd {
namespace cnn {

using bf16 = __nv_bfloat16;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tc_mma_f16;

constexpr int BM = 128;
constexpr int THREADS1 = 416;      // stage-1 kernel: warps 0-3 producers, 4 MMA, 5-8 / 9-12 two epilogue groups (accumulator set 0 / 1)
constexpr int THREADS = 480;       // stages 2..L: + warp 13, the second MMA issuer (window row dy = 1), + warp 14, the proxy-fence helper
constexpr int MMA_WARP2 = 13;
constexpr int FENCE_WARP = 14;
constexpr int PROD = 128;          // producer threads (warps 0-3)
constexpr int MMA_WARP = 4;
constexpr int KSZ = 5;             // kernel size (configs/base.yaml encoder.cnn.kernel_size)
constexpr float kRmsEps = 1e-4f;   // networks.py:212
constexpr int LA = 2;              // cp.async groups in flight per producer thread before it publishes the oldest

// K-major, no swizzle (layout type 0): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48)
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// tcgen05.mma with the descriptors given as (low word, high word): K steps and tile offsets are 32-bit adds on the low word
__device__ __forceinline__ void mma_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum)
      : "memory");
}
// the mbarrier receives this thread's arrival when all of its earlier cp.async operations have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// 16 columns of four accumulators (`stride` columns apart), one wait
__device__ __forceinline__ void tmem_ld16x4(uint32_t taddr, uint32_t stride, float* a, float* b, float* c, float* d) {
  uint32_t r[64];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]), "=r"(r[16 * q + 4]),
          "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]), "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]),
          "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]), "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]), "=r"(r[16 * q + 14]),
          "=r"(r[16 * q + 15])
        : "r"(taddr + (uint32_t)q * stride));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[16 + i]);
    c[i] = __uint_as_float(r[32 + i]);
    d[i] = __uint_as_float(r[48 + i]);
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float silu(float m) { return __fdividef(m, 1.f + __expf(-m)); }

// What a stage writes (all nullable except one of y / embed).
struct StageOut {
  bf16* y;          // next stage's input, NHWC bf16 [N][Hp][Wp][cnext], channels >= cout zero
  float* pool;      // tape: pooled conv output incl. bias, pre-norm, fp32 [N][Hp][Wp][cp]
  uint8_t* arg;     // tape: position (2 dy + dx) of the maximum inside the window [N][Hp][Wp][cp]
  float* embed;     // last stage: fp32 [N][cout * Hp * Wp] flattened in the reference's (C, H, W) order
  const float* bias;  // [64] zero padded
  const float* gain;  // [64] zero padded
  int cout, cp, cnext, HpWp, total;
};

// Epilogue of one tile for the thread that owns TMEM lane `row` of accumulator set at `tm`.  TAPE: also keep the pooled
// pre-norm value and the arg-max position for the backward.  Written for a low instruction count: with one epilogue warp
// per scheduler the code is latency bound at ~8 cycles per instruction.
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool TAPE>
__device__ __forceinline__ void pool_norm_store_t(const StageOut& o, uint32_t tm, int g, const float* s_bias, const float* s_gain,
                                                  uint32_t bar_free, int lane) {
  float p[64];
  uint32_t aw[16];
  if (TAPE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) aw[i] = 0u;
  }
  float ss = 0.f;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    if (c0 < o.cp) {
      float a[16], b[16], c[16], d[16];
      tmem_ld16x4(tm + (uint32_t)c0, (uint32_t)o.cp, b, a, d, c);   // slots 0..3 = members (0,1), (0,0), (1,1), (1,0)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float m01 = fmaxf(a[i], b[i]), m23 = fmaxf(c[i], d[i]);
        const float v = fmaxf(m01, m23) + s_bias[c0 + i];
        if (TAPE) {
          // first maximum in window order (0,0), (0,1), (1,0), (1,1), as torch's max_pool2d backward routes it
          const uint32_t ar = m23 > m01 ? (d[i] > c[i] ? 3u : 2u) : (b[i] > a[i] ? 1u : 0u);
          aw[(c0 + i) >> 2] |= ar << (8 * (i & 3));
        }
        p[c0 + i] = v;
        ss = fmaf(v, v, ss);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) p[c0 + i] = 0.f;
    }
  }
  // the accumulator set is drained: hand it back to the MMA issuer before the (long) store phase
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_free);
  if (g >= o.total) return;
  const float rho = 1.f / sqrtf(ss / (float)o.cout + kRmsEps);
  if (TAPE) {
    float4* dst = reinterpret_cast<float4*>(o.pool + (size_t)g * o.cp);
#pragma unroll
    for (int c = 0; c < 64; c += 4)
      if (c < o.cp) dst[c >> 2] = make_float4(p[c], p[c + 1], p[c + 2], p[c + 3]);
    uint4* da = reinterpret_cast<uint4*>(o.arg + (size_t)g * o.cp);
#pragma unroll
    for (int c = 0; c < 64; c += 16)
      if (c < o.cp) da[c >> 4] = make_uint4(aw[c >> 2], aw[(c >> 2) + 1], aw[(c >> 2) + 2], aw[(c >> 2) + 3]);
  }
  const float hr = 0.5f * rho;
#pragma unroll
  for (int c = 0; c < 64; ++c)
    if ((c & ~15) < o.cp) {
      // SiLU(m) = m sigmoid(m) = h (1 + tanh(h)), h = m / 2, m = p rho gain  (padded channels: gain 0 -> exactly 0)
      const float h = p[c] * (hr * s_gain[c]);
      p[c] = fmaf(h, tanh_fast(h), h);
    }
  if (o.y) {
    uint4* dst = reinterpret_cast<uint4*>(o.y + (size_t)g * o.cnext);
#pragma unroll
    for (int c = 0; c < 64; c += 8)
      if (c < o.cnext)
        dst[c >> 3] = make_uint4(pack2(p[c], p[c + 1]), pack2(p[c + 2], p[c + 3]), pack2(p[c + 4], p[c + 5]), pack2(p[c + 6], p[c + 7]));
  }
  if (o.embed) {
    const int n = g / o.HpWp, rem = g - n * o.HpWp;
    float* dst = o.embed + (size_t)n * o.cout * o.HpWp + rem;
#pragma unroll
    for (int c = 0; c < 64; ++c)
      if (c < o.cout) dst[(size_t)c * o.HpWp] = p[c];
  }
}
__device__ __forceinline__ void pool_norm_store(const StageOut& o, uint32_t tm, int g, const float* s_bias, const float* s_gain,
                                                uint32_t bar_free, int lane) {
  if (o.pool) pool_norm_store_t<true>(o, tm, g, s_bias, s_gain, bar_free, lane);
  else pool_norm_store_t<false>(o, tm, g, s_bias, s_gain, bar_free, lane);
}

// ------------------------------------------------------------------------------------------------ stages 2..L
// Accumulator of window member (dy, dx) sits at TMEM column slot(dy, dx) * cp with slot = 2 dy + (1 - dx): for a view
// (sy, sx) the taps of dx = 1, 0 are kx = sx - 1, sx -- ascending, so the weight tiles of one tap row are adjacent rows
// of a chunk-major matrix and ONE MMA of N = 2 cp (or 4 cp when both tap rows are staged back to back) feeds several
// accumulators from a single read of the view.
struct ConvParams {
  const bf16* x;     // [N][Hin][Win][CIN] bf16 (CIN = template parameter, zero padded channels)
  const bf16* wpk;   // [25 taps][CIN/8][cp][8]: w[co][ci][ky][kx] (reference layout networks.py:203) repacked, zero padded
  StageOut out;
  int Hin, Win, Hp, Wp, tiles;
  int stages;        // ring depth (host: what fits beside the resident weights)
  long long* dbg;    // diagnostic (SD_TRACE_CNN=2): cycle counters of CTA 0; null in production
  int dbg_skip;      // diagnostic: bit dy set = that MMA issuer issues nothing (timing experiments only; results are wrong)
};

constexpr int kConvSmemBudget = 220 * 1024;
// RES: all 25 weight tiles stay in shared memory for the life of the (persistent) CTA, laid out [CIN/8][25 cp rows][16 B];
// a ring stage is just the view.  Otherwise a stage also carries the (up to four) weight tiles of its view as one
// [CIN/8][4 cp rows][16 B] matrix in slot order.
template <int CIN, bool RES>
struct ConvSmem {
  static constexpr int KC = CIN / 8;
  static constexpr int kA = KC * BM * 16;
  static constexpr int kBslot = KC * 4 * 64 * 16;                 // streamed: four tiles at cp = 64
  static constexpr int VPS = RES ? 3 : 1;                         // views per ring stage (one barrier round trip each)
  static constexpr int kStage = RES ? VPS * kA : kA + kBslot;
  static constexpr int kLA = RES ? 4 : 2;                         // cp.async groups in flight per producer thread
  static constexpr int kMaxStages = 16;
  static constexpr int kFixed = 512 /*bias, gain*/ + 8 * (3 * kMaxStages + 4) + 16 + 256;
  static int resident_bytes(int cp) { return RES ? 25 * KC * cp * 16 : 0; }
  static int stages(int cp) {
    int n = (kConvSmemBudget - kFixed - resident_bytes(cp)) / kStage;
    return n > kMaxStages ? kMaxStages : n;
  }
  static int total(int cp) { return kFixed + resident_bytes(cp) + stages(cp) * kStage; }
};

// view order: an interior view (all four window members take it) first, so its k-step 0 initialises every accumulator
__host__ __device__ constexpr int view_of(int v) { return v == 0 ? 14 : (v == 14 ? 0 : v); }

template <int CIN, bool RES>
__global__ void __launch_bounds__(THREADS, 1) conv_pool_kernel(const __grid_constant__ ConvParams P) {
  using L = ConvSmem<CIN, RES>;
  constexpr int KC = L::KC;
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const StageOut& o = P.out;
  const int cp = o.cp;
  const int HpWp = o.HpWp;
  const int resB = RES ? 25 * KC * cp * 16 : 0;
  // layout: [resident weights][ring][bias, gain][barriers]
  const uint32_t ring = base + (uint32_t)resB;
  const int off_const = resB + STAGES * L::kStage;
  float* s_bias = reinterpret_cast<float*>(gbase + off_const);
  float* s_gain = s_bias + 64;
  const uint32_t bar_full = base + (uint32_t)off_const + 512u, bar_empty = bar_full + 8 * L::kMaxStages,
                 bar_ready = bar_empty + 8 * L::kMaxStages, bar_accf = bar_ready + 8 * L::kMaxStages, bar_free = bar_accf + 16,
                 tmem_slot = bar_free + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + off_const + 512 + 8 * (3 * L::kMaxStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) {
    s_bias[threadIdx.x] = o.bias[threadIdx.x];
    s_gain[threadIdx.x] = o.gain[threadIdx.x];
  }
  if (RES) {
    // global [tap][c][row][8] -> shared [c][tap * cp + row][16 B]
    const int n16 = 25 * KC * cp;
    for (int i = threadIdx.x; i < n16; i += THREADS) {
      const int row = i % cp, c = (i / cp) % KC, tap = i / (cp * KC);
      reinterpret_cast<uint4*>(gbase)[(size_t)c * 25 * cp + tap * cp + row] = __ldg(reinterpret_cast<const uint4*>(P.wpk) + i);
    }
    fence_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 2);          // one tcgen05.commit per MMA issuer
      mbar_init(bar_ready + 8 * s, 1);          // the fence helper
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accf + 8 * s, 2);
      mbar_init(bar_free + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (warp < 4) {
    // ---------------------------------------------------------------- software im2col producer: thread = tile row
    const int tid = threadIdx.x;
    int it = 0, s = 0;
    uint32_t eph = 1;                              // parity a wait on empty[s] uses (first lap passes immediately)
    long long t_empty = 0, t_grp = 0, t0 = clock64();
    const bool dbg = P.dbg && blockIdx.x == 0 && tid == 0;
    const int rowpitch = P.Win * CIN;                              // elements between input rows
    const int wrow = tid & 63, wc0 = tid >> 6;                     // streamed weights: this thread's row / first K chunk
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const int g = tile * BM + tid;
      const bool ok = g < o.total;
      const int n = g / HpWp, rem = g - n * HpWp, py = rem / P.Wp, px = rem - py * P.Wp;
      // everything view-specific below is a compile-time constant: the address is base + ((sy-2) Win + (sx-2)) CIN and
      // the SAME-padding test is one bit of two 6-bit masks
      const bf16* src0 = P.x + ((size_t)(n * P.Hin + 2 * py) * P.Win + 2 * px) * CIN;
      uint32_t rmask = 0, cmask = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        rmask |= (uint32_t)(ok && 2 * py + k - 2 >= 0 && 2 * py + k - 2 < P.Hin) << k;
        cmask |= (uint32_t)(2 * px + k - 2 >= 0 && 2 * px + k - 2 < P.Win) << k;
      }
#pragma unroll
      for (int vg = 0; vg < 36 / L::VPS; ++vg, ++it) {
        long long c0 = dbg ? clock64() : 0;
        mbar_wait(bar_empty + 8 * s, eph);
        if (dbg) t_empty += clock64() - c0;
        const uint32_t stg = ring + (uint32_t)(s * L::kStage) + (uint32_t)tid * 16u;
#pragma unroll
        for (int u = 0; u < L::VPS; ++u) {
          constexpr int dummy = 0; (void)dummy;
          const int vv = view_of(vg * L::VPS + u), sy = vv / 6, sx = vv - sy * 6;
          const bool valid = ((rmask >> sy) & (cmask >> sx) & 1u) != 0u;
          const bf16* src = valid ? src0 + (sy - 2) * rowpitch + (sx - 2) * CIN : P.x;
          const uint32_t nbytes = valid ? 16u : 0u;
#pragma unroll
          for (int c = 0; c < KC; ++c) cp_async16(stg + (uint32_t)(u * L::kA + c * BM * 16), src + c * 8, nbytes);
          if (!RES) {
            // slot = 2 dy + (1 - dx) holds tap (sy - dy, sx - dx): rows slot*cp .. of a [KC][4 cp][16 B] matrix
            if (wrow < cp) {
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {
                const int ky = sy - (sl >> 1), kx = sx - (1 - (sl & 1));
                if (ky >= 0 && ky < KSZ && kx >= 0 && kx < KSZ) {
                  const uint4* w = reinterpret_cast<const uint4*>(P.wpk) + (size_t)((ky * KSZ + kx) * KC + wc0) * cp + wrow;
                  const uint32_t dst = ring + (uint32_t)(s * L::kStage) + (uint32_t)(L::kA + (sl * cp + wrow) * 16 + wc0 * 4 * cp * 16);
#pragma unroll
                  for (int c = 0; c < KC / 2; ++c) cp_async16(dst + (uint32_t)(c * 2 * 4 * cp * 16), w + (size_t)c * 2 * cp, 16u);
                }
              }
            }
          }
        }
        c0 = dbg ? clock64() : 0;
        cp_async_arrive(bar_full + 8 * s);
        if (dbg) t_grp += clock64() - c0;
        if (++s == STAGES) { s = 0; eph ^= 1u; }
      }
    }
    if (dbg) { P.dbg[0] = clock64() - t0; P.dbg[1] = t_empty; P.dbg[2] = t_grp; P.dbg[3] = it; }
  } else if (warp == MMA_WARP || warp == MMA_WARP2) {
    // two single-thread MMA issuers, one per window row dy (disjoint accumulators): the issue loop of ONE thread costs
    // ~10 cycles per instruction and was the limiter (profiles/r02_umma_probe.txt: the tensor pipe takes 60 cycles per
    // N <= 128 instruction)
    if (lane == 0) {
      const int dy = warp == MMA_WARP ? 0 : 1;
      const uint64_t dA0 = make_desc_nosw(0, BM * 16, 128);
      const uint32_t ldB = (uint32_t)((RES ? 25 : 4) * cp * 16);          // bytes between K chunks of the weight matrix
      const uint64_t dB0 = make_desc_nosw(0, ldB, 128);
      const uint32_t ahi = (uint32_t)(dA0 >> 32), bhi = (uint32_t)(dB0 >> 32), alo0 = (uint32_t)dA0, blo0 = (uint32_t)dB0;
      const uint32_t kstepB = (2 * ldB) >> 4, cp16 = (uint32_t)cp;            // one row = 16 B = 1 descriptor address unit
      const uint32_t id1 = tc::make_idesc(BM, cp), id2 = tc::make_idesc(BM, 2 * cp);
      const uint32_t blo_res = (base & 0x3FFFFu) >> 4;
      int it = 0, lt = 0, s = 0;
      uint32_t fph = 0;
      long long t_full = 0, t_free = 0, t_fence = 0, t_issue = 0, t_commit = 0, t0 = clock64();
      const bool dbg = P.dbg && blockIdx.x == 0 && dy == 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        const int set = lt & 1;
        long long c0 = dbg ? clock64() : 0;
        if (lt >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((lt >> 1) - 1) & 1u);
        if (dbg) t_free += clock64() - c0;
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(set * 256);
#pragma unroll
        for (int vg = 0; vg < 36 / L::VPS; ++vg, ++it) {
          c0 = dbg ? clock64() : 0;
          mbar_wait(bar_ready + 8 * s, fph);       // data landed AND the helper warp's proxy fence is done
          if (dbg) { t_full += clock64() - c0; c0 = clock64(); }
          tc_fence_after();
          if (dbg) { t_fence += clock64() - c0; c0 = clock64(); }
          const uint32_t stlo = ((ring + (uint32_t)(s * L::kStage)) & 0x3FFFFu) >> 4;
          const uint32_t alo = alo0 + stlo, blo = blo0 + (RES ? blo_res : stlo + (uint32_t)(L::kA >> 4));
#pragma unroll
          for (int u = 0; u < L::VPS; ++u) {
            // view (sy, sx), window row dy: taps (sy - dy, sx - 1 + lo ...), slots [lo, hi) of that row -- all compile time
            const int vv = view_of(vg * L::VPS + u), sy = vv / 6, sx = vv - sy * 6;
            const int lo = sx >= 1 ? 0 : 1, hi = sx <= 4 ? 2 : 1;
            const bool takes = dy == 0 ? sy <= 4 : sy >= 1;
            if (takes && !((P.dbg_skip >> dy) & 1)) {
              const int slot = 2 * dy + lo;
              // RES: row (tap * cp) of the [KC][25 cp] matrix; streamed: row (slot * cp) of the stage's [KC][4 cp] matrix
              const uint32_t boff = RES ? (uint32_t)(((sy - dy) * KSZ + (sx - 1 + lo))) * cp16 : (uint32_t)slot * cp16;
              const uint32_t idesc = (hi - lo) == 2 ? id2 : id1;
#pragma unroll
              for (int kk = 0; kk < CIN / 16; ++kk)
                mma_lh(acc + (uint32_t)slot * (uint32_t)cp, alo + (uint32_t)((u * L::kA + kk * 2 * BM * 16) >> 4), ahi, blo + boff + kk * kstepB, bhi,
                       idesc, (vg | u | kk) == 0 ? 0u : 1u);
            }
          }
          if (dbg) { t_issue += clock64() - c0; c0 = clock64(); }
          tc_commit(bar_empty + 8 * s);
          if (dbg) t_commit += clock64() - c0;
          if (++s == STAGES) { s = 0; fph ^= 1u; }
        }
        tc_commit(bar_accf + 8 * set);
      }
      if (dbg) { P.dbg[4] = clock64() - t0; P.dbg[5] = t_full; P.dbg[6] = t_free; P.dbg[7] = lt; P.dbg[10] = t_fence; P.dbg[11] = t_issue; P.dbg[12] = t_commit; }
    }
  } else if (warp == FENCE_WARP) {
    // The views were written through the generic proxy (cp.async) and the MMA reads them through the async proxy: somebody has
    // to run fence.proxy.async between the two.  It costs ~500 cycles; in a producer it drains that thread's copies in flight,
    // in the issuing thread it sat on the critical path of every stage (30 % of the issue loop).  This warp does nothing else:
    // wait for the stage's data, fence, release the stage to both issuers.
    if (lane == 0) {
      int s = 0;
      uint32_t fph = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int vg = 0; vg < 36 / L::VPS; ++vg) {
          mbar_wait(bar_full + 8 * s, fph);
          fence_async_smem();
          mbar_arrive(bar_ready + 8 * s);
          if (++s == STAGES) { s = 0; fph ^= 1u; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = pooled pixel = TMEM lane
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int set = warp >= 9 ? 1 : 0;             // epilogue group = accumulator set: tiles lt = set, set + 2, ...
    for (int lt = set, tile = blockIdx.x + set * gridDim.x; tile < P.tiles; tile += 2 * gridDim.x, lt += 2) {
      const bool dbg = P.dbg && blockIdx.x == 0 && threadIdx.x == 5 * 32;
      long long c0 = dbg ? clock64() : 0;
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      if (dbg) { P.dbg[8] += clock64() - c0; c0 = clock64(); }
      tc_fence_after();
      pool_norm_store(o, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 256), tile * BM + row, s_bias, s_gain,
                      bar_free + 8 * set, lane);
      if (dbg) P.dbg[9] += clock64() - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------ stage 1 (3 input channels)
// K = 5 ky x 16 (= 5 kx x 3 channels + one zero): the whole 5 x 5 x 3 patch of a conv output is ONE 80-wide K row, built
// by the producer threads from an fp32 patch of the frame staged in shared memory (obs - 0.5 applied on the way,
// networks.py:221; zero padding outside the frame).  A tile is 128 / Wp full pooled rows of one frame.
struct Conv1Params {
  const float* obs;  // [N][Hin][Win][3] fp32 in [0, 1]
  const bf16* wpk;   // [10][cp][8]
  StageOut out;
  int Hin, Win, Hp, Wp, tiles;
  long long* dbg;
};
constexpr int K1C = 10;                           // 16-byte chunks of the K = 80 row
constexpr int kA1 = K1C * BM * 16;                // one accumulator's operand: 20 KB
constexpr int kStage1 = 4 * kA1;                  // 80 KB
constexpr int kPatchFloats = 4096;
struct Conv1Smem {
  static constexpr int STAGES = 2;
  static constexpr int kB = STAGES * kStage1;           // weights: 10 x 64 x 16 B
  static constexpr int kPatch = kB + K1C * 64 * 16;
  static constexpr int kConst = kPatch + kPatchFloats * 4;
  static constexpr int kBar = kConst + 512;
  static constexpr int kTotal = kBar + 8 * (2 * STAGES + 4) + 16 + 128;
};

__global__ void __launch_bounds__(THREADS1, 1) conv1_pool_kernel(const __grid_constant__ Conv1Params P) {
  using L = Conv1Smem;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(gbase + L::kConst);
  float* s_gain = s_bias + 64;
  float* patch = reinterpret_cast<float*>(gbase + L::kPatch);
  const uint32_t bar_full = base + L::kBar, bar_empty = bar_full + 8 * STAGES, bar_accf = bar_empty + 8 * STAGES,
                 bar_free = bar_accf + 16, tmem_slot = bar_free + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + L::kBar + 8 * (2 * STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const StageOut& o = P.out;
  const int cp = o.cp;
  if (threadIdx.x < 64) {
    s_bias[threadIdx.x] = o.bias[threadIdx.x];
    s_gain[threadIdx.x] = o.gain[threadIdx.x];
  }
  for (int i = threadIdx.x; i < K1C * cp; i += THREADS1)
    reinterpret_cast<uint4*>(gbase + L::kB)[i] = reinterpret_cast<const uint4*>(P.wpk)[i];
  fence_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accf + 8 * s, 1);
      mbar_init(bar_free + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const int R = BM / P.Wp;                         // pooled rows per tile
  // patch row: [8 zero floats][Win * 3 floats of the frame row][8 zero floats] (16-byte aligned rows for float4 traffic)
  const int rowf = P.Win * 3, PR = 2 * R + 4, PW = rowf + 16;
  const int tiles_per_frame = o.HpWp / BM;

  if (warp < 4) {
    const int tid = threadIdx.x;
    const int lrow = tid / P.Wp, px = tid - lrow * P.Wp;
    const int r4 = rowf >> 2, n4 = PR * r4;                   // float4 items of a patch
    for (int i = tid; i < PR * 4; i += PROD) {               // the zero margins are written once
      const int r = i >> 2, q = i & 3;
      reinterpret_cast<float4*>(patch + r * PW + (q < 2 ? q * 4 : rowf + q * 4))[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int kPre = 8;                                   // float4 items per thread (PR * Win * 3 / 4 <= 128 * 8)
    float4 pre[kPre];
    int rc[kPre];                                             // (patch row << 16 | float4 column) of this thread's items, -1 = none
#pragma unroll
    for (int k = 0; k < kPre; ++k) {
      const int i = tid + k * PROD;
      rc[k] = i < n4 ? ((i / r4) << 16) | (i % r4) : -1;
    }
    int ftile = 0, fn = 0;                                    // frame / first pooled row of the tile being fetched: no divisions per tile
    auto fetch = [&](int n, int py0) {
      const float* frame = P.obs + (size_t)n * P.Hin * rowf;
#pragma unroll
      for (int k = 0; k < kPre; ++k) {
        pre[k] = make_float4(0.5f, 0.5f, 0.5f, 0.5f);         // out-of-frame rows: 0.5 - 0.5 = the zero padding
        const int iy = 2 * py0 - 2 + (rc[k] >> 16);
        if (rc[k] >= 0 && iy >= 0 && iy < P.Hin) pre[k] = __ldg(reinterpret_cast<const float4*>(frame + (size_t)iy * rowf) + (rc[k] & 0xffff));
      }
    };
    int it = 0;
    long long t_patch = 0, t_empty = 0, t_build = 0, t0 = clock64(), c0 = 0;
    const bool dbg = P.dbg && blockIdx.x == 0 && tid == 0;
    fn = blockIdx.x / tiles_per_frame;
    ftile = blockIdx.x - fn * tiles_per_frame;
    const int dn = gridDim.x / tiles_per_frame, dt = gridDim.x - dn * tiles_per_frame;   // per-step advance of (fn, ftile)
    if (blockIdx.x < P.tiles) fetch(fn, ftile * R);
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      if (dbg) c0 = clock64();
      named_bar(1, PROD);                          // everybody finished reading the previous patch
#pragma unroll
      for (int k = 0; k < kPre; ++k)
        if (rc[k] >= 0)
          reinterpret_cast<float4*>(patch + (rc[k] >> 16) * PW + 8)[rc[k] & 0xffff] =
              make_float4(pre[k].x - 0.5f, pre[k].y - 0.5f, pre[k].z - 0.5f, pre[k].w - 0.5f);
      named_bar(1, PROD);
      fn += dn; ftile += dt;
      if (ftile >= tiles_per_frame) { ftile -= tiles_per_frame; ++fn; }
      if (tile + (int)gridDim.x < P.tiles) fetch(fn, ftile * R);   // in flight while this tile is built
      const int s = it % STAGES;
      if (dbg) { t_patch += clock64() - c0; c0 = clock64(); }
      if (it >= STAGES) mbar_wait(bar_empty + 8 * s, (uint32_t)((it / STAGES) - 1) & 1u);
      if (dbg) { t_empty += clock64() - c0; c0 = clock64(); }
      uint8_t* st = gbase + (size_t)s * kStage1;
      // input pixel (2 px - 2 + j) of patch row (2 lrow + i) starts at float 8 + (2 px - 2 + j) * 3
      const float* prow = patch + (2 * lrow) * PW + 2 + 6 * px;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float v[18];
#pragma unroll
        for (int e = 0; e < 18; e += 2) {
          const float2 t = *reinterpret_cast<const float2*>(prow + i * PW + e);
          v[e] = t.x; v[e + 1] = t.y;
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ky = i - dy;
          if (ky < 0 || ky >= KSZ) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float* u = v + dx * 3;
            uint4* dst = reinterpret_cast<uint4*>(st + (size_t)(dy * 2 + dx) * kA1 + (size_t)(ky * 2) * BM * 16 + tid * 16);
            dst[0] = make_uint4(pack2(u[0], u[1]), pack2(u[2], u[3]), pack2(u[4], u[5]), pack2(u[6], u[7]));
            dst[BM] = make_uint4(pack2(u[8], u[9]), pack2(u[10], u[11]), pack2(u[12], u[13]), pack2(u[14], 0.f));
          }
        }
      }
      fence_async_smem();
      mbar_arrive(bar_full + 8 * s);
      if (dbg) t_build += clock64() - c0;
    }
    if (dbg) { P.dbg[0] = clock64() - t0; P.dbg[1] = t_patch; P.dbg[2] = t_empty; P.dbg[3] = t_build; }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, cp);
      const uint32_t wb = base + L::kB;
      const uint64_t dA0 = make_desc_nosw(0, BM * 16, 128), dB0 = make_desc_nosw(wb, (uint32_t)cp * 16u, 128);
      const uint32_t ahi = (uint32_t)(dA0 >> 32), bhi = (uint32_t)(dB0 >> 32), alo0 = (uint32_t)dA0, blo0 = (uint32_t)dB0;
      const uint32_t kstepB = (uint32_t)(2 * cp * 16) >> 4;
      int it = 0;
      long long t_full = 0, t_free = 0, t0 = clock64(), c0 = 0;
      const bool dbg = P.dbg && blockIdx.x == 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
        const int set = it & 1, s = it % STAGES;
        if (dbg) c0 = clock64();
        if (it >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((it >> 1) - 1) & 1u);
        if (dbg) { t_free += clock64() - c0; c0 = clock64(); }
        mbar_wait(bar_full + 8 * s, (uint32_t)(it / STAGES) & 1u);
        if (dbg) t_full += clock64() - c0;
        tc_fence_after();
        const uint32_t st = base + (uint32_t)s * kStage1;
        const uint32_t alo = alo0 + ((st & 0x3FFFFu) >> 4);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int kk = 0; kk < K1C / 2; ++kk)
            mma_lh(tmem + (uint32_t)(set * 256 + ((q & 2) | (1 - (q & 1))) * cp), alo + (uint32_t)((q * kA1 + kk * 2 * BM * 16) >> 4), ahi,
                   blo0 + (uint32_t)kk * kstepB, bhi, idesc, (uint32_t)(kk != 0));
        tc_commit(bar_empty + 8 * s);
        tc_commit(bar_accf + 8 * set);
      }
      if (dbg) { P.dbg[4] = clock64() - t0; P.dbg[5] = t_full; P.dbg[6] = t_free; P.dbg[7] = it; }
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int set = warp >= 9 ? 1 : 0;             // epilogue group = accumulator set: tiles lt = set, set + 2, ...
    for (int lt = set, tile = blockIdx.x + set * gridDim.x; tile < P.tiles; tile += 2 * gridDim.x, lt += 2) {
      const bool dbg = P.dbg && blockIdx.x == 0 && threadIdx.x == 5 * 32;
      long long c0 = dbg ? clock64() : 0;
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      if (dbg) { P.dbg[8] += clock64() - c0; c0 = clock64(); }
      tc_fence_after();
      pool_norm_store(o, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 256), tile * BM + row, s_bias, s_gain,
                      bar_free + 8 * set, lane);
      if (dbg) P.dbg[9] += clock64() - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// w (cout, cin, 5, 5) fp32 -> [25][cinp/8][cp][8] bf16 (zero padded); bias / gain -> [64] zero padded
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ g, int cout,
                                 int cin, int cp, int cinp, bf16* __restrict__ wpk, float* __restrict__ bias, float* __restrict__ gain) {
  const int total = KSZ * KSZ * cinp * cp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, co = (i >> 3) % cp, c8 = (i >> 3) / cp % (cinp / 8), tap = i / (cinp * cp);
    const int ci = c8 * 8 + e;
    float v = 0.f;
    if (co < cout && ci < cin) v = w[((size_t)co * cin + ci) * (KSZ * KSZ) + tap];
    wpk[i] = __float2bfloat16(v);
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    bias[threadIdx.x] = threadIdx.x < cout ? b[threadIdx.x] : 0.f;
    gain[threadIdx.x] = threadIdx.x < cout ? g[threadIdx.x] : 0.f;
  }
}
// stage 1: w (cout, 3, 5, 5) -> [10][cp][8], K index = ky * 16 + kx * 3 + c (15 of every 16 used)
__global__ void pack_conv1_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ g, int cout,
                                  int cp, bf16* __restrict__ wpk, float* __restrict__ bias, float* __restrict__ gain) {
  const int total = K1C * cp * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, co = (i >> 3) % cp, c8 = (i >> 3) / cp;
    const int k = c8 * 8 + e, ky = k >> 4, r = k & 15;
    float v = 0.f;
    if (co < cout && r < 15) v = w[((size_t)co * 3 + (r % 3)) * (KSZ * KSZ) + ky * KSZ + r / 3];
    wpk[i] = __float2bfloat16(v);
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    bias[threadIdx.x] = threadIdx.x < cout ? b[threadIdx.x] : 0.f;
    gain[threadIdx.x] = threadIdx.x < cout ? g[threadIdx.x] : 0.f;
  }
}

}  // namespace cnn
}  // namespace sd
