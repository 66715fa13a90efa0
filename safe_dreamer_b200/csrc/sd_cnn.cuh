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
// Warp roles (288 threads): 0-3 producers, 4 = TMEM owner + single-thread MMA issuer, 5-8 epilogue (TMEM lane
// quarter = warp & 3).  CTAs are persistent over tiles; two accumulator sets (2 x 256 TMEM columns) let the epilogue
// of tile i overlap the MMAs of tile i+1.  Every wait is bounded (trap instead of hang).
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
constexpr int THREADS = 288;
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// 16 columns of four accumulators (64 columns apart), one wait
__device__ __forceinline__ void tmem_ld16x4(uint32_t taddr, float* a, float* b, float* c, float* d) {
  uint32_t r[64];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]), "=r"(r[16 * q + 4]),
          "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]), "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]),
          "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]), "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]), "=r"(r[16 * q + 14]),
          "=r"(r[16 * q + 15])
        : "r"(taddr + (uint32_t)(q * 64)));
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
__device__ __forceinline__ float silu(float m) { return m / (1.f + __expf(-m)); }

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

// Epilogue of one tile for the thread that owns TMEM lane `row` of accumulator set at `tm`.
__device__ __forceinline__ void pool_norm_store(const StageOut& o, uint32_t tm, int g, const float* s_bias, const float* s_gain,
                                                uint32_t bar_free, int lane) {
  float p[64];
  uint32_t aw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) aw[i] = 0u;
  float ss = 0.f;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    if (c0 < o.cp) {
      float a[16], b[16], c[16], d[16];
      tmem_ld16x4(tm + (uint32_t)c0, a, b, c, d);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float best = a[i];
        uint32_t ar = 0u;
        if (b[i] > best) { best = b[i]; ar = 1u; }
        if (c[i] > best) { best = c[i]; ar = 2u; }
        if (d[i] > best) { best = d[i]; ar = 3u; }
        const float v = best + s_bias[c0 + i];
        p[c0 + i] = v;
        ss = fmaf(v, v, ss);
        aw[(c0 + i) >> 2] |= ar << (8 * (i & 3));
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
  if (o.pool) {
    float4* dst = reinterpret_cast<float4*>(o.pool + (size_t)g * o.cp);
#pragma unroll
    for (int c = 0; c < 64; c += 4)
      if (c < o.cp) dst[c >> 2] = make_float4(p[c], p[c + 1], p[c + 2], p[c + 3]);
  }
  if (o.arg) {
    uint4* dst = reinterpret_cast<uint4*>(o.arg + (size_t)g * o.cp);
#pragma unroll
    for (int c = 0; c < 64; c += 16)
      if (c < o.cp) dst[c >> 4] = make_uint4(aw[c >> 2], aw[(c >> 2) + 1], aw[(c >> 2) + 2], aw[(c >> 2) + 3]);
  }
#pragma unroll
  for (int c = 0; c < 64; ++c) p[c] = silu(p[c] * rho * s_gain[c]);   // padded channels: gain 0 -> exactly 0
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

// ------------------------------------------------------------------------------------------------ stages 2..L
struct ConvParams {
  const bf16* x;     // [N][Hin][Win][CIN] bf16 (CIN = template parameter, zero padded channels)
  const bf16* wpk;   // [25 taps][CIN/8][cp][8]: w[co][ci][ky][kx] (reference layout networks.py:203) repacked, zero padded
  StageOut out;
  int Hin, Win, Hp, Wp, tiles;
};

template <int CIN>
struct ConvSmem {
  static constexpr int KC = CIN / 8;
  static constexpr int kA = KC * BM * 16;
  static constexpr int kB = KC * 64 * 16;         // slot of one weight tile (cp <= 64)
  static constexpr int kStage = kA + 4 * kB;
  static constexpr int STAGES = CIN == 64 ? 4 : (CIN == 32 ? 8 : 12);
  static constexpr int kConst = STAGES * kStage;  // bias[64] + gain[64]
  static constexpr int kBar = kConst + 512;       // full[STAGES] empty[STAGES] acc_full[2] acc_free[2] tmem slot
  static constexpr int kTotal = kBar + 8 * (2 * STAGES + 4) + 16 + 128;
};

template <int CIN>
__global__ void __launch_bounds__(THREADS, 1) conv_pool_kernel(const __grid_constant__ ConvParams P) {
  using L = ConvSmem<CIN>;
  constexpr int STAGES = L::STAGES, KC = L::KC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(gbase + L::kConst);
  float* s_gain = s_bias + 64;
  const uint32_t bar_full = base + L::kBar, bar_empty = bar_full + 8 * STAGES, bar_accf = bar_empty + 8 * STAGES,
                 bar_free = bar_accf + 16, tmem_slot = bar_free + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + L::kBar + 8 * (2 * STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const StageOut& o = P.out;
  if (threadIdx.x < 64) {
    s_bias[threadIdx.x] = o.bias[threadIdx.x];
    s_gain[threadIdx.x] = o.gain[threadIdx.x];
  }
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
  const int cp = o.cp;
  const int HpWp = o.HpWp;

  if (warp < 4) {
    // ---------------------------------------------------------------- software im2col producer: thread = tile row
    const int tid = threadIdx.x;
    const int nB = KC * cp;                       // 16-byte chunks of one weight tile
    int it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const int g = tile * BM + tid;
      const bool ok = g < o.total;
      const int n = g / HpWp, rem = g - n * HpWp, py = rem / P.Wp, px = rem - py * P.Wp;
#pragma unroll 1
      for (int v = 0; v < 36; ++v, ++it) {
        const int sy = v / 6, sx = v - sy * 6;
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(bar_empty + 8 * s, (uint32_t)((it / STAGES) - 1) & 1u);
        const uint32_t st = base + (uint32_t)s * L::kStage;
        const int iy = 2 * py + sy - 2, ix = 2 * px + sx - 2;
        const bool valid = ok && iy >= 0 && iy < P.Hin && ix >= 0 && ix < P.Win;
        const bf16* src = valid ? P.x + ((size_t)(n * P.Hin + iy) * P.Win + ix) * CIN : P.x;
        const uint32_t nbytes = valid ? 16u : 0u;
#pragma unroll
        for (int c = 0; c < KC; ++c) cp_async16(st + (uint32_t)(c * BM * 16 + tid * 16), src + c * 8, nbytes);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ky = sy - (j >> 1), kx = sx - (j & 1);
          if (ky >= 0 && ky < KSZ && kx >= 0 && kx < KSZ) {
            const bf16* w = P.wpk + (size_t)(ky * KSZ + kx) * nB * 8;
            const uint32_t dst = st + (uint32_t)(L::kA + j * L::kB);
            for (int i = tid; i < nB; i += PROD) cp_async16(dst + (uint32_t)i * 16u, w + (size_t)i * 8, 16u);
          }
        }
        cp_async_commit();
        if (it >= LA) {
          cp_async_wait<LA>();
          fence_async_smem();
          mbar_arrive(bar_full + 8 * ((it - LA) % STAGES));
        }
      }
    }
    // drain: publish the last LA stages
    cp_async_wait<0>();
    fence_async_smem();
    for (int k = (it >= LA ? it - LA : 0); k < it; ++k) mbar_arrive(bar_full + 8 * (k % STAGES));
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, cp);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        const int set = lt & 1;
        if (lt >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((lt >> 1) - 1) & 1u);
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(set * 256);
#pragma unroll 1
        for (int v = 0; v < 36; ++v, ++it) {
          const int sy = v / 6, sx = v - sy * 6;
          const int s = it % STAGES;
          mbar_wait(bar_full + 8 * s, (uint32_t)(it / STAGES) & 1u);
          tc_fence_after();
          const uint32_t st = base + (uint32_t)s * L::kStage;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ky = sy - (j >> 1), kx = sx - (j & 1);
            if (ky >= 0 && ky < KSZ && kx >= 0 && kx < KSZ) {
              const uint32_t bt = st + (uint32_t)(L::kA + j * L::kB);
#pragma unroll
              for (int kk = 0; kk < CIN / 16; ++kk)
                tc_mma_f16(acc + (uint32_t)(j * 64), make_desc_nosw(st + (uint32_t)(kk * 2 * BM * 16), BM * 16, 128),
                           make_desc_nosw(bt + (uint32_t)(kk * 2 * cp * 16), (uint32_t)cp * 16u, 128), idesc,
                           (uint32_t)((ky | kx | kk) != 0));
            }
          }
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_accf + 8 * set);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = pooled pixel = TMEM lane
    const int quarter = warp & 3, row = quarter * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
      const int set = lt & 1;
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      pool_norm_store(o, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 256), tile * BM + row, s_bias, s_gain,
                      bar_free + 8 * set, lane);
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

__global__ void __launch_bounds__(THREADS, 1) conv1_pool_kernel(const __grid_constant__ Conv1Params P) {
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
  for (int i = threadIdx.x; i < K1C * cp; i += THREADS)
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
  const int PR = 2 * R + 4, PW = (2 * P.Wp + 4) * 3;   // patch rows / floats per patch row
  const int tiles_per_frame = o.HpWp / BM;

  if (warp < 4) {
    const int tid = threadIdx.x;
    const int lrow = tid / P.Wp, px = tid - lrow * P.Wp;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int n = tile / tiles_per_frame, py0 = (tile - n * tiles_per_frame) * R;
      // stage the frame patch: rows 2 py0 - 2 .. 2 py0 + 2 R + 1, two zero pixels left and right
      named_bar(1, PROD);                          // everybody finished reading the previous patch
      const float* frame = P.obs + (size_t)n * P.Hin * P.Win * 3;
      const int rowf = P.Win * 3;
      for (int i = tid; i < PR * PW; i += PROD) {
        const int r = i / PW, c = i - r * PW;
        const int iy = 2 * py0 - 2 + r, cf = c - 6;
        float v = 0.f;
        if (iy >= 0 && iy < P.Hin && cf >= 0 && cf < rowf) v = __ldg(frame + (size_t)iy * rowf + cf) - 0.5f;
        patch[i] = v;
      }
      named_bar(1, PROD);
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(bar_empty + 8 * s, (uint32_t)((it / STAGES) - 1) & 1u);
      uint8_t* st = gbase + (size_t)s * kStage1;
      const float* prow = patch + (2 * lrow) * PW + (2 * px) * 3;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float v[18];
#pragma unroll
        for (int e = 0; e < 18; ++e) v[e] = prow[i * PW + e];
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
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, cp);
      const uint32_t wb = base + L::kB;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
        const int set = it & 1, s = it % STAGES;
        if (it >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((it >> 1) - 1) & 1u);
        mbar_wait(bar_full + 8 * s, (uint32_t)(it / STAGES) & 1u);
        tc_fence_after();
        const uint32_t st = base + (uint32_t)s * kStage1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int kk = 0; kk < K1C / 2; ++kk)
            tc_mma_f16(tmem + (uint32_t)(set * 256 + q * 64), make_desc_nosw(st + (uint32_t)(q * kA1 + kk * 2 * BM * 16), BM * 16, 128),
                       make_desc_nosw(wb + (uint32_t)(kk * 2 * cp * 16), (uint32_t)cp * 16u, 128), idesc, (uint32_t)(kk != 0));
        tc_commit(bar_empty + 8 * s);
        tc_commit(bar_accf + 8 * set);
      }
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
      const int set = lt & 1;
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      pool_norm_store(o, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 256), tile * BM + row, s_bias, s_gain,
                      bar_free + 8 * set, lane);
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
