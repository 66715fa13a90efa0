// sd_wgrad_tc.cuh -- weight gradients of the posterior scan / batched prior on tcgen05:
//
//     dW[n][k] += sum_r dY[r][n] * X[r][k]        r = the T*B taped rows (1024 at base sizes)
//
// The contraction index is the ROW, the slow index of both row-major operands, so both are MN-major UMMA operands: the
// shared-memory image is [8-element chunk][row][16 B] (no swizzle; 8 rows x 16 B = one core matrix, LBO = 128 B between the
// two 8-row K blocks of a K = 16 step, SBO = the chunk plane), exactly what a coalesced cp.async of a row produces.
// kind::tf32 does NOT take MN-major operands (profiles/micro/umma_tf32_check.cu: K-major exact, MN-major all zeros), so
// fp32 accuracy comes from a two-term bf16 split instead: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa
// bits together), and
//     dY^T X ~ hi^T hi + hi^T lo + lo^T hi          (the dropped lo^T lo term is 2^-16 relative)
// as three kind::f16 MMAs per K step into the same fp32 TMEM accumulator: ~30x tighter than single-pass TF32, which is what
// the reference's own fp32 mode computes with (train.py:38).  split_bf16_kernel writes the hi / lo planes once per operand.
// One CTA owns a 128 (n) x KT (k <= 256) tile over ALL rows (no row slices: the result is complete, summation order
// fixed) and stores it in the reference layout (nn.Linear (N,K) or BlockLinear (O/G, I/G, G) through the (sn, sk) strides);
// the existing vectorised reduce kernel adds it into the caller's gradient tensor.  The 3xTF32 mma.sync kernel (83 TFLOP/s raw) stays as the fallback for
// shapes that are not 16-byte friendly, few rows, or SD_WGRAD_TC=0.
#pragma once
#include "sd_kernels.cuh"
#include "sd_tc.cuh"

namespace sd {
namespace wgtc {

using bf16 = __nv_bfloat16;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;

constexpr int BMN = 128;          // n tile (UMMA M)
constexpr int KT_MAX = 256;       // k tile (UMMA N)
constexpr int RS = 32;            // rows per ring stage (2 K steps of 16)
constexpr int STAGES = 4;         // 4 x 48 KB (2 x 96 KB measured 5 us per stage: too coarse to overlap load, fence and MMA)
constexpr int PROD = 128;
constexpr int PW = PROD / 32;               // producer warps
constexpr int THREADS = 32 * (PW + 1 + 4);  // warps 0-3 producers, 4 MMA issuer + TMEM owner, 5-8 epilogue (8 producer warps measured slower)
constexpr int kPlane = RS * 16 + 16;        // chunk plane: RS rows x 16 B (+ 16: consecutive chunks in different banks)
constexpr int kA = (BMN / 8) * kPlane;      // one of the hi / lo images of the dY tile
constexpr int kB = (KT_MAX / 8) * kPlane;
constexpr int kStage = 2 * kA + 2 * kB;
constexpr int kSmem = STAGES * kStage + 8 * (2 * STAGES + 1) + 16 + 256;

struct Problem {
  const bf16 *a_hi, *a_lo; int lda;          // dY  [R x N] split
  const bf16 *b1_hi, *b1_lo; int ldb1;       // X   [R x K1]
  const bf16 *b2_hi, *b2_lo; int ldb2;       // X2  [R x (K - K1)]
  int K1, K, N;
  float* dW; long long sn, sk;               // element (n, k) is STORED at dW + n * sn + k * sk (a scratch in the gradient's layout)
};
struct Batch {
  int count, R;
  int rows_per_slice;          // multiple of RS; slice s = rows [s * rows_per_slice, ...) -> its own partial image
  long long slice_stride;      // floats between the partial images
  Problem p[kMaxBatch];
};

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// up to three operands per launch: src fp32 [R x C] (row stride ld) -> hi, lo bf16 [R x C] dense;  C % 4 == 0, ld % 4 == 0
struct SplitJob { const float* src; int ld, C; bf16 *hi, *lo; };
struct SplitBatch { int count, R; SplitJob j[3]; };
__global__ void __launch_bounds__(256) split_bf16_kernel(const SplitBatch sb) {
  pdl_prologue();
  long long end[3], total = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) { total += q < sb.count ? (long long)sb.R * (sb.j[q].C >> 2) : 0; end[q] = total; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = i < end[0] ? 0 : (i < end[1] ? 1 : 2);
    const SplitJob& jb = sb.j[q];
    const long long li = i - (q == 0 ? 0 : end[q - 1]);
    const int c4 = jb.C >> 2;
    const int r = (int)(li / c4), c = (int)(li - (long long)r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(jb.src + (size_t)r * jb.ld + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat162 h2[2], l2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16(x[2 * e]), h1 = __float2bfloat16(x[2 * e + 1]);
      h2[e] = __halves2bfloat162(h0, h1);
      l2[e] = __floats2bfloat162_rn(x[2 * e] - __bfloat162float(h0), x[2 * e + 1] - __bfloat162float(h1));
    }
    *reinterpret_cast<uint2*>(jb.hi + (size_t)r * jb.C + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h2[0]), *reinterpret_cast<uint32_t*>(&h2[1]));
    *reinterpret_cast<uint2*>(jb.lo + (size_t)r * jb.C + c) = make_uint2(*reinterpret_cast<uint32_t*>(&l2[0]), *reinterpret_cast<uint32_t*>(&l2[1]));
  }
}

// BlockLinear finish: partial images are dense [slice][g][n][k] (the tile kernel stores them with 16-byte rows); thread (n, k)
// sums the slices in order for all G <= 8 blocks and adds the G contiguous words of the reference layout (O/G, I/G, G).
__global__ void __launch_bounds__(256) block_finish_kernel(const float* __restrict__ part, int slices, int G, int N, int K,
                                                           float* __restrict__ dW) {
  pdl_prologue();
  const long long nk = (long long)N * K, img = nk * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += (long long)gridDim.x * blockDim.x) {
    float* dst = dW + i * G;
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
      for (int sl = 0; sl < slices; ++sl) s += __ldg(part + (size_t)sl * img + (size_t)g * nk + i);
      dst[g] += s;
    }
  }
}

// grid (ceil(N / 128), ceil(K / kt), count * slices); kt = this launch's k tile (multiple of 16, <= 256)
__global__ void __launch_bounds__(THREADS, 1) wgrad_split_tc_kernel(const __grid_constant__ Batch b, int kt) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + STAGES * kStage, bar_empty = bar_full + 8 * STAGES, bar_done = bar_empty + 8 * STAGES,
                 tmem_slot = bar_done + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + STAGES * kStage + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Problem& p = b.p[blockIdx.z % b.count];
  const int slice = blockIdx.z / b.count;
  const int n0 = blockIdx.x * BMN, k0 = blockIdx.y * kt;
  if (n0 >= p.N || k0 >= p.K) return;           // whole CTA leaves before any barrier / TMEM use
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  pdl_prologue();                                // everything above overlaps the previous kernel's tail
  const int row_lo = slice * b.rows_per_slice, row_hi = min(b.R, row_lo + b.rows_per_slice);
  const int nst = (row_hi - row_lo + RS - 1) / RS;

  if (warp < PW) {
    const int tid = threadIdx.x;
    // dY: 16 chunks per row, lane % 16 = chunk: a half warp reads 256 consecutive bytes of one row
    const int ca = tid & 15, ra = tid >> 4;                       // rows ra, ra + PROD / 16, ...
    const bool a_ok = n0 + 8 * ca < p.N;
    const size_t a_off = (size_t)n0 + 8 * ca;
    // X: kt / 8 chunks per row (<= 32): chunk = tid & 31, rows tid >> 5, + PROD / 32, ...
    const int cb = tid & 31, rb = tid >> 5;
    const int kk = k0 + 8 * cb;
    const bool b_ok = 8 * cb < kt && kk < p.K;
    const bool seg1 = kk < p.K1;
    const bf16* bh = seg1 ? p.b1_hi + kk : p.b2_hi + (kk - p.K1);
    const bf16* bl = seg1 ? p.b1_lo + kk : p.b2_lo + (kk - p.K1);
    const size_t b_ld = seg1 ? (size_t)p.ldb1 : (size_t)p.ldb2;
    uint32_t eph = 1;
    // running source pointers of this thread's first row in the stage; a copy costs an add + the cp.async on the fast path
    // (a lone producer warp pays ~10 cycles per instruction: per-copy index arithmetic made the first version 3x slower)
    constexpr int RA = PROD / 16, RB = PROD / 32;                    // row step of the dY / X copies
    const bf16* pah = p.a_hi + (size_t)(row_lo + ra) * p.lda + a_off;
    const bf16* pal = p.a_lo + (size_t)(row_lo + ra) * p.lda + a_off;
    const bf16* pbh = bh + (size_t)(row_lo + rb) * b_ld;
    const bf16* pbl = bl + (size_t)(row_lo + rb) * b_ld;
    const size_t sa_step = (size_t)RA * p.lda, sb_step = (size_t)RB * b_ld;
    for (int st = 0, s = 0; st < nst; ++st) {
      mbar_wait(bar_empty + 8 * s, eph);
      const int r0 = row_lo + st * RS;
      const uint32_t sa = base + (uint32_t)(s * kStage + ca * kPlane + ra * 16), sb = base + (uint32_t)(s * kStage + 2 * kA + cb * kPlane + rb * 16);
      if (r0 + RS <= row_hi && a_ok && b_ok) {
#pragma unroll
        for (int j = 0; j < RS / RA; ++j) {
          cp_async16(sa + (uint32_t)(j * RA * 16), pah + j * sa_step, 16u);
          cp_async16(sa + (uint32_t)(kA + j * RA * 16), pal + j * sa_step, 16u);
        }
#pragma unroll
        for (int j = 0; j < RS / RB; ++j) {
          cp_async16(sb + (uint32_t)(j * RB * 16), pbh + j * sb_step, 16u);
          cp_async16(sb + (uint32_t)(kB + j * RB * 16), pbl + j * sb_step, 16u);
        }
      } else {
#pragma unroll 2
        for (int j = 0; j < RS / RA; ++j) {
          const bool ok = a_ok && r0 + ra + j * RA < row_hi;
          cp_async16(sa + (uint32_t)(j * RA * 16), ok ? pah + j * sa_step : p.a_hi, ok ? 16u : 0u);
          cp_async16(sa + (uint32_t)(kA + j * RA * 16), ok ? pal + j * sa_step : p.a_hi, ok ? 16u : 0u);
        }
#pragma unroll 2
        for (int j = 0; j < RS / RB; ++j) {
          const bool ok = b_ok && r0 + rb + j * RB < row_hi;
          cp_async16(sb + (uint32_t)(j * RB * 16), ok ? pbh + j * sb_step : p.a_hi, ok ? 16u : 0u);
          cp_async16(sb + (uint32_t)(kB + j * RB * 16), ok ? pbl + j * sb_step : p.a_hi, ok ? 16u : 0u);
        }
      }
      pah += (size_t)RS * p.lda; pal += (size_t)RS * p.lda; pbh += (size_t)RS * b_ld; pbl += (size_t)RS * b_ld;
      cp_async_arrive(bar_full + 8 * s);
      if (++s == STAGES) { s = 0; eph ^= 1u; }
    }
  } else if (warp == PW) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BMN, kt) | (1u << 15) | (1u << 16);      // both operands MN-major
      // K block (8 rows) = 128 B contiguous, two per K = 16 step (LBO = 128 B); MN block (8 elements) stride = the chunk plane
      const uint64_t d0 = make_desc_nosw(0, 128, kPlane);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo0 = (uint32_t)d0;
      uint32_t fph = 0;
      for (int st = 0, s = 0; st < nst; ++st) {
        mbar_wait(bar_full + 8 * s, fph);
        fence_async_smem();
        tc_fence_after();
        const uint32_t ah = dlo0 + (((base + (uint32_t)(s * kStage)) & 0x3FFFFu) >> 4), al = ah + (kA >> 4);
        const uint32_t bh = ah + (2 * kA >> 4), bl = bh + (kB >> 4);
#pragma unroll
        for (int q = 0; q < RS / 16; ++q) {           // 16 rows = 256 B further down every chunk plane
          const uint32_t o = (uint32_t)(q * 16);
          mma_bf16(tmem, ah + o, dhi, bh + o, dhi, idesc, (st | q) == 0 ? 0u : 1u);
          mma_bf16(tmem, ah + o, dhi, bl + o, dhi, idesc, 1u);
          mma_bf16(tmem, al + o, dhi, bh + o, dhi, idesc, 1u);
        }
        tc_commit(bar_empty + 8 * s);
        if (++s == STAGES) { s = 0; fph ^= 1u; }
      }
      tc_commit(bar_done);
    }
  } else {
    // epilogue: TMEM lane = n, columns = k.  Plain stores of the finished tile at out[n * sn + k * sk] (`out` = the
    // gradient tensor's layout in a scratch buffer); wgrad_reduce4_kernel then adds the scratch into the caller's tensor
    // with 16-byte accesses.  (Read-modify-write from here was 5x slower for the BlockLinear layout, where the eight
    // blocks' CTAs interleave 4-byte words of the same sectors.)
    const int quarter = warp & 3, n = n0 + quarter * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* dst = p.dW + (size_t)slice * b.slice_stride + (size_t)n * p.sn;
    for (int c0 = 0; c0 < kt; c0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (n < p.N) {
        if (p.sk == 1 && k0 + c0 + 15 < p.K) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(dst + k0 + c0 + 4 * q) =
                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = k0 + c0 + i;
            if (k < p.K) dst[(size_t)k * p.sk] = __uint_as_float(r[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

// ------------------------------------------------------------------------------------------------ all layers in three launches
// The posterior backward has seven weight gradients of 16-64 tiles each: launched one by one, every kernel runs a partial
// wave and pays its own prologue / epilogue latency (3 launches x 7 layers = 0.35 ms on the critical path).  The batched
// form flattens (problem, n tile, k tile, row slice) over ONE grid for the tile kernel, one for the hi / lo split of every
// operand and one for the finish.
constexpr int kMaxProb = 24;      // in0, in1, hid x8, gates x8, obs layers, obs logit (+ spare)
constexpr int kMaxJobs = 24;      // split jobs (three operands per layer)
constexpr int kMaxLayers = 8;
struct MultiBatch {
  int count, R, rows_per_slice, nslices;
  int tile_end[kMaxProb];         // running end of each problem's (n tile, k tile, slice) range
  int kt[kMaxProb], nkt[kMaxProb];
  long long slice_stride[kMaxProb];
  Problem p[kMaxProb];
};
struct SplitMulti { int count, R; long long end[kMaxJobs]; SplitJob j[kMaxJobs]; };
struct FinishLayer { const float* part; float* dW; int G, N, K, slices; long long slice_stride; };
struct FinishMulti { int count; long long end[kMaxLayers]; FinishLayer l[kMaxLayers]; };

__global__ void __launch_bounds__(256) split_multi_kernel(const __grid_constant__ SplitMulti sb) {
  pdl_prologue();
  const long long total = sb.end[sb.count - 1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = 0;
    while (i >= sb.end[q]) ++q;
    const SplitJob& jb = sb.j[q];
    const long long li = i - (q == 0 ? 0 : sb.end[q - 1]);
    const int c4 = jb.C >> 2;
    const int r = (int)(li / c4), c = (int)(li - (long long)r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(jb.src + (size_t)r * jb.ld + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat162 h2[2], l2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16(x[2 * e]), h1 = __float2bfloat16(x[2 * e + 1]);
      h2[e] = __halves2bfloat162(h0, h1);
      l2[e] = __floats2bfloat162_rn(x[2 * e] - __bfloat162float(h0), x[2 * e + 1] - __bfloat162float(h1));
    }
    *reinterpret_cast<uint2*>(jb.hi + (size_t)r * jb.C + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h2[0]), *reinterpret_cast<uint32_t*>(&h2[1]));
    *reinterpret_cast<uint2*>(jb.lo + (size_t)r * jb.C + c) = make_uint2(*reinterpret_cast<uint32_t*>(&l2[0]), *reinterpret_cast<uint32_t*>(&l2[1]));
  }
}

// thread (layer, n, k): sums the row slices in order for all G blocks and adds into the reference layout ((N,K) or (O/G,I/G,G))
__global__ void __launch_bounds__(256) finish_multi_kernel(const __grid_constant__ FinishMulti fb) {
  pdl_prologue();
  const long long total = fb.end[fb.count - 1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = 0;
    while (i >= fb.end[q]) ++q;
    const FinishLayer& L = fb.l[q];
    const long long li = i - (q == 0 ? 0 : fb.end[q - 1]);
    const long long nk = (long long)L.N * L.K;
    float* dst = L.dW + li * L.G;
    // all G x slices loads first (independent), then the sums in slice order: one memory latency instead of G x slices
    float v[8][4], d[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int sl = 0; sl < 4; ++sl)
        v[g][sl] = (g < L.G && sl < L.slices) ? __ldg(L.part + (size_t)sl * L.slice_stride + (size_t)g * nk + li) : 0.f;
      d[g] = g < L.G ? dst[g] : 0.f;
    }
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (g < L.G) dst[g] = d[g] + (((v[g][0] + v[g][1]) + v[g][2]) + v[g][3]);
  }
}

// grid.x = total tiles; the tile body is the same as wgrad_split_tc_kernel's
__global__ void __launch_bounds__(THREADS, 1) wgrad_multi_tc_kernel(const __grid_constant__ MultiBatch b) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + STAGES * kStage, bar_empty = bar_full + 8 * STAGES, bar_done = bar_empty + 8 * STAGES,
                 tmem_slot = bar_done + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + STAGES * kStage + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int pi = 0;
  while ((int)blockIdx.x >= b.tile_end[pi]) ++pi;
  const Problem& p = b.p[pi];
  const int kt = b.kt[pi];
  int lt = (int)blockIdx.x - (pi == 0 ? 0 : b.tile_end[pi - 1]);
  const int slice = lt % b.nslices; lt /= b.nslices;
  const int k0 = (lt % b.nkt[pi]) * kt, n0 = (lt / b.nkt[pi]) * BMN;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  pdl_prologue();
  const int row_lo = slice * b.rows_per_slice, row_hi = min(b.R, row_lo + b.rows_per_slice);
  const int nst = (row_hi - row_lo + RS - 1) / RS;

  if (warp < PW) {
    const int tid = threadIdx.x;
    const int ca = tid & 15, ra = tid >> 4;
    const bool a_ok = n0 + 8 * ca < p.N;
    const size_t a_off = (size_t)n0 + 8 * ca;
    const int cb = tid & 31, rb = tid >> 5;
    const int kk = k0 + 8 * cb;
    const bool b_ok = 8 * cb < kt && kk < p.K;
    const bool seg1 = kk < p.K1;
    const bf16* bh = seg1 ? p.b1_hi + kk : p.b2_hi + (kk - p.K1);
    const bf16* bl = seg1 ? p.b1_lo + kk : p.b2_lo + (kk - p.K1);
    const size_t b_ld = seg1 ? (size_t)p.ldb1 : (size_t)p.ldb2;
    uint32_t eph = 1;
    constexpr int RA = PROD / 16, RB = PROD / 32;
    const bf16* pah = p.a_hi + (size_t)(row_lo + ra) * p.lda + a_off;
    const bf16* pal = p.a_lo + (size_t)(row_lo + ra) * p.lda + a_off;
    const bf16* pbh = bh + (size_t)(row_lo + rb) * b_ld;
    const bf16* pbl = bl + (size_t)(row_lo + rb) * b_ld;
    const size_t sa_step = (size_t)RA * p.lda, sb_step = (size_t)RB * b_ld;
    for (int st = 0, s = 0; st < nst; ++st) {
      mbar_wait(bar_empty + 8 * s, eph);
      const int r0 = row_lo + st * RS;
      const uint32_t sa = base + (uint32_t)(s * kStage + ca * kPlane + ra * 16), sb = base + (uint32_t)(s * kStage + 2 * kA + cb * kPlane + rb * 16);
      if (r0 + RS <= row_hi && a_ok && b_ok) {
#pragma unroll
        for (int j = 0; j < RS / RA; ++j) {
          cp_async16(sa + (uint32_t)(j * RA * 16), pah + j * sa_step, 16u);
          cp_async16(sa + (uint32_t)(kA + j * RA * 16), pal + j * sa_step, 16u);
        }
#pragma unroll
        for (int j = 0; j < RS / RB; ++j) {
          cp_async16(sb + (uint32_t)(j * RB * 16), pbh + j * sb_step, 16u);
          cp_async16(sb + (uint32_t)(kB + j * RB * 16), pbl + j * sb_step, 16u);
        }
      } else {
#pragma unroll 2
        for (int j = 0; j < RS / RA; ++j) {
          const bool ok = a_ok && r0 + ra + j * RA < row_hi;
          cp_async16(sa + (uint32_t)(j * RA * 16), ok ? pah + j * sa_step : p.a_hi, ok ? 16u : 0u);
          cp_async16(sa + (uint32_t)(kA + j * RA * 16), ok ? pal + j * sa_step : p.a_hi, ok ? 16u : 0u);
        }
#pragma unroll 2
        for (int j = 0; j < RS / RB; ++j) {
          const bool ok = b_ok && r0 + rb + j * RB < row_hi;
          cp_async16(sb + (uint32_t)(j * RB * 16), ok ? pbh + j * sb_step : p.a_hi, ok ? 16u : 0u);
          cp_async16(sb + (uint32_t)(kB + j * RB * 16), ok ? pbl + j * sb_step : p.a_hi, ok ? 16u : 0u);
        }
      }
      pah += (size_t)RS * p.lda; pal += (size_t)RS * p.lda; pbh += (size_t)RS * b_ld; pbl += (size_t)RS * b_ld;
      cp_async_arrive(bar_full + 8 * s);
      if (++s == STAGES) { s = 0; eph ^= 1u; }
    }
  } else if (warp == PW) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BMN, kt) | (1u << 15) | (1u << 16);
      const uint64_t d0 = make_desc_nosw(0, 128, kPlane);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo0 = (uint32_t)d0;
      uint32_t fph = 0;
      for (int st = 0, s = 0; st < nst; ++st) {
        mbar_wait(bar_full + 8 * s, fph);
        fence_async_smem();
        tc_fence_after();
        const uint32_t ah = dlo0 + (((base + (uint32_t)(s * kStage)) & 0x3FFFFu) >> 4), al = ah + (kA >> 4);
        const uint32_t bh = ah + (2 * kA >> 4), bl = bh + (kB >> 4);
#pragma unroll
        for (int q = 0; q < RS / 16; ++q) {
          const uint32_t o = (uint32_t)(q * 16);
          mma_bf16(tmem, ah + o, dhi, bh + o, dhi, idesc, (st | q) == 0 ? 0u : 1u);
          mma_bf16(tmem, ah + o, dhi, bl + o, dhi, idesc, 1u);
          mma_bf16(tmem, al + o, dhi, bh + o, dhi, idesc, 1u);
        }
        tc_commit(bar_empty + 8 * s);
        if (++s == STAGES) { s = 0; fph ^= 1u; }
      }
      tc_commit(bar_done);
    }
  } else {
    const int quarter = warp & 3, n = n0 + quarter * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* dst = p.dW + (size_t)slice * b.slice_stride[pi] + (size_t)n * p.sn;
    for (int c0 = 0; c0 < kt; c0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (n < p.N) {
        if (k0 + c0 + 15 < p.K) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(dst + k0 + c0 + 4 * q) =
                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (k0 + c0 + i < p.K) dst[k0 + c0 + i] = __uint_as_float(r[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

}  // namespace wgtc
}  // namespace sd
