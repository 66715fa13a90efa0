// sd_tc.cuh -- tcgen05 (5th-gen tensor core) batched GEMM for the large-row dense layers.
//
//   C[R x N] (fp32) = [A1 | A2][R x K] (bf16) * W[N x K]^T (bf16) + bias        (fp32 accumulate in TMEM)
//
// sm_100a only.  One CTA computes one 128 x BN output tile of one problem of the batch:
//   warp 0    : TMA producer  (cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, kind::f16, M=128,N=BN,K=16)
//   warps 2-9 : epilogue (tcgen05.ld 32x32b -> smem transpose -> +bias -> coalesced float4 stores), two warps per
//               TMEM lane quadrant
// Both operands are K-major (activations row-major [R x K], weights in nn.Linear [N x K] layout), so the
// shared-memory tiles are the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect
// (8-row x 128-byte swizzle atoms, SBO = 1024 B); K advances inside an atom by bumping the descriptor
// start address by 32 B per UMMA_K=16.
// The A operand may be the concatenation of two buffers along K (block input [deter_g | x] of the
// block-GRU, feat = [stoch | deter], obs input [deter | embed]) and every problem has its own column
// offsets, which is how the 8 block-diagonal problems of a BlockLinear share two tensor maps.
//
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sd {
namespace tc {

constexpr int BM = 128;      // UMMA M (rows of the output tile == TMEM lanes)
constexpr int BK = 64;       // K elements per stage: 64 bf16 = 128 B = one swizzle atom row
constexpr int THREADS = 320;  // warp 0 TMA, warp 1 MMA/TMEM, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int kMaxProblems = 8;
constexpr int kMaxMaps = 12;

struct Problem {
  int a1_map, a1_col;  // tensor-map index + starting column of K segment 1
  int a2_map, a2_col;  // K segment 2 (unused when K1 == K)
  int w_map, w_row;    // weight map + first weight row (n offset inside the stacked [G*Npad x K] matrix)
  int K1, K;           // both multiples of 64
  int N;               // logical output columns (store guard); tiles cover ceil(N/BN)*BN
  int ldc;
  float* C;
  float* Cpart;        // split-K: slice s >= 1 writes to Cpart + (s-1)*part_stride (same ldc, no bias)
  const float* bias;   // nullable
  // EPI_GATES (BN = 192 = reset|cand|update of 64 units): GRU gate math in the epilogue (rssm.py:63-75)
  const float* e_in; int e_ld_in;         // deter entering the step
  float* e_out; int e_ld_out;             // deter' fp32
  __nv_bfloat16* e_out_bf; int e_ld_bf;   // deter' bf16 (next tcgen05 operand), nullable
  int e_dg;                               // units per block (weight rows of one gate)
  const float* e_gain;                    // EPI_NORM: RMS scale
  int ksplit;                             // per-problem split-K factor (0 = the batch's); CTAs of slices >= it exit
};
// EPI_NORM (BN = 64, N = 256, cluster of 4 CTAs along N): RMSNorm(1e-4)*gain -> SiLU fused in the epilogue.  The
// four CTAs of a cluster own the four 64-column tiles of the same 128 rows; per-row partial sums of squares are
// exchanged through distributed shared memory (st.shared::cluster + barrier.cluster) so every CTA can normalise
// its tile.  e_gain = RMS scale [N]; e_out / e_out_bf = activation fp32 / bf16.
// EPI_NORMW (BN = 256 = N): the CTA owns whole rows, so RMSNorm -> SiLU needs no cluster: two epilogue warps share a
// TMEM lane quadrant (128 columns each) and exchange their partial row sums through shared memory.  Writes bf16 (and
// optionally fp32) activations; used for the K = 2560 first layers of the heads on 16 384 rows.
enum { EPI_STORE = 0, EPI_GATES = 1, EPI_NORM = 2, EPI_NORMW = 3 };
struct alignas(64) Batch {
  CUtensorMap maps[kMaxMaps];
  Problem p[kMaxProblems];
  int count;
  int R;
  int ksplit;             // K is cut into `ksplit` slices over blockIdx.z / count (partials summed by the consumer)
  long long part_stride;  // floats between partial slices
  long long* timing;      // diagnostic (SD_TRACE=2): clock64 stamps of CTA (0,0,0); null in production
};
#define SD_TC_STAMP(i) do { if (batch.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) batch.timing[i] = clock64(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~seconds of polling, then trap (turns a protocol bug into an error, never a hang).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 26); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30) (ignored for swizzled K-major), SBO>>4 [32,46) = 1024 B, version=1 [46,48), layout=2 [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format=F32 [4,6), a/b_format=BF16 [7,10)/[10,13),
// a/b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN, int NSTAGES>
struct SmemLayout {
  static constexpr int STAGES = NSTAGES;  // skinny-N, long-K tiles are latency bound: deeper TMA ring
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = STAGES * kStage;
  static constexpr int kNormOff = kBarOff + 256;           // EPI_NORM: [8][128] partial sums written by cluster peers
  static constexpr int kTotal = kNormOff + 4096 + 1024;    // barriers + tmem slot + norm slots + 1 KB alignment slack  // barriers + tmem slot, + 1 KB alignment slack
};

template <int BN, int NSTAGES, int EPI = EPI_STORE>
__global__ void __launch_bounds__(THREADS, NSTAGES <= 4 ? 2 : 1) gemm_bf16_tc_kernel(const __grid_constant__ Batch batch) {
  static_assert(BN == 64 || BN == 128 || BN == 192 || BN == 256, "unsupported tile width");
  static_assert(EPI != EPI_GATES || BN == 192, "the gate epilogue uses 192-wide tiles (3 gates x 64 units)");
  static_assert(EPI != EPI_NORM || BN == 64, "the fused-norm epilogue uses four 64-wide tiles per cluster");
  static_assert(EPI != EPI_NORMW || BN == 256, "the wide fused-norm epilogue needs the whole 256-wide row in one tile");
  constexpr int TMEM_COLS = BN == 192 ? 256 : BN;   // allocations are powers of two >= 32 columns
  using L = SmemLayout<BN, NSTAGES>;
  constexpr int STAGES = L::STAGES;
  const int prob = blockIdx.z % batch.count, slice = blockIdx.z / batch.count;
  const Problem pr = batch.p[prob];  // by value: keeps the fields in registers instead of re-reading the param bank
  const int n0 = blockIdx.x * (EPI == EPI_GATES ? 64 : BN);   // gates: first of this tile's 64 units
  const int ks = pr.ksplit > 0 ? pr.ksplit : batch.ksplit;
  if (n0 >= pr.N || slice >= ks) return;  // whole CTA exits before any barrier/TMEM use
  const int m0 = blockIdx.y * BM;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + L::kBarOff;           // STAGES x 8 B
  const uint32_t bar_empty = bar_full + STAGES * 8;      // STAGES x 8 B
  const uint32_t bar_acc = bar_empty + STAGES * 8;       // 8 B
  const uint32_t tmem_slot = bar_acc + 8;                // 4 B
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + L::kBarOff + STAGES * 16 + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SD_TC_STAMP(0);
  const int num_kb_all = pr.K / BK;
  const int kb_per = (num_kb_all + ks - 1) / ks;
  const int kb0 = slice * kb_per;
  const int kb1 = min(num_kb_all, kb0 + kb_per);
  const int num_kb = kb1 > kb0 ? kb1 - kb0 : 0;   // an empty slice just stores zeros

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if (threadIdx.x == 0) SD_TC_STAMP(1);
  // PDL: barrier init / TMEM allocation above overlap the previous kernel.  The weight tiles do not depend on
  // it either, so the producer arms the first ring stages and issues their weight TMA loads BEFORE the wait;
  // activations (A tiles), bias-free outputs etc. are only touched after it.
  const int n_pref = num_kb < STAGES ? num_kb : STAGES;
  if (warp == 0 && lane == 0) {
    const CUtensorMap* mw = &batch.maps[pr.w_map];
    for (int it = 0; it < n_pref; ++it) {
      const uint32_t sb = base + it * L::kStage + L::kABytes;
      mbar_expect_tx(bar_full + it * 8, L::kStage);
      if (EPI == EPI_GATES) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_2d(sb + j * (64 * BK * 2), mw, (kb0 + it) * BK, pr.w_row + j * pr.e_dg + n0, bar_full + it * 8);
      } else {
        tma_load_2d(sb, mw, (kb0 + it) * BK, pr.w_row + n0, bar_full + it * 8);
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) SD_TC_STAMP(2);

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* ma1 = &batch.maps[pr.a1_map];
      const CUtensorMap* ma2 = &batch.maps[pr.a2_map];
      const CUtensorMap* mw = &batch.maps[pr.w_map];
      const int kbA = pr.K1 / BK;  // k-blocks that come from A segment 1
      for (int it = 0; it < num_kb; ++it) {
        const int kb = kb0 + it;
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(bar_empty + s * 8, ph ^ 1u);
        const uint32_t sa = base + s * L::kStage, sb = sa + L::kABytes;
        if (it >= n_pref) mbar_expect_tx(bar_full + s * 8, L::kStage);   // first ring round was armed before the wait
        if (kb < kbA) tma_load_2d(sa, ma1, pr.a1_col + kb * BK, m0, bar_full + s * 8);
        else          tma_load_2d(sa, ma2, pr.a2_col + (kb - kbA) * BK, m0, bar_full + s * 8);
        if (it < n_pref) {
          // weight tile already in flight
        } else if (EPI == EPI_GATES) {  // three 64-row boxes: the reset / cand / update rows of this tile's units
#pragma unroll
          for (int j = 0; j < 3; ++j)
            tma_load_2d(sb + j * (64 * BK * 2), mw, kb * BK, pr.w_row + j * pr.e_dg + n0, bar_full + s * 8);
        } else {
          tma_load_2d(sb, mw, kb * BK, pr.w_row + n0, bar_full + s * 8);
        }
        if (it == 0) SD_TC_STAMP(3);
      }
    }
    if (EPI == EPI_NORM) {  // every thread of the cluster takes part in the epilogue's cluster barrier
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(bar_full + s * 8, ph);
        if (it == 0) SD_TC_STAMP(4);
        if (it == num_kb - 1) SD_TC_STAMP(5);
        tc_fence_after();
        const uint32_t sa = base + s * L::kStage, sb = sa + L::kABytes;
        const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // +32 B per UMMA_K inside the 128 B swizzle row => +2 in the >>4 field
          tc_mma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
        tc_commit(bar_empty + s * 8);  // frees the smem stage once these MMAs have read it
      }
      if (num_kb > 0) tc_commit(bar_acc);  // accumulator complete
    }
    if (EPI == EPI_NORM) {
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  } else {
    // epilogue: warp w owns TMEM lanes [32*(w%4), +32) == output rows m0 + 32*(w%4) + lane
    const int quad = warp & 3;
    if (num_kb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) SD_TC_STAMP(6);
    if (EPI == EPI_GATES) {
      // warp (quad, half) owns rows [32*quad, +32) and units [n0 + 32*half, +32): its reset / cand / update
      // pre-activations are TMEM columns 32*half, 64 + 32*half, 128 + 32*half.
      const int halfg = (warp - 2) >> 2;
      constexpr int SLDG = 33;
      float* stg = reinterpret_cast<float*>(gen_base) + (warp - 2) * (32 * SLDG);
      const int rb = m0 + quad * 32;
      const int u0 = n0 + halfg * 32;
      // stage deter_in[32 rows][32 units] with coalesced loads (lane = unit), then read it row-wise
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr)
        stg[rr * SLDG + lane] = (rb + rr < batch.R) ? pr.e_in[(size_t)(rb + rr) * pr.e_ld_in + u0 + lane] : 0.f;
      __syncwarp();
      float qr[32], qc[32], qu[32];
      const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16);
      tmem_ld32(tb + (uint32_t)(halfg * 32), qr);
      tmem_ld32(tb + (uint32_t)(64 + halfg * 32), qc);
      tmem_ld32(tb + (uint32_t)(128 + halfg * 32), qu);
      float outv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float br = pr.bias ? pr.bias[u0 + j] : 0.f;
        const float bc = pr.bias ? pr.bias[pr.e_dg + u0 + j] : 0.f;
        const float bu = pr.bias ? pr.bias[2 * pr.e_dg + u0 + j] : 0.f;
        // bf16-operand path (tolerance ~1e-2): hardware exp2-based intrinsics instead of the accurate libm
        // routines the fp32 parity path uses; tanh(x) = 1 - 2 / (1 + e^{2x})
        const float reset = __fdividef(1.f, 1.f + __expf(-(qr[j] + br)));
        const float cand = 1.f - __fdividef(2.f, 1.f + __expf(2.f * (reset * (qc[j] + bc))));
        const float upd = __fdividef(1.f, 1.f + __expf(-((qu[j] + bu) - 1.f)));
        outv[j] = upd * cand + (1.f - upd) * stg[lane * SLDG + j];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * SLDG + j] = outv[j];
      __syncwarp();
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        if (rb + rr < batch.R) {
          const float o = stg[rr * SLDG + lane];
          pr.e_out[(size_t)(rb + rr) * pr.e_ld_out + u0 + lane] = o;
          if (pr.e_out_bf) pr.e_out_bf[(size_t)(rb + rr) * pr.e_ld_bf + u0 + lane] = __float2bfloat16(o);
        }
      }
    } else if (EPI == EPI_NORM) {
      // warp (quad, half) owns rows [32*quad, +32) x columns [n0 + 32*half, +32) of this CTA's 64-wide tile
      const int halfn = (warp - 2) >> 2;
      constexpr int SLDN = 33;
      float* stg = reinterpret_cast<float*>(gen_base) + (warp - 2) * (32 * SLDN);
      float* ssq = reinterpret_cast<float*>(gen_base + L::kNormOff);       // [8 partials][128 rows], never touched by TMA
      const int rb = m0 + quad * 32;
      const int c0 = n0 + halfn * 32;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(halfn * 32), v);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] += pr.bias ? pr.bias[c0 + j] : 0.f;
        ss = fmaf(v[j], v[j], ss);
      }
      // partial (cluster rank, half) of row (quad*32 + lane) -> slot of every CTA in the cluster
      uint32_t my_rank;
      asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(my_rank));
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(ssq + (my_rank * 2 + halfn) * 128 + quad * 32 + lane);
#pragma unroll
      for (uint32_t r = 0; r < 4; ++r) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(ss) : "memory");
      }
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      float tot = 0.f;
#pragma unroll
      for (int pidx = 0; pidx < 8; ++pidx) tot += ssq[pidx * 128 + quad * 32 + lane];   // fixed order
      const float rs = 1.f / sqrtf(tot / (float)pr.N + 1e-4f);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float y = (v[j] * rs) * pr.e_gain[c0 + j];
        stg[lane * SLDN + j] = __fdividef(y, 1.f + __expf(-y));
      }
      __syncwarp();
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        if (rb + rr < batch.R) {
          const float o = stg[rr * SLDN + lane];
          if (pr.e_out) pr.e_out[(size_t)(rb + rr) * pr.e_ld_out + c0 + lane] = o;
          if (pr.e_out_bf) pr.e_out_bf[(size_t)(rb + rr) * pr.e_ld_bf + c0 + lane] = __float2bfloat16(o);
        }
      }
    } else if (EPI == EPI_NORMW) {
      // warp (quad, half): rows [32*quad, +32) (thread = row), columns [128*half, +128) of the 256-wide row
      const int halfw = (warp - 2) >> 2;
      float* ssq = reinterpret_cast<float*>(gen_base + L::kNormOff);   // [2 halves][128 rows]
      const int rb = m0 + quad * 32;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(halfw * 128);
      float v[32];
      float ss = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float t = v[j] + (pr.bias ? __ldg(pr.bias + halfw * 128 + c0 + j) : 0.f);
          ss = fmaf(t, t, ss);
        }
      }
      ssq[halfw * 128 + quad * 32 + lane] = ss;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps
      const float tot = ssq[quad * 32 + lane] + ssq[128 + quad * 32 + lane];
      const float rs = 1.f / sqrtf(tot / (float)pr.N + 1e-4f);
      const bool rowok = rb + lane < batch.R;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tmem_ld32(trow + (uint32_t)c0, v);
        const int cb = halfw * 128 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float y = ((v[j] + (pr.bias ? __ldg(pr.bias + cb + j) : 0.f)) * rs) * __ldg(pr.e_gain + cb + j);
          v[j] = __fdividef(y, 1.f + __expf(-y));
        }
        if (rowok) {
          if (pr.e_out_bf) {
            __nv_bfloat16* ob = pr.e_out_bf + (size_t)(rb + lane) * pr.e_ld_bf + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(ob + j) = pk;
            }
          }
          if (pr.e_out) {
            float* of = pr.e_out + (size_t)(rb + lane) * pr.e_ld_out + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(of + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
    } else {
    float* cbase = slice == 0 ? pr.C : pr.Cpart + (long long)(slice - 1) * batch.part_stride;
    const float* bias = slice == 0 ? pr.bias : nullptr;
    // Coalesced, vectorised epilogue.  tcgen05.ld hands each thread one ROW (32 consecutive columns); storing
    // that directly issues 32 scattered 16-byte requests per instruction.  Each warp instead transposes its
    // 32x32 block through a padded staging tile in the (now idle) pipeline shared memory and stores float4s
    // so that every store instruction writes four full 128-byte row segments.  Two warps share each TMEM
    // lane quadrant and split the tile's columns (a lone warp per SMSP is issue-latency bound).
    const int half = (warp - 2) >> 2;                 // 0 or 1
    constexpr int SLD = 36;                           // staging row stride in floats (16 B aligned, conflict free)
    float* stage = reinterpret_cast<float*>(gen_base) + (warp - 2) * (32 * SLD);
    const int row_base = m0 + quad * 32;
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    const bool vec_ok = (pr.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(cbase) & 15) == 0;
#pragma unroll 1
    for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
      float v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stage + lane * SLD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int col = n0 + c0 + c4;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) {
        if (col + 0 < pr.N) bv.x = bias[col + 0];
        if (col + 1 < pr.N) bv.y = bias[col + 1];
        if (col + 2 < pr.N) bv.z = bias[col + 2];
        if (col + 3 < pr.N) bv.w = bias[col + 3];
      }
      float* cptr = cbase + (size_t)(row_base + sub_r) * pr.ldc + col;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + sub_r;
        float4 o = *reinterpret_cast<const float4*>(stage + rr * SLD + c4);
        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
        if (row_base + rr < batch.R) {
          if (vec_ok && col + 3 < pr.N) {
            *reinterpret_cast<float4*>(cptr) = o;
          } else {
            if (col + 0 < pr.N) cptr[0] = o.x;
            if (col + 1 < pr.N) cptr[1] = o.y;
            if (col + 2 < pr.N) cptr[2] = o.z;
            if (col + 3 < pr.N) cptr[3] = o.w;
          }
        }
        cptr += (size_t)4 * pr.ldc;
      }
      __syncwarp();
    }
    }  // EPI_STORE
    if (threadIdx.x == 64) SD_TC_STAMP(7);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SD_TC_STAMP(8);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace tc
}  // namespace sd
