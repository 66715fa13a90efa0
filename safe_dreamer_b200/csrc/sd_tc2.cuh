// sd_tc2.cuh -- first layers of the frozen heads on a CTA PAIR (tcgen05 cta_group::2), sm_100a.
//
//   act_h[R x 256] (bf16) = SiLU(RMSNorm_256(feat[R x K] (bf16) * W_h[256 x K]^T + b_h) * g_h)      h = 0, 1
//
// (networks.py:339-377: the first Linear -> RMSNorm -> SiLU of two MLPHeads that read the same features, dreamer.py:589-596.)
// On the 16 384 imagined rows this layer is the GEMM-shaped part of the head evaluation (K = 2560).  The single-CTA kernel
// (sd_tc.cuh, 128 x 256 tile) pulls 48 KB through the SM's ingest path per 512 tensor-core cycles and is bound by it (40 %
// tensor pipe, profiles/r02c_heads_chain_ncu.txt).  Here two CTAs of a cluster form one UMMA of M = 256:
//   * each CTA loads ITS 128 feature rows (16 KB per k-block) and HALF of each head's weight tile (rows [128*rank, +128):
//     16 KB per head per k-block) -- the pair reads B once instead of twice, and the feature tile is used by both heads;
//   * the leader CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) per head; the hardware
//     reads A rows 0-127 / B rows 0-127 from the leader's shared memory and rows 128-255 from the peer's, at the same
//     offsets; each CTA's TMEM receives the accumulator rows of its own 128 feature rows, head h in columns [256 h, +256);
//   * 48 KB per CTA per 1024 tensor-core cycles: half the ingest per FLOP of the single-CTA tile.
// Pipeline: per CTA a 4-stage ring; the peer's TMA loads signal the LEADER's full barrier (cp.async.bulk.tensor with
// .cta_group::2 and a cluster-mapped mbarrier address), one arrive.expect_tx per CTA; tcgen05.commit.cta_group::2 with
// .multicast::cluster releases the stage in both CTAs and finally publishes the accumulators to both epilogues.
// Epilogue: as EPI_NORMW of sd_tc.cuh (a CTA owns whole 256-wide rows of a head: RMSNorm needs no cluster exchange).
// Every mbarrier wait is bounded (trap instead of hang).
#pragma once
#include "sd_tc.cuh"

namespace sd {
namespace tc2 {

using tc::BK;
using tc::BM;
constexpr int THREADS = 320;   // warp 0 TMA, warp 1 TMEM + MMA issue, warps 2-9 epilogue
constexpr int STAGES = 4;
constexpr int NH = 256;        // width of a head's first layer
constexpr int kABytes = BM * BK * 2;          // 16 KB: this CTA's feature rows
constexpr int kBHalf = (NH / 2) * BK * 2;     // 16 KB: this CTA's half of one head's weight tile
constexpr int kStage = kABytes + 2 * kBHalf;  // 48 KB
constexpr int kBarOff = STAGES * kStage;      // full[STAGES] | empty[STAGES] | acc | tmem slot
constexpr int kNormOff = kBarOff + 256;       // [2 halves][128 rows] partial sums of squares
constexpr int kBgOff = kNormOff + 1024;       // [2 heads][bias | gain][256] fp32
constexpr int kSmem = kBgOff + 4096 + 1024;   // + alignment slack

struct Params {
  CUtensorMap map_a;      // feat bf16 [R x K], box 64 x 128
  CUtensorMap map_w[2];   // head weights bf16 [256 x K], box 64 x 128
  const float* bias[2];
  const float* gain[2];
  __nv_bfloat16* out[2];  // [R x 256], row stride ld_out
  int ld_out;
  int R, K, nheads;       // nheads = 1: a lone head (the second weight tile is never loaded / multiplied)
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive (count 1) + expect `bytes` on an mbarrier of another CTA of the cluster (cluster-mapped address)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion bytes are signalled on an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// SiLU with ONE MUFU op per element: y * sigmoid(y) = y * (0.5 + 0.5 tanh(y / 2)).  The epilogue of a 128 x 256 tile is MUFU
// bound (32 768 elements per head on 16 MUFU lanes per clock); exp + reciprocal would be two.
__device__ __forceinline__ float silu_tanh(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return y * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ void mma_pair_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the mbarrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) heads_first_pair_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
#ifdef SD_TC2_DIAG
  const long long t_start = clock64();
#endif
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;   // same offset in both CTAs of the pair
  uint8_t* gen_base = smem_raw + (base - tc::smem_u32(smem_raw));
  const uint32_t bar_full = base + kBarOff;            // STAGES x 8 B (waited on by the leader only)
  const uint32_t bar_empty = bar_full + STAGES * 8;    // STAGES x 8 B
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  const uint32_t tmem_slot = bar_acc + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + kBarOff + STAGES * 16 + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int m0 = pair * (2 * BM) + (int)rank * BM;     // this CTA's first feature row
  const int num_kb = P.K / BK;
  const int nh = P.nheads;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(bar_full + s * 8, 2);   // one arrive.expect_tx per CTA of the pair
      tc::mbar_init(bar_empty + s * 8, 1);
    }
    tc::mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync();   // both CTAs' barriers are initialised before any remote arrive / TMA completion can reach them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t full_leader0 = map_to_rank(bar_full, 0);   // the leader's full barriers in the cluster window
  if (warp == 0) {
    // producer: arms the leader's full barrier with this CTA's bytes of the stage, loads its feature tile and its halves of
    // the heads' weight tiles (splitting the three boxes over two issuing threads was measured: no difference)
    if (lane == 0) {
      const uint32_t stage_bytes = (uint32_t)(kABytes + nh * kBHalf);
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        tc::mbar_wait(bar_empty + s * 8, ph ^ 1u);
        mbar_expect_tx_cluster(full_leader0 + s * 8, stage_bytes);
        tma_load_2d_pair(base + s * kStage, &P.map_a, it * BK, m0, full_leader0 + s * 8);
        for (int h = 0; h < nh; ++h)
          tma_load_2d_pair(base + s * kStage + kABytes + h * kBHalf, &P.map_w[h], it * BK, (int)rank * (NH / 2), full_leader0 + s * 8);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc(2 * BM, NH);
#ifdef SD_TC2_DIAG
      long long t_wait = 0, t0 = clock64();
#endif
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
#ifdef SD_TC2_DIAG
        const long long w0 = clock64();
#endif
        tc::mbar_wait(bar_full + s * 8, ph);
#ifdef SD_TC2_DIAG
        t_wait += clock64() - w0;
#endif
        tc::tc_fence_after();
        const uint32_t sa = base + s * kStage;
        const uint64_t da = tc::make_desc_sw128(sa);
        for (int h = 0; h < nh; ++h) {
          const uint64_t db = tc::make_desc_sw128(sa + kABytes + h * kBHalf);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            mma_pair_f16(tmem_base + (uint32_t)(h * NH), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
        }
        commit_pair(bar_empty + s * 8);   // frees this stage in both CTAs once the MMAs have read it
      }
      commit_pair(bar_acc);               // accumulators complete: both epilogues
#ifdef SD_TC2_DIAG
      const long long t1 = clock64();
      tc::mbar_wait(bar_acc, 0);
      if (blockIdx.x == 0 || blockIdx.x == 100)
        printf("[tc2 diag] cta %d: mma loop %lld cycles (%d k-blocks, %d heads), of which waiting for full %lld; drain %lld\n",
               (int)blockIdx.x, t1 - t0, num_kb, nh, t_wait, clock64() - t1);
#endif
    }
    __syncwarp();
  } else {
    // epilogue: warp (quad, half) owns rows [32*quad, +32) (thread = row) x columns [128*half, +128) of each head's 256-wide row
    const int quad = warp & 3, halfw = (warp - 2) >> 2;
    float* ssq = reinterpret_cast<float*>(gen_base + kNormOff);   // [2][128]
    // bias / RMS scale of both heads -> shared memory while the main loop runs
    float* s_bg = reinterpret_cast<float*>(gen_base + kBgOff);
    for (int i = threadIdx.x - 64; i < nh * 512; i += THREADS - 64) {
      const int hh = i >> 9, w = (i >> 8) & 1, c = i & 255;
      const float* src = w ? P.gain[hh] : P.bias[hh];
      s_bg[i] = src ? __ldg(src + c) : (w ? 1.f : 0.f);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tc::mbar_wait(bar_acc, 0);
    tc::tc_fence_after();
#ifdef SD_TC2_DIAG
    long long e_t[4] = {clock64(), 0, 0, 0};
#endif
    const int r = quad * 32 + lane;   // this thread's row of the tile
    const int e = warp - 2;           // 0..7
    for (int h = 0; h < nh; ++h) {
      const float* sb = s_bg + h * 512 + halfw * 128;   // bias / gain of this thread's 128 columns (broadcast reads)
      const float* sg = sb + 256;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * NH + halfw * 128);
      float v[32];
      float ss = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tc::tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float t = v[j] + sb[c0 + j];
          ss = fmaf(t, t, ss);
        }
      }
      ssq[halfw * 128 + r] = ss;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps
      const float tot = ssq[r] + ssq[128 + r];
      const float rs = 1.f / sqrtf(tot / (float)NH + 1e-4f);
      // normalised bf16 rows go to a staging tile in the (idle) pipeline stages, 16-byte chunks XOR-swizzled by the row so that
      // the row-per-lane writes here and the chunk-per-lane reads below are both bank-conflict free; the global stores are
      // then whole 512-byte rows per warp instruction instead of 32 scattered 16-byte pieces
      uint8_t* stile = gen_base + h * (BM * NH * 2);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tc::tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = silu_tanh(((v[j] + sb[c0 + j]) * rs) * sg[c0 + j]);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
          const int c = (halfw * 128 + c0 + j) >> 3;   // 16-byte chunk of the 512-byte row
          *reinterpret_cast<uint4*>(stile + r * (NH * 2) + (((c & ~7) | ((c ^ r) & 7)) << 4)) = pk;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // tile staged (and ssq free for the next head)
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int row = e * 16 + i, gr = m0 + row;
        if (gr < P.R) {
          const uint4 val = *reinterpret_cast<const uint4*>(stile + row * (NH * 2) + (((lane & ~7) | ((lane ^ row) & 7)) << 4));
          *reinterpret_cast<uint4*>(P.out[h] + (size_t)gr * P.ld_out + lane * 8) = val;
        }
      }
#ifdef SD_TC2_DIAG
      e_t[1 + h] = clock64();
#endif
    }
#ifdef SD_TC2_DIAG
    if (blockIdx.x == 0 && threadIdx.x == 64)
      printf("[tc2 diag] cta 0 epilogue: head0 %lld cycles, head1 %lld cycles; kernel start -> acc ready %lld\n", e_t[1] - e_t[0],
             nh > 1 ? e_t[2] - e_t[1] : 0ll, e_t[0] - t_start);
#endif
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer may still be reading its TMEM half / the leader's MMAs may still read the peer's shared memory
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace tc2
}  // namespace sd
