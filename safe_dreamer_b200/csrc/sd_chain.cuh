// sd_chain.cuh -- row-tile resident MLP chain on tcgen05 (sm_100a).
//
// On the measured path as the trunk of the frozen heads (reward / cont / value / slow value, dreamer.py:589-596): after the
// wide first layer (F -> 256) a head is hidden layers 256 -> 256 + a last layer 256 -> bins on N*H = 16384 rows, i.e. 128
// row tiles: one launch per head instead of one per layer (head_chain in sd_api.cu).  The imagination variants below are
// opt-in (SD_CHAIN=1): at N = 1024 only 8 CTAs have rows, see profiles/r01b_chain_phase_stamps.txt.
//
// The imagination step is a chain of small dependent dense layers: actor2 -> actor3 -> actor head -> action sample
// -> dyn_in2 (dreamer.py:684, networks.py:339-377, rssm.py:44-48) and img_net_1 -> img_net_logit (rssm.py:119-130).
// Every one of them is 256 -> 256 (or 256 -> <= 512) on the same rows, so launching them as separate kernels is
// pure latency.  Here ONE CTA owns a 128-row tile for the whole chain:
//
//   prologue (16 epilogue warps, warp = row):  x = SiLU(RMSNorm(sum of the split-K slices the preceding wide GEMM left))
//             written as bf16 straight into shared memory in the K-major SWIZZLE_128B layout the UMMA descriptor
//             reads (the activation never goes back to HBM/L2 between layers)
//   per layer: warp 0 streams the layer's bf16 weights through a 4 x 32 KB TMA ring (all weight loads are
//             independent of the preceding kernel, so they start before griddepcontrol.wait); warp 1 issues
//             tcgen05.mma (M=128, N<=256, K=16) into TMEM; the epilogue warps read the accumulators back
//             (tcgen05.ld), add bias, RMSNorm + SiLU (row statistics exchanged through shared memory) and write
//             the next layer's A operand in place
//   final layer: (1) fp32 store of up to 512 outputs (logits), coalesced through a shared-memory transpose, or
//             (2) the actor tail: bounded-normal / one-hot sample with injected noise, action normalisation,
//             dyn_in2 projection + RMSNorm + SiLU -> x2 (bf16 operand of the block-GRU input layer).
// CTAs with blockIdx.x >= n_tiles run a side job instead: RMSNorm + SiLU of other split-K results of the same
// preceding launch (dyn_in0 / dyn_in1), so that one launch finishes everything the block-GRU needs.
#pragma once
#include "sd_kernels.cuh"
#include "sd_tc.cuh"

namespace sd {
namespace chain {

using tc::BK;
using tc::BM;

constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA/TMEM, warps 2-17 epilogue
constexpr int HID = 256;                      // width of every chained activation (K of every chained layer)
constexpr int KBLK = HID / BK;                // 4 k-blocks of 64
constexpr int WST = 4;                        // weight ring stages
constexpr int kUnit = 256 * BK * 2;           // one weight ring slot: 256 rows x 64 k (bf16) = 32 KB
constexpr int kATile = BM * BK * 2;           // one k-block of the activation tile = 16 KB
constexpr int kMaxLayers = 3;                 // up to two hidden layers + the final layer
constexpr int kTailMaxOut = 48;               // actor head outputs (padded to a multiple of 16 for the MMA)
constexpr int TAIL_LD = 49;

constexpr int kOffA = 0;
constexpr int kOffW = KBLK * kATile;                 // 64 KB
constexpr int kOffBar = kOffW + WST * kUnit;         // 192 KB
constexpr int kOffBias = kOffBar + 128;
constexpr int kOffGain = kOffBias + kMaxLayers * 512 * 4;
constexpr int kOffSsq = kOffGain + kMaxLayers * HID * 4;
constexpr int kSmemBytes = kOffSsq + 4 * BM * 4 + 1024;  // + alignment slack

struct Layer {
  int w_map;          // tensor map of the bf16 [npad x 256] weight (box = 64 k x box_rows rows)
  int N;              // logical outputs
  int box_rows;       // 256 (N up to 512, loaded as 256-row halves) or 64 (N <= 64)
  const float* bias;  // [N]
  const float* gain;  // RMS scale [256] (hidden layers only)
};
struct Side {  // y = SiLU(RMSNorm_256(in + parts)) -> bf16
  const float* in; int ld_in;
  const float* parts; int nparts;
  const float* gain;
  __nv_bfloat16* out_bf; int ld_bf;
};
struct Params {
  CUtensorMap maps[kMaxLayers];
  Layer layer[kMaxLayers];
  int n_layers;  // the last one is the final layer
  int R, n_tiles;
  // input of the chain: pre-norm output of the preceding wide layer (slice 0 + nparts split-K slices)
  const float* in; int ld_in;
  const float* parts; int nparts;
  const float* in_gain;
  long long part_stride;
  int fin_mode;  // 1: store fp32; 2: actor tail; 3: TwoHot.mode; 4: sigmoid of logit 0
  float* out; int ld_out;
  // actor tail
  int act_out, A, act_kind;
  float min_std, max_std, unimix;
  const float* noise; int ld_n;
  const float* w2_t; int ldw_2;  // dyn_in2 weight, [A][ldw_2] (n contiguous)
  const float* b2; const float* g2;
  float* aout; float* action; int ld_act; float* abar;
  __nv_bfloat16* x2_bf; int ld_x2;
  Side side[2];
  int n_side;
  // alternative input: the ALREADY normalised + activated bf16 output of the preceding layer (row stride ld_in elements);
  // the prologue then only copies rows into the swizzled A tile (frozen-head trunks, networks.py:339-377)
  const __nv_bfloat16* in_bf;
  // fin_mode 3: TwoHot.mode of the last layer's `N` logits over `bins` (distributions.py:78-98) -> scalar[row];
  // fin_mode 4: sigmoid of logit 0 (the cont head's mean, dreamer.py:592) -> scalar[row].  The logits never leave the CTA.
  const float* bins;
  float* scalar;
  long long* timing;   // diagnostic (SD_TRACE_CHAIN=1): clock64 stamps of CTA 0; null in production
};
#define SD_CH_STAMP(i) do { if (P.timing && blockIdx.x == 0) P.timing[i] = clock64(); } while (0)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// y * sigmoid(y) = y * (0.5 + 0.5 tanh(y / 2)): ONE MUFU op per element (exp + reciprocal are two); the epilogues of the
// 128 x 256 tiles are MUFU bound
__device__ __forceinline__ float silu_fast(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return y * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack_bf8(const float* y) {
  return make_uint4(pack_bf2(y[0], y[1]), pack_bf2(y[2], y[3]), pack_bf2(y[4], y[5]), pack_bf2(y[6], y[7]));
}

// one warp normalises one 256-wide row: lane owns columns [8*lane, +8)
__device__ __forceinline__ void norm_row_256(const float* in, const float* parts, int nparts, long long part_stride,
                                             const float* g, int lane, float* y) {
  const float4 a = *reinterpret_cast<const float4*>(in + lane * 8);
  const float4 b = *reinterpret_cast<const float4*>(in + lane * 8 + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  for (int s = 0; s < nparts; ++s) {  // fixed slice order
    const float4 c = *reinterpret_cast<const float4*>(parts + s * part_stride + lane * 8);
    const float4 d = *reinterpret_cast<const float4*>(parts + s * part_stride + lane * 8 + 4);
    v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w; v[4] += d.x; v[5] += d.y; v[6] += d.z; v[7] += d.w;
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss * (1.f / HID) + kRmsEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = silu_fast((v[i] * rs) * g[i]);
}

__device__ __forceinline__ void side_job(const Params& P, int warp, int lane) {
  float g0[8], g1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g0[i] = P.n_side > 0 ? __ldg(P.side[0].gain + lane * 8 + i) : 0.f;
    g1[i] = P.n_side > 1 ? __ldg(P.side[1].gain + lane * 8 + i) : 0.f;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (P.n_side <= 0) return;
  const int nw = ((int)gridDim.x - P.n_tiles) * (THREADS / 32);
  const int w0 = ((int)blockIdx.x - P.n_tiles) * (THREADS / 32) + warp;
  const int items = P.R * P.n_side;
  for (int it = w0; it < items; it += nw) {
    const int row = it / P.n_side, seg = it - row * P.n_side;
    const Side& s = P.side[seg];
    float g[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = seg ? g1[i] : g0[i];
    norm_row_256(s.in + (size_t)row * s.ld_in, s.parts + (size_t)row * s.ld_in, s.nparts, P.part_stride, g, lane, y);
    *reinterpret_cast<uint4*>(s.out_bf + (size_t)row * s.ld_bf + lane * 8) = pack_bf8(y);
  }
}

__global__ void __launch_bounds__(THREADS, 1) mlp_chain_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x >= P.n_tiles) {
    side_job(P, warp, lane);
    return;
  }
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - tc::smem_u32(smem_raw));
  const uint32_t sA = base + kOffA, sW = base + kOffW;
  const uint32_t bar_full = base + kOffBar;        // WST x 8
  const uint32_t bar_empty = bar_full + WST * 8;   // WST x 8
  const uint32_t bar_acc = bar_empty + WST * 8;    // accumulators of a layer complete
  const uint32_t bar_a = bar_acc + 8;              // A operand of a layer complete (one arrival per epilogue warp)
  const uint32_t tmem_slot = bar_a + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + kOffBar + (2 * WST + 2) * 8);
  float* s_bias = reinterpret_cast<float*>(gbase + kOffBias);  // [layer][512]
  float* s_gain = reinterpret_cast<float*>(gbase + kOffGain);  // [layer][256]
  float* s_ssq = reinterpret_cast<float*>(gbase + kOffSsq);    // [4][128]
  const int m0 = blockIdx.x * BM;
  const int nl = P.n_layers;
  if (threadIdx.x == 64) SD_CH_STAMP(0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WST; ++s) {
      tc::mbar_init(bar_full + s * 8, 1);
      tc::mbar_init(bar_empty + s * 8, 1);
    }
    tc::mbar_init(bar_acc, 1);
    tc::mbar_init(bar_a, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ---------------------------------------------------------------- weight producer (never touches activations)
    if (lane == 0) {
      int q = 0;
      for (int l = 0; l < nl; ++l) {
        const Layer& L = P.layer[l];
        const int halves = L.box_rows == 256 ? (L.N + 255) / 256 : 1;
        const uint32_t bytes = (uint32_t)L.box_rows * BK * 2;
        for (int hf = 0; hf < halves; ++hf)
          for (int kb = 0; kb < KBLK; ++kb, ++q) {
            const int s = q % WST;
            const uint32_t ph = (uint32_t)(q / WST) & 1u;
            tc::mbar_wait(bar_empty + s * 8, ph ^ 1u);
            tc::mbar_expect_tx(bar_full + s * 8, bytes);
            tc::tma_load_2d(sW + s * kUnit, &P.maps[L.w_map], kb * BK, hf * 256, bar_full + s * 8);
          }
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (lane == 0) {
      int q = 0;
      for (int l = 0; l < nl; ++l) {
        const Layer& L = P.layer[l];
        const int halves = L.box_rows == 256 ? (L.N + 255) / 256 : 1;
        tc::mbar_wait(bar_a, (uint32_t)l & 1u);
        tc::tc_fence_after();
        for (int hf = 0; hf < halves; ++hf) {
          const int n_mma = L.box_rows == 256 ? 256 : ((L.N + 15) / 16) * 16;
          const uint32_t idesc = tc::make_idesc(BM, n_mma);
          for (int kb = 0; kb < KBLK; ++kb, ++q) {
            const int s = q % WST;
            const uint32_t ph = (uint32_t)(q / WST) & 1u;
            tc::mbar_wait(bar_full + s * 8, ph);
            tc::tc_fence_after();
            const uint64_t da = tc::make_desc_sw128(sA + kb * kATile), db = tc::make_desc_sw128(sW + s * kUnit);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::tc_mma_f16(tmem_base + (uint32_t)(hf * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                             (kb | k) != 0 ? 1u : 0u);
            tc::tc_commit(bar_empty + s * 8);
          }
        }
        tc::tc_commit(bar_acc);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int e = warp - 2;            // 0..15
    const int quad = warp & 3;         // TMEM lane quadrant this warp may read
    const int colq = e >> 2;           // column quarter
    const int r = quad * 32 + lane;    // tile row of this thread in the TMEM-shaped phases
    // weights (biases, RMS scales) before the PDL wait
    for (int l = 0; l < nl; ++l) {
      const Layer& L = P.layer[l];
      for (int i = e * 32 + lane; i < 512; i += EPI_WARPS * 32) s_bias[l * 512 + i] = (L.bias && i < L.N) ? __ldg(L.bias + i) : 0.f;
      if (L.gain)
        for (int i = e * 32 + lane; i < HID; i += EPI_WARPS * 32) s_gain[l * HID + i] = __ldg(L.gain + i);
    }
    float gin[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gin[i] = P.in_bf ? 0.f : __ldg(P.in_gain + lane * 8 + i);
    if (threadIdx.x == 64) SD_CH_STAMP(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 64) SD_CH_STAMP(2);

    // ---- prologue: A0 = SiLU(RMSNorm(in + parts)) for rows e, e+16, ... (warp = row)
    if (P.in_bf) {   // already normalised bf16 rows: a plain copy into the swizzled tile, the warp's 8 row loads in flight together
      uint4 pk[BM / EPI_WARPS];
#pragma unroll
      for (int i = 0; i < BM / EPI_WARPS; ++i) {
        const int gr = m0 + e + i * EPI_WARPS;
        pk[i] = gr < P.R ? *reinterpret_cast<const uint4*>(P.in_bf + (size_t)gr * P.ld_in + lane * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < BM / EPI_WARPS; ++i) {
        const int rr = e + i * EPI_WARPS;
        *reinterpret_cast<uint4*>(gbase + kOffA + (lane >> 3) * kATile + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4)) = pk[i];
      }
    } else
#pragma unroll 2
    for (int rr = e; rr < BM; rr += EPI_WARPS) {
      const int gr = m0 + rr;
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (gr < P.R) {
        float y[8];
        norm_row_256(P.in + (size_t)gr * P.ld_in, P.parts + (size_t)gr * P.ld_in, P.nparts, P.part_stride, gin, lane, y);
        pk = pack_bf8(y);
      }
      // columns [8*lane, +8) = 16-byte chunk (lane & 7) of k-block (lane >> 3)
      uint8_t* dst = gbase + kOffA + (lane >> 3) * kATile + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = pk;
    }
    fence_async_smem();
    epi_bar();   // also publishes s_bias / s_gain
    if (lane == 0) mbar_arrive(bar_a);
    if (threadIdx.x == 64) SD_CH_STAMP(3);

    // ---- hidden layers
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int l = 0; l < nl - 1; ++l) {
      tc::mbar_wait(bar_acc, (uint32_t)l & 1u);
      tc::tc_fence_after();
      if (threadIdx.x == 64) SD_CH_STAMP(4 + 2 * l);
      const float* bs = s_bias + l * 512 + colq * 64;
      const float* gs = s_gain + l * HID + colq * 64;
      float v[32];
      float ss = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tc::tmem_ld32(trow + (uint32_t)(colq * 64 + hh * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float t = v[j] + bs[hh * 32 + j];
          ss = fmaf(t, t, ss);
        }
      }
      s_ssq[colq * BM + r] = ss;
      epi_bar();
      const float tot = ((s_ssq[r] + s_ssq[BM + r]) + s_ssq[2 * BM + r]) + s_ssq[3 * BM + r];
      const float rs = 1.f / sqrtf(tot * (1.f / HID) + kRmsEps);
      uint8_t* arow = gbase + kOffA + colq * kATile + r * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tc::tmem_ld32(trow + (uint32_t)(colq * 64 + hh * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = silu_fast(((v[j] + bs[hh * 32 + j]) * rs) * gs[hh * 32 + j]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(arow + (((hh * 4 + c) ^ (r & 7)) << 4)) = pack_bf8(v + c * 8);
      }
      tc::tc_fence_before();
      fence_async_smem();
      epi_bar();   // every warp is done with s_ssq and TMEM before anything is overwritten
      if (lane == 0) mbar_arrive(bar_a);
      if (threadIdx.x == 64) SD_CH_STAMP(5 + 2 * l);
    }

    // ---- final layer
    const int lf = nl - 1;
    const Layer& LF = P.layer[lf];
    tc::mbar_wait(bar_acc, (uint32_t)lf & 1u);
    tc::tc_fence_after();
    if (threadIdx.x == 64) SD_CH_STAMP(8);
    // from here on the A tile and the weight ring are free: use them as scratch
    if (P.fin_mode == 1) {
      const int npad = ((LF.N + 255) / 256) * 256;   // 256 or 512
      const int cpq = npad / 4;                      // columns per quarter
      constexpr int SLD = 36;
      float* stage = reinterpret_cast<float*>(gbase + kOffW) + e * (32 * SLD);
      const float* bs = s_bias + lf * 512;
      const int row_base = m0 + quad * 32;
      const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
      const bool vec_ok = (P.ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out) & 15) == 0;
#pragma unroll 1
      for (int c0 = colq * cpq; c0 < (colq + 1) * cpq; c0 += 32) {
        float v[32];
        tc::tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * SLD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int col = c0 + c4;
        const float4 bv = make_float4(bs[col], bs[col + 1], bs[col + 2], bs[col + 3]);
        float* cptr = P.out + (size_t)(row_base + sub_r) * P.ld_out + col;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + sub_r;
          float4 o = *reinterpret_cast<const float4*>(stage + rr * SLD + c4);
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          if (row_base + rr < P.R) {
            if (vec_ok && col + 3 < LF.N) {
              *reinterpret_cast<float4*>(cptr) = o;
            } else {
              if (col + 0 < LF.N) cptr[0] = o.x;
              if (col + 1 < LF.N) cptr[1] = o.y;
              if (col + 2 < LF.N) cptr[2] = o.z;
              if (col + 3 < LF.N) cptr[3] = o.w;
            }
          }
          cptr += (size_t)4 * P.ld_out;
        }
        __syncwarp();
      }
    } else if (P.fin_mode == 3 || P.fin_mode == 4) {
      // ---- scalar heads: the tile's logits (+ bias) go to shared memory (the A tile and the weight ring are free), then a
      // warp per row evaluates TwoHot.mode with the operation order of twohot_mode_kernel (softmax, then the reference's
      // symmetric pairing sum_j (p[m-1-j] b[m-1-j] + p[m+1+j] b[m+1+j]) + p[m] b[m]) or the sigmoid of logit 0
      constexpr int RLD = 257;
      float* rows = reinterpret_cast<float*>(gbase + kOffA);   // [128][RLD]
      const float* bs = s_bias + lf * 512;
      {
        float v[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tc::tmem_ld32(trow + (uint32_t)(colq * 64 + hh * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) rows[r * RLD + colq * 64 + hh * 32 + j] = v[j] + bs[colq * 64 + hh * 32 + j];
        }
      }
      epi_bar();
      const int n = LF.N;
      for (int rr = e; rr < BM; rr += EPI_WARPS) {
        const int gr = m0 + rr;
        if (gr >= P.R) break;   // warp-uniform
        float* lp = rows + rr * RLD;
        if (P.fin_mode == 4) {
          if (lane == 0) P.scalar[gr] = sigmoidf_(lp[0]);
          continue;
        }
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, lp[j]);
        m = warp_max(m);
        float sm = 0.f;
        for (int j = lane; j < n; j += 32) {   // the exponentials are computed once and replace the logits in the row buffer
          const float ex = expf(lp[j] - m);
          lp[j] = ex;
          sm += ex;
        }
        sm = warp_sum(sm);
        __syncwarp();
        float acc = 0.f;
        if (n & 1) {
          const int mid = (n - 1) / 2;
          for (int j = lane; j < mid; j += 32) {
            const float lo = (lp[mid - 1 - j] / sm) * __ldg(P.bins + mid - 1 - j);
            const float hi = (lp[mid + 1 + j] / sm) * __ldg(P.bins + mid + 1 + j);
            acc += lo + hi;
          }
          acc = warp_sum(acc);
          acc += (lp[mid] / sm) * __ldg(P.bins + mid);
        } else {
          const int hn = n / 2;
          for (int j = lane; j < hn; j += 32) {
            const float lo = (lp[hn - 1 - j] / sm) * __ldg(P.bins + hn - 1 - j);
            const float hi = (lp[hn + j] / sm) * __ldg(P.bins + hn + j);
            acc += lo + hi;
          }
          acc = warp_sum(acc);
        }
        if (lane == 0) P.scalar[gr] = acc;
      }
    } else {
      // ---- actor tail (dreamer.py:684, distributions.py:217-231, rssm.py:44,48)
      float* s_o = reinterpret_cast<float*>(gbase + kOffW);   // [128][TAIL_LD] head outputs
      float* s_ab = s_o + BM * TAIL_LD;                        // [128][33] normalised action
      float* s_w2 = s_ab + BM * 33;                            // [A][256]
      float* s_b2 = s_w2 + 32 * HID;                           // [256]
      float* s_g2 = s_b2 + HID;                                // [256]
      const int A = P.A;
      for (int i = e * 32 + lane; i < A * HID; i += EPI_WARPS * 32) {
        const int a = i >> 8, n = i & 255;
        s_w2[i] = __ldg(P.w2_t + (size_t)a * P.ldw_2 + n);
      }
      for (int i = e * 32 + lane; i < HID; i += EPI_WARPS * 32) {
        s_b2[i] = __ldg(P.b2 + i);
        s_g2[i] = __ldg(P.g2 + i);
      }
      const int gr = m0 + r;
      if (colq == 0) {
        float v[32];
        const float* bs = s_bias + lf * 512;
        tc::tmem_ld32(trow, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) s_o[r * TAIL_LD + j] = v[j] + bs[j];
        if (P.act_out > 32) {
          tc::tmem_ld32(trow + 32u, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) s_o[r * TAIL_LD + 32 + j] = v[j] + bs[32 + j];
        }
        __syncwarp();
        const float* o = s_o + r * TAIL_LD;
        if (gr < P.R) {
          if (P.aout)
            for (int j = 0; j < P.act_out; ++j) P.aout[(size_t)gr * P.act_out + j] = o[j];
          if (P.act_kind == 0) {
            for (int a = 0; a < A; ++a) {
              const float std = (P.max_std - P.min_std) * sigmoidf_(o[A + a] + 2.f) + P.min_std;
              const float act = tanhf(o[a]) + std * P.noise[(size_t)gr * P.ld_n + a];
              const float ab = act / fmaxf(fabsf(act), 1.f);
              P.action[(size_t)gr * P.ld_act + a] = act;
              P.abar[(size_t)gr * A + a] = ab;
              s_ab[r * 33 + a] = ab;
            }
          } else {
            float lg[32], uu[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              lg[k] = k < A ? o[k] : 0.f;
              uu[k] = k < A ? P.noise[(size_t)gr * P.ld_n + k] : 0.5f;
            }
            const int best = sample_category(lg, uu, A, P.unimix, nullptr);
            for (int k = 0; k < A; ++k) {
              const float av = (k == best) ? 1.f : 0.f;
              P.action[(size_t)gr * P.ld_act + k] = av;
              P.abar[(size_t)gr * A + k] = av;
              s_ab[r * 33 + k] = av;
            }
          }
        } else {
          for (int a = 0; a < A; ++a) s_ab[r * 33 + a] = 0.f;
        }
      }
      epi_bar();
      if (threadIdx.x == 64) SD_CH_STAMP(9);
      // dyn_in2 (A -> 256) + RMSNorm + SiLU; thread = (row, 64-column quarter)
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = s_b2[colq * 64 + j];
      for (int a = 0; a < A; ++a) {
        const float ab = s_ab[r * 33 + a];
        const float* w = s_w2 + a * HID + colq * 64;
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = fmaf(ab, w[j], acc[j]);
      }
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) ss = fmaf(acc[j], acc[j], ss);
      s_ssq[colq * BM + r] = ss;
      epi_bar();
      const float tot = ((s_ssq[r] + s_ssq[BM + r]) + s_ssq[2 * BM + r]) + s_ssq[3 * BM + r];
      const float rs = 1.f / sqrtf(tot * (1.f / HID) + kRmsEps);
      if (gr < P.R) {
        __nv_bfloat16* xo = P.x2_bf + (size_t)gr * P.ld_x2 + colq * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = silu_fast((acc[c * 8 + j] * rs) * s_g2[colq * 64 + c * 8 + j]);
          *reinterpret_cast<uint4*>(xo + c * 8) = pack_bf8(y);
        }
      }
    }
  }
  if (threadIdx.x == 64) SD_CH_STAMP(10);
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) SD_CH_STAMP(11);
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace chain
}  // namespace sd
