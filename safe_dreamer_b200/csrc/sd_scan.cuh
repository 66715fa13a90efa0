// sd_scan.cuh -- persistent, weight-stationary posterior scan (RSSM.observe, rssm.py:140-178) for small batches.
//
// At B <= 16 rows a posterior step is ~170 MFLOP against 21 MB of fp32 weights: launching it as nine dependent
// kernels per step (and re-reading every weight from L2 each step) is pure latency.  This kernel runs ALL T steps
// in one launch on 128 CTAs (32 clusters of 4, one CTA per SM):
//
//   * every CTA copies its slices of the fp32 weights into shared memory ONCE (~144-176 KB per CTA, 21 MB in
//     total across the grid) and keeps them there for the whole scan;
//   * a step is five phases; activations (16 rows) travel between CTAs through L2, never through HBM.  Where every CTA
//     consumes every CTA's output (after P1 and P2) the phases are separated by a grid barrier (one L2 atomic + acquire
//     spin); the three boundaries with few producers (P3 -> P4 / next hidden layer, P4 -> P5, P5 -> P1) are flagged
//     hand-offs: value + tag in ONE 8-byte store, polled by the consumer (see ll_store / ll_load4 / ll_prewait):
//       P1  block-GRU hidden layer   h_pre[g] = W_hid[g] [d_g | x0 | x1 | x2]         (rssm.py:52-61)   128 tiles of 16 columns
//       P2  gate projection + gates  d' = GRU(W_gru[g] SiLU(RMSNorm_2048(h_pre)))     (rssm.py:63-75)   128 tiles of 16 units
//       P3  the two K = 2048 layers that read d':  obs_net_0 (deter part, + the precomputed embed part) and dyn_in0
//           of the NEXT step; K split over the 4 CTAs of a cluster, partial tiles reduced through DSMEM in rank order
//       P4  logits = W_logit SiLU(RMSNorm(v_obs)) + unimix Gumbel sample                (rssm.py:171-177) one tile per 16 classes
//       P5  dyn_in1 of the next step as a gather-sum of the sampled one-hot rows       (rssm.py:47)      16 tiles
//     The first half of the NEXT step's hidden layer (deter + x0 columns) runs in the P4 / P5 slot of the CTAs that do not
//     sample; for the 32 sampling CTAs it is computed by helper CTAs 48..79 (hid_first_half_pair), so that the step's
//     critical path is P3 -> P4 -> P5 -> P1 only.
//   * everything that does not depend on the recurrence is computed before the scan: the embed part of obs_net_0 for
//     all (b, t) (one batched GEMM) and x2 = SiLU(RMSNorm(dyn_in2(action_t))) (obs_prep_kernel).
// All arithmetic is fp32 FMA with fixed reduction orders (k-group tree inside a warp, warps in order, cluster ranks
// in order): deterministic, and within the parity tolerances of the layer-by-layer path.  With a tape (stride 1)
// the per-step buffers are the backward tape itself, so sd_observe_bwd consumes the result unchanged.
#pragma once
#include "sd_kernels.cuh"

namespace sd {
namespace scan {

constexpr int THREADS = 512;          // 16 warps: four per scheduler (every phase is latency bound)
constexpr int NCTA = 128;
constexpr int CLUSTER = 4;
constexpr int HW = 256;            // U == Dg == 256 (checked on the host)
constexpr int KC = 512;            // K elements staged at once
constexpr int ALD = KC + 4;        // padded row stride of the staged activations
// shared memory carve-up (floats)
constexpr int kW1 = 0;                         // P1: [1024][16]
constexpr int kW2 = kW1 + 1024 * 16;           // P2: 3 x [256][16]
constexpr int kW3 = kW2 + 3 * 256 * 16;        // P3: [512][16]
constexpr int kW45 = kW3 + 512 * 16;           // P4: [256][16] (CTAs < SK/16) | P5: [SK][16] (CTAs 32..47)
constexpr int kAs = kW45 + 512 * 16;           // [16][ALD]; after the FMA loop of a tile it is the [16 warps][256] reduction buffer
constexpr int kSlots = kAs + 16 * ALD;         // [4 ranks][256] (cluster leader)
constexpr int kGain = kSlots + CLUSTER * 256;  // RMS scales: g_in0 | g_in1 | g_hid[block] | g_obs (4 x 256)
constexpr int kSmemFloats = kGain + 4 * 256;
constexpr int kSmemBytes = kSmemFloats * 4;

struct Params {
  int B, T, D, SK, S, K, G, E, A;
  float unimix;
  // packed fp32 weights ([K][ldw], n contiguous), biases, RMS scales
  const float *w_in0, *b_in0, *g_in0; int ld_in0;
  const float *w_in1, *b_in1, *g_in1; int ld_in1;
  const float *w_hid, *b_hid, *g_hid; int ld_hid;   // [G][Dg + 3U][ld]
  const float *w_gru, *b_gru; int ld_gru;           // [G][Dg][ld]
  const float *w_obs, *b_obs, *g_obs; int ld_obs;   // [D + E][ld] (rows < D are used here)
  const float *w_lg, *b_lg; int ld_lg;              // [U][ld]
  // inputs
  const float *init_stoch, *init_deter;
  const uint8_t* is_first;   // (B, T)
  const float* u;            // (B, T, SK)
  const float* eproj;        // (B*T, U), rows (b, t): embed part of obs_net_0, no bias
  const float* x2;           // (T*B, U) step-major
  // outputs
  float *stochs, *deters, *logits;   // (B, T, .)
  // per-step buffers: step t lives at base + t * step * B * width (step = 0: reused, 1: backward tape)
  float *zin, *din, *vin, *x, *hpre, *h, *q, *lg, *vobs, *o;
  int step;
  int ll;              // 0: five grid barriers per step; 1: flagged hand-offs instead of the barriers after P3 / P4 / P5; 2: + helper CTAs (SD_SCAN_LL)
  float* ssq_h;        // [128 tiles][16 rows]
  unsigned int* idx;   // [16][S]: (tag << 8) | class index of the step's sample
  float2* ll_sa;       // [32 sampling CTAs][256]: first-half hidden-layer tiles computed for them by the helper CTAs 48..79
  float2 *ll_x0, *ll_vobs, *ll_x1;   // [16][256] {value, tag} pairs: flagged hand-offs P3 -> next hidden layer, P3 -> P4, P5 -> P1
  unsigned int* bar;   // grid barrier counter, zeroed before the launch
  long long* timing;   // diagnostic (SD_TRACE_SCAN=1): clock64 stamps of CTA 0 / 40 during step 2; null in production
};
#define SD_SC_KSTAMP(i) do { if (P.timing && threadIdx.x == 0 && blockIdx.x == 0) P.timing[32 + (i)] = clock64(); } while (0)
#define SD_SC_STAMP(i) do { if (P.timing && t == 2 && tid == 0 && (cta == 0 || cta == 40)) P.timing[(cta ? 16 : 0) + (i)] = clock64(); } while (0)

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Grid barrier: monotonically increasing counter; bounded spin (a protocol bug traps instead of hanging the GPU).
// Arrive = red.release (orders this CTA's earlier writes, which thread 0 observed through bar.sync); the poll is a
// RELAXED load (an acquire load would invalidate L1 on every iteration) followed by one acquire fence.
// (A hierarchical variant -- cluster barrier, one atomic per cluster, cluster barrier -- measured slower: 3.3k vs
//  2.7k cycles per barrier.)
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned int target = epoch * (unsigned int)NCTA;
    unsigned int v = 0;
#pragma unroll 1
    for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
    }
    if (v < target) __trap();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

// Flagged hand-offs ("LL": value and tag travel in ONE 8-byte store, so the consumer may poll the data itself -- no fence,
// no counter).  Three of the five phase boundaries of a step have few producers (the 16 P3 cluster leaders, the 32 sampling
// CTAs, the 16 P5 CTAs): there a grid barrier costs a store-ack wait + an atomic + a poll (~2.7k cycles); polling the
// tagged data is one L2 hop behind the producer's store.  Tag = step + 1 of the producing step; the host zeroes the
// buffers before every launch (also inside a replayed graph), so a stale tag can never match.  Re-use is safe because the
// two remaining grid barriers (after P1 and P2) of the next step lie between any consumer's read and the producer's next write.
__device__ __forceinline__ void ll_store(float2* p, float v, unsigned int tag) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// One lane per warp waits for the tag of ONE element before the whole warp polls its data: 80+ CTAs spinning with every
// thread on the same 32 KB (measured) saturate the L2 slices that hold it and slow the producers down.
__device__ __forceinline__ void ll_prewait(const void* tagword, unsigned int tag, bool lane_polls) {
  if (lane_polls) {
    unsigned int v = 0;
#pragma unroll 1
    for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(tagword) : "memory");
      if (v == tag) break;
    }
    if (v != tag) __trap();
  }
  __syncwarp();
}
// 4 consecutive values (p is 32-byte aligned: the lane's float4 of the plain layout); bounded spin, trap on a protocol bug
__device__ __forceinline__ float4 ll_load4(const float2* p, unsigned int tag) {
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  bool ok = false;
#pragma unroll 1
  for (unsigned int spin = 0; spin < (1u << 22); ++spin) {
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p) : "memory");
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p + 2) : "memory");
    ok = (a1 == tag) && (a3 == tag) && (b1 == tag) && (b3 == tag);
    if (ok) break;
  }
  if (!ok) __trap();
  return make_float4(__uint_as_float(a0), __uint_as_float(a2), __uint_as_float(b0), __uint_as_float(b2));
}

// copy W[k0 .. k0+K)[col0 .. col0+16) (global, row stride ld) into the swizzled k-quad-major layout:
// element (k, n) -> dst[(k/4)*64 + ((k%4) ^ ((k/4)&1))*16 + n]  (conflict-free LDS.128 for two adjacent k-quads)
__device__ __forceinline__ void stage_w(float* dst, const float* w, int ld, int k0, int K, int col0, bool swz) {
  for (int i = threadIdx.x; i < K * 4; i += THREADS) {
    const int k = i >> 2, n4 = i & 3;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(k0 + k) * ld + col0 + n4 * 4));
    const int kq = k >> 2, j = k & 3;
    const int jj = swz ? (j ^ (kq & 1)) : j;
    *reinterpret_cast<float4*>(dst + kq * 64 + jj * 16 + n4 * 4) = v;
  }
}

// NG 16x16 output tiles that share the A operand: s_g(row, col) = sum_k A_s[row][k] * W_g[k][col], W_g = W + g*wstride,
// over the nkq staged k-quads (nkq*4 = K elements, a multiple of 128).
// The 16 batch rows are exactly the M of mma.sync.m16n8k8, so the contraction runs on the tensor pipe with the
// 3xTF32 split (x = hi + lo, hi = tf32(x), lo = tf32(x - hi); hi*hi + hi*lo + lo*hi in fp32 accumulators): the dropped
// lo*lo term is 2^-22 relative, i.e. fp32-class accuracy -- the SIMT version of this tile was shared-memory bandwidth
// bound (2048 LDS wavefronts per K=512, 3.3k cycles; this one needs ~1/8 of the shared-memory traffic).
// Warp w owns the k-range [w*K/16, (w+1)*K/16); the 16 per-warp partial tiles are summed in warp order through shared
// memory (fixed order => deterministic).  The reduction buffer ALIASES the A tile (every thread is past its last A
// read at the first barrier); callers must barrier before they overwrite A_s again.
// out[g] = element (row = tid >> 4, col = tid & 15) of tile g, valid for tid < 256.  Deliberately not inlined.
template <int NG>
__device__ __noinline__ void tile_product(float* A_s, const float* W, int wstride, int nkq, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;   // mma fragment coordinates: groupID, threadID_in_group
  float acc[NG][2][4];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[g][nt][c] = 0.f;
  const int ksteps = nkq / (2 * (THREADS / 32));   // k-steps of 8 per warp
#pragma unroll 1
  for (int s = 0; s < ksteps; ++s) {
    const int k0 = (warp * ksteps + s) * 8;
    // A fragment: a0 (gq, tq), a1 (gq+8, tq), a2 (gq, tq+4), a3 (gq+8, tq+4)
    uint32_t ah[4], al[4];
    split_tf32(A_s[gq * ALD + k0 + tq], ah[0], al[0]);
    split_tf32(A_s[(gq + 8) * ALD + k0 + tq], ah[1], al[1]);
    split_tf32(A_s[gq * ALD + k0 + tq + 4], ah[2], al[2]);
    split_tf32(A_s[(gq + 8) * ALD + k0 + tq + 4], ah[3], al[3]);
    // B fragment of n-tile nt: b0 (k = tq, n = gq), b1 (k = tq + 4, n = gq); W is [k/4][k%4][16]
    const float* w0 = W + ((k0 >> 2) * 64) + tq * 16 + gq;   // k0 + tq      -> k-quad k0/4, j = tq
    const float* w1 = w0 + 64;                                // k0 + tq + 4  -> next k-quad
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(w0[g * wstride + nt * 8], bh0, bl0);
        split_tf32(w1[g * wstride + nt * 8], bh1, bl1);
        mma_tf32(acc[g][nt], al, bh0, bh1);   // small terms first
        mma_tf32(acc[g][nt], ah, bl0, bl1);
        mma_tf32(acc[g][nt], ah, bh0, bh1);
      }
  }
  float* red = A_s;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    __syncthreads();   // every thread is done reading A_s / the previous reduction
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      // C fragment: c0 (gq, 2tq), c1 (gq, 2tq+1), c2 (gq+8, 2tq), c3 (gq+8, 2tq+1)
      float* r0 = red + warp * 256 + gq * 16 + nt * 8 + 2 * tq;
      *reinterpret_cast<float2*>(r0) = make_float2(acc[g][nt][0], acc[g][nt][1]);
      *reinterpret_cast<float2*>(r0 + 8 * 16) = make_float2(acc[g][nt][2], acc[g][nt][3]);
    }
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x < 256) {
#pragma unroll
      for (int w16 = 0; w16 < THREADS / 32; ++w16) s += red[w16 * 256 + threadIdx.x];
    }
    out[g] = s;
  }
}
__device__ __forceinline__ float sum16(float v) {   // over the 16 lanes that share a tile row
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float sq4(const float4& v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }

// y = SiLU((v * rs) * g) on 4 consecutive columns (same op order as normact_kernel); not inlined (code size)
__device__ __noinline__ float4 normact4(float4 v, float rs, float4 gg) {
  float4 y;
  y.x = siluf_((v.x * rs) * gg.x); y.y = siluf_((v.y * rs) * gg.y);
  y.z = siluf_((v.z * rs) * gg.z); y.w = siluf_((v.w * rs) * gg.w);
  return y;
}
// Load mapping: one warp per batch row (lrow = tid >> 5); lane ls owns columns 128*i + 4*ls, i = 0, 1 of a 256-wide row
// (512 B coalesced loads, conflict-free float4 stores into the A tile).
__device__ __forceinline__ void normact8(const float4 (&v)[2], float rs, const float* g_s, int ls, float4 (&y)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) y[i] = normact4(v[i], rs, *reinterpret_cast<const float4*>(g_s + i * 128 + ls * 4));
}

// First half of the block-GRU hidden layer of step tt: s_a = W_hid[g][:, 0:512] . [keep * d_g | x0], x0 = SiLU(RMSNorm(v_in0)).
// Neither operand depends on the sample of the previous step (v_in0 is produced two phases earlier, in P3), so this
// runs in the otherwise idle P4 / P5 slot and only its 16x16 partial tile (one float per thread) is carried into P1.
// v0 comes from the plain buffer (step 0: written by the host-side launches) or, when v0ll is set, from the flagged hand-off.
__device__ __noinline__ float hid_first_half(const float* dsrc, float keep, const float* v0src, const float2* v0ll, unsigned int tag,
                                             bool lok, float* A_s, const float* W1, const float* G_s, int lrow, int ls,
                                             float* din_t, float* x_t) {
  float4 dv[2], v0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = i * 128 + ls * 4;
    dv[i] = lok ? ldcg4(dsrc + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!v0ll) v0[i] = lok ? ldcg4(v0src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (v0ll) {
    ll_prewait(reinterpret_cast<const unsigned int*>(v0ll) + 1, tag, lok && ls == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) v0[i] = lok ? ll_load4(v0ll + i * 128 + ls * 4, tag) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float ss0 = warp_sum(sq4(v0[0]) + sq4(v0[1]));
  const float rs0 = 1.f / sqrtf(ss0 / (float)HW + kRmsEps);
  float4 x0[2];
  normact8(v0, rs0, G_s, ls, x0);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    dv[i].x *= keep; dv[i].y *= keep; dv[i].z *= keep; dv[i].w *= keep;
    *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = dv[i];
    *reinterpret_cast<float4*>(A_s + lrow * ALD + HW + i * 128 + ls * 4) = x0[i];
    if (din_t) *reinterpret_cast<float4*>(din_t + i * 128 + ls * 4) = dv[i];   // backward tape: masked deter input
    if (x_t) *reinterpret_cast<float4*>(x_t + i * 128 + ls * 4) = x0[i];       // backward tape: x0
  }
  __syncthreads();
  float s;
  tile_product<1>(A_s, W1, 0, KC / 4, &s);
  return s;
}

// Helper CTAs (48 .. 48 + SK/16): their own first half, then the first half of sampling CTA `cta - 48` with that CTA's weight
// slice (kept in the helper's otherwise unused W45 region).  The sampling CTAs run P4 right after P3, so their own
// hid_first_half would sit on the step's critical path (P3 -> P4 -> hid_first_half -> P1: 16 k cycles, everyone else is
// done after 12 k); the helpers have that slack.  Same operands, same tile_product: bit-identical to the CTA's own result.
__device__ __noinline__ float hid_first_half_pair(const float* dsrc, const float* dsrc_p, float keep, const float2* v0ll, unsigned int tag,
                                                  bool lok, bool rok, float* A_s, const float* W1, const float* Wp, const float* G_s,
                                                  int lrow, int ls, float* din_t, float* din_p, float* x_p, float2* sa_ll) {
  float4 dv[2], dp[2], v0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = i * 128 + ls * 4;
    dv[i] = lok ? ldcg4(dsrc + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    dp[i] = lok ? ldcg4(dsrc_p + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ll_prewait(reinterpret_cast<const unsigned int*>(v0ll) + 1, tag, lok && ls == 0);
#pragma unroll
  for (int i = 0; i < 2; ++i) v0[i] = lok ? ll_load4(v0ll + i * 128 + ls * 4, tag) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float ss0 = warp_sum(sq4(v0[0]) + sq4(v0[1]));
  const float rs0 = 1.f / sqrtf(ss0 / (float)HW + kRmsEps);
  float4 x0[2];
  normact8(v0, rs0, G_s, ls, x0);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    dv[i].x *= keep; dv[i].y *= keep; dv[i].z *= keep; dv[i].w *= keep;
    dp[i].x *= keep; dp[i].y *= keep; dp[i].z *= keep; dp[i].w *= keep;
    *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = dv[i];
    *reinterpret_cast<float4*>(A_s + lrow * ALD + HW + i * 128 + ls * 4) = x0[i];
    if (din_t) *reinterpret_cast<float4*>(din_t + i * 128 + ls * 4) = dv[i];   // backward tape: masked deter input (own block)
    if (din_p) *reinterpret_cast<float4*>(din_p + i * 128 + ls * 4) = dp[i];   // ... and the sampling CTA's block
    if (x_p) *reinterpret_cast<float4*>(x_p + i * 128 + ls * 4) = x0[i];       // backward tape: x0 (written by CTA 0 otherwise)
  }
  __syncthreads();
  float s;
  tile_product<1>(A_s, W1, 0, KC / 4, &s);
  __syncthreads();   // the reduction buffer aliases the A tile: restage both halves
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = dp[i];
    *reinterpret_cast<float4*>(A_s + lrow * ALD + HW + i * 128 + ls * 4) = x0[i];
  }
  __syncthreads();
  float sp;
  tile_product<1>(A_s, Wp, 0, KC / 4, &sp);
  if (rok) ll_store(sa_ll + threadIdx.x, sp, tag);
  return s;
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(THREADS, 1) observe_scan_kernel(const Params P) {
  extern __shared__ __align__(16) float sm[];
  float* W1 = sm + kW1;
  float* W2 = sm + kW2;
  float* W3 = sm + kW3;
  float* W45 = sm + kW45;
  float* A_s = sm + kAs;
  float* slots = sm + kSlots;
  float* G_s = sm + kGain;
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int B = P.B, T = P.T, D = P.D, SK = P.SK;
  // load mapping (all 512 threads): one warp per batch row
  const int lrow = tid >> 5, ls = tid & 31;
  const bool lok = lrow < B;
  // tile mapping (threads < 256): output element (row, col) of the CTA's 16x16 tile
  const bool et = tid < 256;
  const int row = (tid >> 4) & 15, col = tid & 15;
  const bool rok = et && row < B;
  const int g = cta >> 4, jt = cta & 15;                       // P1 / P2: block and 16-column (16-unit) tile
  const int p3 = cta >> 6, j3 = (cta & 63) >> 2, r3 = cta & 3; // P3: problem, column tile, k-slice (= cluster rank)
  const int n4tiles = SK / 16;
  const bool do4 = cta < n4tiles, do5 = cta >= 32 && cta < 48;
  const int j5 = cta - 32;
  // first-half hidden-layer tiles of the sampling CTAs are computed by helper CTAs (flagged hand-off mode only)
  const bool offload = P.ll >= 2;
  const bool helper = offload && cta >= 48 && cta - 48 < n4tiles;
  const int pc = cta - 48, gp = (pc >> 4) & 7, jtp = pc & 15;   // the helper's sampling CTA, its block and column tile
  const size_t sstep = (size_t)P.step * B;   // rows between consecutive steps of the per-step buffers

  SD_SC_KSTAMP(0);
  // ---------------------------------------------------------------- one-time: weights -> shared memory
  stage_w(W1, P.w_hid + (size_t)g * (4 * HW) * P.ld_hid, P.ld_hid, 0, 4 * HW, jt * 16, false);
  for (int gate = 0; gate < 3; ++gate)
    stage_w(W2 + gate * HW * 16, P.w_gru + (size_t)g * HW * P.ld_gru, P.ld_gru, 0, HW, gate * HW + jt * 16, false);
  stage_w(W3, p3 == 0 ? P.w_in0 : P.w_obs, p3 == 0 ? P.ld_in0 : P.ld_obs, r3 * KC, KC, j3 * 16, false);
  if (do4) stage_w(W45, P.w_lg, P.ld_lg, 0, HW, cta * 16, false);
  if (do5) stage_w(W45, P.w_in1, P.ld_in1, 0, SK, j5 * 16, false);
  if (helper) stage_w(W45, P.w_hid + (size_t)gp * (4 * HW) * P.ld_hid, P.ld_hid, 0, 2 * HW, jtp * 16, false);
  // RMS scales -> shared memory, per-thread biases -> registers (the acquire of every grid barrier invalidates L1, so
  // anything re-read from global each phase would pay an L2 round trip on the critical path)
  if (et) {
    G_s[tid] = __ldg(P.g_in0 + tid);
    G_s[256 + tid] = __ldg(P.g_in1 + tid);
    G_s[512 + tid] = __ldg(P.g_hid + g * HW + tid);
    G_s[768 + tid] = __ldg(P.g_obs + tid);
  }
  const float bias_h = __ldg(P.b_hid + g * HW + jt * 16 + col);
  const float bias_qr = __ldg(P.b_gru + (size_t)g * 3 * HW + jt * 16 + col);
  const float bias_qc = __ldg(P.b_gru + (size_t)g * 3 * HW + HW + jt * 16 + col);
  const float bias_qu = __ldg(P.b_gru + (size_t)g * 3 * HW + 2 * HW + jt * 16 + col);
  const float bias_3 = __ldg((p3 == 0 ? P.b_in0 : P.b_obs) + j3 * 16 + col);
  const float bias_lg = do4 ? __ldg(P.b_lg + cta * 16 + col) : 0.f;
  const float bias_5 = do5 ? __ldg(P.b_in1 + j5 * 16 + col) : 0.f;
  unsigned int epoch = 0;
  __syncthreads();
  SD_SC_KSTAMP(1);
  // first half of step 0's hidden layer (v_in0 of step 0 comes from the host-side launches)
  float s_a = hid_first_half(P.init_deter + (size_t)lrow * D + g * HW, (lok && P.is_first[(size_t)lrow * T]) ? 0.f : 1.f,
                             P.vin + (size_t)lrow * (3 * HW), nullptr, 0u, lok, A_s, W1, G_s, lrow, ls,
                             (P.step && lok && jt == 0) ? P.din + (size_t)lrow * D + g * HW : nullptr,
                             (P.step && lok && cta == 0) ? P.x + (size_t)lrow * (3 * HW) : nullptr);

  SD_SC_KSTAMP(2);
  for (int t = 0; t < T; ++t) {
    if (t == 1) SD_SC_KSTAMP(3);
    if (t == T / 2) SD_SC_KSTAMP(4);
    // reset masks (rssm.py:161-165) in both mappings
    const float keep_t = (rok && P.is_first[(size_t)row * T + t]) ? 0.f : 1.f;
    const float keep_n = (t + 1 < T && rok && P.is_first[(size_t)row * T + t + 1]) ? 0.f : 1.f;
    const float lkeep_n = (t + 1 < T && lok && P.is_first[(size_t)lrow * T + t + 1]) ? 0.f : 1.f;
    float* vin_t = P.vin + (size_t)t * sstep * (3 * HW);
    const unsigned int tag = (unsigned int)t + 1u;   // of everything this step hands over through the flagged buffers
    float* hpre_t = P.hpre + (size_t)t * sstep * D;
    float* vobs_t = P.vobs + (size_t)t * sstep * HW;
    // pure inputs of this step, fetched now so that their latency is hidden behind P1 / P2
    const float ep_t = (r3 == 0 && p3 == 1 && rok) ? __ldg(P.eproj + ((size_t)row * T + t) * HW + j3 * 16 + col) : 0.f;
    const float uu_t = (do4 && rok) ? __ldg(P.u + ((size_t)row * T + t) * SK + cta * 16 + col) : 0.5f;

    SD_SC_STAMP(0);
    // ================================================================ P1: hidden layer of the block GRU (second half)
    {
      float4 v1[2], xv2[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = i * 128 + ls * 4;
        if (t == 0 || !P.ll) v1[i] = lok ? ldcg4(vin_t + (size_t)lrow * (3 * HW) + HW + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv2[i] = lok ? __ldg(reinterpret_cast<const float4*>(P.x2 + ((size_t)t * B + lrow) * HW + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (t > 0 && P.ll) {   // produced by P5 of the previous step (tag t)
        ll_prewait(reinterpret_cast<const unsigned int*>(P.ll_x1 + lrow * HW) + 1, tag - 1u, lok && ls == 0);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          v1[i] = lok ? ll_load4(P.ll_x1 + lrow * HW + i * 128 + ls * 4, tag - 1u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float ss1 = warp_sum(sq4(v1[0]) + sq4(v1[1]));
      const float rs1 = 1.f / sqrtf(ss1 / (float)HW + kRmsEps);
      float4 x1[2];
      normact8(v1, rs1, G_s + 256, ls, x1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = x1[i];
        *reinterpret_cast<float4*>(A_s + lrow * ALD + HW + i * 128 + ls * 4) = xv2[i];
      }
      if (P.step && lok && cta == 0) {   // backward tape: x1 (x0 by hid_first_half, x2 by obs_prep_kernel)
        float* xt = P.x + ((size_t)t * sstep + lrow) * (3 * HW) + HW + ls * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(xt + i * 128) = x1[i];
      }
      __syncthreads();
      float s2;
      tile_product<1>(A_s, W1 + (KC / 4) * 64, 0, KC / 4, &s2);
      if (do4 && offload && t > 0 && et) {   // this CTA's first half was computed by helper CTA 48 + cta during the previous step
        uint32_t v0 = 0, v1 = 0;
        ll_prewait(reinterpret_cast<const unsigned int*>(P.ll_sa + cta * 256 + tid) + 1, tag - 1u, rok && (tid & 31) == 0);
        if (rok) {
          const float2* src = P.ll_sa + cta * 256 + tid;
#pragma unroll 1
          for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
            asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "l"(src) : "memory");
            if (v1 == tag - 1u) break;
          }
          if (v1 != tag - 1u) __trap();
        }
        s_a = __uint_as_float(v0);
      }
      if (et) {
        const int n = g * HW + jt * 16 + col;
        const float hp = (s_a + s2) + bias_h;
        if (rok) hpre_t[(size_t)row * D + n] = hp;
        const float ssr = sum16(hp * hp);
        if (col == 0) P.ssq_h[row * NCTA + cta] = ssr;
      }
    }
    SD_SC_STAMP(1);
    grid_sync(P.bar, epoch);
    SD_SC_STAMP(2);

    // ================================================================ P2: gate projection + GRU gates
    {
      float4 hv[2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        hv[i] = lok ? ldcg4(hpre_t + (size_t)lrow * D + g * HW + i * 128 + ls * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 pq = ldcg4(P.ssq_h + lrow * NCTA + ls * 4);   // 128 tile partials per row, 4 per lane
      const float tot = warp_sum((pq.x + pq.y) + (pq.z + pq.w));
      const float rs = 1.f / sqrtf(tot / (float)D + kRmsEps);
      float4 hh[2];
      normact8(hv, rs, G_s + 512, ls, hh);
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = hh[i];
      if (P.step && lok && jt == 0) {
        float* ht = P.h + ((size_t)t * sstep + lrow) * D + g * HW + ls * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(ht + i * 128) = hh[i];
      }
      const int n = g * HW + jt * 16 + col;   // unit
      const float* dsrc = t == 0 ? P.init_deter + (size_t)row * D : P.deters + ((size_t)row * T + (t - 1)) * D;
      const float dprev = rok ? keep_t * ldcg(dsrc + n) : 0.f;
      __syncthreads();
      float q3[3];
      tile_product<3>(A_s, W2, HW * 16, HW / 4, q3);   // reset | cand | update tiles share the A operand
      if (rok) {
        const float qr = q3[0] + bias_qr;
        const float qc = q3[1] + bias_qc;
        const float qu = q3[2] + bias_qu;
        if (P.step) {
          float* qt = P.q + ((size_t)t * sstep + row) * (3 * D) + (size_t)g * 3 * HW + jt * 16 + col;
          qt[0] = qr; qt[HW] = qc; qt[2 * HW] = qu;
        }
        const float reset = sigmoidf_(qr);
        const float cand = tanhf(reset * qc);
        const float upd = sigmoidf_(qu - 1.f);
        P.deters[((size_t)row * T + t) * D + n] = upd * cand + (1.f - upd) * dprev;
      }
    }
    SD_SC_STAMP(3);
    grid_sync(P.bar, epoch);
    SD_SC_STAMP(4);

    // ================================================================ P3: obs_net_0 (deter part) and next step's dyn_in0
    {
      const float* dsrc = P.deters + ((size_t)lrow * T + t) * D + r3 * KC + ls * 4;
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = lok ? ldcg4(dsrc + i * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = a[i];
      __syncthreads();
      SD_SC_STAMP(11);
      float s;
      tile_product<1>(A_s, W3, 0, KC / 4, &s);
      SD_SC_STAMP(12);
      if (et) {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(slots + r3 * 256 + tid);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
      }
      cluster_sync_all();
      SD_SC_STAMP(13);
      if (r3 == 0 && rok) {
        const int n = j3 * 16 + col;
        const float extra = bias_3 + ep_t;   // bias (+ embed part of obs_net_0)
        const float tot = ((slots[tid] + slots[256 + tid]) + slots[512 + tid]) + slots[768 + tid];
        if (p3 == 0) {
          if (t + 1 < T) {
            const float v = extra + keep_n * tot;
            ll_store(P.ll_x0 + row * HW + n, v, tag);
            P.vin[(size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW) + n] = v;   // backward tape
          }
        } else {
          ll_store(P.ll_vobs + row * HW + n, tot + extra, tag);
          vobs_t[(size_t)row * HW + n] = tot + extra;   // backward tape
        }
      }
    }
    SD_SC_STAMP(5);
    if (!P.ll) grid_sync(P.bar, epoch);
    SD_SC_STAMP(6);   // (no grid barrier: P4 and the next hidden layer poll the flagged hand-offs of the P3 leaders)

    // ================================================================ P4: logits + sample (other CTAs: first half of the next hidden layer)
    if (helper && t + 1 < T)
      s_a = hid_first_half_pair(P.deters + ((size_t)lrow * T + t) * D + g * HW, P.deters + ((size_t)lrow * T + t) * D + gp * HW, lkeep_n,
                                P.ll_x0 + lrow * HW, tag, lok, rok, A_s, W1, W45, G_s, lrow, ls,
                                (P.step && lok && jt == 0) ? P.din + ((size_t)(t + 1) * sstep + lrow) * D + g * HW : nullptr,
                                (P.step && lok && jtp == 0) ? P.din + ((size_t)(t + 1) * sstep + lrow) * D + gp * HW : nullptr,
                                (P.step && lok && pc == 0) ? P.x + ((size_t)(t + 1) * sstep + lrow) * (3 * HW) : nullptr,
                                P.ll_sa + pc * 256);
    else if (!do4 && t + 1 < T)
      s_a = hid_first_half(P.deters + ((size_t)lrow * T + t) * D + g * HW, lkeep_n,
                           P.vin + (size_t)(t + 1) * sstep * (3 * HW) + (size_t)lrow * (3 * HW), P.ll ? P.ll_x0 + lrow * HW : nullptr, tag,
                           lok, A_s, W1, G_s, lrow, ls,
                           (P.step && lok && jt == 0) ? P.din + ((size_t)(t + 1) * sstep + lrow) * D + g * HW : nullptr,
                           (P.step && lok && cta == 0) ? P.x + ((size_t)(t + 1) * sstep + lrow) * (3 * HW) : nullptr);
    if (do4) {
      float4 vv[2];
      if (P.ll) {
        ll_prewait(reinterpret_cast<const unsigned int*>(P.ll_vobs + lrow * HW) + 1, tag, lok && ls == 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) vv[i] = lok ? ll_load4(P.ll_vobs + lrow * HW + i * 128 + ls * 4, tag) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) vv[i] = lok ? ldcg4(vobs_t + (size_t)lrow * HW + i * 128 + ls * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float ss = warp_sum(sq4(vv[0]) + sq4(vv[1]));
      const float rs = 1.f / sqrtf(ss / (float)HW + kRmsEps);
      float4 oo[2];
      normact8(vv, rs, G_s + 768, ls, oo);
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(A_s + lrow * ALD + i * 128 + ls * 4) = oo[i];
      if (P.step && lok && cta == 0) {
        float* ot = P.o + ((size_t)t * sstep + lrow) * HW + ls * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(ot + i * 128) = oo[i];
      }
      __syncthreads();
      float lgv;
      tile_product<1>(A_s, W45, 0, HW / 4, &lgv);
      if (et) {   // warps 0-7: one thread per (row, class)
        lgv += bias_lg;
        const int n = cta * 16 + col;
        const int Kc = P.K, kcls = col % Kc;
        int best;
        if (Kc == 16) best = sample_group<16>(lgv, uu_t, true, kcls, Kc, P.unimix, nullptr);
        else if (Kc == 8) best = sample_group<8>(lgv, uu_t, true, kcls, Kc, P.unimix, nullptr);
        else if (Kc == 4) best = sample_group<4>(lgv, uu_t, true, kcls, Kc, P.unimix, nullptr);
        else best = sample_group<2>(lgv, uu_t, true, kcls, Kc, P.unimix, nullptr);
        if (rok) {
          if (kcls == 0)   // the hand-off to P5 first: the output / tape stores below are off the critical path
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(P.idx + row * P.S + n / Kc), "r"((tag << 8) | (unsigned int)best) : "memory");
          const float oh = (kcls == best) ? 1.f : 0.f;
          const size_t off = ((size_t)row * T + t) * SK + n;
          P.stochs[off] = oh;
          P.logits[off] = lgv;
          if (P.step) {
            P.lg[((size_t)t * sstep + row) * SK + n] = lgv;
            if (t + 1 < T) P.zin[((size_t)(t + 1) * sstep + row) * SK + n] = keep_n * oh;
          }
        }
      }
      __syncthreads();   // the reduction buffer of the logit tile aliases A_s, which hid_first_half overwrites next
    }
    SD_SC_STAMP(7);
    if (!P.ll) grid_sync(P.bar, epoch);
    SD_SC_STAMP(8);   // (no grid barrier: the P5 CTAs poll the tagged indices)

    // ================================================================ P5: next step's dyn_in1 (gather-sum of one-hot rows)
    if (do4 && !offload && t + 1 < T)
      s_a = hid_first_half(P.deters + ((size_t)lrow * T + t) * D + g * HW, lkeep_n,
                           P.vin + (size_t)(t + 1) * sstep * (3 * HW) + (size_t)lrow * (3 * HW), P.ll ? P.ll_x0 + lrow * HW : nullptr, tag,
                           lok, A_s, W1, G_s, lrow, ls,
                           (P.step && lok && jt == 0) ? P.din + ((size_t)(t + 1) * sstep + lrow) * D + g * HW : nullptr,
                           (P.step && lok && cta == 0) ? P.x + ((size_t)(t + 1) * sstep + lrow) * (3 * HW) : nullptr);
    if (do5 && t + 1 < T && et) {   // whole warps (et = tid < 256)
      {   // pre-wait: one lane per row spins on the row's first index word
        unsigned int v = 0;
        if (P.ll && P.S != 32 && rok && col == 0) {
#pragma unroll 1
          for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.idx + row * P.S) : "memory");
            if ((v >> 8) == tag) break;
          }
          if ((v >> 8) != tag) __trap();
        }
        __syncwarp();
      }
      SD_SC_STAMP(14);
      if (P.S == 32 && P.ll) {
        // every thread polls TWO of its row's 32 index words (one L2 round trip), the 16 lanes of the row exchange them by shuffle
        const int lane = tid & 31;
        unsigned int w0 = tag << 8, w1 = tag << 8;
        if (rok) {
          const unsigned int* src = P.idx + row * 32 + col * 2;
          bool ok = false;
#pragma unroll 1
          for (unsigned int spin = 0; spin < (1u << 22) && !ok; ++spin) {
            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "l"(src) : "memory");
            ok = ((w0 >> 8) == tag) && ((w1 >> 8) == tag);
          }
          if (!ok) __trap();
        }
        __syncwarp();
        SD_SC_STAMP(15);
        float v = 0.f;
        const int Kc = P.K;
#pragma unroll
        for (int sI = 0; sI < 32; ++sI) {   // ascending s: same summation order as the general path
          const unsigned int wsel = __shfl_sync(0xffffffffu, (sI & 1) ? w1 : w0, (lane & 16) | (sI >> 1));
          v += W45[(sI * Kc + (int)(wsel & 0xffu)) * 16 + col];
        }
        if (rok) {
          const int n = j5 * 16 + col;
          const float v1n = bias_5 + keep_n * v;
          ll_store(P.ll_x1 + row * HW + n, v1n, tag);
          P.vin[(size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW) + HW + n] = v1n;   // backward tape
        }
      } else if (rok) {
        float v = 0.f;
        const int Kc = P.K;
        for (int s0 = 0; s0 < P.S; s0 += 32) {   // all loads of a chunk in flight together; re-poll the chunk until every tag matches
          unsigned int w[32];
          bool ok = false;
#pragma unroll 1
          for (unsigned int spin = 0; spin < (P.ll ? (1u << 22) : 1u) && !ok; ++spin) {
            ok = true;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              w[j] = tag << 8;
              if (s0 + j < P.S) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w[j]) : "l"(P.idx + row * P.S + s0 + j) : "memory");
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) ok = ok && ((w[j] >> 8) == tag);
          }
          if (!ok && P.ll) __trap();
          SD_SC_STAMP(15);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (s0 + j < P.S) v += W45[((s0 + j) * Kc + (int)(w[j] & 0xffu)) * 16 + col];
        }
        const int n = j5 * 16 + col;
        const float v1n = bias_5 + keep_n * v;
        ll_store(P.ll_x1 + row * HW + n, v1n, tag);
        P.vin[(size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW) + HW + n] = v1n;   // backward tape
      }
    }
    SD_SC_STAMP(9);
    if (!P.ll && t + 1 < T) grid_sync(P.bar, epoch);
    __syncthreads();   // (no grid barrier: P1 of the next step polls the flagged x1; A_s is free once every thread is past hid_first_half)
    SD_SC_STAMP(10);
  }
  SD_SC_KSTAMP(5);
}

// Everything of the posterior scan that only depends on the inputs (rssm.py:44,48,161-165): per (t, b)
//   keep = !is_first, abar = keep * a / max(|a|, 1), v2 = W_in2 abar + b, x2 = SiLU(RMSNorm(v2) * g).
// One warp per row; U <= 256.  Tape outputs (nullable): ain (T*B, A), keep (T*B), vin[:, 2U:3U], x[:, 2U:3U].
__global__ void obs_prep_kernel(const float* __restrict__ action, const uint8_t* __restrict__ is_first, int B, int T, int A,
                                int U, const float* __restrict__ w2_t, int ldw, const float* __restrict__ b2,
                                const float* __restrict__ g2, float* x2_all, float* ain, float* keep, float* vin, float* x) {
  pdl_prologue();
  const int wrow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wrow >= B * T) return;
  const int t = wrow / B, b = wrow - t * B;
  const float kp = is_first[(size_t)b * T + t] ? 0.f : 1.f;
  float ab = 0.f;
  if (lane < A) {
    const float v = kp * action[((size_t)b * T + t) * A + lane];
    ab = v / fmaxf(fabsf(v), 1.f);
    if (ain) ain[(size_t)wrow * A + lane] = ab;
  }
  if (keep && lane == 0) keep[wrow] = kp;
  float vv[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = lane + 32 * i;
    float v = 0.f;
    for (int a = 0; a < A; ++a) {
      const float aa = __shfl_sync(0xffffffffu, ab, a);   // every lane takes part in the shuffle
      if (n < U) v = fmaf(aa, __ldg(w2_t + (size_t)a * ldw + n), v);
    }
    if (n < U) {
      v += __ldg(b2 + n);
      if (vin) vin[(size_t)wrow * 3 * U + 2 * U + n] = v;
    }
    vv[i] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss / (float)U + kRmsEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = lane + 32 * i;
    if (n < U) {
      const float y = siluf_((vv[i] * rs) * __ldg(g2 + n));
      x2_all[(size_t)wrow * U + n] = y;
      if (x) x[(size_t)wrow * 3 * U + 2 * U + n] = y;
    }
  }
}

}  // namespace scan
}  // namespace sd
