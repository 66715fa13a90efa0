// sd_scan.cuh -- persistent, weight-stationary posterior scan (RSSM.observe, rssm.py:140-178) for small batches.
//
// At B <= 16 rows a posterior step is ~170 MFLOP against 21 MB of fp32 weights: launching it as nine dependent
// kernels per step (and re-reading every weight from L2 each step) is pure latency.  This kernel runs ALL T steps
// in one launch on 128 CTAs (32 clusters of 4, one CTA per SM):
//
//   * every CTA copies its slices of the fp32 weights into shared memory ONCE (~144-176 KB per CTA, 21 MB in
//     total across the grid) and keeps them there for the whole scan;
//   * a step is five phases separated by grid barriers (one L2 atomic + acquire spin each); activations (16 rows)
//     travel between CTAs through L2 (ld.global.cg), never through HBM:
//       P1  block-GRU hidden layer   h_pre[g] = W_hid[g] [d_g | x0 | x1 | x2]         (rssm.py:52-61)   128 tiles of 16 columns
//       P2  gate projection + gates  d' = GRU(W_gru[g] SiLU(RMSNorm_2048(h_pre)))     (rssm.py:63-75)   128 tiles of 16 units
//       P3  the two K = 2048 layers that read d':  obs_net_0 (deter part, + the precomputed embed part) and dyn_in0
//           of the NEXT step; K split over the 4 CTAs of a cluster, partial tiles reduced through DSMEM in rank order
//       P4  logits = W_logit SiLU(RMSNorm(v_obs)) + unimix Gumbel sample                (rssm.py:171-177) one tile per 16 classes
//       P5  dyn_in1 of the next step as a gather-sum of the sampled one-hot rows       (rssm.py:47)      16 tiles
//   * everything that does not depend on the recurrence is computed before the scan: the embed part of obs_net_0 for
//     all (b, t) (one batched GEMM) and x2 = SiLU(RMSNorm(dyn_in2(action_t))) (obs_prep_kernel).
// All arithmetic is fp32 FMA with fixed reduction orders (k-group tree inside a warp, warps in order, cluster ranks
// in order): deterministic, and within the parity tolerances of the layer-by-layer path.  With a tape (stride 1)
// the per-step buffers are the backward tape itself, so sd_observe_bwd consumes the result unchanged.
#pragma once
#include "sd_kernels.cuh"

namespace sd {
namespace scan {

constexpr int THREADS = 256;
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
constexpr int kAs = kW45 + 512 * 16;           // [16][ALD]
constexpr int kRed = kAs + 16 * ALD;           // [8 warps][256]
constexpr int kSlots = kRed + 8 * 256;         // [4 ranks][256] (cluster leader)
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
  float* ssq_h;        // [128 tiles][16 rows]
  int* idx;            // [16][S]
  unsigned int* bar;   // grid barrier counter, zeroed before the launch
  long long* timing;   // diagnostic (SD_TRACE_SCAN=1): clock64 stamps of CTA 0 / 40 during step 2; null in production
};
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
// over the nkq staged k-quads.  Thread (kg = tid >> 4, ty = (tid >> 2) & 3, tx = tid & 3) owns rows 4ty..4ty+3 x
// columns 4tx..4tx+3 for the k-quads kg, kg + 16, ...; the operands of the next k-quad are fetched from shared memory
// while the current one is multiplied (two warps per scheduler cannot hide the LDS latency otherwise).  The 16
// k-group partials are summed by ONE shuffle (the two k-groups of a warp), then the 8 warps in order through shared
// memory (fixed order => deterministic).  out[g] = element (row = tid >> 4, col = tid & 15) of tile g.
// Deliberately not inlined: the persistent kernel runs every phase once per step, its code has to stay small.
template <int NG>
__device__ __noinline__ void tile_product(const float* A_s, const float* W, int wstride, int nkq, float* red, float* out) {
  const int tx = threadIdx.x & 3, ty = (threadIdx.x >> 2) & 3, kg = threadIdx.x >> 4;
  float acc[NG][4][4];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[g][r][c] = 0.f;
  const float* ap = A_s + (ty * 4) * ALD;
  float4 xa[4], wa[NG][4];
  {
    const int kq = kg < nkq ? kg : 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) xa[r] = *reinterpret_cast<const float4*>(ap + r * ALD + kq * 4);
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) wa[g][j] = *reinterpret_cast<const float4*>(W + g * wstride + kq * 64 + j * 16 + tx * 4);
  }
#pragma unroll 1
  for (int kq = kg; kq < nkq; kq += 16) {
    float4 x[4], w[NG][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = xa[r];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[g][j] = wa[g][j];
    const int kn = kq + 16 < nkq ? kq + 16 : kq;   // prefetch (re-reads the last one harmlessly)
#pragma unroll
    for (int r = 0; r < 4; ++r) xa[r] = *reinterpret_cast<const float4*>(ap + r * ALD + kn * 4);
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) wa[g][j] = *reinterpret_cast<const float4*>(W + g * wstride + kn * 64 + j * 16 + tx * 4);
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[g][r][0] = fmaf(x[r].x, w[g][0].x, acc[g][r][0]); acc[g][r][1] = fmaf(x[r].x, w[g][0].y, acc[g][r][1]);
        acc[g][r][2] = fmaf(x[r].x, w[g][0].z, acc[g][r][2]); acc[g][r][3] = fmaf(x[r].x, w[g][0].w, acc[g][r][3]);
        acc[g][r][0] = fmaf(x[r].y, w[g][1].x, acc[g][r][0]); acc[g][r][1] = fmaf(x[r].y, w[g][1].y, acc[g][r][1]);
        acc[g][r][2] = fmaf(x[r].y, w[g][1].z, acc[g][r][2]); acc[g][r][3] = fmaf(x[r].y, w[g][1].w, acc[g][r][3]);
        acc[g][r][0] = fmaf(x[r].z, w[g][2].x, acc[g][r][0]); acc[g][r][1] = fmaf(x[r].z, w[g][2].y, acc[g][r][1]);
        acc[g][r][2] = fmaf(x[r].z, w[g][2].z, acc[g][r][2]); acc[g][r][3] = fmaf(x[r].z, w[g][2].w, acc[g][r][3]);
        acc[g][r][0] = fmaf(x[r].w, w[g][3].x, acc[g][r][0]); acc[g][r][1] = fmaf(x[r].w, w[g][3].y, acc[g][r][1]);
        acc[g][r][2] = fmaf(x[r].w, w[g][3].z, acc[g][r][2]); acc[g][r][3] = fmaf(x[r].w, w[g][3].w, acc[g][r][3]);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[g][r][c] += __shfl_xor_sync(0xffffffffu, acc[g][r][c], 16);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    __syncthreads();   // red may still be read by the previous reduction
    if (lane < 16) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(red + warp * 256 + (ty * 4 + r) * 16 + tx * 4) =
            make_float4(acc[g][r][0], acc[g][r][1], acc[g][r][2], acc[g][r][3]);
    }
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 256 + threadIdx.x];
    out[g] = s;
  }
}
__device__ __forceinline__ float sum16(float v) {   // over the 16 lanes that share a row
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
// the thread's i-th float4 of a 256-wide row sits at columns 64*i + 4*seg (conflict-free stores, 256 B coalesced loads)
__device__ __forceinline__ void normact16(const float4 (&v)[4], float rs, const float* g_s, int seg, float4 (&y)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = normact4(v[i], rs, *reinterpret_cast<const float4*>(g_s + i * 64 + seg * 4));
}

// First half of the block-GRU hidden layer of step tt: s_a = W_hid[g][:, 0:512] . [keep * d_g | x0], x0 = SiLU(RMSNorm(v_in0)).
// Neither operand depends on the sample of the previous step (v_in0 is produced two phases earlier, in P3), so this
// runs in the otherwise idle P4 / P5 slot and only its 16x16 partial tile (one float per thread) is carried into P1.
__device__ __noinline__ float hid_first_half(const float* dsrc, float keep, const float* v0src, bool rok, float* A_s,
                                             const float* W1, float* red, const float* G_s, int row, int seg, float* din_t,
                                             float* x_t) {
  float4 dv[4], v0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = i * 64 + seg * 4;
    dv[i] = rok ? ldcg4(dsrc + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    v0[i] = rok ? ldcg4(v0src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float ss0 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) ss0 += sq4(v0[i]);
  ss0 = sum16(ss0);
  const float rs0 = 1.f / sqrtf(ss0 / (float)HW + kRmsEps);
  float4 x0[4];
  normact16(v0, rs0, G_s, seg, x0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dv[i].x *= keep; dv[i].y *= keep; dv[i].z *= keep; dv[i].w *= keep;
    *reinterpret_cast<float4*>(A_s + row * ALD + i * 64 + seg * 4) = dv[i];
    *reinterpret_cast<float4*>(A_s + row * ALD + HW + i * 64 + seg * 4) = x0[i];
    if (din_t) *reinterpret_cast<float4*>(din_t + i * 64 + seg * 4) = dv[i];   // backward tape: masked deter input
    if (x_t) *reinterpret_cast<float4*>(x_t + i * 64 + seg * 4) = x0[i];       // backward tape: x0
  }
  __syncthreads();
  float s;
  tile_product<1>(A_s, W1, 0, KC / 4, red, &s);
  return s;
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(THREADS, 1) observe_scan_kernel(const Params P) {
  extern __shared__ __align__(16) float sm[];
  float* W1 = sm + kW1;
  float* W2 = sm + kW2;
  float* W3 = sm + kW3;
  float* W45 = sm + kW45;
  float* A_s = sm + kAs;
  float* red = sm + kRed;
  float* slots = sm + kSlots;
  float* G_s = sm + kGain;
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int row = tid >> 4, seg = tid & 15, col = tid & 15;
  const int B = P.B, T = P.T, D = P.D, SK = P.SK;
  const bool rok = row < B;
  const int g = cta >> 4, jt = cta & 15;                       // P1 / P2: block and 16-column (16-unit) tile
  const int p3 = cta >> 6, j3 = (cta & 63) >> 2, r3 = cta & 3; // P3: problem, column tile, k-slice (= cluster rank)
  const int n4tiles = SK / 16;
  const bool do4 = cta < n4tiles, do5 = cta >= 32 && cta < 48;
  const int j5 = cta - 32;
  const size_t sstep = (size_t)P.step * B;   // rows between consecutive steps of the per-step buffers

  // ---------------------------------------------------------------- one-time: weights -> shared memory
  stage_w(W1, P.w_hid + (size_t)g * (4 * HW) * P.ld_hid, P.ld_hid, 0, 4 * HW, jt * 16, false);
  for (int gate = 0; gate < 3; ++gate)
    stage_w(W2 + gate * HW * 16, P.w_gru + (size_t)g * HW * P.ld_gru, P.ld_gru, 0, HW, gate * HW + jt * 16, false);
  stage_w(W3, p3 == 0 ? P.w_in0 : P.w_obs, p3 == 0 ? P.ld_in0 : P.ld_obs, r3 * KC, KC, j3 * 16, false);
  if (do4) stage_w(W45, P.w_lg, P.ld_lg, 0, HW, cta * 16, false);
  if (do5) stage_w(W45, P.w_in1, P.ld_in1, 0, SK, j5 * 16, false);
  // RMS scales -> shared memory, per-thread biases -> registers (the acquire of every grid barrier invalidates L1, so
  // anything re-read from global each phase would pay an L2 round trip on the critical path)
  G_s[tid] = __ldg(P.g_in0 + tid);
  G_s[256 + tid] = __ldg(P.g_in1 + tid);
  G_s[512 + tid] = __ldg(P.g_hid + g * HW + tid);
  G_s[768 + tid] = __ldg(P.g_obs + tid);
  const float bias_h = __ldg(P.b_hid + g * HW + jt * 16 + col);
  const float bias_qr = __ldg(P.b_gru + (size_t)g * 3 * HW + jt * 16 + col);
  const float bias_qc = __ldg(P.b_gru + (size_t)g * 3 * HW + HW + jt * 16 + col);
  const float bias_qu = __ldg(P.b_gru + (size_t)g * 3 * HW + 2 * HW + jt * 16 + col);
  const float bias_3 = __ldg((p3 == 0 ? P.b_in0 : P.b_obs) + j3 * 16 + col);
  const float bias_lg = do4 ? __ldg(P.b_lg + cta * 16 + col) : 0.f;
  const float bias_5 = do5 ? __ldg(P.b_in1 + j5 * 16 + col) : 0.f;
  unsigned int epoch = 0;
  __syncthreads();
  // first half of step 0's hidden layer (v_in0 of step 0 comes from the host-side launches)
  float s_a = hid_first_half(P.init_deter + (size_t)row * D + g * HW, (rok && P.is_first[(size_t)row * T]) ? 0.f : 1.f,
                             P.vin + (size_t)row * (3 * HW), rok, A_s, W1, red, G_s, row, seg,
                             (P.step && rok && jt == 0) ? P.din + (size_t)row * D + g * HW : nullptr,
                             (P.step && rok && cta == 0) ? P.x + (size_t)row * (3 * HW) : nullptr);

  for (int t = 0; t < T; ++t) {
    const float keep_t = (rok && P.is_first[(size_t)row * T + t]) ? 0.f : 1.f;
    const float keep_n = (t + 1 < T && rok && P.is_first[(size_t)row * T + t + 1]) ? 0.f : 1.f;
    float* vin_t = P.vin + (size_t)t * sstep * (3 * HW);
    float* hpre_t = P.hpre + (size_t)t * sstep * D;
    float* vobs_t = P.vobs + (size_t)t * sstep * HW;
    // pure inputs of this step, fetched now so that their latency is hidden behind P1 / P2
    const float ep_t = (r3 == 0 && p3 == 1 && rok) ? __ldg(P.eproj + ((size_t)row * T + t) * HW + j3 * 16 + col) : 0.f;
    const float uu_t = (do4 && rok) ? __ldg(P.u + ((size_t)row * T + t) * SK + cta * 16 + col) : 0.5f;

    SD_SC_STAMP(0);
    // ================================================================ P1: hidden layer of the block GRU
    {
      float4 v1[4], xv2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = i * 64 + seg * 4;
        v1[i] = rok ? ldcg4(vin_t + (size_t)row * (3 * HW) + HW + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv2[i] = rok ? __ldg(reinterpret_cast<const float4*>(P.x2 + ((size_t)t * B + row) * HW + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float ss1 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss1 += sq4(v1[i]);
      ss1 = sum16(ss1);
      const float rs1 = 1.f / sqrtf(ss1 / (float)HW + kRmsEps);
      float4 x1[4];
      normact16(v1, rs1, G_s + 256, seg, x1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(A_s + row * ALD + i * 64 + seg * 4) = x1[i];
        *reinterpret_cast<float4*>(A_s + row * ALD + HW + i * 64 + seg * 4) = xv2[i];
      }
      if (P.step && rok && cta == 0) {   // backward tape: x1 (x0 by hid_first_half, x2 by obs_prep_kernel)
        float* xt = P.x + ((size_t)t * sstep + row) * (3 * HW) + HW + seg * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(xt + i * 64) = x1[i];
      }
      __syncthreads();
      float s2;
      tile_product<1>(A_s, W1 + (KC / 4) * 64, 0, KC / 4, red, &s2);
      const float s = s_a + s2;
      const int n = g * HW + jt * 16 + col;
      const float hp = s + bias_h;
      if (rok) hpre_t[(size_t)row * D + n] = hp;
      const float ssr = sum16(hp * hp);
      if (col == 0) P.ssq_h[row * NCTA + cta] = ssr;
    }
    SD_SC_STAMP(1);
    grid_sync(P.bar, epoch);
    SD_SC_STAMP(2);

    // ================================================================ P2: gate projection + GRU gates
    {
      float4 hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        hv[i] = rok ? ldcg4(hpre_t + (size_t)row * D + g * HW + i * 64 + seg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) tot += ldcg(P.ssq_h + row * NCTA + seg * 8 + i);   // 128 tile partials per row
      tot = sum16(tot);
      const float rs = 1.f / sqrtf(tot / (float)D + kRmsEps);
      float4 hh[4];
      normact16(hv, rs, G_s + 512, seg, hh);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(A_s + row * ALD + i * 64 + seg * 4) = hh[i];
      if (P.step && rok && jt == 0) {
        float* ht = P.h + ((size_t)t * sstep + row) * D + g * HW + seg * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(ht + i * 64) = hh[i];
      }
      const int n = g * HW + jt * 16 + col;   // unit
      const float* dsrc = t == 0 ? P.init_deter + (size_t)row * D : P.deters + ((size_t)row * T + (t - 1)) * D;
      const float dprev = rok ? keep_t * ldcg(dsrc + n) : 0.f;
      __syncthreads();
      float q3[3];
      tile_product<3>(A_s, W2, HW * 16, HW / 4, red, q3);   // reset | cand | update tiles share the A operand
      const float qr = q3[0] + bias_qr;
      const float qc = q3[1] + bias_qc;
      const float qu = q3[2] + bias_qu;
      if (rok) {
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
      const float* dsrc = P.deters + ((size_t)row * T + t) * D + r3 * KC + seg * 4;
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = rok ? ldcg4(dsrc + i * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = j3 * 16 + col;
      const float extra = bias_3 + ep_t;   // leader: bias (+ embed part of obs_net_0)
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(A_s + row * ALD + i * 64 + seg * 4) = a[i];
      __syncthreads();
      SD_SC_STAMP(11);
      float s;
      tile_product<1>(A_s, W3, 0, KC / 4, red, &s);
      SD_SC_STAMP(12);
      {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(slots + r3 * 256 + tid);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
      }
      cluster_sync_all();
      SD_SC_STAMP(13);
      if (r3 == 0 && rok) {
        const float tot = ((slots[tid] + slots[256 + tid]) + slots[512 + tid]) + slots[768 + tid];
        if (p3 == 0) {
          if (t + 1 < T) P.vin[(size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW) + n] = extra + keep_n * tot;
        } else {
          vobs_t[(size_t)row * HW + n] = tot + extra;
        }
      }
    }
    SD_SC_STAMP(5);
    grid_sync(P.bar, epoch);
    SD_SC_STAMP(6);

    // ================================================================ P4: logits + sample (other CTAs: first half of the next hidden layer)
    if (!do4 && t + 1 < T) s_a = hid_first_half(P.deters + ((size_t)row * T + t) * D + g * HW, keep_n, P.vin + (size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW),
                           rok, A_s, W1, red, G_s, row, seg,
                           (P.step && rok && jt == 0) ? P.din + ((size_t)(t + 1) * sstep + row) * D + g * HW : nullptr,
                           (P.step && rok && cta == 0) ? P.x + ((size_t)(t + 1) * sstep + row) * (3 * HW) : nullptr);
    if (do4) {
      float4 vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vv[i] = rok ? ldcg4(vobs_t + (size_t)row * HW + i * 64 + seg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = cta * 16 + col;
      const float uu = uu_t;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += sq4(vv[i]);
      ss = sum16(ss);
      const float rs = 1.f / sqrtf(ss / (float)HW + kRmsEps);
      float4 oo[4];
      normact16(vv, rs, G_s + 768, seg, oo);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(A_s + row * ALD + i * 64 + seg * 4) = oo[i];
      if (P.step && rok && cta == 0) {
        float* ot = P.o + ((size_t)t * sstep + row) * HW + seg * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(ot + i * 64) = oo[i];
      }
      __syncthreads();
      float lgv;
      tile_product<1>(A_s, W45, 0, HW / 4, red, &lgv);
      lgv += bias_lg;
      const int Kc = P.K, kcls = col % Kc;
      int best;
      if (Kc == 16) best = sample_group<16>(lgv, uu, true, kcls, Kc, P.unimix, nullptr);
      else if (Kc == 8) best = sample_group<8>(lgv, uu, true, kcls, Kc, P.unimix, nullptr);
      else if (Kc == 4) best = sample_group<4>(lgv, uu, true, kcls, Kc, P.unimix, nullptr);
      else best = sample_group<2>(lgv, uu, true, kcls, Kc, P.unimix, nullptr);
      if (rok) {
        const float oh = (kcls == best) ? 1.f : 0.f;
        const size_t off = ((size_t)row * T + t) * SK + n;
        P.stochs[off] = oh;
        P.logits[off] = lgv;
        if (P.step) {
          P.lg[((size_t)t * sstep + row) * SK + n] = lgv;
          if (t + 1 < T) P.zin[((size_t)(t + 1) * sstep + row) * SK + n] = keep_n * oh;
        }
        if (kcls == 0) P.idx[row * P.S + n / Kc] = best;
      }
    }
    SD_SC_STAMP(7);
    grid_sync(P.bar, epoch);
    SD_SC_STAMP(8);

    // ================================================================ P5: next step's dyn_in1 (gather-sum of one-hot rows)
    if (do4 && t + 1 < T) s_a = hid_first_half(P.deters + ((size_t)row * T + t) * D + g * HW, keep_n, P.vin + (size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW),
                           rok, A_s, W1, red, G_s, row, seg,
                           (P.step && rok && jt == 0) ? P.din + ((size_t)(t + 1) * sstep + row) * D + g * HW : nullptr,
                           (P.step && rok && cta == 0) ? P.x + ((size_t)(t + 1) * sstep + row) * (3 * HW) : nullptr);
    if (do5 && t + 1 < T) {
      float v = 0.f;
      if (rok) {
        const int Kc = P.K;
        for (int s = 0; s < P.S; ++s) {
          const int id = __ldcg(P.idx + row * P.S + s);
          v += W45[(s * Kc + id) * 16 + col];
        }
      }
      const int n = j5 * 16 + col;
      if (rok) P.vin[(size_t)(t + 1) * sstep * (3 * HW) + (size_t)row * (3 * HW) + HW + n] = bias_5 + keep_n * v;
    }
    SD_SC_STAMP(9);
    if (t + 1 < T) grid_sync(P.bar, epoch);
    SD_SC_STAMP(10);
  }
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
