// sd_cnn_bwd.cuh -- backward of the CNN encoder stages (autograd of world_model/networks.py:192-234) for sm_100a.
//
// Per stage, in reverse order:
//   1. norm_pool_bwd_kernel  (SIMT, warp per pooled pixel, lane = channel): SiLU / RMSNorm backward on the pooled pre-norm
//      map kept by the forward, RMS-scale and bias gradient partials, and the max-pool scatter -- writes dy, the gradient
//      of the conv output, as a dense bf16 NHWC map (zeros off the arg-max).
//   2. conv_dgrad_kernel     (tcgen05): dx = conv of dy with the flipped, transposed taps: 25 shifted views of dy, one
//      accumulator of N = Cin columns, same chunk-major no-swizzle operands / cp.async producer as the forward.
//   3. conv_wgrad_kernel     (tcgen05, MN-major operands): dW[tap][ci][co] = sum over pixels x[p + tap][ci] dy[p][co].  The
//      K dimension is the pixel index, so the SAME [chunk][pixel][16 B] shared-memory image the forward uses as a K-major
//      operand is read as an MN-major one (LBO = 128 B between 8-pixel K blocks, SBO = 2 KB between 8-channel MN blocks);
//      128 / Cin views are stacked into one M = 128 operand.  Each CTA keeps its tap groups' accumulators in TMEM over all
//      of its pixel tiles and writes one fp32 partial; wgrad_reduce_kernel sums the partials in a fixed order into the
//      reference layout (Cout, Cin, 5, 5), accumulating into the caller's gradient tensors.
#pragma once
#include "sd_cnn.cuh"

namespace sd {
namespace cnn {

// sum of p[0], p[stride], ... in a fixed association (sixteen interleaved running sums, then a fixed tree): the same bits on
// every run, with sixteen loads in flight instead of one
__device__ __forceinline__ float ordered_sum(const float* __restrict__ p, int n, size_t stride) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 0.f;
  int b = 0;
  for (; b + 16 <= n; b += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldg(p + (size_t)(b + k) * stride);
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] += v[k];
  }
  for (int k = 0; b < n; ++b, ++k) a[k] += __ldg(p + (size_t)b * stride);
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int k = 0; k < w; ++k) a[k] += a[k + w];
  return a[0];
}

// the same for four adjacent columns at once (16-byte loads)
__device__ __forceinline__ float4 ordered_sum4(const float* __restrict__ p, int n, size_t stride) {
  float4 a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int b = 0;
  for (; b + 8 <= n; b += 8) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(p + (size_t)(b + k) * stride));
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k].x += v[k].x; a[k].y += v[k].y; a[k].z += v[k].z; a[k].w += v[k].w; }
  }
  for (int k = 0; b < n; ++b, ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)b * stride));
    a[k].x += v.x; a[k].y += v.y; a[k].z += v.z; a[k].w += v.w;
  }
#pragma unroll
  for (int w = 4; w > 0; w >>= 1)
#pragma unroll
    for (int k = 0; k < w; ++k) { a[k].x += a[k + w].x; a[k].y += a[k + w].y; a[k].z += a[k + w].z; a[k].w += a[k + w].w; }
  return a[0];
}

// ------------------------------------------------------------------------------------------------ 1. norm / pool backward
struct NormBwdParams {
  const float* pool;      // [total][cp] pooled pre-norm values (forward tape)
  const uint8_t* arg;     // [total][cp]
  const float* gain;      // [64] zero padded
  const float* dout;      // gradient of the stage output: fp32 [total][ldo] NHWC, or (embed != 0) [N][cout * HpWp] in (C,H,W) order
  bf16* dy;               // [N][2 Hp][2 Wp][cp] bf16, fully written
  float* partial;         // [gridDim.x][2][64]: per-CTA sums of d(gain), d(bias)
  int total, cp, cout, ldo, embed, HpWp, Wp;
};

// CPL = channels per lane (cp <= 32 * CPL)
template <int CPL>
__global__ void __launch_bounds__(256) norm_pool_bwd_kernel(const NormBwdParams P) {
  __shared__ float red[8][2][64];
  constexpr int U = 4;                         // pixels per warp iteration: all of their loads are issued before any math
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * CPL;
  const bool act = c0 < P.cp;
  float g[CPL], dg[CPL], db[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { g[j] = act ? P.gain[c0 + j] : 0.f; dg[j] = 0.f; db[j] = 0.f; }
  const float inv_c = 1.f / (float)P.cout;
  const int W = 2 * P.Wp, Hp = P.HpWp / P.Wp;
  for (int px0 = (blockIdx.x * 8 + warp) * U; px0 < P.total; px0 += gridDim.x * 8 * U) {
    float p[U][CPL], d[U][CPL];
    uint32_t ar[U][CPL];
    int nn_[U], rem_[U];
    const int n_first = px0 / P.HpWp, rem_first = px0 - n_first * P.HpWp;   // one division per U pixels
    const int py_first = rem_first / P.Wp, pc_first = rem_first - py_first * P.Wp;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int px = px0 + u;
      const bool ok = act && px < P.total;
      nn_[u] = n_first; rem_[u] = rem_first + u;
      if (rem_[u] >= P.HpWp) { rem_[u] -= P.HpWp; ++nn_[u]; }
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        p[u][j] = 0.f; d[u][j] = 0.f; ar[u][j] = 0u;
        if (ok) {
          p[u][j] = __ldg(P.pool + (size_t)px * P.cp + c0 + j);
          ar[u][j] = __ldg(P.arg + (size_t)px * P.cp + c0 + j);
          if (c0 + j < P.cout)
            d[u][j] = P.embed ? __ldg(P.dout + (size_t)nn_[u] * P.cout * P.HpWp + (size_t)(c0 + j) * P.HpWp + rem_[u])
                              : __ldg(P.dout + (size_t)px * P.ldo + c0 + j);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int px = px0 + u;
      if (px >= P.total) break;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) ss = fmaf(p[u][j], p[u][j], ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rho = rsqrtf(ss * inv_c + kRmsEps);
      float nn[CPL], dn[CPL], dot = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        nn[j] = p[u][j] * rho;
        const float m = nn[j] * g[j];
        const float sg = __fdividef(1.f, 1.f + __expf(-m));
        const float dm = d[u][j] * (sg * (1.f + m * (1.f - sg)));
        dg[j] = fmaf(dm, nn[j], dg[j]);
        dn[j] = dm * g[j];
        dot = fmaf(dn[j], nn[j], dot);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dot *= inv_c;
      int py, pxx;
      if (P.Wp >= U) {                 // U consecutive pixels cross at most one row end: no division per pixel
        py = py_first; pxx = pc_first + u;
        if (pxx >= P.Wp) { pxx -= P.Wp; ++py; }
        if (rem_first + u >= P.HpWp) { py = 0; pxx = rem_first + u - P.HpWp; }
      } else {
        py = rem_[u] / P.Wp; pxx = rem_[u] - py * P.Wp;
      }
      bf16* base = P.dy + ((size_t)(nn_[u] * 2 * Hp + 2 * py) * W + 2 * pxx) * P.cp + c0;
      float v[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) { v[j] = rho * (dn[j] - nn[j] * dot); db[j] += v[j]; }
      if (act) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          bf16* dst = base + ((size_t)(q >> 1) * W + (q & 1)) * P.cp;
          if (CPL == 2) *reinterpret_cast<uint32_t*>(dst) = pack2(ar[u][0] == (uint32_t)q ? v[0] : 0.f, ar[u][CPL - 1] == (uint32_t)q ? v[CPL - 1] : 0.f);
          else dst[0] = __float2bfloat16(ar[u][0] == (uint32_t)q ? v[0] : 0.f);
        }
      }
    }
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) { red[warp][0][c0 + j] = dg[j]; red[warp][1][c0 + j] = db[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
    float s = 0.f;
    if (c < P.cp)
      for (int w = 0; w < 8; ++w) s += red[w][k][c];
    P.partial[((size_t)blockIdx.x * 2 + k) * 64 + c] = s;
  }
}

// Thread-per-pixel variant (the default): a thread keeps its pixel's CP channels in registers, so the row statistics need
// no shuffles and the per-pixel scalar work (index arithmetic, rsqrt) is done once per pixel instead of once per lane; only
// the two per-channel sums (d gain, d bias) cross lanes, as one warp reduction per channel and 32 pixels.  ~6x fewer warp
// instructions per pixel than the lane-per-channel kernel above, which is kept for cp = 16 / odd shapes.
template <int CP>
__global__ void __launch_bounds__(128) norm_pool_bwd_px_kernel(const NormBwdParams P) {
  __shared__ float red[4][2][64];
  __shared__ float s_g[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) s_g[threadIdx.x] = P.gain[threadIdx.x];
  __syncthreads();
  constexpr int NA = (CP + 31) / 32;
  float acc_g[NA], acc_b[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) { acc_g[k] = 0.f; acc_b[k] = 0.f; }
  const float inv_c = 1.f / (float)P.cout;
  const int W = 2 * P.Wp, Hp = P.HpWp / P.Wp;
  for (int px0 = (blockIdx.x * 4 + warp) * 32; px0 < P.total; px0 += gridDim.x * 128) {
    const int px = px0 + lane;
    const bool ok = px < P.total;
    float p[CP], d[CP];
    if (ok) {
      const float4* src = reinterpret_cast<const float4*>(P.pool + (size_t)px * CP);
#pragma unroll
      for (int c = 0; c < CP; c += 4) { const float4 t = __ldg(src + (c >> 2)); p[c] = t.x; p[c + 1] = t.y; p[c + 2] = t.z; p[c + 3] = t.w; }
    } else {
#pragma unroll
      for (int c = 0; c < CP; ++c) p[c] = 0.f;
    }
    const int n = px / P.HpWp, rem = px - n * P.HpWp;
    if (ok && P.embed) {
      const float* src = P.dout + (size_t)n * P.cout * P.HpWp + rem;     // (C, H, W) order: consecutive lanes = consecutive addresses
#pragma unroll
      for (int c = 0; c < CP; ++c) d[c] = c < P.cout ? __ldg(src + (size_t)c * P.HpWp) : 0.f;
    } else if (ok) {
      const float4* src = reinterpret_cast<const float4*>(P.dout + (size_t)px * P.ldo);
#pragma unroll
      for (int c = 0; c < CP; c += 4) { const float4 t = __ldg(src + (c >> 2)); d[c] = t.x; d[c + 1] = t.y; d[c + 2] = t.z; d[c + 3] = t.w; }
    } else {
#pragma unroll
      for (int c = 0; c < CP; ++c) d[c] = 0.f;
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) ss = fmaf(p[c], p[c], ss);
    const float rho = rsqrtf(ss * inv_c + kRmsEps);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float nn = p[c] * rho, g = s_g[c];
      const float m = nn * g;
      const float sg = __fdividef(1.f, 1.f + __expf(-m));
      const float dm = d[c] * (sg * (1.f + m * (1.f - sg)));
      float t = dm * nn;                                   // d gain contribution: summed over the warp's 32 pixels
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == (c & 31)) acc_g[c >> 5] += t;
      d[c] = dm * g;                                       // dn
      p[c] = nn;
      dot = fmaf(d[c], nn, dot);
    }
    dot *= inv_c;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float v = rho * (d[c] - p[c] * dot);           // gradient of the pooled pre-norm value = of the bias
      float t = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == (c & 31)) acc_b[c >> 5] += t;
      p[c] = v;
    }
    if (ok) {
      uint32_t ar[CP / 4];
      const uint4* asrc = reinterpret_cast<const uint4*>(P.arg + (size_t)px * CP);
#pragma unroll
      for (int c = 0; c < CP / 16; ++c) { const uint4 t = __ldg(asrc + c); ar[4 * c] = t.x; ar[4 * c + 1] = t.y; ar[4 * c + 2] = t.z; ar[4 * c + 3] = t.w; }
      const int py = rem / P.Wp, pxx = rem - py * P.Wp;
      bf16* base = P.dy + ((size_t)(n * 2 * Hp + 2 * py) * W + 2 * pxx) * CP;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4* dst = reinterpret_cast<uint4*>(base + ((size_t)(q >> 1) * W + (q & 1)) * CP);
#pragma unroll
        for (int c = 0; c < CP; c += 8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ((ar[(c + e) >> 2] >> (8 * ((c + e) & 3))) & 0xffu) == (uint32_t)q ? p[c + e] : 0.f;
          dst[c >> 3] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NA; ++k)
    if (lane + 32 * k < 64) { red[warp][0][lane + 32 * k] = acc_g[k]; red[warp][1][lane + 32 * k] = acc_b[k]; }
  if (NA == 1) { red[warp][0][lane + 32] = 0.f; red[warp][1][lane + 32] = 0.f; }
  __syncthreads();
  {
    const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
    float s = 0.f;
    if (c < CP)
      for (int w = 0; w < 4; ++w) s += red[w][k][c];
    P.partial[((size_t)blockIdx.x * 2 + k) * 64 + c] = s;
  }
}

// Split-pixel variant (the default for cp = 32 / 48 / 64): TPP = 2 or 4 adjacent lanes share a pixel, each with CPT = cp / TPP
// (16 or 12) of its channels in registers.  Row statistics cost one or two xor-shuffles, the per-channel sums (d gain,
// d bias) are per-thread running sums over all of the thread's pixels (reduced across lanes once, at the end), every load
// and store of a pixel group is one contiguous run, and ~90 registers leave 20+ warps per SM to cover the memory latency
// (the thread-per-pixel kernel above runs at 12 warps per SM and 2.1 TB/s; ncu: profiles/r02_norm_pool_bwd_ncu.txt).
template <int CP, int TPP>
__global__ void __launch_bounds__(256) norm_pool_bwd_split_kernel(const NormBwdParams P) {
  constexpr int CPT = CP / TPP, PPW = 32 / TPP;          // channels per thread, pixels per warp
  static_assert(CPT % 4 == 0 && CP % TPP == 0, "channel split");
  __shared__ float red[8][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int part = lane % TPP, sub = lane / TPP, c0 = part * CPT;
  float g[CPT], acc_g[CPT], acc_b[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) { g[j] = P.gain[c0 + j]; acc_g[j] = 0.f; acc_b[j] = 0.f; }
  const float inv_c = 1.f / (float)P.cout;
  const int W = 2 * P.Wp, Hp = P.HpWp / P.Wp;
  for (int px0 = (blockIdx.x * 8 + warp) * PPW; px0 < P.total; px0 += gridDim.x * 8 * PPW) {
    const int px = px0 + sub;
    const bool ok = px < P.total;
    float p[CPT], d[CPT];
    uint32_t ar[CPT / 4];
    const int n = px / P.HpWp, rem = px - n * P.HpWp;
    if (ok) {
      const float4* src = reinterpret_cast<const float4*>(P.pool + (size_t)px * CP + c0);
#pragma unroll
      for (int j = 0; j < CPT; j += 4) { const float4 t = __ldg(src + (j >> 2)); p[j] = t.x; p[j + 1] = t.y; p[j + 2] = t.z; p[j + 3] = t.w; }
      if (P.embed) {
        const float* e = P.dout + (size_t)n * P.cout * P.HpWp + rem;
#pragma unroll
        for (int j = 0; j < CPT; ++j) d[j] = c0 + j < P.cout ? __ldg(e + (size_t)(c0 + j) * P.HpWp) : 0.f;
      } else {
        const float4* dsrc = reinterpret_cast<const float4*>(P.dout + (size_t)px * P.ldo + c0);
#pragma unroll
        for (int j = 0; j < CPT; j += 4) { const float4 t = __ldg(dsrc + (j >> 2)); d[j] = t.x; d[j + 1] = t.y; d[j + 2] = t.z; d[j + 3] = t.w; }
      }
      const uint32_t* asrc = reinterpret_cast<const uint32_t*>(P.arg + (size_t)px * CP + c0);
#pragma unroll
      for (int j = 0; j < CPT / 4; ++j) ar[j] = __ldg(asrc + j);
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j) { p[j] = 0.f; d[j] = 0.f; }
#pragma unroll
      for (int j = 0; j < CPT / 4; ++j) ar[j] = 0u;
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) ss = fmaf(p[j], p[j], ss);
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rho = rsqrtf(ss * inv_c + kRmsEps);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const float nn = p[j] * rho, m = nn * g[j];
      const float sg = __fdividef(1.f, 1.f + __expf(-m));
      const float dm = d[j] * (sg * (1.f + m * (1.f - sg)));
      acc_g[j] = fmaf(dm, nn, acc_g[j]);
      d[j] = dm * g[j];
      p[j] = nn;
      dot = fmaf(d[j], nn, dot);
    }
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    dot *= inv_c;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { p[j] = rho * (d[j] - p[j] * dot); acc_b[j] += p[j]; }
    if (ok) {
      const int py = rem / P.Wp, pxx = rem - py * P.Wp;
      bf16* base = P.dy + ((size_t)(n * 2 * Hp + 2 * py) * W + 2 * pxx) * CP + c0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[CPT / 2];
#pragma unroll
        for (int j = 0; j < CPT; j += 2) {
          const float v0 = ((ar[j >> 2] >> (8 * (j & 3))) & 0xffu) == (uint32_t)q ? p[j] : 0.f;
          const float v1 = ((ar[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 0xffu) == (uint32_t)q ? p[j + 1] : 0.f;
          w[j >> 1] = pack2(v0, v1);
        }
        uint2* dst = reinterpret_cast<uint2*>(base + ((size_t)(q >> 1) * W + (q & 1)) * CP);
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) dst[j] = make_uint2(w[2 * j], w[2 * j + 1]);
      }
    }
  }
  // lanes with the same `part` hold sums of the same channels: fold them, then one smem pass over the 8 warps
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
#pragma unroll
    for (int o = TPP; o < 32; o <<= 1) {
      acc_g[j] += __shfl_xor_sync(0xffffffffu, acc_g[j], o);
      acc_b[j] += __shfl_xor_sync(0xffffffffu, acc_b[j], o);
    }
  }
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) { red[warp][0][c0 + j] = acc_g[j]; red[warp][1][c0 + j] = acc_b[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
    float s = 0.f;
    if (c < CP)
      for (int w = 0; w < 8; ++w) s += red[w][k][c];
    P.partial[((size_t)blockIdx.x * 2 + k) * 64 + c] = s;
  }
}

// sums the per-CTA partials in order and ACCUMULATES into the caller's gradient tensors (nullable)
__global__ void norm_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int cout, float* __restrict__ d_gain,
                                       float* __restrict__ d_bias) {
  const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
  if (c >= cout) return;
  const float s = ordered_sum(partial + (size_t)k * 64 + c, nblocks, 128);
  float* dst = k == 0 ? d_gain : d_bias;
  if (dst) dst[c] += s;
}

// ------------------------------------------------------------------------------------------------ 2. dgrad
// dx[n][iy][ix][ci] = sum_{ky,kx,co} dy[n][iy - ky + 2][ix - kx + 2][co] w[co][ci][ky][kx]
struct DgradParams {
  const bf16* dy;    // [N][H][W][CK] bf16 (CK = template parameter = the stage's cp)
  const bf16* wT;    // [25 taps][CK/8][cinp][8]: w[co = k][ci = row][ky][kx], zero padded
  float* dx;         // [N][H][W][ldx] fp32; the first `cin` channels are written ... up to ldx (zero padded by the weights)
  int H, W, total, tiles, cinp, ldx, nwrite, stages;
};

template <int CK, bool RES>
struct DgradSmem {
  static constexpr int KC = CK / 8;
  static constexpr int kA = KC * BM * 16;
  static constexpr int kBt = KC * 64 * 16;                         // streamed: one tap tile at cinp = 64
  static constexpr int VPS = RES ? 5 : 2;                          // taps per ring stage
  static constexpr int NG = (25 + VPS - 1) / VPS;
  static constexpr int kStage = RES ? VPS * kA : VPS * (kA + kBt);
  static constexpr int kMaxStages = 8;
  static constexpr int kFixed = 8 * (2 * kMaxStages + 4) + 16 + 256;
  static int resident_bytes(int cinp) { return RES ? 25 * KC * cinp * 16 : 0; }
  static int stages(int cinp) {
    int n = (kConvSmemBudget - kFixed - resident_bytes(cinp)) / kStage;
    return n > kMaxStages ? kMaxStages : n;
  }
  static int total(int cinp) { return kFixed + resident_bytes(cinp) + stages(cinp) * kStage; }
};

constexpr int DG_THREADS = 416;   // warps 0-3 producers, 4 MMA, 5-8 / 9-12 two epilogue groups

template <int CK, bool RES>
__global__ void __launch_bounds__(DG_THREADS, 1) conv_dgrad_kernel(const __grid_constant__ DgradParams P) {
  using L = DgradSmem<CK, RES>;
  constexpr int KC = L::KC;
  const int STAGES = P.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int cinp = P.cinp;
  const int resB = RES ? 25 * KC * cinp * 16 : 0;
  const uint32_t ring = base + (uint32_t)resB;
  const int off_bar = resB + STAGES * L::kStage;
  const uint32_t bar_full = base + (uint32_t)off_bar, bar_empty = bar_full + 8 * L::kMaxStages, bar_accf = bar_empty + 8 * L::kMaxStages,
                 bar_free = bar_accf + 16, tmem_slot = bar_free + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 8 * (2 * L::kMaxStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (RES) {
    const int n16 = 25 * KC * cinp;      // global [tap][c][row][8] -> shared [c][tap * cinp + row][16 B]
    for (int i = threadIdx.x; i < n16; i += DG_THREADS) {
      const int row = i % cinp, c = (i / cinp) % KC, tap = i / (cinp * KC);
      reinterpret_cast<uint4*>(gbase)[(size_t)c * 25 * cinp + tap * cinp + row] = __ldg(reinterpret_cast<const uint4*>(P.wT) + i);
    }
    fence_async_smem();
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const int HW = P.H * P.W;

  if (warp < 4) {
    const int tid = threadIdx.x;
    int s = 0;
    uint32_t eph = 1;
    const int rowpitch = P.W * CK;
    const int wrow = tid & 63, wc0 = tid >> 6;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const int g = tile * BM + tid;
      const bool ok = g < P.total;
      const int n = g / HW, rem = g - n * HW, iy = rem / P.W, ix = rem - iy * P.W;
      const bf16* src0 = P.dy + ((size_t)(n * P.H + iy) * P.W + ix) * CK;
      uint32_t rmask = 0, cmask = 0;   // bit k: row iy + 2 - k / column ix + 2 - k is inside the map
#pragma unroll
      for (int k = 0; k < KSZ; ++k) {
        rmask |= (uint32_t)(ok && iy + 2 - k >= 0 && iy + 2 - k < P.H) << k;
        cmask |= (uint32_t)(ix + 2 - k >= 0 && ix + 2 - k < P.W) << k;
      }
#pragma unroll
      for (int vg = 0; vg < L::NG; ++vg) {
        mbar_wait(bar_empty + 8 * s, eph);
        const uint32_t stg = ring + (uint32_t)(s * L::kStage);
#pragma unroll
        for (int u = 0; u < L::VPS; ++u) {
          const int t = vg * L::VPS + u;
          if (t >= 25) break;
          const int ky = t / KSZ, kx = t - ky * KSZ;
          const bool valid = ((rmask >> ky) & (cmask >> kx) & 1u) != 0u;
          const bf16* src = valid ? src0 + (2 - ky) * rowpitch + (2 - kx) * CK : P.dy;
          const uint32_t nbytes = valid ? 16u : 0u;
          const uint32_t va = stg + (uint32_t)(u * (RES ? L::kA : L::kA + L::kBt));
#pragma unroll
          for (int c = 0; c < KC; ++c) cp_async16(va + (uint32_t)(c * BM * 16 + tid * 16), src + c * 8, nbytes);
          if (!RES) {
            if (wrow < cinp) {
              const uint4* w = reinterpret_cast<const uint4*>(P.wT) + (size_t)(t * KC + wc0) * cinp + wrow;
              const uint32_t dst = va + (uint32_t)(L::kA + wrow * 16 + wc0 * cinp * 16);
#pragma unroll
              for (int c = 0; c < KC / 2; ++c) cp_async16(dst + (uint32_t)(c * 2 * cinp * 16), w + (size_t)c * 2 * cinp, 16u);
            }
          }
        }
        cp_async_arrive(bar_full + 8 * s);
        if (++s == STAGES) { s = 0; eph ^= 1u; }
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, cinp);
      const uint64_t dA0 = make_desc_nosw(0, BM * 16, 128);
      const uint32_t ldB = (uint32_t)((RES ? 25 : 1) * cinp * 16);
      const uint64_t dB0 = make_desc_nosw(0, ldB, 128);
      const uint32_t ahi = (uint32_t)(dA0 >> 32), bhi = (uint32_t)(dB0 >> 32), alo0 = (uint32_t)dA0, blo0 = (uint32_t)dB0;
      const uint32_t kstepB = (2 * ldB) >> 4, blo_res = (base & 0x3FFFFu) >> 4;
      int lt = 0, s = 0;
      uint32_t fph = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        const int set = lt & 1;
        if (lt >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((lt >> 1) - 1) & 1u);
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(set * 64);
#pragma unroll
        for (int vg = 0; vg < L::NG; ++vg) {
          mbar_wait(bar_full + 8 * s, fph);
          fence_async_smem();
          tc_fence_after();
          const uint32_t stlo = ((ring + (uint32_t)(s * L::kStage)) & 0x3FFFFu) >> 4;
#pragma unroll
          for (int u = 0; u < L::VPS; ++u) {
            const int t = vg * L::VPS + u;
            if (t >= 25) break;
            const uint32_t alo = alo0 + stlo + (uint32_t)((u * (RES ? L::kA : L::kA + L::kBt)) >> 4);
            const uint32_t blo = blo0 + (RES ? blo_res + (uint32_t)(t * cinp) : stlo + (uint32_t)((u * (L::kA + L::kBt) + L::kA) >> 4));
#pragma unroll
            for (int kk = 0; kk < CK / 16; ++kk)
              mma_lh(acc, alo + (uint32_t)((kk * 2 * BM * 16) >> 4), ahi, blo + kk * kstepB, bhi, idesc, (t | kk) == 0 ? 0u : 1u);
          }
          tc_commit(bar_empty + 8 * s);
          if (++s == STAGES) { s = 0; fph ^= 1u; }
        }
        tc_commit(bar_accf + 8 * set);
      }
    }
  } else {
    // epilogue: thread = input pixel; fp32 NHWC store of the first nwrite channels (row stride ldx)
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int set = warp >= 9 ? 1 : 0;
    for (int lt = set, tile = blockIdx.x + set * gridDim.x; tile < P.tiles; tile += 2 * gridDim.x, lt += 2) {
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t tm = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 64);
      float v[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16)
        if (c0 < cinp) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(tm + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) v[c0 + i] = __uint_as_float(r[i]);
        }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * set);
      const int g = tile * BM + row;
      if (g < P.total) {
        float* dst = P.dx + (size_t)g * P.ldx;
        if ((P.nwrite & 3) == 0 && (P.ldx & 3) == 0) {
#pragma unroll
          for (int c = 0; c < 64; c += 4)
            if (c < P.nwrite) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c < P.nwrite) dst[c] = v[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
  }
}

// ------------------------------------------------------------------------------------------------ 2b. dgrad, patch-resident
// The 25 shifted views of dy overlap almost completely, so for maps with H % 16 == 0 and W % 8 == 0 a tile is a 16 x 8 pixel
// block and ONE (16+4) x (8+4) halo patch of dy is staged per tile ([chunk][240 patch pixels][16 B], zero filled outside the
// map); the view of tap (ky, kx) is the patch read through a shifted descriptor: 8 consecutive pixels of a patch row are
// one 128-byte core matrix, SBO = the patch row pitch (192 B), LBO = the chunk plane (3840 B), start += ((4-ky) 12 + 4-kx) 16.
// 12x less shared-memory ingest than staging every view (23 KB instead of 300 KB per tile at 48 channels).
constexpr int PT_W = 8, PT_H = 16, PP_W = PT_W + 4, PP_H = PT_H + 4, PP = PP_W * PP_H;   // 240 patch pixels
constexpr int PPL = PP + 1;   // chunk plane in 16-byte units: + 1 so that the chunks of one pixel fall into different banks
template <int CK, bool RES>
struct DgradPatchSmem {
  static constexpr int KC = CK / 8;
  static constexpr int kPatch = (KC * PPL * 16 + 127) / 128 * 128;
  static constexpr int kBt = KC * 64 * 16;
  static constexpr int VPS = 2;                                    // streamed weight tiles per ring stage
  static constexpr int NG = (25 + VPS - 1) / VPS;
  static constexpr int kStage = VPS * kBt;
  static constexpr int STAGES = RES ? 0 : 6;
  static constexpr int kFixed = 8 * (2 * 6 + 8) + 16 + 256;
  static int resident_bytes(int cinp) { return RES ? 25 * KC * cinp * 16 : 0; }
  static int total(int cinp) { return kFixed + resident_bytes(cinp) + 2 * kPatch + STAGES * kStage; }
};

template <int CK, bool RES>
__global__ void __launch_bounds__(DG_THREADS, 1) conv_dgrad_patch_kernel(const __grid_constant__ DgradParams P) {
  using L = DgradPatchSmem<CK, RES>;
  constexpr int KC = L::KC, STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int cinp = P.cinp;
  const int resB = RES ? 25 * KC * cinp * 16 : 0;
  const uint32_t patch0 = base + (uint32_t)resB, ring = patch0 + 2 * L::kPatch;
  const int off_bar = resB + 2 * L::kPatch + STAGES * L::kStage;
  const uint32_t bar_full = base + (uint32_t)off_bar, bar_empty = bar_full + 48, bar_pf = bar_empty + 48, bar_pe = bar_pf + 16,
                 bar_accf = bar_pe + 16, bar_free = bar_accf + 16, tmem_slot = bar_free + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 160);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (RES) {
    const int n16 = 25 * KC * cinp;
    for (int i = threadIdx.x; i < n16; i += DG_THREADS) {
      const int row = i % cinp, c = (i / cinp) % KC, tap = i / (cinp * KC);
      reinterpret_cast<uint4*>(gbase)[(size_t)c * 25 * cinp + tap * cinp + row] = __ldg(reinterpret_cast<const uint4*>(P.wT) + i);
    }
    fence_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 6; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_pf + 8 * s, PROD);
      mbar_init(bar_pe + 8 * s, 1);
      mbar_init(bar_accf + 8 * s, 1);
      mbar_init(bar_free + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const int tx_n = P.W / PT_W, tiles_per_img = tx_n * (P.H / PT_H);

  if (warp < 4) {
    const int tid = threadIdx.x;
    constexpr int NI = (KC * PP + PROD - 1) / PROD;       // patch items (pixel, chunk) per thread, chunk fastest
    int it_dst[NI], it_src[NI], it_rc[NI];                // destination, source offset from the patch origin, (row << 8 | column); -1 = none
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * PROD;
      const int c = i % KC, pp = i / KC, pr = pp / PP_W, pc = pp - pr * PP_W;
      it_dst[k] = (c * PPL + pr * PP_W + pc) * 16;
      it_src[k] = (pr * P.W + pc) * CK + c * 8;
      it_rc[k] = i < KC * PP ? (pr << 8) | pc : -1;
    }
    const int wrow = tid & 63, wc0 = tid >> 6;
    int s = 0, lt = 0;
    uint32_t eph = 1;
    const int ty_n = P.H / PT_H;
    auto load_patch = [&](int tile, int pb) {
      const int n = tile / tiles_per_img, r = tile - n * tiles_per_img, ty = r / tx_n, tx = r - ty * tx_n;
      const int gy0 = ty * PT_H - 2, gx0 = tx * PT_W - 2;
      const int rlo = gy0 < 0 ? -gy0 : 0, rhi = min(PP_H, P.H - gy0), clo = gx0 < 0 ? -gx0 : 0, chi = min(PP_W, P.W - gx0);
      const uint32_t rmask = ((1u << rhi) - 1u) & ~((1u << rlo) - 1u), cmask = ((1u << chi) - 1u) & ~((1u << clo) - 1u);
      const bf16* org = P.dy + (((long long)n * P.H + gy0) * P.W + gx0) * CK;   // may lie before the map: only valid items are read
      const uint32_t dst0 = patch0 + (uint32_t)(pb * L::kPatch);
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        if (it_rc[k] < 0) continue;
        const bool valid = ((rmask >> (it_rc[k] >> 8)) & (cmask >> (it_rc[k] & 0xff)) & 1u) != 0u;
        cp_async16(dst0 + (uint32_t)it_dst[k], valid ? org + it_src[k] : P.dy, valid ? 16u : 0u);
      }
      cp_async_arrive(bar_pf + 8 * pb);
    };
    (void)ty_n;
    if (blockIdx.x < P.tiles) load_patch(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
      const int nt = tile + gridDim.x;
      if (RES) {
        if (nt < P.tiles) {
          mbar_wait(bar_pe + 8 * ((lt + 1) & 1), (uint32_t)((((lt + 1) >> 1) & 1) ^ 1));
          load_patch(nt, (lt + 1) & 1);
        }
      } else {
#pragma unroll 1
        for (int vg = 0; vg < L::NG; ++vg) {
          if (vg == 4 && nt < P.tiles) {           // next patch goes out while this tile's weights still stream
            mbar_wait(bar_pe + 8 * ((lt + 1) & 1), (uint32_t)((((lt + 1) >> 1) & 1) ^ 1));
            load_patch(nt, (lt + 1) & 1);
          }
          mbar_wait(bar_empty + 8 * s, eph);
          if (wrow < cinp) {
#pragma unroll
            for (int u = 0; u < L::VPS; ++u) {
              const int t = vg * L::VPS + u;
              if (t < 25) {
                const uint4* w = reinterpret_cast<const uint4*>(P.wT) + (size_t)(t * KC + wc0) * cinp + wrow;
                const uint32_t dst = ring + (uint32_t)(s * L::kStage + u * L::kBt + wrow * 16 + wc0 * cinp * 16);
#pragma unroll
                for (int c = 0; c < KC / 2; ++c) cp_async16(dst + (uint32_t)(c * 2 * cinp * 16), w + (size_t)c * 2 * cinp, 16u);
              }
            }
          }
          cp_async_arrive(bar_full + 8 * s);
          if (++s == STAGES) { s = 0; eph ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, cinp);
      const uint64_t dA0 = make_desc_nosw(0, PPL * 16, PP_W * 16);
      const uint32_t ldB = (uint32_t)((RES ? 25 : 1) * cinp * 16);
      const uint64_t dB0 = make_desc_nosw(0, ldB, 128);
      const uint32_t ahi = (uint32_t)(dA0 >> 32), bhi = (uint32_t)(dB0 >> 32), alo0 = (uint32_t)dA0, blo0 = (uint32_t)dB0;
      const uint32_t kstepB = (2 * ldB) >> 4, blo_res = (base & 0x3FFFFu) >> 4;
      int lt = 0, s = 0;
      uint32_t fph = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        const int set = lt & 1, pb = lt & 1;
        if (lt >= 2) mbar_wait(bar_free + 8 * set, (uint32_t)((lt >> 1) - 1) & 1u);
        mbar_wait(bar_pf + 8 * pb, (uint32_t)(lt >> 1) & 1u);
        fence_async_smem();
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(set * 64);
        const uint32_t plo = alo0 + (((patch0 + (uint32_t)(pb * L::kPatch)) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int vg = 0; vg < L::NG; ++vg) {
          uint32_t stlo = 0;
          if (!RES) {
            mbar_wait(bar_full + 8 * s, fph);
            fence_async_smem();
            tc_fence_after();
            stlo = ((ring + (uint32_t)(s * L::kStage)) & 0x3FFFFu) >> 4;
          }
#pragma unroll
          for (int u = 0; u < L::VPS; ++u) {
            const int t = vg * L::VPS + u;
            if (t >= 25) break;
            const int ky = t / KSZ, kx = t - ky * KSZ;
            const uint32_t alo = plo + (uint32_t)((4 - ky) * PP_W + (4 - kx));
            const uint32_t blo = blo0 + (RES ? blo_res + (uint32_t)(t * cinp) : stlo + (uint32_t)((u * L::kBt) >> 4));
#pragma unroll
            for (int kk = 0; kk < CK / 16; ++kk)
              mma_lh(acc, alo + (uint32_t)(kk * 2 * PPL), ahi, blo + kk * kstepB, bhi, idesc, (t | kk) == 0 ? 0u : 1u);
          }
          if (!RES) {
            tc_commit(bar_empty + 8 * s);
            if (++s == STAGES) { s = 0; fph ^= 1u; }
          }
        }
        tc_commit(bar_pe + 8 * pb);
        tc_commit(bar_accf + 8 * set);
      }
    }
  } else {
    // epilogue: TMEM lane r = tile pixel (r / 8, r % 8); fp32 NHWC store of the first nwrite channels (row stride ldx)
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int set = warp >= 9 ? 1 : 0;
    for (int lt = set, tile = blockIdx.x + set * gridDim.x; tile < P.tiles; tile += 2 * gridDim.x, lt += 2) {
      mbar_wait(bar_accf + 8 * set, (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t tm = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 64);
      float v[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16)
        if (c0 < cinp) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(tm + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) v[c0 + i] = __uint_as_float(r[i]);
        }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * set);
      const int n = tile / tiles_per_img, rr = tile - n * tiles_per_img, ty = rr / tx_n, tx = rr - ty * tx_n;
      const size_t g = ((size_t)n * P.H + ty * PT_H + (row >> 3)) * P.W + tx * PT_W + (row & 7);
      float* dst = P.dx + g * P.ldx;
      if ((P.nwrite & 3) == 0 && (P.ldx & 3) == 0) {
#pragma unroll
        for (int c = 0; c < 64; c += 4)
          if (c < P.nwrite) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < P.nwrite) dst[c] = v[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
  }
}

// w (cout, cin, 5, 5) fp32 -> wT [25][ckp/8][cinp][8] bf16: element (tap, k = co, row = ci)
__global__ void pack_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int ckp, int cinp, bf16* __restrict__ wT) {
  const int total = KSZ * KSZ * ckp * cinp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, ci = (i >> 3) % cinp, c8 = (i >> 3) / cinp % (ckp / 8), tap = i / (ckp * cinp);
    const int co = c8 * 8 + e;
    float v = 0.f;
    if (co < cout && ci < cin) v = w[((size_t)co * cin + ci) * (KSZ * KSZ) + tap];
    wT[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------ 3. wgrad
// instruction descriptor with both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) { return tc::make_idesc(M, N) | (1u << 15) | (1u << 16); }
// MN-major, no swizzle, [chunk][128 pixels][16 B] image: 8-pixel K blocks 128 B apart (LBO), 8-channel MN blocks 2 KB apart (SBO)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) { return make_desc_nosw(saddr, 128, BM * 16); }

struct WgradParams {
  const bf16* x;       // stage input [N][H][W][CX] bf16
  const bf16* dy;      // [N][H][W][cp] bf16
  float* partial;      // [gridDim.x][taps_padded][CX][cp] fp32 (this half's tap range only is written)
  int H, W, total, tiles, cp;
  int gph;             // tap groups per CTA (TMEM columns = gph * cp <= 512)
  int taps_padded;     // halves * gph * (128 / CX)
  long long* dbg;      // diagnostic (SD_TRACE_CNN=2)
};

constexpr int WG_THREADS = 288;    // warps 0-3 producers, 4 MMA, 5-8 epilogue (once, at the end)
constexpr int WG_STAGES = 5;
constexpr int kWgStage = 32 * 1024;   // one tap group: 128 (tap, channel) rows x 128 pixels
constexpr int kWgDy = 16 * 1024;      // dy tile slot (cp <= 64)
constexpr int kWgSmem = WG_STAGES * kWgStage + 2 * kWgDy + 8 * (2 * WG_STAGES + 6) + 16 + 256;

template <int CX>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams P) {
  constexpr int XC = CX / 8, TPG = 128 / CX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t dyb = base + WG_STAGES * kWgStage;
  const uint32_t bar_full = dyb + 2 * kWgDy, bar_empty = bar_full + 8 * WG_STAGES, bar_dyf = bar_empty + 8 * WG_STAGES,
                 bar_dye = bar_dyf + 16, bar_done = bar_dye + 16, tmem_slot = bar_done + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + WG_STAGES * kWgStage + 2 * kWgDy + 8 * (2 * WG_STAGES + 6));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cp = P.cp, DC = cp / 8;
  const int tap0 = blockIdx.y * P.gph * TPG;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_dyf + 8 * s, PROD);
      mbar_init(bar_dye + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
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
  const int HW = P.H * P.W;

  if (warp < 4) {
    const int tid = threadIdx.x;
    int s = 0, lt = 0;
    uint32_t eph = 1;
    const int rowpitch = P.W * CX;
    const bool dbg = P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
    long long t_empty = 0, t_dye = 0, c0 = 0, t0 = clock64();
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
      const int g = tile * BM + tid;
      const bool ok = g < P.total;
      const int n = g / HW, rem = g - n * HW, oy = rem / P.W, ox = rem - oy * P.W;
      // dy tile (the N operand of every group of this pixel tile)
      if (dbg) c0 = clock64();
      mbar_wait(bar_dye + 8 * (lt & 1), (uint32_t)(((lt >> 1) & 1) ^ 1));
      if (dbg) t_dye += clock64() - c0;
      {
        const bf16* src = ok ? P.dy + (size_t)g * cp : P.dy;
        const uint32_t dst = dyb + (uint32_t)((lt & 1) * kWgDy + tid * 16);
        for (int c = 0; c < DC; ++c) cp_async16(dst + (uint32_t)(c * BM * 16), src + c * 8, ok ? 16u : 0u);
        cp_async_arrive(bar_dyf + 8 * (lt & 1));
      }
      const bf16* src0 = P.x + ((size_t)(n * P.H + oy) * P.W + ox) * CX;
      uint32_t rmask = 0, cmask = 0;   // bit k: row oy + k - 2 / column ox + k - 2 inside the map
#pragma unroll
      for (int k = 0; k < KSZ; ++k) {
        rmask |= (uint32_t)(ok && oy + k - 2 >= 0 && oy + k - 2 < P.H) << k;
        cmask |= (uint32_t)(ox + k - 2 >= 0 && ox + k - 2 < P.W) << k;
      }
#pragma unroll 1
      for (int j = 0; j < P.gph; ++j) {
        if (dbg) c0 = clock64();
        mbar_wait(bar_empty + 8 * s, eph);
        if (dbg) t_empty += clock64() - c0;
        const uint32_t stg = base + (uint32_t)(s * kWgStage + tid * 16);
#pragma unroll
        for (int u = 0; u < TPG; ++u) {
          const int t = tap0 + j * TPG + u;
          const int ky = t / KSZ, kx = t - ky * KSZ;
          const bool valid = t < 25 && ((rmask >> ky) & (cmask >> kx) & 1u) != 0u;
          const bf16* src = valid ? src0 + (ky - 2) * rowpitch + (kx - 2) * CX : P.x;
#pragma unroll
          for (int c = 0; c < XC; ++c) cp_async16(stg + (uint32_t)((u * XC + c) * BM * 16), src + c * 8, valid ? 16u : 0u);
        }
        cp_async_arrive(bar_full + 8 * s);
        if (++s == WG_STAGES) { s = 0; eph ^= 1u; }
      }
    }
    if (dbg) { P.dbg[0] = clock64() - t0; P.dbg[1] = t_empty; P.dbg[2] = t_dye; P.dbg[3] = lt; }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(BM, cp);
      const uint64_t d0 = make_desc_mn(0);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo0 = (uint32_t)d0;
      int s = 0, lt = 0;
      uint32_t fph = 0;
      const bool dbg = P.dbg && blockIdx.x == 0 && blockIdx.y == 0;
      long long t_full = 0, t_dyf = 0, c0 = 0, t0 = clock64();
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        if (dbg) c0 = clock64();
        mbar_wait(bar_dyf + 8 * (lt & 1), (uint32_t)(lt >> 1) & 1u);
        if (dbg) t_dyf += clock64() - c0;
        const uint32_t blo = dlo0 + (((dyb + (uint32_t)((lt & 1) * kWgDy)) & 0x3FFFFu) >> 4);
#pragma unroll 1
        for (int j = 0; j < P.gph; ++j) {
          if (dbg) c0 = clock64();
          mbar_wait(bar_full + 8 * s, fph);
          if (dbg) t_full += clock64() - c0;
          fence_async_smem();
          tc_fence_after();
          const uint32_t alo = dlo0 + (((base + (uint32_t)(s * kWgStage)) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int kk = 0; kk < BM / 16; ++kk)     // K step = 16 pixels = two 128-byte K blocks
            mma_lh(tmem + (uint32_t)(j * cp), alo + (uint32_t)(kk * 16), dhi, blo + (uint32_t)(kk * 16), dhi, idesc, (lt | kk) == 0 ? 0u : 1u);
          tc_commit(bar_empty + 8 * s);
          if (++s == WG_STAGES) { s = 0; fph ^= 1u; }
        }
        tc_commit(bar_dye + 8 * (lt & 1));
      }
      tc_commit(bar_done);
      if (dbg) { P.dbg[4] = clock64() - t0; P.dbg[5] = t_full; P.dbg[6] = t_dyf; }
    }
  } else {
    // one epilogue at the very end: row r of group j = (tap tap0 + j TPG + r / CX, input channel r % CX)
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const long long e0 = clock64();
    mbar_wait(bar_done, 0);
    const long long e1 = clock64();
    tc_fence_after();
    const bool any = blockIdx.x < P.tiles;   // a CTA without pixel tiles never ran an MMA: its accumulators are undefined
    for (int j = 0; j < P.gph; ++j) {
      const int t = tap0 + j * TPG + row / CX, ci = row % CX;
      float* dst = P.partial + (((size_t)blockIdx.x * P.taps_padded + t) * CX + ci) * cp;
      for (int c0 = 0; c0 < cp; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * cp + c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + c0 + 4 * q) =
              any ? make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 5 * 32) { P.dbg[8] = e1 - e0; P.dbg[9] = clock64() - e1; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------ 3b. wgrad, patch-resident
// Same idea as the patch dgrad: a tile is a 16 x 8 pixel block (or two 8 x 8 maps) and the x views of all taps are windows
// of ONE halo patch of x, read through shifted MN-major descriptors (K block = 8 pixels of a tile row, LBO = the patch row
// pitch 192 B; MN block = 8 channels, SBO = the chunk plane).  An M = 128 operand stacks TPG = 128 / CX vertically
// adjacent taps (ky0 .. ky0 + TPG - 1, same kx): MN block b must sit at start + b * PLANE, so the patch is staged TPG
// times, copy j at + j * (XC * PLANE - 192): through the same descriptor, copy j shows the window one patch row further
// down.  Taps with ky > 4 are padding (they read zero rows; their accumulator rows are never stored).  Per tile this moves
// ~80 KB into shared memory instead of ~240 KB, and every cp.async instruction reads consecutive 16-byte pieces (the old
// thread-per-pixel gather touched 32 cache lines per instruction and was bound by that).
constexpr int WGP_ISSUERS = 4;                       // power of two
constexpr int WGP_THREADS = 32 * (4 + WGP_ISSUERS + 4);   // warps 0-3 producers, 4-7 MMA issuers, 8-11 epilogue
template <int CX>
struct WgradPatchCfg {
  static constexpr int XC = CX / 8, TPG = 128 / CX, NKG = (KSZ + TPG - 1) / TPG;     // tap groups per kx
  static constexpr int NGRP = KSZ * NKG;                                              // 15 (CX = 64) / 10 (CX = 32)
  static constexpr int SBR(int sh) { return sh + TPG + 4; }                           // patch rows per sub-block (zero rows below the halo)
  static constexpr int kDyPlane = BM * 16 + 16;
  static constexpr int kDy = 8 * kDyPlane;
  static int plane(int sh, int nsub) { return nsub * (sh + TPG + 4) * PP_W * 16 + 16; }
  static int patch_bytes(int sh, int nsub) { return ((TPG - 1) * (XC * plane(sh, nsub) - PP_W * 16) + XC * plane(sh, nsub) + 127) / 128 * 128; }
  static int smem(int sh, int nsub) { return 2 * patch_bytes(sh, nsub) + 2 * kDy + 8 * 12 + 16 + 256; }
};

struct WgradPatchParams {
  const bf16* x;       // [N][H][W][CX]
  const bf16* dy;      // [N][H][W][cp]
  float* partial;      // [gridDim.x][taps_padded][CX][cp]: tap index here = group * TPG + j, see wgrad_patch_reduce_kernel
  int H, W, tiles, cp, sh, nsub;   // tile = nsub sub-blocks of sh rows x 8 columns (16 x 1 block or 8 x 2 maps)
  int gph;             // groups per CTA (blockIdx.y picks the range)
  int ngroups_padded;  // gridDim.y * gph
  int patch_bytes;     // one patch buffer: (TPG - 1) * copy stride + XC * plane, rounded up to 128
  long long* dbg;      // diagnostic (SD_TRACE_CNN=2)
};

template <int CX>
__global__ void __launch_bounds__(WGP_THREADS, 1) conv_wgrad_patch_kernel(const __grid_constant__ WgradPatchParams P) {
  using Cfg = WgradPatchCfg<CX>;
  constexpr int XC = Cfg::XC, TPG = Cfg::TPG, NKG = Cfg::NKG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int sh = P.sh, nsub = P.nsub, sbr = sh + TPG + 4;
  const int plane = nsub * sbr * PP_W * 16 + 16;               // + 16: consecutive chunks land in different banks
  const int copy_stride = XC * plane - PP_W * 16;
  const uint32_t dyb = base + 2 * (uint32_t)P.patch_bytes;
  const uint32_t bar_pf = dyb + 2 * Cfg::kDy, bar_pe = bar_pf + 16, bar_done = bar_pe + 16, tmem_slot = bar_done + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + 2 * P.patch_bytes + 2 * Cfg::kDy + 48);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cp = P.cp, DC = cp / 8;
  // zero rows / padding are never written by the loads: clear both buffers once
  for (int i = threadIdx.x; i < 2 * P.patch_bytes / 16; i += WGP_THREADS) reinterpret_cast<uint4*>(gbase)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_pf + 8 * s, PROD);
      mbar_init(bar_pe + 8 * s, WGP_ISSUERS);
    }
    mbar_init(bar_done, WGP_ISSUERS);
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
  const int tx_n = P.W / PT_W;
  const int tiles_per_img = nsub == 1 ? tx_n * (P.H / sh) : tx_n;     // nsub == 2: a tile covers the same column block of two maps
  const int g0 = blockIdx.y * P.gph;

  if (warp < 4) {
    const int tid = threadIdx.x;
    const int prow = sh + 4, npx = nsub * prow * PP_W;                 // loaded patch pixels
    const int nitems = npx * XC, ndy = BM * DC;
    // this thread's items do not depend on the tile: (destination, source offset from the tile origin, patch row / column)
    constexpr int NI = CX == 64 ? 18 : 9, ND = 8;        // (2 x 12 x 12 patch pixels) x XC chunks / 128 threads
    int x_dst[NI], x_src[NI], x_rc[NI], d_dst[ND], d_src[ND];
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * PROD;
      const int c = i % XC, pp = i / XC, sb = pp / (prow * PP_W), q = pp - sb * (prow * PP_W), pr = q / PP_W, pc = q - pr * PP_W;
      x_dst[k] = c * plane + ((sb * sbr + pr) * PP_W + pc) * 16;
      x_src[k] = ((sb * P.H + pr) * P.W + pc) * CX + c * 8;
      x_rc[k] = i < nitems ? (pr << 8) | pc : -1;
    }
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const int i = tid + k * PROD;
      const int c = i % DC, p = i / DC, sb = p / (sh * PT_W), q = p - sb * (sh * PT_W), y = q >> 3, xx = q & 7;
      d_dst[k] = i < ndy ? c * Cfg::kDyPlane + p * 16 : -1;
      d_src[k] = ((sb * P.H + y) * P.W + xx) * cp + c * 8;
    }
    int lt = 0;
    const bool dbg = P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
    long long t_pe = 0, c0 = 0, t0 = clock64();
    // tile -> (first map, row block, column block), advanced by gridDim.x per step without dividing
    const int tyn = nsub == 1 ? P.H / sh : 1;
    int tx = blockIdx.x % tx_n, ty = (blockIdx.x / tx_n) % tyn, tn = blockIdx.x / (tx_n * tyn);
    const int dtx = gridDim.x % tx_n, dty = (gridDim.x / tx_n) % tyn, dtn = gridDim.x / (tx_n * tyn);
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
      const int pb = lt & 1;
      if (dbg) c0 = clock64();
      mbar_wait(bar_pe + 8 * pb, (uint32_t)(((lt >> 1) & 1) ^ 1));
      if (dbg) t_pe += clock64() - c0;
      const int n0 = tn * nsub;
      const int gy0 = ty * sh - 2, gx0 = tx * PT_W - 2;
      // bit k of rmask / cmask: patch row / column k lies inside the map (rows [max(0,-gy0), min(prow, H - gy0)))
      const int rlo = gy0 < 0 ? -gy0 : 0, rhi = min(prow, P.H - gy0), clo = gx0 < 0 ? -gx0 : 0, chi = min(PP_W, P.W - gx0);
      const uint32_t rmask = ((1u << rhi) - 1u) & ~((1u << rlo) - 1u), cmask = ((1u << chi) - 1u) & ~((1u << clo) - 1u);
      const uint32_t pdst = base + (uint32_t)(pb * P.patch_bytes);
      const bf16* xorg = P.x + (((long long)n0 * P.H + gy0) * P.W + gx0) * CX;   // may point before the map: only valid items are read
      // x patch: item = (patch pixel, chunk), chunk fastest: a warp instruction reads 512 consecutive bytes of a row
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        if (x_rc[k] < 0) continue;
        const bool valid = ((rmask >> (x_rc[k] >> 8)) & (cmask >> (x_rc[k] & 0xff)) & 1u) != 0u;
        const bf16* src = valid ? xorg + x_src[k] : P.x;
#pragma unroll
        for (int j = 0; j < TPG; ++j) cp_async16(pdst + (uint32_t)(x_dst[k] + j * copy_stride), src, valid ? 16u : 0u);
      }
      // dy tile: row p = tile pixel (p / 8, p % 8) of sub-block p / (8 sh)
      const bf16* dorg = P.dy + (((size_t)n0 * P.H + ty * sh) * P.W + tx * PT_W) * cp;
#pragma unroll
      for (int k = 0; k < ND; ++k)
        if (d_dst[k] >= 0) cp_async16(dyb + (uint32_t)(pb * Cfg::kDy + d_dst[k]), dorg + d_src[k], 16u);
      cp_async_arrive(bar_pf + 8 * pb);
      tx += dtx; ty += dty; tn += dtn;
      if (tx >= tx_n) { tx -= tx_n; ++ty; }
      if (ty >= tyn) { ty -= tyn; ++tn; }
    }
    if (dbg) { P.dbg[0] = clock64() - t0; P.dbg[1] = t_pe; P.dbg[3] = lt; }
  } else if (warp < 4 + WGP_ISSUERS) {
    // WGP_ISSUERS single-thread MMA issuers: issuer q owns groups j = q, q + WGP_ISSUERS, ... (disjoint accumulators).  One
    // thread cannot issue a tcgen05.mma more often than every ~60 cycles whatever its size (profiles/r02_umma_probe.txt),
    // and an N <= 64 MMA needs only half of that on the tensor pipe
    if (lane == 0) {
      const int q = warp - 4;
      const uint32_t idesc = make_idesc_mn(BM, cp);
      const uint64_t dA0 = make_desc_nosw(0, PP_W * 16, (uint32_t)plane), dB0 = make_desc_nosw(0, 128, Cfg::kDyPlane);
      const uint32_t ahi = (uint32_t)(dA0 >> 32), bhi = (uint32_t)(dB0 >> 32), alo0 = (uint32_t)dA0, blo0 = (uint32_t)dB0;
      // per K step (two tile rows): patch pixel offset of its first row; per group: (ky0, kx) -> ky0 * 12 + kx.  Nothing in the
      // issue loop below divides: one thread issues every MMA and each of its instructions costs ~10 cycles
      uint32_t rowoff[BM / 16];
#pragma unroll
      for (int kk = 0; kk < BM / 16; ++kk) {
        const int r2 = 2 * kk, sb = r2 / sh, y = r2 - sb * sh;
        rowoff[kk] = (uint32_t)((sb * sbr + y) * PP_W);
      }
      const int kx_first = g0 / NKG, kg_first = g0 - kx_first * NKG;
      int lt = 0;
      const bool dbg = P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && q == 0;
      long long t_pf = 0, c0 = 0, t0 = clock64();
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++lt) {
        const int pb = lt & 1;
        if (dbg) c0 = clock64();
        mbar_wait(bar_pf + 8 * pb, (uint32_t)(lt >> 1) & 1u);
        if (dbg) t_pf += clock64() - c0;
        fence_async_smem();
        tc_fence_after();
        const uint32_t plo = alo0 + (((base + (uint32_t)(pb * P.patch_bytes)) & 0x3FFFFu) >> 4);
        const uint32_t blo = blo0 + (((dyb + (uint32_t)(pb * Cfg::kDy)) & 0x3FFFFu) >> 4);
        const uint32_t first = lt == 0 ? 0u : 1u;
        int kx = kx_first, kg = kg_first;
        uint32_t acc = tmem;
#pragma unroll 1
        for (int j = 0; j < P.gph && kx < KSZ; ++j, acc += (uint32_t)cp) {
          if ((j & (WGP_ISSUERS - 1)) == q) {
            const uint32_t a = plo + (uint32_t)(kg * TPG * PP_W + kx);
#pragma unroll
            for (int kk = 0; kk < BM / 16; ++kk) mma_lh(acc, a + rowoff[kk], ahi, blo + (uint32_t)(kk * 16), bhi, idesc, kk == 0 ? first : 1u);
          }
          if (++kg == NKG) { kg = 0; ++kx; }
        }
        tc_commit(bar_pe + 8 * pb);
      }
      tc_commit(bar_done);
      if (dbg) { P.dbg[4] = clock64() - t0; P.dbg[5] = t_pf; }
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const bool any = blockIdx.x < P.tiles;
    for (int j = 0; j < P.gph; ++j) {
      const int gi = g0 + j;
      if (gi / NKG >= KSZ) break;
      // accumulator row = (stacked tap row / CX, channel row % CX); slot = group * TPG + stacked tap
      float* dst = P.partial + (((size_t)blockIdx.x * P.ngroups_padded * TPG + (size_t)gi * TPG + row / CX) * CX + row % CX) * cp;
      for (int c0 = 0; c0 < cp; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * cp + c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + c0 + 4 * q) =
              any ? make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// partial slot (group gi = kx * NKG + kg, stacked tap j) holds tap (ky = kg * TPG + j, kx); dw (cout, cin, 5, 5) += sum_b
__global__ void wgrad_patch_reduce_kernel(const float* __restrict__ partial, int nblocks, int slots_padded, int tpg, int nkg, int cx, int cp,
                                          int cout, int cin, float* __restrict__ dw) {
  const int total4 = KSZ * nkg * tpg * cx * cp / 4;          // cp % 16 == 0: four adjacent output channels per thread
  const size_t stride = (size_t)slots_padded * cx * cp;
  for (int i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += gridDim.x * blockDim.x) {
    const int i = i4 * 4;
    const int co = i % cp, ci = (i / cp) % cx, slot = i / (cp * cx);
    const int gi = slot / tpg, j = slot - gi * tpg, kx = gi / nkg, ky = (gi - kx * nkg) * tpg + j;
    if (co >= cout || ci >= cin || ky >= KSZ) continue;
    const float4 s = ordered_sum4(partial + i, nblocks, stride);
    float* dst = dw + ((size_t)co * cin + ci) * (KSZ * KSZ) + ky * KSZ + kx;
    const size_t os = (size_t)cin * (KSZ * KSZ);
    dst[0] += s.x;
    if (co + 1 < cout) dst[os] += s.y;
    if (co + 2 < cout) dst[2 * os] += s.z;
    if (co + 3 < cout) dst[3 * os] += s.w;
  }
}

// dw (cout, cin, 5, 5) += sum over CTAs of partial[b][tap][ci][co]; fixed order = run-to-run identical
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nblocks, int taps_padded, int cx, int cp, int cout, int cin,
                                    float* __restrict__ dw) {
  // threads walk the partial's own order (co fastest): coalesced reads; the 25 * cin * cout scattered writes are few
  const int total = KSZ * KSZ * cx * cp;
  const size_t stride = (size_t)taps_padded * cx * cp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cp, ci = (i / cp) % cx, tap = i / (cp * cx);
    if (co >= cout || ci >= cin) continue;
    float s = ordered_sum(partial + i, nblocks, stride);
    dw[((size_t)co * cin + ci) * (KSZ * KSZ) + tap] += s;
  }
}

// ------------------------------------------------------------------------------------------------ 3'. wgrad of stage 1
// M rows = the forward's K index k = ky * 16 + kx * 3 + c (80 used of 128), K = 128 conv pixels (BM / W full rows of one
// frame), N = cp.  The im2col operand is built from an fp32 patch of the frame exactly like the forward's.
struct Wgrad1Params {
  const float* obs;    // [N][H][W][3]
  const bf16* dy;      // [N][H][W][cp]
  float* partial;      // [gridDim.x][128][cp]
  int H, W, total, tiles, cp;
};
constexpr int kW1A = 16 * BM * 16;      // 32 KB: [16 chunks][128 px][16 B], chunks 10..15 stay zero
constexpr int W1_THREADS = 32 * 13;     // warps 0-3 / 4-7: two producer groups (tiles alternate), 8 MMA, 9-12 epilogue
constexpr int W1_MMA_WARP = 8;
constexpr int kW1Smem = 2 * kW1A + 2 * kWgDy + 2 * kPatchFloats * 4 + 8 * 10 + 16 + 256;

__global__ void __launch_bounds__(W1_THREADS, 1) conv1_wgrad_kernel(const __grid_constant__ Wgrad1Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t dyb = base + 2 * kW1A;
  const int off_bar = 2 * kW1A + 2 * kWgDy + 2 * kPatchFloats * 4;
  const uint32_t bar_full = base + (uint32_t)off_bar, bar_empty = bar_full + 16, bar_done = bar_empty + 16, tmem_slot = bar_done + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 48);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cp = P.cp, DC = cp / 8;
  for (int i = threadIdx.x; i < 2 * kW1A / 16; i += W1_THREADS) reinterpret_cast<uint4*>(gbase)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, 2 * PROD);   // each producer thread: one plain arrive (operand built) + one cp.async arrive (dy landed)
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W1_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const int R = BM / P.W;                          // conv rows per tile
  const int rowf = P.W * 3, PR = R + 4, PW = rowf + 16;
  const int tiles_per_frame = P.H / R;

  if (warp < 8) {
    // producer group grp builds the operand of tiles it = grp, grp + 2, ... in stage grp from its own patch buffer
    const int grp = warp >> 2, tid = threadIdx.x & (PROD - 1);
    float* patch = reinterpret_cast<float*>(gbase + 2 * kW1A + 2 * kWgDy) + grp * kPatchFloats;
    const int lrow = tid / P.W, ox = tid - lrow * P.W;
    const int r4 = rowf >> 2, n4 = PR * r4;
    for (int i = tid; i < PR * 4; i += PROD) {
      const int r = i >> 2, q = i & 3;
      reinterpret_cast<float4*>(patch + r * PW + (q < 2 ? q * 4 : rowf + q * 4))[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int NI = 4;                            // float4 patch items per thread ((R + 4) rows x W * 3 / 4 <= 512)
    int rc[NI];
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * PROD;
      rc[k] = i < n4 ? ((i / r4) << 16) | (i % r4) : -1;
    }
    const int s = grp;
    int fn = (blockIdx.x + grp * gridDim.x) / tiles_per_frame, ft = (blockIdx.x + grp * gridDim.x) - fn * tiles_per_frame;
    const int dn = (2 * gridDim.x) / tiles_per_frame, dt = (2 * gridDim.x) - dn * tiles_per_frame;
    int use = 0;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < P.tiles; tile += 2 * gridDim.x, ++use) {
      const int oy0 = ft * R;
      // frame rows first: their latency overlaps the wait for the stage
      float4 pre[NI];
      const float* frame = P.obs + (size_t)fn * P.H * rowf;
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        pre[k] = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
        const int iy = oy0 - 2 + (rc[k] >> 16);
        if (rc[k] >= 0 && iy >= 0 && iy < P.H) pre[k] = __ldg(reinterpret_cast<const float4*>(frame + (size_t)iy * rowf) + (rc[k] & 0xffff));
      }
      if (use >= 1) mbar_wait(bar_empty + 8 * s, (uint32_t)(use - 1) & 1u);
      {
        const size_t g = (size_t)tile * BM + tid;
        const uint32_t dst = dyb + (uint32_t)(s * kWgDy + tid * 16);
        for (int c = 0; c < DC; ++c) cp_async16(dst + (uint32_t)(c * BM * 16), P.dy + g * cp + c * 8, 16u);
        cp_async_arrive(bar_full + 8 * s);
      }
      named_bar(1 + grp, PROD);                    // the previous tile's build has finished reading the patch
#pragma unroll
      for (int k = 0; k < NI; ++k)
        if (rc[k] >= 0)
          reinterpret_cast<float4*>(patch + (rc[k] >> 16) * PW + 8)[rc[k] & 0xffff] =
              make_float4(pre[k].x - 0.5f, pre[k].y - 0.5f, pre[k].z - 0.5f, pre[k].w - 0.5f);
      named_bar(1 + grp, PROD);
      uint8_t* st = gbase + (size_t)s * kW1A;
      // input pixel (ox - 2 + j) of patch row (lrow + ky) starts at float 8 + (ox - 2 + j) * 3 = 2 + 3 ox + 3 j
      const float* prow = patch + lrow * PW + 2 + 3 * ox;
#pragma unroll
      for (int ky = 0; ky < KSZ; ++ky) {
        float u[16];
#pragma unroll
        for (int e = 0; e < 15; ++e) u[e] = prow[ky * PW + e];
        u[15] = 0.f;
        uint4* dst = reinterpret_cast<uint4*>(st + (size_t)(ky * 2) * BM * 16 + tid * 16);
        dst[0] = make_uint4(pack2(u[0], u[1]), pack2(u[2], u[3]), pack2(u[4], u[5]), pack2(u[6], u[7]));
        dst[BM] = make_uint4(pack2(u[8], u[9]), pack2(u[10], u[11]), pack2(u[12], u[13]), pack2(u[14], u[15]));
      }
      fence_async_smem();
      mbar_arrive(bar_full + 8 * s);
      fn += dn; ft += dt;
      if (ft >= tiles_per_frame) { ft -= tiles_per_frame; ++fn; }
    }
  } else if (warp == W1_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(BM, cp);
      const uint64_t d0 = make_desc_mn(0);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo0 = (uint32_t)d0;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(bar_full + 8 * s, (uint32_t)(it >> 1) & 1u);
        fence_async_smem();
        tc_fence_after();
        const uint32_t alo = dlo0 + (((base + (uint32_t)(s * kW1A)) & 0x3FFFFu) >> 4);
        const uint32_t blo = dlo0 + (((dyb + (uint32_t)(s * kWgDy)) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk) mma_lh(tmem, alo + (uint32_t)(kk * 16), dhi, blo + (uint32_t)(kk * 16), dhi, idesc, (it | kk) == 0 ? 0u : 1u);
        tc_commit(bar_empty + 8 * s);
      }
      tc_commit(bar_done);
    }
  } else {
    const int quarter = warp & 3, row = quarter * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const bool any = blockIdx.x < P.tiles;
    float* dst = P.partial + ((size_t)blockIdx.x * BM + row) * cp;
    for (int c0 = 0; c0 < cp; c0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(dst + c0 + 4 * q) =
            any ? make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]))
                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W1_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
  }
}

// dw1 (cout, 3, 5, 5) += sum_b partial[b][ky * 16 + kx * 3 + c][co]
__global__ void wgrad1_reduce_kernel(const float* __restrict__ partial, int nblocks, int cp, int cout, float* __restrict__ dw) {
  const int total = 80 * cp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cp, k = i / cp, ky = k >> 4, r = k & 15;
    if (co >= cout || r >= 15) continue;
    float s = ordered_sum(partial + i, nblocks, (size_t)BM * cp);
    dw[((size_t)co * 3 + (r % 3)) * (KSZ * KSZ) + ky * KSZ + r / 3] += s;
  }
}

}  // namespace cnn
}  // namespace sd
