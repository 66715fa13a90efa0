// sd_bwd.cuh -- row-wise kernels of the reverse-time backward scans (K2: observe, K4: imagine).
//
// Math: SURVEY.md Appendix A (checked there against autograd in fp64); restated in
// oracle/rssm_oracle.py:{normact_bwd, sample_bwd, deter_step_bwd} which the tests compare against.
// Dense transposes (dgrad) reuse gemm_f32_kernel with the k-contiguous weight copy; weight gradients are
// ONE batched contraction over all T*B rows after the scan (wgrad_f32_kernel), fed by the per-step
// "d-tape" these kernels write.  All reductions have a fixed order (deterministic gradients).
#pragma once
#include "sd_kernels.cuh"

namespace sd {

// d(out)/d(v) of out = SiLU(RMSNorm_width(v) * w)  (networks.py:325-327 / rssm.py:16-31).
//   dv   = rho * (dn - n * mean(dn * n)),  dn = dm * w,  dm = dout * silu'(m),  m = n * w,  n = v * rho
//   dmn  = dm * n  (per-row contribution to the RMS scale gradient; column-summed after the scan)
struct NormActBwdP {
  const float* dout; int ld_dout;  // grad w.r.t. the activation output
  int nsum; long long sum_stride;  // dout[c] = sum_{g < nsum} dout[g*sum_stride + c] (block-diagonal consumers); 0/1 = plain
  const float* v;    int ld_v;     // saved pre-norm values
  const float* w;                  // [width]
  float* dv;  int ld_dv;
  __nv_bfloat16* dv_bf;            // nullable bf16 copy of dv (same row stride) for the tcgen05 dgrad
  float* dmn; int ld_dmn;          // nullable (dgrad-only)
  int width;
};
struct NormActBwdBatch {
  int count;
  NormActBwdP p[4];
};
__global__ void __launch_bounds__(256) normact_bwd_kernel(const NormActBwdBatch b) {
  __shared__ float sh[32];
  const NormActBwdP& p = b.p[blockIdx.y];
  const size_t row = blockIdx.x;
  const float* v = p.v + row * p.ld_v;
  const float* dout = p.dout + row * p.ld_dout;
  float vv[8], dn[8], nn[8], gw[8], dsl[8], dy[8];
  float ss = 0.f;
  // Everything that only needs the forward tape (saved pre-norm values, RMS scale) is done BEFORE the PDL wait: the row
  // statistic (one block reduction), n = v * rho and silu'(m).  After the wait only `dout` is fetched -- all slices of all
  // eight elements in flight together -- followed by the second reduction.
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    vv[i] = (c < p.width) ? __ldg(v + c) : 0.f;
    gw[i] = (c < p.width) ? __ldg(p.w + c) : 0.f;
    ss = fmaf(vv[i], vv[i], ss);
  }
  ss = block_sum(ss, sh);
  const float rho = 1.f / sqrtf(ss / (float)p.width + kRmsEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float n = vv[i] * rho;
    const float m = n * gw[i];
    const float sg = sigmoidf_(m);
    nn[i] = n;
    dsl[i] = sg * (1.f + m * (1.f - sg));
  }
  pdl_prologue();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    dy[i] = (c < p.width) ? dout[c] : 0.f;
  }
  for (int g = 1; g < p.nsum; ++g) {   // fixed order => deterministic; the eight loads of a slice are independent
    const float* ds = dout + g * p.sum_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < p.width) dy[i] += ds[c];
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    dn[i] = 0.f;
    if (c < p.width) {
      const float dm = dy[i] * dsl[i];
      if (p.dmn) p.dmn[row * p.ld_dmn + c] = dm * nn[i];
      dn[i] = dm * gw[i];
      dot = fmaf(dn[i], nn[i], dot);
    } else {
      nn[i] = 0.f;
    }
  }
  dot = block_sum(dot, sh) / (float)p.width;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < p.width) {
      const float o = rho * (dn[i] - nn[i] * dot);
      p.dv[row * p.ld_dv + c] = o;
      if (p.dv_bf) p.dv_bf[row * p.ld_dv + c] = __float2bfloat16(o);
    }
  }
}

// Backward of the straight-through unimix Gumbel sample w.r.t. the raw logits (oracle: sample_bwd),
// recomputing p = softmax(logit) and y = softmax(l + g) from the saved logits and uniforms.
//   gz = carry (+ upstream): grad w.r.t. the sampled one-hot;  d_logit = upstream_logit + p*(dp - <dp,p>)
template <int GS>
__global__ void sample_bwd_kernel(const float* __restrict__ logits, int ld_l, const float* __restrict__ u, int ld_u,
                                  const float* __restrict__ gz_a, int ld_a, const float* __restrict__ gz_b, int ld_b,
                                  const float* __restrict__ up_logit, int ld_ul, int R, int S, int K, float unimix,
                                  float* d_logit, int ld_d, __nv_bfloat16* d_logit_bf, const float* __restrict__ a_scale) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long cat = t / GS;
  const int k = (int)(t % GS);
  const bool in_range = cat < (long long)R * S;
  const size_t row = in_range ? (size_t)(cat / S) : 0;
  const int sidx = in_range ? (int)(cat - (long long)row * S) : 0;
  const bool valid = in_range && k < K;
  const int col = sidx * K + k;
  // Everything that only needs the forward tape (saved logits, uniforms) is done BEFORE the PDL wait: the
  // re-sampling softmaxes y and p are most of this kernel's arithmetic.
  float lg = 0.f, uu = 0.5f;
  if (valid) {
    lg = __ldg(logits + row * ld_l + col);
    uu = __ldg(u + row * ld_u + col);
  }
  float y;
  (void)sample_group<GS>(lg, uu, valid, k, K, unimix, &y);
  if (!valid) y = 0.f;
  const float m = group_max<GS>(valid ? lg : -INFINITY);
  const float e = valid ? expf(lg - m) : 0.f;
  const float s = group_sum<GS>(e);
  const float p = e / s;
  const float pt = p * (1.f - unimix) + unimix / (float)K;
  pdl_prologue();
  float gz = 0.f;
  if (valid) {
    if (gz_a) gz += gz_a[row * ld_a + col] * (a_scale ? a_scale[row] : 1.f);  // a_scale: reset cut of the later step
    if (gz_b) gz += gz_b[row * ld_b + col];
  }
  // dl = y * (gz - <gz, y>);  d(log p~) = dl - p~ * sum(dl);  dp = d(log p~)/p~ * (1-unimix);  d_logit = p*(dp - <dp,p>)
  const float gy = group_sum<GS>(gz * y);
  const float dl = y * (gz - gy);
  const float sdl = group_sum<GS>(dl);
  const float dlp = dl - pt * sdl;
  const float dp = valid ? dlp / pt * (1.f - unimix) : 0.f;
  const float dpp = group_sum<GS>(dp * p);
  if (valid) {
    float out = p * (dp - dpp);
    if (up_logit) out += up_logit[row * ld_ul + col];
    d_logit[row * ld_d + col] = out;
    if (d_logit_bf) d_logit_bf[row * ld_d + col] = __float2bfloat16(out);
  }
}

// Backward of the GRU-style gates (rssm.py:63-75): given g = d(deter'), the saved gate pre-activations q
// (R, 3D) [g][reset|cand|update][Dg] and the step's input deter, emit dq (same layout) and the direct
// path dd = g * (1 - update).
// g = ga + gb + gc (gb, gc nullable): carry from step t+1, upstream d_deters[:, t] and the gradient coming back
// through the posterior / prior net, summed here instead of in a separate kernel.
// When `ga2` / `dxin` are given the carry is assembled here instead of by carry_kernel:
//   carry = (ga + ga2 + dxin[row][g][o]) * a_scale[row]   (d(deter_in) pieces of step t+1 and its reset cut).
__global__ void gates_bwd_kernel(const float* __restrict__ ga, int ld_a, const float* __restrict__ gb, int ld_b,
                                 const float* __restrict__ gc, int ld_c, const float* __restrict__ ga2,
                                 const float* __restrict__ dxin, int G, int Kb, const float* __restrict__ a_scale,
                                 const float* __restrict__ q,
                                 const float* __restrict__ deter_in, int ld_in, float* dq, __nv_bfloat16* dq_bf, float* dd,
                                 int R, int D, int Dg) {
  const long long total = (long long)R * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // first element's taped operands (gate pre-activations, step input) are fetched before the PDL wait
  float pr_ = 0.f, pc_ = 0.f, pu_ = 0.f, pd_ = 0.f;
  if (i < total) {
    const int d = (int)(i % D);
    const size_t row = (size_t)(i / D);
    const int gi = d / Dg, o = d - gi * Dg;
    const size_t qo = row * 3 * D + (size_t)gi * 3 * Dg + o;
    pr_ = __ldg(q + qo); pc_ = __ldg(q + qo + Dg); pu_ = __ldg(q + qo + 2 * Dg);
    pd_ = __ldg(deter_in + row * ld_in + d);
  }
  // ... and so are the gate activations recomputed from them
  const float pRg = sigmoidf_(pr_), pC = tanhf(pRg * pc_), pUu = sigmoidf_(pu_ - 1.f);
  pdl_prologue();
  for (bool first = true; i < total; i += stride, first = false) {
    const int d = (int)(i % D);
    const size_t row = (size_t)(i / D);
    const int gi = d / Dg, o = d - gi * Dg;
    const size_t qo = row * 3 * D + (size_t)gi * 3 * Dg + o;
    const float r = first ? pr_ : q[qo], c = first ? pc_ : q[qo + Dg], uu = first ? pu_ : q[qo + 2 * Dg];
    const float din = first ? pd_ : deter_in[row * ld_in + d];
    const float Rg = first ? pRg : sigmoidf_(r), C = first ? pC : tanhf(Rg * c), Uu = first ? pUu : sigmoidf_(uu - 1.f);
    float carry = ga ? ga[row * ld_a + d] : 0.f;
    if (ga2) carry += ga2[row * D + d];
    if (dxin) carry += dxin[(row * G + gi) * (size_t)Kb + o];
    if (a_scale) carry *= a_scale[row];
    const float gd = carry + (gb ? gb[row * ld_b + d] : 0.f) + (gc ? gc[row * ld_c + d] : 0.f);
    const float dUu = gd * (C - din);
    const float dC = gd * Uu;
    const float dtn = dC * (1.f - C * C);
    const float o_r = (dtn * c) * Rg * (1.f - Rg), o_c = dtn * Rg, o_u = dUu * Uu * (1.f - Uu);
    dq[qo] = o_r;
    dq[qo + Dg] = o_c;
    dq[qo + 2 * Dg] = o_u;
    if (dq_bf) {
      dq_bf[qo] = __float2bfloat16(o_r);
      dq_bf[qo + Dg] = __float2bfloat16(o_c);
      dq_bf[qo + 2 * Dg] = __float2bfloat16(o_u);
    }
    dd[row * D + d] = gd * (1.f - Uu);
  }
}

// End of a reverse step: carry_d = (dd + d_din0) * keep, carry_z = dz * keep, keep = 1 - is_first saved by the
// forward (rssm.py:161-165).
// `extra_z/extra_d` (nullable) are added before masking (imagination: grads through the actor's feat input).
__global__ void carry_kernel(const float* __restrict__ dd, const float* __restrict__ d_din0,
                             const float* __restrict__ dz, const float* __restrict__ keep_mask,
                             const float* __restrict__ extra, int ld_x, const float* __restrict__ extra2, int ld_x2,
                             const float* __restrict__ dxin, int G, int Dg, int Kb,
                             int R, int SK, int D, float* carry_z, float* carry_d) {
  pdl_prologue();
  const int W = SK + D;
  const long long total = (long long)R * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t row = (size_t)(i / W);
    const int c = (int)(i - (long long)row * W);
    const float keep = keep_mask ? keep_mask[row] : 1.f;
    const float ex = (extra ? extra[row * ld_x + c] : 0.f) + (extra2 ? extra2[row * ld_x2 + c] : 0.f);
    if (c < SK) carry_z[row * SK + c] = ((dz ? dz[row * SK + c] : 0.f) + ex) * keep;
    else {
      const int d = c - SK;
      // dxin (R, G, Kb): gradient of the block input [deter_g | x]; its deter part belongs to unit d = g*Dg + o
      const float dh = dxin ? dxin[(row * G + d / Dg) * (size_t)Kb + (d % Dg)] : 0.f;
      carry_d[row * D + d] = ((dd ? dd[row * D + d] : 0.f) + (d_din0 ? d_din0[row * D + d] : 0.f) + dh + ex) * keep;
    }
  }
}

// out[c] += sum_r in[r][c]: column sums for bias / RMS-scale gradients in a fixed order.
// Block = 32 columns x 32 row lanes; each lane sums rows lane, lane+32, ... then the 32 lane partials are
// added in ascending lane order => deterministic.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ in, int ld, int R, int W, float* out) {
  pdl_prologue();
  __shared__ float sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < W)
    for (int r = threadIdx.y; r < R; r += 32) s += in[(size_t)r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += sh[j][threadIdx.x];
    out[c] += t;
  }
}

// The same column sums for up to 16 matrices of R rows in ONE launch (blockIdx.y = matrix): the bias / RMS-scale gradients of
// all layers after the reverse scan.  Eight row loads per lane are in flight at a time; the additions keep colsum_kernel's
// order (rows lane, lane+32, ... then lanes ascending), so the results are bit-identical to 12 separate launches, which cost
// 10 us each on the critical path (8-CTA grids, 32 dependent loads per lane).
struct ColsumP { const float* in; int ld; int W; float* out; };
struct ColsumBatch { int count; int R; ColsumP p[16]; };
__global__ void __launch_bounds__(1024) colsum_batch_kernel(const ColsumBatch b) {
  pdl_prologue();
  __shared__ float sh[32][33];
  const ColsumP& p = b.p[blockIdx.y];
  if ((int)blockIdx.x * 32 >= p.W) return;   // block-uniform
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < p.W) {
    for (int r0 = threadIdx.y; r0 < b.R; r0 += 32 * 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + 32 * j;
        v[j] = r < b.R ? p.in[(size_t)r * p.ld + c] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + 32 * j < b.R) s += v[j];
    }
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < p.W) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += sh[j][threadIdx.x];
    p.out[c] += t;
  }
}

// Backward of the bounded-normal actor sample (distributions.py:217-222): action = tanh(mean) + std*eps,
// std = (max-min)*sigmoid(sraw+2)+min.  d_act = upstream + d(abar)/max(|a|,1) (rssm.py:44, detached clip).
__global__ void actor_sample_bwd_kernel(const float* __restrict__ out, const float* __restrict__ eps, int ld_n,
                                        const float* __restrict__ action, int ld_act, const float* __restrict__ d_up,
                                        int ld_up, const float* __restrict__ d_abar, int R, int A, float min_std,
                                        float max_std, float* d_out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * A) return;
  const size_t row = i / A;
  const int a = i - (int)row * A;
  const float act = action[row * ld_act + a];
  float da = d_up ? d_up[row * ld_up + a] : 0.f;
  if (d_abar) da += d_abar[row * A + a] / fmaxf(fabsf(act), 1.f);
  const float mean = out[row * 2 * A + a], sraw = out[row * 2 * A + A + a];
  const float tm = tanhf(mean);
  const float sg = sigmoidf_(sraw + 2.f);
  d_out[row * 2 * A + a] = da * (1.f - tm * tm);
  d_out[row * 2 * A + A + a] = da * eps[row * ld_n + a] * (max_std - min_std) * sg * (1.f - sg);
}

}  // namespace sd
