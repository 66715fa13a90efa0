// sd_kernels.cuh -- fp32 SIMT kernels and the fused row-wise kernels of the RSSM hot path.
//
// Everything here computes in fp32 with IEEE division / sqrt and the accurate expf/logf/tanhf
// (the library is built WITHOUT --use_fast_math): this is the parity path that reproduces the
// reference's sampled category indices bit for bit except at fp32 near-ties.  The bf16 tensor-core
// GEMM that replaces gemm_f32_kernel for large row counts lives in sd_tc.cuh; all row-wise kernels
// below are shared by both paths (they optionally emit a bf16 copy for the next tcgen05 operand).
//
// Reference math: world_model/rssm.py:36-75 (Deter), :158-195 (obs_step / img_step / prior),
// world_model/distributions.py:16-36 (OneHotDist), :78-98 (TwoHot.mode), :217-222 (bounded_normal),
// world_model/dreamer.py:694-707 (_lambda_return).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sd {

constexpr float kRmsEps = 1e-4f;

// Programmatic dependent launch prologue: wait until every prerequisite grid has completed and its writes are
// visible, then let the next kernel in the stream get scheduled (it will block in its own wait until this grid
// is done).  Triggering only AFTER the wait keeps the look-ahead at depth one: a kernel that runs early can
// overlap with its immediate predecessor only, which is what makes the weight prefetch in gemm_f32_kernel
// safe (weights are written by pack_weight_kernel, which never triggers early).  No-ops without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_wait();
  pdl_trigger();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.f + expf(-x)); }

// 3xTF32 on mma.sync (m16n8k8): x = hi + lo with hi = tf32(x), lo = tf32(x - hi); A*B ~ hi*hi + hi*lo + lo*hi accumulated
// in fp32.  The dropped lo*lo term is 2^-22 relative, so the result is fp32-class (all fp32 parity tolerances hold),
// while the 16-row skinny tiles of the small-batch path stop being shared-memory / FMA-issue bound.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


// Lane-per-class variant of sample_category for the hot kernels: the group of `GS` lanes (GS = pow2 >= K)
// that shares a category reduces with shuffles.  Same operation sequence per element as above.
template <int GS>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int GS>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// returns the first-max class index of y = softmax(l + g) for this lane's category (valid on all lanes)
template <int GS>
__device__ __forceinline__ int sample_group(float lg, float u, bool valid, int k, int K, float unimix, float* y_out) {
  const float NEG = -INFINITY;
  const float m = group_max<GS>(valid ? lg : NEG);
  float e = valid ? expf(lg - m) : 0.f;
  const float s = group_sum<GS>(e);
  float l = valid ? logf((e / s) * (1.f - unimix) + unimix / (float)K) : NEG;
  const float m2 = group_max<GS>(l);
  const float s2 = group_sum<GS>(valid ? expf(l - m2) : 0.f);
  const float lse = m2 + logf(s2);
  float z = valid ? (l - lse) + (-logf(-logf(u))) : NEG;
  const float m3 = group_max<GS>(z);
  const float ez = valid ? expf(z - m3) : 0.f;
  const float s3 = group_sum<GS>(ez);
  const float y = valid ? ez / s3 : NEG;
  if (y_out) *y_out = y;
  // arg-first-max over the group: max value, ties -> smallest index
  float bv = y;
  int bi = valid ? k : 0x7fffffff;
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}

// ------------------------------------------------------------------------------------------------
// Batched skinny GEMM:  C[R x N] = [A | A2][R x K] * Wt[K x N] + bias, fp32.
// A CTA owns a 16-row x 16-column output tile: the 16 rows are the M of mma.sync.m16n8k8 (3xTF32 split, fp32-class
// accuracy), K is split over the 8 warps and the partial tiles are summed in warp order through shared memory
// (fixed order => deterministic).  Built for the
// latency-bound regime of the posterior scan (R = 16): the grid spreads N over many SMs and every
// thread does <= ~2k FMAs.  Rows are tiled by 16, so it is also the fp32 parity path for large R.
// ------------------------------------------------------------------------------------------------
struct GemmP {
  const float* A;    // [R x K1], row stride lda
  const float* A2;   // [R x (K-K1)], row stride lda2 (second K segment; may be null when K1 == K)
  const float* Wt;   // [K x ldw] (n contiguous, ldw % 16 == 0, zero padded)
  const float* bias; // [N] or null
  float* C;          // [R x N], row stride ldc (nullable for the fused epilogues when no tape is kept)
  int lda, lda2, K1, K, ldw, ldc, N;
  // fused epilogues (K <= 512 only, i.e. no split-K cluster along y):
  //  EPI_GATES : the problem is one block of the gate projection, N = Dg; a cluster of 3 CTAs computes the
  //              reset / cand / update tiles of the same 16 units and the leader applies the GRU gate math
  //              (rssm.py:63-75) writing deter' (e_out); C receives q in the reference layout [r|c|u].
  //  EPI_SAMPLE: N = S*Kc logits; each 16-column tile holds 16/Kc whole categories; the CTA samples them
  //              (distributions.py:16-36) from uniforms e_in, writes the one-hot (e_out), the logits (e_out2)
  //              and C (raw logits for the backward tape).
  int epi;
  const float* e_in; int e_ld_in;
  float* e_out; int e_ld_out;
  float* e_out2; int e_ld_out2;
  int e_k;           // Dg (gates) or Kc (sample)
  float e_f;         // unimix (sample)
  // fused A-operand prologue (reverse scans; single k-slice, K <= 256, 16-byte aligned rows):
  //  PRE_NORMBWD: A is d(out) of a [Linear -> RMSNorm -> SiLU] layer; the CTA turns it into d(pre-norm) on the fly
  //               (dv = rho*(dn - n*mean(dn*n)), dn = dout*silu'(m)*w) from the saved pre-norm values
  //               pre_v and the RMS scale pre_w, instead of a separate normact_bwd launch.  The CTA of column tile 0 of a
  //               problem with pre_write also stores dv / dm*n (the d-tape the weight-gradient pass reads).
  int pre, pre_write;
  const float* pre_v; int pre_ldv;
  const float* pre_w;
  float* pre_dv; int pre_lddv;
  float* pre_dmn; int pre_lddmn;
  // fused epilogue of the reverse posterior scan (single k-slice):
  //  EPI_GATESBWD: the problem is the deter' column range of the dgrad of obs_net_0 at step t (C = gradient coming back
  //              through the posterior net, N = D); the thread that owns output element (row, n) continues with
  //              gates_bwd_kernel of step t for unit n (the gate pre-activations and the step's input deter are forward-tape
  //              data: fetched, and the gate activations recomputed, BEFORE the PDL wait):
  //              x0/xi0 carry dd (nullable), x1/xi1 upstream d(deter) (nullable), x2 carry t_din0 (nullable, stride D),
  //              x3 carry block-input gradient dxin (nullable; xi2 = G, xi3 = Kb), x4 reset cut (nullable), x5 gate
  //              pre-activations q (stride 3D), x6/xi4 the step's input deter, y0 dq out, y1 dd out, e_k = Dg; D = N.
  const float *x0, *x1, *x2, *x3, *x4, *x5, *x6;
  float *y0, *y1;
  int xi0, xi1, xi2, xi3, xi4;
};
enum { PRE_NONE = 0, PRE_NORMBWD = 1 };
enum { EPI_STORE = 0, EPI_GATES = 1, EPI_SAMPLE = 2, EPI_GATESBWD = 3 };
constexpr int kMaxBatch = 8;
struct GemmBatch {
  int count;
  int R;
  long long* timing;   // diagnostic (SD_TRACE_TC): clock64 stamps of CTA 0; null in production
  GemmP p[kMaxBatch];
};
#define SD_G_STAMP(i) do { if (b.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) b.timing[i] = clock64(); } while (0)

constexpr int GB_KC = 512;           // largest K slice one CTA handles (staged in shared memory)
constexpr int GB_XLD = GB_KC + 4;    // padded row stride of the staged activations
constexpr int GB_RLD = 16 * 16 + 4;  // padded stride of the split-K reduction buffer
constexpr int GB_MAXSPLIT = 8;       // cluster size along K (portable maximum)
constexpr int GB_SMEM = (16 * GB_XLD) * 4 + GB_MAXSPLIT * 256 * 4;   // staged activations (reused as the [8][256] reduction buffer) + cluster slots

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = (ceil(N/16), ksplit, problems * row_tiles), cluster = (1, ksplit, 1).
// K is cut into `ksplit` slices of `kslice` (<= 512) elements, one per CTA of the cluster: every CTA does ONE
// global round trip (its weight quads + its activation slice are all in flight together), reduces its 64
// thread-group partials through shared memory, and ships its 16x16 partial tile to the cluster leader through
// distributed shared memory; the leader adds the slices in rank order (deterministic) and stores.
// BWDEPI selects the variant that carries the fused reverse-scan epilogue (EPI_GATESBWD): a separate
// instantiation, so that the plain kernel keeps its register count (120: two CTAs per SM).
template <bool BWDEPI>
__device__ __forceinline__ void gemm_f32_body(const GemmBatch& b, int ksplit, int kslice) {
  SD_G_STAMP(0);
  const int prob = blockIdx.z % b.count, rtile = blockIdx.z / b.count;
  const GemmP& p = b.p[prob];
  const bool gates = p.epi == EPI_GATES;   // cluster = (3,1,1): blockIdx.x = tile*3 + gate
  const int gate = gates ? (int)(blockIdx.x % 3) : 0;
  const int n0 = gates ? (int)(blockIdx.x / 3) * 16 : blockIdx.x * 16;
  const bool active = n0 < p.N;           // whole clusters are inactive together
  const int r0 = rtile * 16;
  const int rank = gates ? gate : (int)blockIdx.y;   // == cluster rank
  const int wcol = gates ? gate * p.e_k + n0 : n0;   // first weight column of this CTA's tile
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* slots = smem + (GB_SMEM / 4 - GB_MAXSPLIT * 256);   // leader: one 16x16 tile per rank
  const int tid = threadIdx.x, tx = tid & 3, ty = tid >> 2;
  const int kc = gates ? 0 : rank * kslice;
  const int kend = min(p.K, kc + kslice);
  float s = 0.f;
  // Weight prefetch BEFORE the PDL wait: the packed weights do not depend on the preceding kernel, so their
  // L2 round trip overlaps its execution; only the activations have to wait.  They are fetched straight into the
  // mma.sync B-fragment layout: warp w owns the k-steps (of 8) w, w + 8, ... of this CTA's slice; lane (gq, tq) holds
  // W[k0 + tq][n] and W[k0 + tq + 4][n] for n = 8*nt + gq.
  const int lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  constexpr int KSTEPS = GB_KC / 8 / 8;   // k-steps per warp when the slice is full
  float wb[KSTEPS][2][2];
#pragma unroll
  for (int i = 0; i < KSTEPS; ++i) {
    const int k0 = kc + (i * 8 + warp) * 8;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + tq + 4 * h;
        wb[i][nt][h] = (active && k < kend) ? __ldg(p.Wt + (size_t)k * p.ldw + wcol + nt * 8 + gq) : 0.f;
      }
  }
  // bias of this thread's output element(s) (packed weights too)
  const int pre_n = n0 + (tid & 15);
  float pb0 = 0.f, pb1 = 0.f, pb2 = 0.f;
  if (p.bias && active && pre_n < p.N) {
    pb0 = __ldg(p.bias + pre_n);
    if (gates) { pb1 = __ldg(p.bias + p.e_k + pre_n); pb2 = __ldg(p.bias + 2 * p.e_k + pre_n); }
  }
  // PRE_NORMBWD: the saved pre-norm values and the RMS scale are forward-tape data: fetch them and finish everything
  // that only depends on them (rho, n, silu') before the PDL wait.  Lane layout: 16 lanes per row, float4 i of a lane
  // sits at columns 64*i + 4*(tid & 15).
  const int prow = tid >> 4, pseg = tid & 15;
  float4 pn[4], pc[4], pre_w4[4];   // n = v*rho, c = silu'(m) and the RMS scale w per element
  float prho = 0.f;
  if (p.pre == PRE_NORMBWD && active) {
    float4 pv[4], pw[4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 64 + pseg * 4;
      const bool ok = c < p.K && r0 + prow < b.R;
      pv[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.pre_v + (size_t)(r0 + prow) * p.pre_ldv + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pw[i] = c < p.K ? __ldg(reinterpret_cast<const float4*>(p.pre_w + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += pv[i].x * pv[i].x + pv[i].y * pv[i].y + pv[i].z * pv[i].z + pv[i].w * pv[i].w;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    prho = 1.f / sqrtf(ss / (float)p.K + kRmsEps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float vv[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w}, ww[4] = {pw[i].x, pw[i].y, pw[i].z, pw[i].w};
      float nn[4], cc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        nn[e] = vv[e] * prho;
        const float m = nn[e] * ww[e];
        const float sg = sigmoidf_(m);
        cc[e] = sg * (1.f + m * (1.f - sg));
      }
      pn[i] = make_float4(nn[0], nn[1], nn[2], nn[3]);
      pc[i] = make_float4(cc[0], cc[1], cc[2], cc[3]);
      pre_w4[i] = pw[i];
    }
  }
  // fused backward epilogues: forward-tape operands of this thread's output element and everything that only depends on them
  const int erow = r0 + (tid >> 4), ecol = n0 + (tid & 15);
  const bool eok = active && erow < b.R && ecol < p.N;
  float ef0 = 0.f, ef1 = 0.f, ef2 = 0.f, ef3 = 0.f, ef4 = 0.f;
  if (BWDEPI && p.epi == EPI_GATESBWD) {   // ef0 = reset gate, ef1 = candidate, ef2 = update gate, ef3 = raw cand pre-activation, ef4 = input deter
    if (eok) {
      const int Dg = p.e_k, gi = ecol / Dg, o = ecol - gi * Dg;
      const size_t qo = (size_t)erow * 3 * p.N + (size_t)gi * 3 * Dg + o;
      const float qr = __ldg(p.x5 + qo), qu = __ldg(p.x5 + qo + 2 * Dg);
      ef3 = __ldg(p.x5 + qo + Dg);
      ef4 = __ldg(p.x6 + (size_t)erow * p.xi4 + ecol);
      ef0 = sigmoidf_(qr);
      ef1 = tanhf(ef0 * ef3);
      ef2 = sigmoidf_(qu - 1.f);
    }
  }
  pdl_prologue();
  SD_G_STAMP(1);
  if (active && kc < kend) {
    // all global loads of this CTA are issued before anything is stored to shared memory
    // (weights were prefetched into w[][] before the PDL wait, see below)
    // activation slice -> shared memory.  Fast path: 16-byte loads/stores (thread owns 8 float4 = 2 rows x 4
    // quads... precisely: quad q = tid & 127 of rows 2*i + (tid >> 7)); needs 16 B aligned rows and a segment
    // boundary on a multiple of 4.  Otherwise scalar (thread owns columns tid and tid+256 of all 16 rows).
    const bool vec = ((p.lda & 3) == 0) && ((p.K1 & 3) == 0) && ((kc & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
                     (p.K1 == p.K || (((p.lda2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A2) & 15) == 0)));
    if (p.pre == PRE_NORMBWD) {
      // dv = rho * (dn - n * mean(dn * n)), dn = dm * w, dm = dout * silu'(m); dm * n is the RMS-scale gradient term
      float4 dn4[4];
      float dot = 0.f;
      const bool wr = p.pre_write && blockIdx.x == 0 && r0 + prow < b.R;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * 64 + pseg * 4;
        const bool ok = c < p.K && r0 + prow < b.R;
        const float4 dy = ok ? __ldg(reinterpret_cast<const float4*>(p.A + (size_t)(r0 + prow) * p.lda + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 dm = make_float4(dy.x * pc[i].x, dy.y * pc[i].y, dy.z * pc[i].z, dy.w * pc[i].w);
        if (wr && ok && p.pre_dmn)
          *reinterpret_cast<float4*>(p.pre_dmn + (size_t)(r0 + prow) * p.pre_lddmn + c) =
              make_float4(dm.x * pn[i].x, dm.y * pn[i].y, dm.z * pn[i].z, dm.w * pn[i].w);
        dn4[i] = make_float4(dm.x * pre_w4[i].x, dm.y * pre_w4[i].y, dm.z * pre_w4[i].z, dm.w * pre_w4[i].w);
        dot = fmaf(dn4[i].x, pn[i].x, dot); dot = fmaf(dn4[i].y, pn[i].y, dot);
        dot = fmaf(dn4[i].z, pn[i].z, dot); dot = fmaf(dn4[i].w, pn[i].w, dot);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dot /= (float)p.K;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * 64 + pseg * 4;
        const float4 o4 = make_float4(prho * (dn4[i].x - pn[i].x * dot), prho * (dn4[i].y - pn[i].y * dot),
                                      prho * (dn4[i].z - pn[i].z * dot), prho * (dn4[i].w - pn[i].w * dot));
        *reinterpret_cast<float4*>(xs + prow * GB_XLD + c) = (c < p.K) ? o4 : make_float4(0.f, 0.f, 0.f, 0.f);
        if (wr && c < p.K) *reinterpret_cast<float4*>(p.pre_dv + (size_t)(r0 + prow) * p.pre_lddv + c) = o4;
      }
      // columns [256, 512) of the staged slice are never read: kend <= 256 on this path
    } else if (vec) {
      float4 v4[8];
      const int q = tid & 127, rsel = tid >> 7;
      const int kk = kc + q * 4;
      const bool in1 = kk < p.K1;
      const float* src = in1 ? p.A + kk : p.A2 + (kk - p.K1);
      const size_t ld = in1 ? p.lda : p.lda2;
      const bool kok = kk < kend;   // kend and K1 are multiples of 4 on this path or the tail quad is zero-padded below
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = r0 + 2 * i + rsel;
        v4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kok && row < b.R) {
          if (kk + 3 < kend) {
            v4[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * ld));
          } else {  // ragged tail of the slice
            const float* s1 = src + (size_t)row * ld;
            v4[i].x = s1[0];
            if (kk + 1 < kend) v4[i].y = s1[1];
            if (kk + 2 < kend) v4[i].z = s1[2];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(xs + (2 * i + rsel) * GB_XLD + q * 4) = v4[i];
    } else {
      float v[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int kk = kc + hf * 256 + tid;
        const bool in1 = kk < p.K1;
        const float* src = in1 ? p.A + kk : p.A2 + (kk - p.K1);
        const size_t ld = in1 ? p.lda : p.lda2;
        const bool kok = kk < kend;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int row = r0 + r;
          v[r * 2 + hf] = (kok && row < b.R) ? __ldg(src + (size_t)row * ld) : 0.f;
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int r = 0; r < 16; ++r) xs[r * GB_XLD + hf * 256 + tid] = v[r * 2 + hf];
    }
    __syncthreads();
    SD_G_STAMP(2);
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
#pragma unroll
    for (int i = 0; i < KSTEPS; ++i) {
      const int kl = (i * 8 + warp) * 8;   // offset inside the staged slice (zero filled beyond kend)
      if (kc + kl < kend) {                // warp-uniform
        uint32_t ah[4], al[4];
        split_tf32(xs[gq * GB_XLD + kl + tq], ah[0], al[0]);
        split_tf32(xs[(gq + 8) * GB_XLD + kl + tq], ah[1], al[1]);
        split_tf32(xs[gq * GB_XLD + kl + tq + 4], ah[2], al[2]);
        split_tf32(xs[(gq + 8) * GB_XLD + kl + tq + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(wb[i][nt][0], bh0, bl0);
          split_tf32(wb[i][nt][1], bh1, bl1);
          mma_tf32(acc[nt], al, bh0, bh1);   // small terms first
          mma_tf32(acc[nt], ah, bl0, bl1);
          mma_tf32(acc[nt], ah, bh0, bh1);
        }
      }
    }
    __syncthreads();
    SD_G_STAMP(3);
    float* red = smem;   // [8 warps][256]: C fragment c0 (gq, 2tq), c1 (gq, 2tq+1), c2 (gq+8, 2tq), c3 (gq+8, 2tq+1)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* r0p = red + warp * 256 + gq * 16 + nt * 8 + 2 * tq;
      *reinterpret_cast<float2*>(r0p) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(r0p + 8 * 16) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 256 + tid];   // fixed warp order => deterministic
    SD_G_STAMP(4);
  }
  const int row = r0 + (tid >> 4), cc = tid & 15, n = n0 + cc;
  if (gates) {
    // slots of the leader: [gate][256]; every CTA ships its full-K tile, the leader applies the gate math
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(slots + gate * 256 + tid);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
    cluster_sync_all();
    if (gate != 0 || !active || row >= b.R || n >= p.N) return;
    const int Dg = p.e_k;
    const float qr = slots[tid] + pb0;
    const float qc = slots[256 + tid] + pb1;
    const float qu = slots[512 + tid] + pb2;
    if (p.C) {
      float* q = p.C + (size_t)row * p.ldc;
      q[n] = qr; q[Dg + n] = qc; q[2 * Dg + n] = qu;
    }
    const float reset = sigmoidf_(qr);
    const float cand = tanhf(reset * qc);
    const float upd = sigmoidf_(qu - 1.f);
    p.e_out[(size_t)row * p.e_ld_out + n] = upd * cand + (1.f - upd) * p.e_in[(size_t)row * p.e_ld_in + n];
    return;
  }
  if (ksplit > 1) {
    // ship the partial tile into slot `rank` of the leader's shared memory (DSMEM), then cluster barrier
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(slots + rank * 256 + tid);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
    cluster_sync_all();
    SD_G_STAMP(5);
    if (rank != 0) return;
    s = 0.f;
    for (int r = 0; r < ksplit; ++r) s += slots[r * 256 + tid];
  }
  const bool ok = active && row < b.R && n < p.N;
  const float val = s + pb0;
  if (p.epi == EPI_SAMPLE) {
    // 16 consecutive threads own the 16 logits of one (row, tile): 16/Kc whole categories
    const int Kc = p.e_k;
    const int kcls = cc % Kc;
    const float uu = ok ? p.e_in[(size_t)row * p.e_ld_in + n] : 0.5f;
    int best;
    if (Kc == 16) best = sample_group<16>(val, uu, ok, kcls, Kc, p.e_f, nullptr);
    else if (Kc == 8) best = sample_group<8>(val, uu, ok, kcls, Kc, p.e_f, nullptr);
    else if (Kc == 4) best = sample_group<4>(val, uu, ok, kcls, Kc, p.e_f, nullptr);
    else best = sample_group<2>(val, uu, ok, kcls, Kc, p.e_f, nullptr);
    if (ok) {
      if (p.C) p.C[(size_t)row * p.ldc + n] = val;
      if (p.e_out2) p.e_out2[(size_t)row * p.e_ld_out2 + n] = val;
      p.e_out[(size_t)row * p.e_ld_out + n] = (kcls == best) ? 1.f : 0.f;
    }
    return;
  }
  if (BWDEPI && p.epi == EPI_GATESBWD) {   // gates_bwd_kernel on unit n of this row; `val` is the gc term
    if (!ok) return;
    if (p.C) p.C[(size_t)row * p.ldc + n] = val;
    const int D = p.N, Dg = p.e_k, gi = n / Dg, o = n - gi * Dg;
    const size_t qo = (size_t)row * 3 * D + (size_t)gi * 3 * Dg + o;
    const float Rg = ef0, Cc = ef1, Uu = ef2, c = ef3, din = ef4;
    float carry = p.x0 ? p.x0[(size_t)row * p.xi0 + n] : 0.f;
    if (p.x2) carry += p.x2[(size_t)row * D + n];
    if (p.x3) carry += p.x3[((size_t)row * p.xi2 + gi) * (size_t)p.xi3 + o];
    if (p.x4) carry *= p.x4[row];
    const float gd = carry + (p.x1 ? p.x1[(size_t)row * p.xi1 + n] : 0.f) + val;
    const float dUu = gd * (Cc - din);
    const float dC = gd * Uu;
    const float dtn = dC * (1.f - Cc * Cc);
    p.y0[qo] = (dtn * c) * Rg * (1.f - Rg);
    p.y0[qo + Dg] = dtn * Rg;
    p.y0[qo + 2 * Dg] = dUu * Uu * (1.f - Uu);
    p.y1[(size_t)row * D + n] = gd * (1.f - Uu);
    return;
  }
  if (ok) p.C[(size_t)row * p.ldc + n] = val;
  SD_G_STAMP(6);
}
__global__ void __launch_bounds__(256, 2) gemm_f32_kernel(const GemmBatch b, int ksplit, int kslice) { gemm_f32_body<false>(b, ksplit, kslice); }
__global__ void __launch_bounds__(256, 2) gemm_f32_bwdepi_kernel(const GemmBatch b, int ksplit, int kslice) { gemm_f32_body<true>(b, ksplit, kslice); }

// ------------------------------------------------------------------------------------------------
// Weight-gradient GEMM (contraction over rows):  dW[n][k] (+)= sum_r dY[r][n] * X[r][k].
// Classic 64x64 tile, 4x4 per thread; output addressed with arbitrary strides so it can write the
// reference layouts directly (nn.Linear (N,K); BlockLinear (O/G, I/G, G) via sn = I*G/G.., sk = G).
// X may be a 2-segment concat like the forward operand.
// ------------------------------------------------------------------------------------------------
struct WgradP {
  const float* dY; int ldy;   // [R x N]
  const float* X;  int ldx;   // [R x K1]
  const float* X2; int ldx2;  // [R x (K-K1)]
  int K1, K, N;
  float* dW; long long sn, sk; // partial slice s lives at dW + s*slice_stride; element (n,k) at n*sn + k*sk
  long long slice_stride;
};
struct WgradBatch {
  int count;
  int R;
  int rows_per_slice;          // rows are cut into gridDim.z / count slices (fixed assignment => deterministic)
  int slice0;                  // first slice this launch covers (a launch may handle a sub-range of the slices)
  WgradP p[kMaxBatch];
};

// Each CTA owns a 64x64 (n,k) tile of one row slice; partial tiles are summed in slice order by
// wgrad_reduce_kernel.  The contraction over the rows runs on mma.sync.m16n8k8 with the 3xTF32 split (fp32-class
// accuracy): M = 64 output features (4 m-tiles), N = 64 input features (8 n-tiles), K = 16 rows per staged chunk.
// Warp w owns m-tile (w & 3) and the n-tiles 4*(w >> 2) .. +3.  Next chunk's loads are issued before the current
// chunk's MMAs (register prefetch).
__global__ void __launch_bounds__(256) wgrad_f32_kernel(const WgradBatch b) {
  pdl_prologue();
  const int prob = blockIdx.z % b.count, slice = b.slice0 + blockIdx.z / b.count;
  const WgradP& p = b.p[prob];
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  if (n0 >= p.N || k0 >= p.K) return;
  constexpr int LD = 64 + 8;   // row r of a chunk at bank offset 8r: the fragment reads (4 rows x 8 columns) are conflict free
  __shared__ __align__(16) float ys[16][LD];
  __shared__ __align__(16) float xs[16][LD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  const int mt = warp & 3, nh = warp >> 2;
  const int r_begin = slice * b.rows_per_slice;
  const int r_end = min(b.R, r_begin + b.rows_per_slice);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // element e (0..3) of this thread: row (tid>>6) + 4e of the chunk, column tid&63
  const int lc = tid & 63, lr = tid >> 6;
  const bool nok = n0 + lc < p.N;
  const int kk = k0 + lc;
  const bool kok = kk < p.K;
  const bool in1 = kk < p.K1;
  const float* xsrc = in1 ? p.X + kk : p.X2 + (kk - p.K1);
  const size_t xld = in1 ? p.ldx : p.ldx2;
  float yv[4], xv[4];
  auto load = [&](int rc) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = rc + lr + 4 * e;
      const bool rok = row < r_end;
      yv[e] = (rok && nok) ? __ldg(p.dY + (size_t)row * p.ldy + n0 + lc) : 0.f;
      xv[e] = (rok && kok) ? __ldg(xsrc + (size_t)row * xld) : 0.f;
    }
  };
  load(r_begin);
  for (int rc = r_begin; rc < r_end; rc += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ys[lr + 4 * e][lc] = yv[e];
      xs[lr + 4 * e][lc] = xv[e];
    }
    __syncthreads();
    if (rc + 16 < r_end) load(rc + 16);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      // A[m][kr] = dY[row kr][feature m]: a0 (gq, tq), a1 (gq+8, tq), a2 (gq, tq+4), a3 (gq+8, tq+4)
      uint32_t ah[4], al[4];
      split_tf32(ys[ks * 8 + tq][mt * 16 + gq], ah[0], al[0]);
      split_tf32(ys[ks * 8 + tq][mt * 16 + gq + 8], ah[1], al[1]);
      split_tf32(ys[ks * 8 + tq + 4][mt * 16 + gq], ah[2], al[2]);
      split_tf32(ys[ks * 8 + tq + 4][mt * 16 + gq + 8], ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        // B[kr][n] = X[row kr][feature n]: b0 (k = tq, n = gq), b1 (k = tq + 4, n = gq)
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(xs[ks * 8 + tq][nh * 32 + nt * 8 + gq], bh0, bl0);
        split_tf32(xs[ks * 8 + tq + 4][nh * 32 + nt * 8 + gq], bh1, bl1);
        mma_tf32(acc[nt], al, bh0, bh1);   // small terms first
        mma_tf32(acc[nt], ah, bl0, bl1);
        mma_tf32(acc[nt], ah, bh0, bh1);
      }
    }
    __syncthreads();
  }
  float* out = p.dW + (long long)slice * p.slice_stride;
  // C fragment: c0 (gq, 2tq), c1 (gq, 2tq+1), c2 (gq+8, 2tq), c3 (gq+8, 2tq+1)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + mt * 16 + gq + (c >> 1) * 8, kq = k0 + nh * 32 + nt * 8 + 2 * tq + (c & 1);
      if (n < p.N && kq < p.K) out[n * p.sn + kq * p.sk] = acc[nt][c];
    }
}
// dst[i] += sum_s partial[s*stride + i], s ascending.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, long long stride, int slices, long long numel,
                                    float* dst) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += partial[k * stride + i];
    dst[i] += s;
  }
}

// Same sum, 16 bytes per thread and all S slice loads of an element group in flight together (the run-time loop above
// issues them one L2/DRAM round trip at a time: 34 us per launch in the end-to-end profile).  Same order => same bits.
template <int S>
__global__ void __launch_bounds__(256) wgrad_reduce4_kernel(const float4* __restrict__ partial, long long stride4,
                                                            long long numel4, float4* dst) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v[S];
#pragma unroll
    for (int k = 0; k < S; ++k) v[k] = __ldg(partial + k * stride4 + i);
    float4 d = dst[i];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < S; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
    dst[i] = d;
  }
}

// ------------------------------------------------------------------------------------------------
// Weight repacking: reference layout -> Wt[g][k][ldw] fp32 (forward operand, n contiguous),
// Wn[g][n][ldk] fp32 (dgrad operand, k contiguous) and bf16 copies of both for the tcgen05 path.
// Source element (g,n,k) sits at src[g*s_g + n*s_n + k*s_k]:
//   nn.Linear (N,K): s_g=0, s_n=K, s_k=1;   BlockLinear (O/G, I/G, G): s_g=1, s_n=(I/G)*G, s_k=G.
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, int G, int N, int K, long long s_g,
                                   long long s_n, long long s_k, float* wt, int ldw, float* wn, int ldk,
                                   __nv_bfloat16* wn_bf, __nv_bfloat16* wt_bf) {
  pdl_wait();   // no early trigger: consumers may prefetch packed weights before their own wait
  const long long total = (long long)G * N * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int n = (int)((i / K) % N);
    const int g = (int)(i / ((long long)K * N));
    const float v = src[g * s_g + n * s_n + k * s_k];
    if (wt) wt[((size_t)g * K + k) * ldw + n] = v;
    if (wn) wn[((size_t)g * N + n) * ldk + k] = v;
    if (wn_bf) wn_bf[((size_t)g * N + n) * ldk + k] = __float2bfloat16(v);
    if (wt_bf) wt_bf[((size_t)g * K + k) * ldw + n] = __float2bfloat16(v);
  }
}

// All layers of one module in ONE launch (sd_set_weights is on the critical path of every training step: the weights
// change once per update): a table of layers in the kernel parameters, a contiguous range of blocks per layer.  Writes
// the four packed layouts (fp32 Wt / Wn, bf16 [g][npad][K] and its transpose [g][kpad][N]) and copies bias / RMS scale.
struct PackEntry {
  const float* src; int G, N, K; long long s_g, s_n, s_k;
  float* wt; int ldw; float* wn; int ldk;
  __nv_bfloat16* w_bf; int npad;
  __nv_bfloat16* wT_bf; int kpad;
  const float* bias_src; float* bias_dst; int nbias;
  const float* gain_src; float* gain_dst; int ngain;
  int blk0, nblk;
  int mode;   // 0: element-wise (any strides); 1: nn.Linear source, 64 x 64 tiles; 2: BlockLinear source with G = 8, 8 x 64 x 8 tiles
};
constexpr int kMaxPack = 16;
struct PackTable {
  int n;
  PackEntry e[kMaxPack];
};
__global__ void __launch_bounds__(256) pack_module_kernel(const PackTable t) {
  pdl_wait();   // no early trigger: consumers may prefetch packed weights before their own wait
  int ei = 0;
  while (ei + 1 < t.n && (int)blockIdx.x >= t.e[ei + 1].blk0) ++ei;
  const PackEntry& e = t.e[ei];
  const int lb = blockIdx.x - e.blk0;
  const long long total = (long long)e.G * e.N * e.K;
  if (lb == 0) {
    if (e.bias_src && e.bias_dst)
      for (int i = threadIdx.x; i < e.nbias; i += 256) e.bias_dst[i] = e.bias_src[i];
    if (e.gain_src && e.gain_dst)
      for (int i = threadIdx.x; i < e.ngain; i += 256) e.gain_dst[i] = e.gain_src[i];
  }
  // Tiled paths: a tile of the source is read ONCE with coalesced loads into shared memory and written to the k-contiguous
  // layouts (Wn, bf16 [n][K]) and, transposed, to the n-contiguous ones (Wt, bf16 [k][N]) with coalesced stores; 32-bit
  // index arithmetic.  (The element-wise path below costs two strided passes and 64-bit divisions per element: 0.20 ms
  // for the six modules of the base model, on the critical path of every training step; tiled: see profiles/refresh_time.py.)
  __shared__ float tile[64 * 65];
  const int tid = threadIdx.x;
  if (e.mode == 1) {   // source (N, K), k fastest
    const int tk = (e.K + 63) >> 6;
    const int n0 = (lb / tk) << 6, k0 = (lb % tk) << 6;
    {
      const int kk = tid & 63, k = k0 + kk;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int nl = (tid >> 6) + 4 * i, n = n0 + nl;
        const float v = (n < e.N && k < e.K) ? __ldg(e.src + (size_t)n * e.K + k) : 0.f;
        tile[nl * 65 + kk] = v;
        if (n < e.N && k < e.K) {
          if (e.wn) e.wn[(size_t)n * e.ldk + k] = v;
          if (e.w_bf) e.w_bf[(size_t)n * e.K + k] = __float2bfloat16(v);
        }
      }
    }
    __syncthreads();
    {
      const int nn = tid & 63, n = n0 + nn;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int kl = (tid >> 6) + 4 * i, k = k0 + kl;
        if (n < e.N && k < e.K) {
          const float v = tile[nn * 65 + kl];
          if (e.wt) e.wt[(size_t)k * e.ldw + n] = v;
          if (e.wT_bf) e.wT_bf[(size_t)k * e.N + n] = __float2bfloat16(v);
        }
      }
    }
    return;
  }
  if (e.mode == 2) {   // source (N, K, G = 8), block index fastest; tile = 8 n x 64 k x 8 g, shared memory [g][n][k] (65-float rows)
    const int tk = (e.K + 63) >> 6;
    const int n0 = (lb / tk) << 3, k0 = (lb % tk) << 6;
    const int kw = min(64, e.K - k0);
    for (int nl = 0; nl < 8; ++nl) {
      const int n = n0 + nl;
      if (n >= e.N) break;
      const float* row = e.src + ((size_t)n * e.K + k0) * 8;   // kw * 8 contiguous floats
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + 256 * j, kl = idx >> 3, g = idx & 7;
        if (kl < kw) tile[(g * 8 + nl) * 65 + kl] = __ldg(row + idx);
      }
    }
    __syncthreads();
    {
      const int kk = tid & 63, k = k0 + kk;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int rs = (tid >> 6) + 4 * i, g = rs >> 3, nl = rs & 7, n = n0 + nl;
        if (n < e.N && k < e.K) {
          const float v = tile[rs * 65 + kk];
          if (e.wn) e.wn[((size_t)g * e.N + n) * e.ldk + k] = v;
          if (e.w_bf) e.w_bf[((size_t)g * e.npad + n) * e.K + k] = __float2bfloat16(v);
        }
      }
    }
    {
      const int nl = tid & 7, n = n0 + nl;
      for (int g = 0; g < 8; ++g)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int kl = hf * 32 + (tid >> 3), k = k0 + kl;
          if (n < e.N && k < e.K) {
            const float v = tile[(g * 8 + nl) * 65 + kl];
            if (e.wt) e.wt[((size_t)g * e.K + k) * e.ldw + n] = v;
            if (e.wT_bf) e.wT_bf[((size_t)g * e.kpad + k) * e.N + n] = __float2bfloat16(v);
          }
        }
    }
    return;
  }
  // two passes so that every WRITE is coalesced: k fastest for the k-contiguous layouts, n fastest for the n-contiguous
  // ones (the strided reads of the second pass hit L2: a module's weights are a few MB)
  for (long long i = lb * 256ll + threadIdx.x; i < total; i += e.nblk * 256ll) {
    const int k = (int)(i % e.K);
    const int n = (int)((i / e.K) % e.N);
    const int g = (int)(i / ((long long)e.K * e.N));
    const float v = e.src[g * e.s_g + n * e.s_n + k * e.s_k];
    if (e.wn) e.wn[((size_t)g * e.N + n) * e.ldk + k] = v;
    if (e.w_bf) e.w_bf[((size_t)g * e.npad + n) * e.K + k] = __float2bfloat16(v);
  }
  for (long long i = lb * 256ll + threadIdx.x; i < total; i += e.nblk * 256ll) {
    const int n = (int)(i % e.N);
    const int k = (int)((i / e.N) % e.K);
    const int g = (int)(i / ((long long)e.K * e.N));
    const float v = e.src[g * e.s_g + n * e.s_n + k * e.s_k];
    if (e.wt) e.wt[((size_t)g * e.K + k) * e.ldw + n] = v;
    if (e.wT_bf) e.wT_bf[((size_t)g * e.kpad + k) * e.N + n] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Row-wise helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Deterministic block-wide sum (every thread gets the result); blockDim.x <= 1024.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += sh[i];
  return s;
}

// y = SiLU(RMSNorm_width(v) * w): one CTA per (row, segment).  rssm.py:16-31,106-130; networks.py:325-327.
struct NormActP {
  float* in; int ld_in;         // segment base (row 0), row stride; rewritten with the summed value when nparts > 0
  const float* parts; int nparts; long long part_stride;  // split-K partial slices 1..nparts (same ld as `in`)
  const float* w;               // [width]
  float* out; int ld_out;       // nullable
  __nv_bfloat16* out_bf; int ld_bf; // nullable
  int width;
  int no_writeback;             // nparts > 0: do not rewrite `in` with the summed value (no backward tape wanted)
};
struct NormActBatch {
  int count;
  NormActP p[4];
};
__global__ void __launch_bounds__(256) normact_kernel(const NormActBatch b) {
  __shared__ float sh[32];
  const NormActP& p = b.p[blockIdx.y];
  const size_t row = blockIdx.x;
  float* in = p.in + row * p.ld_in;
  float v[8], g[8];
  // the RMS scale is a packed weight: fetch it before the PDL wait (off the post-reduction critical path)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    g[i] = (c < p.width) ? __ldg(p.w + c) : 0.f;
  }
  pdl_prologue();
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = (c < p.width) ? in[c] : 0.f;
  }
  for (int s = 0; s < p.nparts; ++s) {   // fixed slice order => deterministic; the 8 loads of a slice are independent
    const float* ps = p.parts + s * p.part_stride + row * p.ld_in;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < p.width) v[i] += ps[c];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    if (p.nparts > 0 && !p.no_writeback && c < p.width) in[c] = v[i];
    ss = fmaf(v[i], v[i], ss);
  }
  ss = block_sum(ss, sh);
  const float rs = 1.f / sqrtf(ss / (float)p.width + kRmsEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < p.width) {
      const float y = siluf_((v[i] * rs) * g[i]);
      if (p.out) p.out[row * p.ld_out + c] = y;
      if (p.out_bf) p.out_bf[row * p.ld_bf + c] = __float2bfloat16(y);
    }
  }
}

// Same for 256-wide segments (the U / units layers), one WARP per (row, segment): lane owns 8 consecutive columns, the
// slice-0 value and every split-K partial are fetched as 16-byte loads that are all in flight together (the CTA-per-row
// kernel above walks the partial slices one L2 round trip at a time and needs two block barriers), the row statistic is a
// shuffle reduction, outputs are 16-byte stores.  Measured by ablation on B200 (N = 1024 imagination scan): the
// 3-segment norm after the wide feat layers cost 10.5 us per step on the critical path with the CTA-per-row kernel.
__global__ void __launch_bounds__(256) normact256_warp_kernel(const NormActBatch b, int R) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool live = item < R * b.count;
  const int row = live ? item / b.count : 0, seg = live ? item - row * b.count : 0;
  const NormActP& p = b.p[seg];
  // the RMS scale is a packed weight: fetch it before the PDL wait
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.w + lane * 8));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.w + lane * 8 + 4));
  pdl_prologue();
  if (!live) return;
  const size_t off = (size_t)row * p.ld_in + lane * 8;
  float4 a = *reinterpret_cast<const float4*>(p.in + off), c = *reinterpret_cast<const float4*>(p.in + off + 4);
  float4 pa[7], pc[7];
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    if (s < p.nparts) {
      pa[s] = *reinterpret_cast<const float4*>(p.parts + s * p.part_stride + off);
      pc[s] = *reinterpret_cast<const float4*>(p.parts + s * p.part_stride + off + 4);
    }
  }
#pragma unroll
  for (int s = 0; s < 7; ++s) {   // fixed slice order => deterministic
    if (s < p.nparts) {
      a.x += pa[s].x; a.y += pa[s].y; a.z += pa[s].z; a.w += pa[s].w;
      c.x += pc[s].x; c.y += pc[s].y; c.z += pc[s].z; c.w += pc[s].w;
    }
  }
  if (p.nparts > 0 && !p.no_writeback) {   // the summed pre-norm value is the backward tape
    *reinterpret_cast<float4*>(p.in + off) = a;
    *reinterpret_cast<float4*>(p.in + off + 4) = c;
  }
  float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss / 256.f + kRmsEps);
  float y[8] = {siluf_((a.x * rs) * g0.x), siluf_((a.y * rs) * g0.y), siluf_((a.z * rs) * g0.z), siluf_((a.w * rs) * g0.w),
                siluf_((c.x * rs) * g1.x), siluf_((c.y * rs) * g1.y), siluf_((c.z * rs) * g1.z), siluf_((c.w * rs) * g1.w)};
  if (p.out) {
    float* o = p.out + (size_t)row * p.ld_out + lane * 8;
    *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (p.out_bf) {
    __nv_bfloat162 q0 = __floats2bfloat162_rn(y[0], y[1]), q1 = __floats2bfloat162_rn(y[2], y[3]);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(y[4], y[5]), q3 = __floats2bfloat162_rn(y[6], y[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
    pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
    *reinterpret_cast<uint4*>(p.out_bf + (size_t)row * p.ld_bf + lane * 8) = pk;
  }
}

// GRU-style gates (rssm.py:63-75): q (R, 3D) laid out [g][reset|cand|update][D/G].
__global__ void gates_kernel(const float* __restrict__ q, const float* __restrict__ deter_in, int ld_in,
                             float* __restrict__ deter_out, int ld_out, __nv_bfloat16* out_bf, int ld_bf,
                             int R, int D, int Dg) {
  pdl_prologue();
  const long long total = (long long)R * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t row = (size_t)(i / D);
    const int g = d / Dg, o = d - g * Dg;
    const float* qr = q + row * 3 * D + (size_t)g * 3 * Dg + o;
    const float reset = sigmoidf_(qr[0]);
    const float cand = tanhf(reset * qr[Dg]);
    const float upd = sigmoidf_(qr[2 * Dg] - 1.f);
    const float out = upd * cand + (1.f - upd) * deter_in[row * ld_in + d];
    deter_out[row * ld_out + d] = out;
    if (out_bf) out_bf[row * ld_bf + d] = __float2bfloat16(out);
  }
}

// OneHotDist(logit, unimix).rsample() with injected uniforms (distributions.py:16-36 +
// F.gumbel_softmax(hard=True)); K <= 32 classes held in registers, one thread per category.
// Returns the first-max index of y = softmax(l + g); y[] is left in yv (for the backward).
__device__ __forceinline__ int sample_category(const float* lg, const float* u, int K, float unimix, float* yv) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) m = fmaxf(m, lg[k]);
  float e[32];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) { e[k] = expf(lg[k] - m); s += e[k]; }
  const float uni = unimix / (float)K;
  float m2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) { e[k] = logf((e[k] / s) * (1.f - unimix) + uni); m2 = fmaxf(m2, e[k]); }
  float s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) s2 += expf(e[k] - m2);
  const float lse = m2 + logf(s2);
  float m3 = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) { e[k] = (e[k] - lse) + (-logf(-logf(u[k]))); m3 = fmaxf(m3, e[k]); }
  float s3 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) { e[k] = expf(e[k] - m3); s3 += e[k]; }
  int best = 0;
  float bv = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) {
      const float y = e[k] / s3;
      if (yv) yv[k] = y;
      if (y > bv) { bv = y; best = k; }
    }
  return best;
}

// Sample all S categories of R rows.  logits (R, S*K) row stride ld_l; u (R, S*K) row stride ld_u.
// Writes exact one-hot fp32 (row stride ld_o), optional bf16 copy, optional copy of the logits
// (the `logits` output of observe) and the indices.  One lane per class.
template <int GS>
__global__ void sample_kernel(const float* __restrict__ logits, int ld_l, const float* __restrict__ u, int ld_u,
                              int R, int S, int K, float unimix, float* stoch, int ld_o, __nv_bfloat16* stoch_bf,
                              int ld_bf, float* logit_copy, int ld_c, int* idx_out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long cat = t / GS;          // (row, s)
  const int k = (int)(t % GS);
  const bool in_range = cat < (long long)R * S;
  const size_t row = in_range ? (size_t)(cat / S) : 0;
  const int sidx = in_range ? (int)(cat - (long long)row * S) : 0;
  const bool valid = in_range && k < K;
  float lg = 0.f, uu = 0.5f;
  if (valid) uu = __ldg(u + row * ld_u + sidx * K + k);   // injected noise: independent of the preceding kernel
  pdl_prologue();
  if (valid) lg = logits[row * ld_l + sidx * K + k];
  const int best = sample_group<GS>(lg, uu, valid, k, K, unimix, nullptr);
  if (valid) {
    const float v = (k == best) ? 1.f : 0.f;
    if (stoch) stoch[row * ld_o + sidx * K + k] = v;
    if (stoch_bf) stoch_bf[row * ld_bf + sidx * K + k] = __float2bfloat16(v);
    if (logit_copy) logit_copy[row * ld_c + sidx * K + k] = lg;
    if (idx_out && k == 0) idx_out[cat] = best;
  }
}

// obs_step prologue (rssm.py:161-165 + :44): zero stoch/deter/action where is_first, normalise the
// action magnitude, and stage the three step inputs contiguously (they are also the backward tape).
__global__ void prep_obs_kernel(const float* __restrict__ stoch, int ld_s, const float* __restrict__ deter, int ld_d,
                                const float* __restrict__ action, int ld_a, const uint8_t* __restrict__ is_first,
                                int ld_f, int R, int SK, int D, int A, float* zin, float* din, float* ain,
                                float* keep_out, float* araw_out) {
  pdl_prologue();
  const int W = SK + D + A;
  const long long total = (long long)R * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t row = (size_t)(i / W);
    const int c = (int)(i - (long long)row * W);
    const bool rs = is_first ? (is_first[row * ld_f] != 0) : false;
    if (c == 0 && keep_out) keep_out[row] = rs ? 0.f : 1.f;
    if (c < SK) {
      zin[row * SK + c] = rs ? 0.f : stoch[row * ld_s + c];
    } else if (c < SK + D) {
      const int d = c - SK;
      din[row * D + d] = rs ? 0.f : deter[row * ld_d + d];
    } else {
      const int a = c - SK - D;
      const float v = rs ? 0.f : action[row * ld_a + a];
      ain[row * A + a] = v / fmaxf(fabsf(v), 1.f);
      if (araw_out) araw_out[row * A + a] = v;
    }
  }
}

// (B, T, W) -> (T, B, W): puts user-layout tensors into the step-major tape layout.
__global__ void bt_to_tb_kernel(const float* __restrict__ in, float* out, int B, int T, int W) {
  pdl_prologue();
  const long long total = (long long)B * T * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long bt = i / W;
    const int t = (int)(bt % T), b = (int)(bt / T);
    out[((size_t)t * B + b) * W + w] = in[i];
  }
}

// fp32 -> bf16 copy of a strided (R x W) matrix.
__global__ void cast_bf16_kernel(const float* __restrict__ in, int ld_in, __nv_bfloat16* out, int ld_out, int R,
                                 int W) {
  pdl_prologue();
  const long long total = (long long)R * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t row = (size_t)(i / W);
    const int c = (int)(i - (long long)row * W);
    out[row * ld_out + c] = __float2bfloat16(in[row * ld_in + c]);
  }
}
// strided (R x W) fp32 copy.
__global__ void copy_f32_kernel(const float* __restrict__ in, int ld_in, float* out, int ld_out, int R, int W) {
  pdl_prologue();
  const long long total = (long long)R * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t row = (size_t)(i / W);
    const int c = (int)(i - (long long)row * W);
    out[row * ld_out + c] = in[row * ld_in + c];
  }
}

// Actor sampling (dreamer.py:684).  Continuous (distributions.py:217-222): out (R, 2A) = [mean | std_raw],
// action = tanh(mean) + ((max-min)*sigmoid(std_raw+2)+min) * eps.  Discrete (distributions.py:230-231):
// one OneHotDist over A classes with injected uniforms.  Also emits the magnitude-normalised action
// (rssm.py:44) that dyn_in2 consumes.
__global__ void actor_sample_kernel(const float* __restrict__ out, int R, int A, int act_kind, float min_std,
                                    float max_std, float unimix, const float* __restrict__ noise, int ld_n,
                                    float* action, int ld_act, float* abar) {
  pdl_prologue();
  if (act_kind == 0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * A) return;
    const size_t row = i / A;
    const int a = i - (int)row * A;
    const float mean = out[row * 2 * A + a], sraw = out[row * 2 * A + A + a];
    const float std = (max_std - min_std) * sigmoidf_(sraw + 2.f) + min_std;
    const float v = tanhf(mean) + std * noise[row * ld_n + a];
    action[row * ld_act + a] = v;
    abar[row * A + a] = v / fmaxf(fabsf(v), 1.f);
  } else {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    float lg[32], uu[32];
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < A) { lg[k] = out[(size_t)row * A + k]; uu[k] = noise[(size_t)row * ld_n + k]; }
    const int best = sample_category(lg, uu, A, unimix, nullptr);
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < A) {
        const float v = (k == best) ? 1.f : 0.f;
        action[(size_t)row * ld_act + k] = v;
        abar[(size_t)row * A + k] = v;  // |v| <= 1: normalisation is the identity
      }
  }
}

// Fused actor tail for the imagination rollout (dreamer.py:684 + rssm.py:44,48), one warp per row:
//   out = W_last * a3 + b  (units -> act_out)          [networks.py:374-377]
//   action = tanh(mean) + std * eps  |  one-hot Gumbel sample                [distributions.py:217-231]
//   abar = action / max(|action|, 1)
//   v2 = W_in2 * abar + b_in2  (A -> U), the pre-norm input projection of the action.
// Replaces two skinny GEMM launches and the sampling kernel; fp32 throughout.
constexpr int kTailMaxOut = 36;
// dynamic shared memory: W_last [act_out][units] | b_last [act_out] | W_in2^T [A][U] | b_in2 [U]
__host__ __device__ inline size_t actor_tail_smem(int act_out, int units, int A, int U) {
  return ((size_t)act_out * units + act_out + (size_t)A * U + U) * sizeof(float);
}
__global__ void __launch_bounds__(256) actor_tail_kernel(const float* __restrict__ a3, int ld_a3, int units,
                                                         const float* __restrict__ wl_n, int ldk_l,
                                                         const float* __restrict__ bl, int act_out, int A, int act_kind,
                                                         float min_std, float max_std, float unimix,
                                                         const float* __restrict__ noise, int ld_n,
                                                         const float* __restrict__ w2_t, int ldw_2,
                                                         const float* __restrict__ b2, int U, int R, float* aout,
                                                         float* action, int ld_act, float* abar, float* v2, int ld_v2,
                                                         const float* __restrict__ g2 = nullptr,
                                                         __nv_bfloat16* x2_bf = nullptr, int ld_x2 = 0) {
  extern __shared__ __align__(16) float tsm[];
  float* wl_s = tsm;                          // [act_out][units]
  float* bl_s = wl_s + act_out * units;       // [act_out]
  float* w2_s = bl_s + act_out;               // [A][U]
  float* b2_s = w2_s + A * U;                 // [U]
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // packed weights and the injected noise do not depend on the preceding kernel: stage them before the PDL wait
  for (int i = threadIdx.x; i < act_out * units; i += blockDim.x) {
    const int j = i / units, k = i - j * units;
    wl_s[i] = __ldg(wl_n + (size_t)j * ldk_l + k);
  }
  for (int i = threadIdx.x; i < A * U; i += blockDim.x) {
    const int a = i / U, n = i - a * U;
    w2_s[i] = __ldg(w2_t + (size_t)a * ldw_2 + n);
  }
  for (int i = threadIdx.x; i < act_out; i += blockDim.x) bl_s[i] = bl[i];
  for (int i = threadIdx.x; i < U; i += blockDim.x) b2_s[i] = b2[i];
  const float nz = (row < R && lane < A) ? noise[(size_t)row * ld_n + lane] : 0.5f;
  pdl_prologue();
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane + 32 * i;
    x[i] = (row < R && k < units) ? a3[(size_t)row * ld_a3 + k] : 0.f;
  }
  __syncthreads();
  if (row >= R) return;
  float acc[kTailMaxOut];
#pragma unroll
  for (int j = 0; j < kTailMaxOut; ++j) {
    acc[j] = 0.f;
    if (j < act_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        if (k < units) acc[j] = fmaf(x[i], wl_s[j * units + k], acc[j]);
      }
      acc[j] = warp_sum(acc[j]) + bl_s[j];
    }
  }
  if (aout && lane == 0) {
#pragma unroll
    for (int j = 0; j < kTailMaxOut; ++j)
      if (j < act_out) aout[(size_t)row * act_out + j] = acc[j];
  }
  // action of this lane (lane < A)
  float act = 0.f;
  if (act_kind == 0) {
    float mean = 0.f, sraw = 0.f;
#pragma unroll
    for (int j = 0; j < kTailMaxOut; ++j) {
      if (j == lane) mean = acc[j];
      if (j == lane + A) sraw = acc[j];
    }
    if (lane < A) {
      const float std = (max_std - min_std) * sigmoidf_(sraw + 2.f) + min_std;
      act = tanhf(mean) + std * nz;
    }
  } else {
    float lg = 0.f;
#pragma unroll
    for (int j = 0; j < kTailMaxOut; ++j)
      if (j == lane) lg = acc[j];
    const bool valid = lane < A;
    const int best = sample_group<32>(lg, nz, valid, lane, A, unimix, nullptr);
    act = (valid && lane == best) ? 1.f : 0.f;
  }
  const float ab = act / fmaxf(fabsf(act), 1.f);
  if (lane < A) {
    action[(size_t)row * ld_act + lane] = act;
    abar[(size_t)row * A + lane] = ab;
  }
  // v2[n] = sum_a W2t[a][n] * abar[a] + b2[n]
  if (x2_bf != nullptr && U <= 256) {
    // x2 = SiLU(RMSNorm(v2) * g2) (rssm.py:48) finished here: the row is already in this warp's registers
    float vv[8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = lane + 32 * i;
      float v = 0.f;
      if (n < U) {
        for (int a = 0; a < A; ++a) v = fmaf(__shfl_sync(0xffffffffu, ab, a), w2_s[a * U + n], v);
        v += b2_s[n];
        if (v2) v2[(size_t)row * ld_v2 + n] = v;
      }
      vv[i] = v;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float rs = 1.f / sqrtf(ss / (float)U + kRmsEps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = lane + 32 * i;
      if (n < U) x2_bf[(size_t)row * ld_x2 + n] = __float2bfloat16(siluf_((vv[i] * rs) * __ldg(g2 + n)));
    }
    return;
  }
  for (int n = lane; n < U; n += 32) {
    float v = 0.f;
    for (int a = 0; a < A; ++a) v = fmaf(__shfl_sync(0xffffffffu, ab, a), w2_s[a * U + n], v);
    v2[(size_t)row * ld_v2 + n] = v + b2_s[n];
  }
}

// Register-resident variant of the fused tail for the 13-launch imagination step (units = U = 256, x2 output only).
// The sizes are compile-time, so every loop unrolls and the AO row reductions interleave; all weights (W_last, W_in2,
// b, g2) sit in REGISTERS, fetched before the PDL wait, so after the wait the warp only loads its activation row: no
// shared-memory staging, no block barrier.  Same arithmetic order as actor_tail_kernel (results are bit-identical).
// Measured by ablation (B200, N = 1024): the shared-memory tail cost 12 us per step on the critical path.
template <int AO, int AA>
__global__ void __launch_bounds__(256) actor_tail_x2_kernel(const float* __restrict__ a3, int ld_a3,
                                                            const float* __restrict__ wl_n, int ldk_l,
                                                            const float* __restrict__ bl, int act_kind, float min_std,
                                                            float max_std, float unimix, const float* __restrict__ noise,
                                                            int ld_n, const float* __restrict__ w2_t, int ldw_2,
                                                            const float* __restrict__ b2, int R, float* action, int ld_act,
                                                            float* abar, const float* __restrict__ g2,
                                                            __nv_bfloat16* x2_bf, int ld_x2) {
  static_assert(AO <= 32 && AA <= 32, "one lane per output");
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float wl[AO][8], w2[AA][8], b2r[8], g2r[8], blr[AO];
#pragma unroll
  for (int j = 0; j < AO; ++j) {
    blr[j] = __ldg(bl + j);
#pragma unroll
    for (int i = 0; i < 8; ++i) wl[j][i] = __ldg(wl_n + (size_t)j * ldk_l + lane + 32 * i);
  }
#pragma unroll
  for (int a = 0; a < AA; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) w2[a][i] = __ldg(w2_t + (size_t)a * ldw_2 + lane + 32 * i);
#pragma unroll
  for (int i = 0; i < 8; ++i) { b2r[i] = __ldg(b2 + lane + 32 * i); g2r[i] = __ldg(g2 + lane + 32 * i); }
  const float nz = (row < R && lane < AA) ? noise[(size_t)row * ld_n + lane] : 0.5f;
  pdl_prologue();
  if (row >= R) return;
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a3[(size_t)row * ld_a3 + lane + 32 * i];
  float acc[AO];
#pragma unroll
  for (int j = 0; j < AO; ++j) {
    acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j] = fmaf(x[i], wl[j][i], acc[j]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int j = 0; j < AO; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
#pragma unroll
  for (int j = 0; j < AO; ++j) acc[j] += blr[j];
  float act = 0.f;
  if (act_kind == 0) {
    float mean = 0.f, sraw = 0.f;
#pragma unroll
    for (int j = 0; j < AO; ++j) {
      if (j == lane) mean = acc[j];
      if (j == lane + AA) sraw = acc[j];
    }
    if (lane < AA) {
      const float std = (max_std - min_std) * sigmoidf_(sraw + 2.f) + min_std;
      act = tanhf(mean) + std * nz;
    }
  } else {
    float lg = 0.f;
#pragma unroll
    for (int j = 0; j < AO; ++j)
      if (j == lane) lg = acc[j];
    const bool valid = lane < AA;
    const int best = sample_group<32>(lg, nz, valid, lane, AA, unimix, nullptr);
    act = (valid && lane == best) ? 1.f : 0.f;
  }
  const float ab = act / fmaxf(fabsf(act), 1.f);
  if (lane < AA) {
    action[(size_t)row * ld_act + lane] = act;
    abar[(size_t)row * AA + lane] = ab;
  }
  float abv[AA];
#pragma unroll
  for (int a = 0; a < AA; ++a) abv[a] = __shfl_sync(0xffffffffu, ab, a);
  float vv[8], ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < AA; ++a) v = fmaf(abv[a], w2[a][i], v);
    v += b2r[i];
    vv[i] = v;
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss / 256.f + kRmsEps);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    x2_bf[(size_t)row * ld_x2 + lane + 32 * i] = __float2bfloat16(siluf_((vv[i] * rs) * g2r[i]));
}

// Replay latent write-back (utils/buffer.py:44-53, call site dreamer.py:450): row i of the freshly inferred posterior
// (stoch (S,K) one-hot, deter (D)) goes to storage slot (time_idx[i], env_idx[i]).  The one-hot is stored as S class
// indices (uint8: 32 B per row instead of 2 KB); an optional fp32 one-hot mirror keeps the reference's layout.
// Duplicate slots inside one call (overlapping sampled slices): the row with the LARGEST i wins, as sequential assignment
// would, so the result is deterministic.  Out-of-range indices are skipped and counted in *n_bad.  One CTA per row.
__global__ void __launch_bounds__(256) latent_writeback_kernel(const long long* __restrict__ env_idx,
                                                               const long long* __restrict__ time_idx, int R,
                                                               const float* __restrict__ stoch,
                                                               const float* __restrict__ deter, int S, int K, int D,
                                                               long long n_time, long long n_env, uint8_t* store_idx,
                                                               float* store_stoch, float* store_deter, int* n_bad) {
  pdl_prologue();
  const int row = blockIdx.x;
  const long long e = env_idx[row], t = time_idx[row];
  if (e < 0 || e >= n_env || t < 0 || t >= n_time) {
    if (threadIdx.x == 0 && n_bad) atomicAdd(n_bad, 1);
    return;
  }
  int later = 0;
  for (int j = row + 1 + threadIdx.x; j < R; j += blockDim.x) later |= (env_idx[j] == e && time_idx[j] == t);
  if (__syncthreads_or(later)) return;   // a later row owns this slot
  const size_t slot = (size_t)t * n_env + e;
  const float* dsrc = deter + (size_t)row * D;
  float* ddst = store_deter + slot * D;
  if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(dsrc) | reinterpret_cast<uintptr_t>(ddst)) & 15) == 0) {
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x)
      reinterpret_cast<float4*>(ddst)[i] = reinterpret_cast<const float4*>(dsrc)[i];
  } else {
    for (int i = threadIdx.x; i < D; i += blockDim.x) ddst[i] = dsrc[i];
  }
  const float* ssrc = stoch + (size_t)row * S * K;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {   // first arg-max over the K classes
    float bv = ssrc[s * K];
    int bi = 0;
    for (int c = 1; c < K; ++c) {
      const float v = ssrc[s * K + c];
      if (v > bv) { bv = v; bi = c; }
    }
    store_idx[slot * S + s] = (uint8_t)bi;
  }
  if (store_stoch) {
    float* sdst = store_stoch + slot * (size_t)S * K;
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) sdst[i] = ssrc[i];
  }
}

// Read side (utils/buffer.py:40: `initial` = stored latents of the context step): gather rows from the storage, decoding the
// class indices to exact one-hots.  One CTA per row; rows with an out-of-range index are zero-filled and counted.
__global__ void __launch_bounds__(256) latent_gather_kernel(const long long* __restrict__ env_idx,
                                                            const long long* __restrict__ time_idx, int R, int S, int K,
                                                            int D, long long n_time, long long n_env,
                                                            const uint8_t* __restrict__ store_idx,
                                                            const float* __restrict__ store_deter, float* stoch,
                                                            float* deter, int* n_bad) {
  pdl_prologue();
  const int row = blockIdx.x;
  const long long e = env_idx[row], t = time_idx[row];
  const bool bad = e < 0 || e >= n_env || t < 0 || t >= n_time;
  if (bad && threadIdx.x == 0 && n_bad) atomicAdd(n_bad, 1);
  const size_t slot = bad ? 0 : (size_t)t * n_env + e;
  float* ddst = deter + (size_t)row * D;
  const float* dsrc = store_deter + slot * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) ddst[i] = bad ? 0.f : dsrc[i];
  float* sdst = stoch + (size_t)row * S * K;
  for (int i = threadIdx.x; i < S * K; i += blockDim.x) {
    const int s = i / K, c = i - s * K;
    sdst[i] = (!bad && store_idx[slot * S + s] == c) ? 1.f : 0.f;
  }
}

// TwoHot.mode (distributions.py:78-98): softmax over `bins` logits, then the reference's symmetric
// pairing sum_j (p[m-1-j]*b[m-1-j] + p[m+1+j]*b[m+1+j]) + p[m]*b[m].  One warp per row.
__global__ void twohot_mode_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ bins, int n,
                                   int R, float* out) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* lp = logits + (size_t)warp * ld;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, lp[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(lp[j] - m);
  s = warp_sum(s);
  float acc = 0.f;
  if (n & 1) {
    const int mid = (n - 1) / 2;
    for (int j = lane; j < mid; j += 32) {
      const float lo = (expf(lp[mid - 1 - j] - m) / s) * bins[mid - 1 - j];
      const float hi = (expf(lp[mid + 1 + j] - m) / s) * bins[mid + 1 + j];
      acc += lo + hi;
    }
    acc = warp_sum(acc);
    acc += (expf(lp[mid] - m) / s) * bins[mid];
  } else {
    const int h = n / 2;
    for (int j = lane; j < h; j += 32) {
      const float lo = (expf(lp[h - 1 - j] - m) / s) * bins[h - 1 - j];
      const float hi = (expf(lp[h + j] - m) / s) * bins[h + j];
      acc += lo + hi;
    }
    acc = warp_sum(acc);
  }
  if (lane == 0) out[warp] = acc;
}

// TwoHot.log_prob (distributions.py:100-129), forward and backward w.r.t. the logits.  One warp per row.
//   below = #(bins <= t) - 1, above = n - #(bins > t), both clamped to [0, n-1]; equal -> weights (1, 1)/2, else the
//   distances to the two neighbouring bins; log_prob = w_below * log_softmax[below] + w_above * log_softmax[above];
//   d(log_prob)/d(logit_j) = mixed_j - softmax_j  (mixed = the two-hot target, sums to one).
__device__ __forceinline__ void twohot_target(const float* __restrict__ bins, int n, float t, int lane, int& below, int& above,
                                              float& wb, float& wa) {
  int le = 0, gt = 0;
  for (int j = lane; j < n; j += 32) {
    const float bj = __ldg(bins + j);
    le += (bj <= t) ? 1 : 0;
    gt += (bj > t) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    le += __shfl_xor_sync(0xffffffffu, le, o);
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
  }
  below = min(max(le - 1, 0), n - 1);
  above = min(max(n - gt, 0), n - 1);
  float db = 1.f, da = 1.f;
  if (below != above) {
    db = fabsf(__ldg(bins + below) - t);
    da = fabsf(__ldg(bins + above) - t);
  }
  const float total = db + da;
  wb = da / total;
  wa = db / total;
}
__global__ void twohot_logprob_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ bins, int n,
                                      const float* __restrict__ target, int R, float* out) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* lp = logits + (size_t)warp * ld;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, lp[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(lp[j] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  int below, above;
  float wb, wa;
  twohot_target(bins, n, target[warp], lane, below, above, wb, wa);
  if (lane == 0) {
    // mixed_target * log_pred summed over the bins: only `below` and `above` are non-zero (added in index order)
    const float lb = wb * (lp[below] - lse), la = wa * (lp[above] - lse);
    out[warp] = below == above ? (wb + wa) * (lp[below] - lse) : lb + la;
  }
}
__global__ void twohot_logprob_bwd_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ bins, int n,
                                          const float* __restrict__ target, const float* __restrict__ g, int R, float* d_logits,
                                          int ld_d) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* lp = logits + (size_t)warp * ld;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, lp[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(lp[j] - m);
  s = warp_sum(s);
  int below, above;
  float wb, wa;
  twohot_target(bins, n, target[warp], lane, below, above, wb, wa);
  const float gr = g ? g[warp] : 1.f;
  for (int j = lane; j < n; j += 32) {
    float mixed = 0.f;
    if (j == below) mixed += wb;
    if (j == above) mixed += wa;
    d_logits[(size_t)warp * ld_d + j] = gr * (mixed - expf(lp[j] - m) / s);
  }
}

// cont head mean = sigmoid(logit) (distributions.py:238-239, Bernoulli.mean).
// ---- backward of the imagined head evaluation (attack shape: d(lambda-return) / d(feats), frozen weights)
// TwoHot.mode = sum_j softmax(l)_j * bins_j (distributions.py:78-98): d l_j = p_j (bins_j - mode) d_mode.  In place over the
// logits (row stride ld); one warp per row.
__global__ void twohot_mode_bwd_kernel(float* __restrict__ logits, int ld, const float* __restrict__ bins, int n, int R,
                                       const float* __restrict__ mode, const float* __restrict__ d_mode) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  float* lp = logits + (size_t)warp * ld;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, lp[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(lp[j] - m);
  s = warp_sum(s);
  const float mo = mode[warp], g = d_mode[warp];
  for (int j = lane; j < n; j += 32) lp[j] = (expf(lp[j] - m) / s) * (bins[j] - mo) * g;
}
// Bernoulli mean = sigmoid(l): d l = c (1 - c) d_c; writes the gradient over the logit (column 0 of a row of stride ld).
__global__ void sigmoid_bwd_kernel(float* __restrict__ logit, int ld, const float* __restrict__ c, const float* __restrict__ d_c, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) logit[(size_t)i * ld] = c[i] * (1.f - c[i]) * d_c[i];
}
// Reverse of imag_weight_ret_kernel's recursion (dreamer.py:694-707 made differentiable: the attack's objective):
//   ret_i = r_{i+1} + live_{i+1} ((1 - lamb) v_{i+1} + lamb ret_{i+1}),  live = c * disc,  ret_{H-1} := v_{H-1}
// a_i = total gradient reaching ret_i.  Optional direct cotangents of reward / cont / value are added.
__global__ void lambda_return_bwd_kernel(int N, int H, const float* __restrict__ reward, const float* __restrict__ cont,
                                         const float* __restrict__ value, float disc, float lamb, const float* __restrict__ d_ret,
                                         const float* __restrict__ g_reward, const float* __restrict__ g_cont,
                                         const float* __restrict__ g_value, float* d_reward, float* d_cont, float* d_value) {
  pdl_prologue();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t o = (size_t)n * H;
  // forward values of ret (needed by d_cont), recomputed: ret_{i+1} for i = H-2 .. 0
  float retv[64];
  retv[H - 1] = value[o + H - 1];
  for (int i = H - 2; i >= 0; --i) {
    const float live = cont[o + i + 1] * disc;
    retv[i] = reward[o + i + 1] + live * ((1.f - lamb) * value[o + i + 1] + lamb * retv[i + 1]);
  }
  d_reward[o] = g_reward ? g_reward[o] : 0.f;
  d_cont[o] = g_cont ? g_cont[o] : 0.f;
  d_value[o] = g_value ? g_value[o] : 0.f;
  float a = 0.f;
  for (int i = 0; i <= H - 2; ++i) {
    a = (d_ret ? d_ret[(size_t)n * (H - 1) + i] : 0.f) + a;    // a_i = g_i + carried
    const float c = cont[o + i + 1], v = value[o + i + 1];
    const float dr = a, dc = a * disc * ((1.f - lamb) * v + lamb * retv[i + 1]);
    float dv = a * c * disc * (1.f - lamb);
    const float carry = a * c * disc * lamb;                   // into ret_{i+1}
    if (i + 1 == H - 1) dv += carry;                           // ret_{H-1} is value_{H-1}
    d_reward[o + i + 1] = dr + (g_reward ? g_reward[o + i + 1] : 0.f);
    d_cont[o + i + 1] = dc + (g_cont ? g_cont[o + i + 1] : 0.f);
    d_value[o + i + 1] = dv + (g_value ? g_value[o + i + 1] : 0.f);
    a = carry;
  }
}
// out (+)= in, contiguous
__global__ void accum_kernel(const float4* __restrict__ in, float4* out, long long n4, int first) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = in[i];
    if (!first) { const float4 b = out[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    out[i] = a;
  }
}
__global__ void sigmoid_kernel(const float* __restrict__ in, int ld, float* out, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sigmoidf_(in[(size_t)i * ld]);
}

// weight = cumprod(cont*disc) (dreamer.py:596) and the lambda-return reverse scan
// (dreamer.py:694-707).  One thread per row; T <= 128.
__global__ void lambda_return_kernel(int N, int T, const float* __restrict__ last, const float* __restrict__ term,
                                     const float* __restrict__ reward, const float* __restrict__ value,
                                     const float* __restrict__ boot, float disc, float lamb, float* out) {
  pdl_prologue();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t o = (size_t)n * T;
  float nxt = boot[o + T - 1];
  for (int i = T - 2; i >= 0; --i) {
    const float live = (1.f - term[o + i + 1]) * disc;
    const float cont = (1.f - (last ? last[o + i + 1] : 0.f)) * lamb;
    const float interm = reward[o + i + 1] + (1.f - cont) * live * boot[o + i + 1];
    nxt = interm + live * cont * nxt;
    out[(size_t)n * (T - 1) + i] = nxt;
  }
}
__global__ void imag_weight_ret_kernel(int N, int H, const float* __restrict__ reward, const float* __restrict__ cont,
                                       const float* __restrict__ value, float disc, float lamb, float* weight,
                                       float* ret) {
  pdl_prologue();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t o = (size_t)n * H;
  if (weight) {
    float w = 1.f;
    for (int i = 0; i < H; ++i) {
      w *= cont[o + i] * disc;
      weight[o + i] = w;
    }
  }
  if (ret) {
    float nxt = value[o + H - 1];
    for (int i = H - 2; i >= 0; --i) {
      const float live = (1.f - (1.f - cont[o + i + 1])) * disc;  // term = 1 - cont, as the reference computes it
      const float c = lamb;                                      // last = 0
      const float interm = reward[o + i + 1] + (1.f - c) * live * value[o + i + 1];
      nxt = interm + live * c * nxt;
      ret[(size_t)n * (H - 1) + i] = nxt;
    }
  }
}

// ReturnEMA.__call__ (networks.py:416-422): q05 / q95 of the flattened returns (torch.quantile, linear interpolation),
// EMA(alpha) of the two, offset = ema[0], scale = max(ema[1] - ema[0], 1).
// One CTA.  An order statistic is found by an MSB-first radix select (4 passes of 8 bits over order-preserving uint32
// keys, shared-memory histogram) -- exact, any n, no sort.  The rank / interpolation arithmetic follows torch in fp32:
// rank = q * (n - 1), below = floor, above = ceil, lerp(a, b, w) = w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w).
__device__ __forceinline__ uint32_t f32_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(1024) return_ema_kernel(const float* __restrict__ x, long long n, float a32, float b32,
                                                         float* ema, float* offset, float* scale) {
  pdl_prologue();
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sel_prefix;
  __shared__ long long sel_k;
  __shared__ float stat[4];
  const float qs[2] = {0.05f, 0.95f};
  for (int s = 0; s < 4; ++s) {
    const float rank = qs[s >> 1] * (float)(n - 1);
    long long kth = (long long)((s & 1) ? ceilf(rank) : floorf(rank));
    if (kth < 0) kth = 0;
    if (kth > n - 1) kth = n - 1;
    unsigned int prefix = 0, mask = 0;
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t key = f32_key(x[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        long long kk = kth;
        int d = 0;
        for (; d < 255; ++d) {
          if (kk < (long long)hist[d]) break;
          kk -= hist[d];
        }
        sel_prefix = prefix | ((unsigned int)d << shift);
        sel_k = kk;
      }
      __syncthreads();
      prefix = sel_prefix;
      kth = sel_k;
      mask |= 255u << shift;
      __syncthreads();
    }
    if (threadIdx.x == 0) stat[s] = key_f32(prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float qv[2];
    for (int j = 0; j < 2; ++j) {
      const float rank = qs[j] * (float)(n - 1);
      const float w = rank - floorf(rank);
      const float a = stat[2 * j], b = stat[2 * j + 1];
      const float d = __fsub_rn(b, a);
      qv[j] = (w < 0.5f) ? __fadd_rn(a, __fmul_rn(w, d)) : __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, w)));
    }
    const float e0 = __fadd_rn(__fmul_rn(a32, qv[0]), __fmul_rn(b32, ema[0]));
    const float e1 = __fadd_rn(__fmul_rn(a32, qv[1]), __fmul_rn(b32, ema[1]));
    ema[0] = e0; ema[1] = e1;
    if (offset) *offset = e0;
    if (scale) *scale = fmaxf(__fsub_rn(e1, e0), 1.f);
  }
}


// RSSM.kl_loss values + unimix entropies (rssm.py:222-230; distributions.py:266-271; dreamer.py:575-576).
// One thread per (row, category); per-row sums are reduced in a fixed order by the last stage.
__global__ void kl_entropy_kernel(const float* __restrict__ post, const float* __restrict__ prior, int R, int S, int K,
                                  float unimix, float* kl_sk, float* ent_post_sk, float* ent_prior_sk) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const float* a = post + (size_t)i * K;
  const float* b = prior + (size_t)i * K;
  float ma = -INFINITY, mb = -INFINITY;
  for (int k = 0; k < K; ++k) { ma = fmaxf(ma, a[k]); mb = fmaxf(mb, b[k]); }
  float sa = 0.f, sb = 0.f;
  for (int k = 0; k < K; ++k) { sa += expf(a[k] - ma); sb += expf(b[k] - mb); }
  const float lsa = ma + logf(sa), lsb = mb + logf(sb);
  float kl = 0.f;
  for (int k = 0; k < K; ++k) kl += (expf(a[k] - ma) / sa) * ((a[k] - lsa) - (b[k] - lsb));
  kl_sk[i] = kl;
  const float uni = unimix / (float)K;
  for (int which = 0; which < 2; ++which) {
    const float* x = which ? b : a;
    const float mx = which ? mb : ma, sx = which ? sb : sa;
    float* dst = which ? ent_prior_sk : ent_post_sk;
    if (!dst) continue;
    float l[32];
    float m2 = -INFINITY;
    for (int k = 0; k < K; ++k) { l[k] = logf((expf(x[k] - mx) / sx) * (1.f - unimix) + uni); m2 = fmaxf(m2, l[k]); }
    float s2 = 0.f;
    for (int k = 0; k < K; ++k) s2 += expf(l[k] - m2);
    const float lse = m2 + logf(s2);
    float ent = 0.f;
    for (int k = 0; k < K; ++k) { const float ll = l[k] - lse; ent -= expf(ll) * ll; }
    dst[i] = ent;
  }
}
__global__ void kl_finish_kernel(const float* __restrict__ kl_sk, const float* __restrict__ ep_sk,
                                 const float* __restrict__ eq_sk, int R, int S, float free_nats, float* dyn, float* rep,
                                 float* ent_post, float* ent_prior) {
  pdl_prologue();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float k = 0.f, ep = 0.f, eq = 0.f;
  for (int s = 0; s < S; ++s) {
    k += kl_sk[(size_t)r * S + s];
    if (ent_post) ep += ep_sk[(size_t)r * S + s];
    if (ent_prior) eq += eq_sk[(size_t)r * S + s];
  }
  const float v = fmaxf(k, free_nats);
  if (dyn) dyn[r] = v;
  if (rep) rep[r] = v;
  if (ent_post) ent_post[r] = ep;
  if (ent_prior) ent_prior[r] = eq;
}

// Backward of RSSM.kl_loss (rssm.py:222-230): rep = clip(KL(post || sg(prior)).sum_s, free) sends gradient to the posterior
// logits only, dyn = clip(KL(sg(post) || prior).sum_s, free) to the prior logits only; the clip passes gradient where the
// row's KL >= free (torch.clamp's mask).  Per category: dKL/d post_j = p_j ((log p_j - log q_j) - KL_s),
// dKL/d prior_j = q_j - p_j.  One thread per (row, category); kl_sk = per-category KL from kl_entropy_kernel.
__global__ void kl_grad_kernel(const float* __restrict__ post, const float* __restrict__ prior, const float* __restrict__ kl_sk,
                               int R, int S, int K, float free_nats, const float* __restrict__ g_dyn,
                               const float* __restrict__ g_rep, float* d_post, float* d_prior) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int r = i / S;
  float tot = 0.f;
  for (int s = 0; s < S; ++s) tot += kl_sk[(size_t)r * S + s];   // same order as kl_finish_kernel
  const float act = tot >= free_nats ? 1.f : 0.f;
  const float gd = act * (g_dyn ? g_dyn[r] : 1.f), gr = act * (g_rep ? g_rep[r] : 1.f);
  const float* a = post + (size_t)i * K;
  const float* b = prior + (size_t)i * K;
  float ma = -INFINITY, mb = -INFINITY;
  for (int k = 0; k < K; ++k) { ma = fmaxf(ma, a[k]); mb = fmaxf(mb, b[k]); }
  float sa = 0.f, sb = 0.f;
  for (int k = 0; k < K; ++k) { sa += expf(a[k] - ma); sb += expf(b[k] - mb); }
  const float lsa = ma + logf(sa), lsb = mb + logf(sb);
  const float kls = kl_sk[i];
  for (int k = 0; k < K; ++k) {
    const float lp = a[k] - lsa, lq = b[k] - lsb;
    const float pk = expf(lp), qk = expf(lq);
    if (d_post) d_post[(size_t)i * K + k] = gr * pk * ((lp - lq) - kls);
    if (d_prior) d_prior[(size_t)i * K + k] = gd * (qk - pk);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused multi-tensor AGC + LaProp step (utils/optim/agc.py:15-60, utils/optim/laprop.py:46-118).
// The reference runs two foreach norms, four foreach ops and then a Python loop of ~8 kernels per tensor; here the whole
// update of every tensor is three launches over a device-resident tensor table:
//   opt_norm_kernel     per 8192-element chunk: sum p^2, sum g^2                        (fixed tree => deterministic)
//   opt_finalize_kernel per tensor: sums the chunk partials in order, scale = 1 / max(||g|| / (clip * max(||p||, pmin)), 1)
//                       (1 when clip <= 0), raises found_inf when a gradient norm is not finite
//   opt_update_kernel   g' = g * scale * inv_scale;  v = beta2 v + (1 - beta2) g'^2;  m = beta1 m + (1 - beta1) lr g' / (sqrt(v / bc2) + eps);
//                       p -= step_size * m  (then p -= wd * p); skipped entirely when found_inf is set
// ------------------------------------------------------------------------------------------------
struct OptTensor {
  float* p; float* g; float* m; float* v;
  long long n;
  int blk0, nblk;
};
constexpr int kOptChunk = 8192;
__device__ __forceinline__ int opt_find(const OptTensor* __restrict__ t, int count, int b) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid].blk0 <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256) opt_norm_kernel(const OptTensor* __restrict__ t, int count, float* partial) {
  pdl_prologue();
  __shared__ float sh[32];
  const int ti = opt_find(t, count, blockIdx.x);
  const OptTensor e = t[ti];
  const long long i0 = (long long)(blockIdx.x - e.blk0) * kOptChunk;
  const long long i1 = min(e.n, i0 + kOptChunk);
  float sp = 0.f, sg = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float pv = e.p[i], gv = e.g[i];
    sp = fmaf(pv, pv, sp);
    sg = fmaf(gv, gv, sg);
  }
  sp = block_sum(sp, sh);
  sg = block_sum(sg, sh);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sp; partial[2 * blockIdx.x + 1] = sg; }
}
// inv_scale: GradScaler's unscale factor.  The reference unscales the gradients first (scaler.unscale_, dreamer.py:422) and
// clips the UNSCALED gradients (dreamer.py:432), so the gradient norm that enters the clip factor is ||g|| * inv_scale.
__global__ void __launch_bounds__(256) opt_finalize_kernel(const OptTensor* __restrict__ t, int count, const float* __restrict__ partial,
                                                           float clip, float pmin, float inv_scale, float* scale, int* found_inf) {
  pdl_prologue();
  for (int ti = threadIdx.x; ti < count; ti += 256) {
    const OptTensor e = t[ti];
    float sp = 0.f, sg = 0.f;
    for (int b = 0; b < e.nblk; ++b) { sp += partial[2 * (e.blk0 + b)]; sg += partial[2 * (e.blk0 + b) + 1]; }
    float s = 1.f;
    if (clip > 0.f) {
      const float upper = fmaxf(sqrtf(sp), pmin) * clip;
      s = 1.f / fmaxf((sqrtf(sg) * inv_scale) / upper, 1.f);
    }
    scale[ti] = s;
    if (found_inf && !isfinite(sg)) atomicOr(found_inf, 1);
  }
}
__global__ void __launch_bounds__(256) opt_update_kernel(const OptTensor* __restrict__ t, int count, const float* __restrict__ scale,
                                                         const int* __restrict__ found_inf, float inv_scale, float beta1, float beta2,
                                                         float omb2, float lr_term, float step_size, float bc2, float eps, float wd,
                                                         int write_grads) {
  pdl_prologue();
  if (found_inf && *found_inf) return;   // GradScaler semantics: a non-finite gradient skips the whole step
  const int ti = opt_find(t, count, blockIdx.x);
  const OptTensor e = t[ti];
  const float gs = scale[ti];
  const long long i0 = (long long)(blockIdx.x - e.blk0) * kOptChunk;
  const long long i1 = min(e.n, i0 + kOptChunk);
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float g = __fmul_rn(__fmul_rn(e.g[i], inv_scale), gs);   // unscale, then clip (dreamer.py:422,432)
    if (write_grads) e.g[i] = g;                    // scaler.unscale_ / clip_grad_agc_ both work on p.grad in place
    const float v = __fadd_rn(__fmul_rn(e.v[i], beta2), __fmul_rn(__fmul_rn(omb2, g), g));
    e.v[i] = v;
    const float denom = __fadd_rn(sqrtf(v / bc2), eps);
    const float s = g / denom;
    const float m = __fadd_rn(__fmul_rn(e.m[i], beta1), __fmul_rn(lr_term, s));
    e.m[i] = m;
    float pv = __fadd_rn(e.p[i], -__fmul_rn(step_size, m));
    if (wd != 0.f) pv = __fadd_rn(pv, -__fmul_rn(wd, pv));
    e.p[i] = pv;
  }
}
__global__ void __launch_bounds__(256) opt_scale_grads_kernel(const OptTensor* __restrict__ t, int count, const float* __restrict__ scale) {
  pdl_prologue();
  const int ti = opt_find(t, count, blockIdx.x);
  const OptTensor e = t[ti];
  const float gs = scale[ti];
  const long long i0 = (long long)(blockIdx.x - e.blk0) * kOptChunk;
  const long long i1 = min(e.n, i0 + kOptChunk);
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) e.g[i] = __fmul_rn(e.g[i], gs);
}

// ------------------------------------------------------------------------------------------------
// Barlow-twins redundancy loss between projected latents x1 (N, E) and detached embeddings x2 (N, E)
// (dreamer.py:525-532): columns standardised with the unbiased std (+1e-8), c = x1n^T x2n / N,
// loss = sum_i (c_ii - 1)^2 + lambd * sum_{i != j} c_ij^2, and its gradient w.r.t. x1.
// The two E x E x N contractions run on the fp32 skinny GEMM above (3xTF32 tensor tiles); these kernels do the column
// statistics, the standardisation (writing x1n transposed so that both GEMM operands are K-major), the loss / dL/dc
// epilogue and the backward of the standardisation.  All reductions have a fixed order.
// ------------------------------------------------------------------------------------------------
// mean / unbiased std of every column of x (R x W); block = 32 columns x 32 row lanes (same shape as colsum_kernel).
__global__ void __launch_bounds__(1024) col_meanstd_kernel(const float* __restrict__ x, int ld, int R, int W, float* mean, float* stdv) {
  pdl_prologue();
  __shared__ float sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < W)
    for (int r = threadIdx.y; r < R; r += 32) s += x[(size_t)r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float mu = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) mu += sh[j][threadIdx.x];
  mu /= (float)R;
  __syncthreads();
  float q = 0.f;
  if (c < W)
    for (int r = threadIdx.y; r < R; r += 32) { const float d = x[(size_t)r * ld + c] - mu; q = fmaf(d, d, q); }
  sh[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += sh[j][threadIdx.x];
    mean[c] = mu;
    stdv[c] = sqrtf(t / (float)(R - 1));
  }
}
// xn = (x - mean) / (std + 1e-8): row-major copy xn (nullable) and transposed copy xnT (W x R, nullable) through a 32x32 tile.
__global__ void __launch_bounds__(1024) standardise_kernel(const float* __restrict__ x, int ld, int R, int W,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv,
                                                           float* xn, float* xnT) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
  float v = 0.f;
  if (c < W && r < R) {
    v = (x[(size_t)r * ld + c] - mean[c]) / (stdv[c] + 1e-8f);
    if (xn) xn[(size_t)r * W + c] = v;
  }
  tile[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  const int ct = blockIdx.x * 32 + threadIdx.y, rt = blockIdx.y * 32 + threadIdx.x;
  if (xnT && ct < W && rt < R) xnT[(size_t)ct * R + rt] = tile[threadIdx.x][threadIdx.y];
}
// craw = x1n^T x2n (E x E).  c = craw / N; per-block partial sums of the loss; dcT[j][i] = dL/d(craw)[i][j].
__global__ void __launch_bounds__(256) barlow_loss_kernel(const float* __restrict__ craw, int E, int N, float lambd, float* partial,
                                                          float* dcT) {
  pdl_prologue();
  __shared__ float sh[32];
  const long long idx = blockIdx.x * 256ll + threadIdx.x;
  float term = 0.f;
  if (idx < (long long)E * E) {
    const int i = (int)(idx / E), j = (int)(idx - (long long)i * E);
    const float c = craw[idx] / (float)N;
    float g;
    if (i == j) { term = (c - 1.f) * (c - 1.f); g = 2.f * (c - 1.f); }
    else { term = lambd * (c * c); g = 2.f * lambd * c; }
    dcT[(size_t)j * E + i] = g / (float)N;
  }
  term = block_sum(term, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = term;
}
__global__ void sum_in_order_kernel(const float* __restrict__ partial, int n, float* out) {
  pdl_prologue();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = s;
  }
}
// Backward of the standardisation: dx = (g - mean(g)) / s - xn * sum(g * xn) / ((N - 1) * s) * (s - 1e-8) ... written with
// sigma = std, s = sigma + 1e-8:  dx_r = (g_r - gbar) / s - (x_r - mu) * sum_q(g_q xn_q) / ((N - 1) sigma s).
// Column reductions as in col_meanstd_kernel; g = d(loss)/d(xn) (R x W), xnT is the transposed standardised copy.
__global__ void __launch_bounds__(1024) standardise_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xnT,
                                                               const float* __restrict__ stdv, int R, int W, float* dx, int ld_dx) {
  pdl_prologue();
  __shared__ float sh[32][33];
  __shared__ float sh2[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float sg = 0.f, sgx = 0.f;
  if (c < W)
    for (int r = threadIdx.y; r < R; r += 32) {
      const float gv = g[(size_t)r * W + c];
      sg += gv;
      sgx = fmaf(gv, xnT[(size_t)c * R + r], sgx);
    }
  sh[threadIdx.y][threadIdx.x] = sg;
  sh2[threadIdx.y][threadIdx.x] = sgx;
  __syncthreads();
  float tg = 0.f, tgx = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { tg += sh[j][threadIdx.x]; tgx += sh2[j][threadIdx.x]; }
  if (c < W) {
    const float sigma = stdv[c], s = sigma + 1e-8f;
    const float gbar = tg / (float)R;
    // (x_r - mu) = xn_r * s
    const float k2 = tgx / ((float)(R - 1) * sigma);
    for (int r = threadIdx.y; r < R; r += 32)
      dx[(size_t)r * ld_dx + c] = (g[(size_t)r * W + c] - gbar) / s - xnT[(size_t)c * R + r] * k2;
  }
}

}  // namespace sd
