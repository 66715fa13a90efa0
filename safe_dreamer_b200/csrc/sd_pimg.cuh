// sd_pimg.cuh -- persistent, team-resident imagination scan on tcgen05 (sm_100a).
//
// Dreamer._imagine (dreamer.py:673-692): H iterations of  feat = [stoch | deter] -> action = actor(feat).rsample()
// -> stoch, deter = img_step(stoch, deter, action)  (rssm.py:36-75, 180-195; networks.py:339-377).
//
// ONE launch runs all H iterations.  A TEAM of 16 CTAs owns a group of 128 rows for the whole rollout (rows-stationary;
// 1024 rows = 8 teams = 128 SMs, one CTA per SM).  Two kinds of phases alternate inside an iteration:
//
//   column-split (wide contractions: the CTAs of a team split the OUTPUT COLUMNS, every CTA streams the whole activation)
//     P7    deter (K = 2048) -> dyn_in0 | img_net_0 | actor_0[:, 512:]     16 columns of each per CTA (N = 48)
//     ZIN   stoch (K = 512)  -> actor_0[:, :512] (on top of P7's partial) | dyn_in1
//     HID   [deter_g | x0 | x1 | x2] (K = 1024) -> 128 of the 2048 hidden columns (CTA c: block c/2, half c%2)
//     GRU   h_g (K = 256) -> reset | cand | update of the same 128 units -> GRU gates -> deter'
//   row-split (the 256-wide MLP chains: CTA c owns rows [8c, 8c+8) of the group and ALL columns, so RMSNorm is local and
//   a chain needs no synchronisation at all; the MMAs are swapped: weights are the M operand, the 8 rows the N operand)
//     img chain    RMSNorm/SiLU(o0) -> img_net_1 -> logits -> unimix + Gumbel arg-max -> stoch'
//     actor chain  RMSNorm/SiLU(a0) -> actor_1 -> actor_2 -> head -> action sample -> dyn_in2 -> x2
//   The column-split producers hand over PRE-norm values (fp32); the row-split consumer normalises them (it sees whole
//   rows), which removes every row-statistics exchange except the 2048-wide one of the hidden layer.
//
// Per CTA: 2 threads stream weight groups and 2 threads stream activation chunks (TMA, 3 x 32 KB rings each; a ring entry
// is 2..4 boxes under ONE mbarrier because a producer iteration -- wait + expect_tx + first box -- costs ~850 cycles
// whatever the box size: profiles/r02_tma_issue_probe.txt); one thread issues tcgen05.mma into fixed TMEM column ranges;
// warps 0-7 are the epilogue.  All roles walk the SAME static schedule (`walk`), so the rings cannot get out of step; the
// weight ring never depends on data and runs ahead across phase boundaries.
// Teams are NOT thread-block clusters: a B200 holds only 7 co-resident 16-CTA clusters with this shared-memory footprint
// (profiles/r02_cluster_probe.txt), one short of the 8 teams of the base shape, so the 7 team synchronisations of an
// iteration go through L2 (red.release.gpu / ld.relaxed + fence.acq_rel.gpu on per-team counters, one 128-byte line each).
// Every wait is bounded (trap instead of hang).
#pragma once
#include "sd_kernels.cuh"
#include "sd_tc.cuh"

namespace sd {
namespace pimg {

using tc::smem_u32;

constexpr int CL = 16;          // CTAs per team
constexpr int BM = 128, BK = 64;
constexpr int NEPI = 16;        // epilogue warps (four per TMEM lane quadrant)
constexpr int NPW = 2, NPA = 2; // weight / activation TMA threads (different warps)
constexpr int W_WARP = NEPI, A_WARP = W_WARP + NPW, M_WARP = A_WARP + NPA;
constexpr int THREADS = 32 * (M_WARP + 1);
constexpr int NA = 3, NW = 3;          // ring depths
constexpr int kSlab = BM * BK * 2;     // 16 KB: one k-block of an activation chunk
constexpr int kEntry = 2 * kSlab;      // 32 KB ring entry: an activation chunk (2 slabs) or a weight group (<= 32 KB of boxes)
constexpr int ROWS = BM / CL;          // 8 rows of the group per CTA in the row-split chains

// architecture the kernel is specialised for (configs/base.yaml:117-127,252-276); the host checks it
constexpr int D = 2048, U = 256, SK = 512, KC = 16, G = 8, DG = 256, F = SK + D;
// bf16 exchange buffer [rows][ACT_LD]: x = [x0 | x1 | x2], h (TMA operands of the column-split phases)
constexpr int CX = 0, CH = 768, ACT_LD = 2816;
// fp32 pre-norm hand-over buffer [rows][RAW_LD]
constexpr int RX0 = 0, RO0 = 256, RA0 = 512, RX1 = 768, RAW_LD = 1024;
// TMEM columns
constexpr uint32_t T_HID = 0, T_GRU = 128, T_X0 = 128, T_O0 = 144, T_A0 = 160, T_X1 = 176, T_CH = 192;

enum { X_D = 0, X_P7, X_Z, X_ZIN, X_X2, X_H, NX };
enum { ACC_P7 = 0, ACC_CH, ACC_ZIN, ACC_HID, ACC_GRU, NACC };
enum { W_P7 = 0, W_Z, W_A1, W_A2, W_I1, W_LG, W_HID, W_GRU, NWMAP };
enum { A_BIG = 0, A_ACT = 1 };

constexpr int kTailFloats = 5120;      // W_last (act_out x 256) + W_in2 (A x 256) + b2 + g2 when act_out + A <= 18
constexpr int kOffA = 0;
constexpr int kOffW = kOffA + NA * kEntry;
constexpr int kOffB = kOffW + NW * kEntry;             // chain operand: [4 k-blocks][16 rows][128 B], also fp32 a2 [8][256]
constexpr int kOffTail = kOffB + 4 * 2048;
constexpr int kOffConst = kOffTail + kTailFloats * 4;  // per-CTA bias slices of the column-split phases
constexpr int kConstFloats = 4 * 16 + 128 + 128 + 384;
constexpr int kOffRed = kOffConst + kConstFloats * 4;  // [16 warps][4 rows] + [4][128]
constexpr int kOffBar = kOffRed + (NEPI * (ROWS / 2) + 4 * BM) * 4;
constexpr int kNumBar = 2 * NA + 2 * NW + NACC + 1;
constexpr int kSmemBytes = kOffBar + kNumBar * 8 + 16 + 1024;   // + tmem slot + alignment slack
static_assert(kSmemBytes <= 232448, "shared-memory budget");

constexpr int kFlagStride = 32;               // one 128-byte line per counter
__host__ __device__ constexpr size_t flags_per_team() { return (size_t)(NX + 1) * kFlagStride; }
__host__ __device__ constexpr size_t ssq_per_team() { return (size_t)CL * BM; }

struct Params {
  CUtensorMap ma[2];        // A_BIG: big_bf [N][H*F]; A_ACT: exchange buffer [N][ACT_LD]   (box 64 x 128)
  CUtensorMap mw[NWMAP];    // weight maps (box 64 x {48, 32, 128, 128, 128, 128, 128, 128})
  int N, H, ngroups;
  float* feats;             // [N][H][F] fp32
  float* actions;           // [N][H][A] fp32
  __nv_bfloat16* big_bf;    // [N][H][F]
  __nv_bfloat16* act;       // [N][ACT_LD]
  float* raw;               // [N][RAW_LD]
  const float* u;           // [N][H][SK] uniforms of the prior samples
  const float* act_noise;   // [N][H][A]
  const float *b_in0, *g_in0, *b_in1, *g_in1, *b_in2, *g_in2, *b_hid, *g_hid, *b_gru, *b_i0, *g_i0, *b_i1, *g_i1, *b_lg;
  const float *b_a0, *g_a0, *b_a1, *g_a1, *b_a2, *g_a2, *b_last;
  const float* w_last; int ldk_last;    // [act_out][ldk] fp32
  const float* w_in2; int ldw_in2;      // [A][ldw] fp32 (n contiguous)
  int A, act_out, act_kind, tail_in_smem;
  unsigned int* flags;      // [teams][NX + 1][32] monotonic counters (zeroed before the launch)
  float* ssq;               // [teams][16 ranks][128 rows] partial sums of squares of the hidden layer
  float min_std, max_std, act_unimix, unimix;
  long long* timing;        // diagnostic: clock64 stamps of CTA 0 (null in production)
};

// ------------------------------------------------------------------------------------------------ primitives
// bounded waits: ~2 s of polling, then trap
__device__ __forceinline__ void wait_local(uint32_t bar, uint32_t parity) {
  if (tc::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i) if (tc::mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// team counter: every CTA of the team adds 1 per completed exchange; waiters poll for 16 x (completions so far)
__device__ __forceinline__ void flag_signal(unsigned int* f) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned int* f, unsigned int target) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  if (v < target) {
    const long long t0 = clock64();
#pragma unroll 1
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// one MUFU op each (tanh.approx.f32, |rel err| ~ 2^-11: below the bf16 rounding of every value these feed)
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_fast(float y) { return y * sigmoid_fast(y); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack_bf8(const float* y) {
  return make_uint4(pack_bf2(y[0], y[1]), pack_bf2(y[2], y[3]), pack_bf2(y[4], y[5]), pack_bf2(y[6], y[7]));
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------ the schedule
// One walk = everything one CTA's producer / MMA roles do, in issue order.  Visitor callbacks:
//   v.wait(x)                       activation producers: exchange x of this iteration must have completed
//   v.a(map, col, row)              next activation chunk: 2 slabs (128 rows x 64 k) at columns col, col + 64
//   v.w(map, k0, row, nrows, nk)    next weight group: nk boxes of nrows rows at k-blocks k0, k0 + 64, ...
//   v.mma(ai, wi, wrow, n, tcol, acc)   slab ai of the current chunk x rows [wrow, wrow+n) of box wi of the current group
//   v.aw(wmap, k0, row)             a weight group (2 boxes of 128 rows) loaded into the ACTIVATION ring: during the row-split
//                                   chains that ring is idle, so alternating the chain weights between the two rings doubles
//                                   the bytes in flight
//   v.mma_sw(wi, kb, tcol, acc, in_a)   swapped: box wi (128 weight rows) of the current group (weight ring, or activation
//                                   ring when in_a) x k-block kb of the chain operand
//   v.bwait()                       the chain operand in shared memory has been (re)written by the epilogue warps
//   v.ra() / v.rw()                 last use of the current chunk / group
//   v.done(bar)                     accumulator `bar` is complete
// one weight group of a swapped chain layer: feature tile q / 2, k-blocks 2 (q % 2) and 2 (q % 2) + 1
template <class V>
__device__ __forceinline__ void chain_group(V& v, int wmap, int q) {
  const bool in_a = (q & 1) != 0;
  if (in_a) v.aw(wmap, (q & 1) * 2 * BK, (q >> 1) * 128);
  else      v.w(wmap, (q & 1) * 2 * BK, (q >> 1) * 128, 128, 2);
  v.mma_sw(0, (q & 1) * 2, T_CH + (q >> 1) * 16, (q & 1) != 0, in_a);
  v.mma_sw(1, (q & 1) * 2 + 1, T_CH + (q >> 1) * 16, true, in_a);
  if (in_a) v.ra(); else v.rw();
}

template <class V>
__device__ __forceinline__ void walk(const Params& P, int rank, V& v) {
  const int g = rank >> 1, hf = rank & 1;
  const int hrow = g * DG + hf * 128;
  for (int grp = (int)(blockIdx.x / CL); grp < P.ngroups; grp += (int)(gridDim.x / CL)) {
    const int row0 = grp * BM;
    for (int i = 0; i < P.H; ++i) {
      const bool next = i + 1 < P.H;
      // ---- P7: deter_i -> x0 | o0 | a0 (deter part); this CTA's own deter block also feeds its hidden columns
      if (i > 0) v.wait(X_D);
      for (int c = 0; c < D / (2 * BK); ++c) {
        v.w(W_P7, c * 2 * BK, rank * 48, 48, 2);
        v.a(A_BIG, i * F + SK + c * 2 * BK, row0);
        v.mma(0, 0, 0, 48, T_X0, c > 0);
        v.mma(1, 1, 0, 48, T_X0, true);
        v.rw();
        if (next && (c >> 1) == g) {   // this CTA's own deter block: first quarter of its hidden-layer contraction
          v.w(W_HID, (c & 1) * 2 * BK, hrow, 128, 2);
          v.mma(0, 0, 0, 128, T_HID, (c & 1) != 0);
          v.mma(1, 1, 0, 128, T_HID, true);
          v.rw();
        }
        v.ra();
      }
      v.done(ACC_P7);
      // ---- img chain (row-split, swapped): img_net_1, logits
      if (i > 0) {
        v.bwait();
        for (int q = 0; q < 4; ++q) chain_group(v, W_I1, q);   // 2 feature tiles x 2 k-block pairs
        v.done(ACC_CH);
        v.bwait();
        for (int q = 0; q < 8; ++q) chain_group(v, W_LG, q);   // 4 feature tiles (512 logits) x 2 k-block pairs
        v.done(ACC_CH);
      }
      // ---- ZIN: stoch_i -> a0 (on top of the deter part) | x1
      v.wait(X_Z);
      for (int c = 0; c < SK / (2 * BK); ++c) {
        if ((c & 1) == 0) v.w(W_Z, c * 2 * BK, rank * 32, 32, 4);
        v.a(A_BIG, i * F + c * 2 * BK, row0);
        for (int s = 0; s < 2; ++s) {
          v.mma(s, (c & 1) * 2 + s, 0, 16, T_A0, true);
          v.mma(s, (c & 1) * 2 + s, 16, 16, T_X1, c > 0 || s > 0);
        }
        if (c & 1) v.rw();
        v.ra();
      }
      v.done(ACC_ZIN);
      if (next)     // hidden layer, x0 columns (normalised by the img-chain owners, published with X_Z)
        for (int c = 0; c < 2; ++c) {
          v.w(W_HID, DG + c * 2 * BK, hrow, 128, 2);
          v.a(A_ACT, CX + c * 2 * BK, row0);
          v.mma(0, 0, 0, 128, T_HID, true);
          v.mma(1, 1, 0, 128, T_HID, true);
          v.rw();
          v.ra();
        }
      // ---- actor chain (row-split, swapped): actor_1, actor_2 (head + sampling + dyn_in2 run in the epilogue warps)
      for (int layer = 0; layer < 2; ++layer) {
        v.bwait();
        for (int q = 0; q < 4; ++q) chain_group(v, layer == 0 ? W_A1 : W_A2, q);
        v.done(ACC_CH);
      }
      if (!next) continue;
      // ---- hidden layer, x1 | x2 columns (published by the actor-chain owners with X_X2)
      v.wait(X_X2);
      for (int c = 0; c < 4; ++c) {
        v.w(W_HID, DG + U + c * 2 * BK, hrow, 128, 2);
        v.a(A_ACT, CX + U + c * 2 * BK, row0);
        v.mma(0, 0, 0, 128, T_HID, true);
        v.mma(1, 1, 0, 128, T_HID, true);
        v.rw();
        v.ra();
      }
      v.done(ACC_HID);
      // ---- gate projection of this CTA's 128 units: reset | cand | update
      v.wait(X_H);
      for (int c = 0; c < 2; ++c) {
        v.a(A_ACT, CH + g * DG + c * 2 * BK, row0);
        for (int j = 0; j < 3; ++j) {
          v.w(W_GRU, c * 2 * BK, g * 3 * DG + j * DG + hf * 128, 128, 2);
          v.mma(0, 0, 0, 128, T_GRU + j * 128, c > 0);
          v.mma(1, 1, 0, 128, T_GRU + j * 128, true);
          v.rw();
        }
        v.ra();
      }
      v.done(ACC_GRU);
    }
  }
}

struct Bars {
  uint32_t a_full, a_empty, w_full, w_empty, acc, b_ready;
};

struct AProducer {
  const Params& P; Bars b; uint32_t sA; const unsigned int* flags; uint32_t me; uint32_t cnt = 0;
  unsigned int xcnt[NX] = {0, 0, 0, 0, 0, 0};
  __device__ __forceinline__ void wait(int x) {
    xcnt[x] += CL;
    flag_wait(flags + x * kFlagStride, xcnt[x]);
    fence_async_global();   // the TMA (async proxy) loads below read what the peers' generic-proxy stores wrote
  }
  __device__ __forceinline__ void a(int map, int col, int row) {
    if (cnt % NPA == me) {
      const uint32_t s = cnt % NA, ph = (cnt / NA) & 1u;
      wait_local(b.a_empty + s * 8, ph ^ 1u);
      tc::mbar_expect_tx(b.a_full + s * 8, kEntry);
      tc::tma_load_2d(sA + s * kEntry, &P.ma[map], col, row, b.a_full + s * 8);
      tc::tma_load_2d(sA + s * kEntry + kSlab, &P.ma[map], col + BK, row, b.a_full + s * 8);
    }
    ++cnt;
  }
  __device__ __forceinline__ void aw(int wmap, int k0, int row) {
    if (cnt % NPA == me) {
      const uint32_t s = cnt % NA, ph = (cnt / NA) & 1u;
      wait_local(b.a_empty + s * 8, ph ^ 1u);
      tc::mbar_expect_tx(b.a_full + s * 8, kEntry);
      tc::tma_load_2d(sA + s * kEntry, &P.mw[wmap], k0, row, b.a_full + s * 8);
      tc::tma_load_2d(sA + s * kEntry + kSlab, &P.mw[wmap], k0 + BK, row, b.a_full + s * 8);
    }
    ++cnt;
  }
  __device__ __forceinline__ void w(int, int, int, int, int) {}
  __device__ __forceinline__ void mma(int, int, int, int, uint32_t, bool) {}
  __device__ __forceinline__ void mma_sw(int, int, uint32_t, bool, bool) {}
  __device__ __forceinline__ void bwait() {}
  __device__ __forceinline__ void ra() {}
  __device__ __forceinline__ void rw() {}
  __device__ __forceinline__ void done(int) {}
};
struct WProducer {
  const Params& P; Bars b; uint32_t sW; uint32_t me; uint32_t cnt = 0;
  __device__ __forceinline__ void wait(int) {}
  __device__ __forceinline__ void a(int, int, int) {}
  __device__ __forceinline__ void w(int map, int k0, int row, int nrows, int nk) {
    if (cnt % NPW == me) {
      const uint32_t s = cnt % NW, ph = (cnt / NW) & 1u;
      wait_local(b.w_empty + s * 8, ph ^ 1u);
      const uint32_t box = (uint32_t)nrows * BK * 2;
      tc::mbar_expect_tx(b.w_full + s * 8, box * (uint32_t)nk);
      for (int j = 0; j < nk; ++j)
        tc::tma_load_2d(sW + s * kEntry + j * box, &P.mw[map], k0 + j * BK, row, b.w_full + s * 8);
    }
    ++cnt;
  }
  __device__ __forceinline__ void aw(int, int, int) {}
  __device__ __forceinline__ void mma(int, int, int, int, uint32_t, bool) {}
  __device__ __forceinline__ void mma_sw(int, int, uint32_t, bool, bool) {}
  __device__ __forceinline__ void bwait() {}
  __device__ __forceinline__ void ra() {}
  __device__ __forceinline__ void rw() {}
  __device__ __forceinline__ void done(int) {}
};
struct MmaIssuer {
  Bars b; uint32_t sA, sW, sB, tmem; long long* timing;
  uint32_t acnt = 0, wcnt = 0, bcnt = 0, cur_a = 0, cur_w = 0, wbox = 0;
  __device__ __forceinline__ void wait(int) {}
  __device__ __forceinline__ void a(int, int, int) {
    cur_a = acnt % NA;
    wait_local(b.a_full + cur_a * 8, (acnt / NA) & 1u);
    ++acnt;
    tc::tc_fence_after();
  }
  __device__ __forceinline__ void aw(int, int, int) { a(0, 0, 0); }
  __device__ __forceinline__ void w(int, int, int, int nrows, int) {
    cur_w = wcnt % NW;
    wait_local(b.w_full + cur_w * 8, (wcnt / NW) & 1u);
    ++wcnt;
    wbox = (uint32_t)nrows * BK * 2;
    tc::tc_fence_after();
  }
  __device__ __forceinline__ void bwait() {
    wait_local(b.b_ready, bcnt & 1u);
    ++bcnt;
    tc::tc_fence_after();
  }
  __device__ __forceinline__ void mma(int ai, int wi, int wrow, int n, uint32_t tcol, bool accum) {
    const uint64_t da = tc::make_desc_sw128(sA + cur_a * kEntry + ai * kSlab);
    const uint64_t db = tc::make_desc_sw128(sW + cur_w * kEntry + wi * wbox + wrow * 128);
    const uint32_t idesc = tc::make_idesc(BM, n);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k)
      tc::tc_mma_f16(tmem + tcol, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (accum || k > 0) ? 1u : 0u);
  }
  __device__ __forceinline__ void mma_sw(int wi, int kb, uint32_t tcol, bool accum, bool in_a) {
    const uint64_t da = tc::make_desc_sw128((in_a ? sA + cur_a * kEntry : sW + cur_w * kEntry) + wi * kSlab);   // weights: the M operand
    const uint64_t db = tc::make_desc_sw128(sB + kb * 2048);                    // 16 rows x 64 k of the chain operand
    const uint32_t idesc = tc::make_idesc(BM, 16);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k)
      tc::tc_mma_f16(tmem + tcol, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (accum || k > 0) ? 1u : 0u);
  }
  __device__ __forceinline__ void ra() { tc::tc_commit(b.a_empty + cur_a * 8); }
  __device__ __forceinline__ void rw() { tc::tc_commit(b.w_empty + cur_w * 8); }
  __device__ __forceinline__ void done(int bar) { tc::tc_commit(b.acc + bar * 8); }
};

// ------------------------------------------------------------------------------------------------ epilogue helpers
// 16 epilogue warps: warp w may read TMEM lanes [32 (w % 4), +32) ("quad"); q4 = w / 4 picks the column quarter.
// The epilogue code is a chain of short dependent phases with 4 warps per scheduler, i.e. latency bound: splitting
// every phase 16 ways (instead of 8) roughly halves it (profiles/r02_pimg_phase_stamps.txt).
struct Epi {
  const Params& P;
  Bars b;
  uint8_t* gbase;       // generic pointer to the aligned shared-memory base
  uint32_t tmem;        // this thread's TMEM lane base (lane quadrant << 16)
  int rank, row, q4, tid, warp, lane;   // row: 0..127 inside the group
  unsigned int* flags;  // this team's counters
  float* ssq;           // this team's row-statistics slots
  uint32_t apar = 0;
  unsigned int scnt = 0, xcnt_p7 = 0, xcnt_zin = 0;

  __device__ __forceinline__ float* consts() const { return reinterpret_cast<float*>(gbase + kOffConst); }
  __device__ __forceinline__ float* red() const { return reinterpret_cast<float*>(gbase + kOffRed); }
  __device__ __forceinline__ void wait_acc(int a) {
    wait_local(b.acc + a * 8, (apar >> a) & 1u);
    apar ^= 1u << a;
    tc::tc_fence_after();
  }
  // all-to-all exchange of one per-row partial sum between the 16 CTAs of the team (through L2)
  __device__ __forceinline__ float exchange(float p, bool sender) {
    if (sender) ssq[rank * BM + row] = p;
    epi_bar();
    scnt += CL;
    if (tid == 0) {
      flag_signal(flags + NX * kFlagStride);
      flag_wait(flags + NX * kFlagStride, scnt);
    }
    epi_bar();
    float q[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) q[r] = __ldcg(ssq + r * BM + row);   // written by other SMs: read through L2
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) t += q[r];                            // fixed order: deterministic
    return t;
  }
  // publish this CTA's part of exchange x to the team
  __device__ __forceinline__ void signal(int x) {
    fence_async_global();       // generic-proxy global stores -> visible to the peers' TMA (async proxy) loads
    tc::tc_fence_before();      // TMEM reads of this phase are complete before anybody may overwrite the columns
    epi_bar();
    if (tid == 0) flag_signal(flags + x * kFlagStride);
  }
  __device__ __forceinline__ void wait_flag(int x, unsigned int& cnt) {
    cnt += CL;
    if (tid == 0) flag_wait(flags + x * kFlagStride, cnt);
    epi_bar();
  }
  // the chain operand in shared memory is complete: hand it to the MMA thread
  __device__ __forceinline__ void publish_b() {
    fence_async_smem();
    tc::tc_fence_before();
    epi_bar();
    if (tid == 0) mbar_arrive(b.b_ready);
  }
};

// Column-split hand-over: this CTA's 16 columns of a 256-wide layer, bias added, PRE-norm, fp32 (thread = row).
__device__ __forceinline__ void store_raw(Epi& e, int grow, uint32_t tcol, int cb, int dst) {
  const float* cs = e.consts();
  float v[16];
  tmem_ld16(e.tmem + tcol, v);
  if (grow < e.P.N) {
    float4* o = reinterpret_cast<float4*>(e.P.raw + (size_t)grow * RAW_LD + dst + e.rank * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q] + cs[cb + 4 * q], v[4 * q + 1] + cs[cb + 4 * q + 1], v[4 * q + 2] + cs[cb + 4 * q + 2], v[4 * q + 3] + cs[cb + 4 * q + 3]);
  }
}

// Row-split prologue: one warp normalises one row of a pre-norm 256-wide hand-over (lane owns columns [8 lane, +8)):
// y = SiLU(RMSNorm(v) * gain) (rssm.py:17-24 etc.).
__device__ __forceinline__ void norm_row(const float* rawrow, const float* __restrict__ gain, int lane, bool ok, float* y) {
  float v[8];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gain + lane * 8)), gb = __ldg(reinterpret_cast<const float4*>(gain + lane * 8 + 4));
  if (ok) {
    const float4 a = ldcg4(rawrow + lane * 8), c = ldcg4(rawrow + lane * 8 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
  }
  const float g8[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) ss = fmaf(v[q], v[q], ss);
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss * (1.f / 256.f) + kRmsEps);
#pragma unroll
  for (int q = 0; q < 8; ++q) y[q] = silu_fast((v[q] * rs) * g8[q]);
}
// write columns [8 lane, +8) of chain-operand row w (K-major SWIZZLE_128B: k-block lane / 8, 16-byte chunk (lane % 8) ^ (w % 8))
__device__ __forceinline__ void put_b_row(uint8_t* sB, int w, int lane, const float* y) {
  *reinterpret_cast<uint4*>(sB + (lane >> 3) * 2048 + w * 128 + (((lane & 7) ^ (w & 7)) << 4)) = pack_bf8(y);
}

// Swapped chain layer epilogue: thread = (output feature f = TMEM lane of feature tile mt, row half rh): rows
// [4 rh, 4 rh + 4) sit in TMEM columns [tcol + 4 rh, +4).  v[r] = acc + bias[f]; per-row RMSNorm over the 256 features
// (the 8 warps that share rh); returns the activations in v.
constexpr int RH = ROWS / 2;
__device__ __forceinline__ void chain_norm(Epi& e, uint32_t tcol, int f, int rh, float bf, float gf, float* v) {
  {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(e.tmem + tcol + (uint32_t)(rh * RH)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < RH; ++i) v[i] = __uint_as_float(r[i]);
  }
  float sq[RH];
#pragma unroll
  for (int r = 0; r < RH; ++r) { v[r] += bf; sq[r] = v[r] * v[r]; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < RH; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], off);
  float* rd = e.red();
  if (e.lane == 0)
#pragma unroll
    for (int r = 0; r < RH; ++r) rd[e.warp * RH + r] = sq[r];
  epi_bar();
#pragma unroll
  for (int r = 0; r < RH; ++r) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += rd[((w >> 2) * 8 + rh * 4 + (w & 3)) * RH + r];   // warps (quad, q4 = 2 mt + rh): fixed order
    const float rs = 1.f / sqrtf(t * (1.f / 256.f) + kRmsEps);
    v[r] = silu_fast((v[r] * rs) * gf);
  }
}
// activations of feature f for rows [4 rh, +4) -> chain operand (bf16, element (row r, k = f))
__device__ __forceinline__ void put_b_col(uint8_t* sB, int f, int rh, const float* v) {
  uint8_t* base = sB + (f >> 6) * 2048 + (f & 7) * 2;
  const int ch = (f & 63) >> 3;
#pragma unroll
  for (int i = 0; i < RH; ++i) {
    const int r = rh * RH + i;
    *reinterpret_cast<__nv_bfloat16*>(base + r * 128 + ((ch ^ r) << 4)) = __float2bfloat16(v[i]);
  }
}

// order-preserving float -> uint key (larger float <=> larger key)
__device__ __forceinline__ uint32_t fkey(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Actor head + action sample + dyn_in2 -> x2 for one row (one warp; lane owns columns [8 lane, +8)).
// networks.py:374-377, distributions.py:217-231, rssm.py:44,48.  W: shared (kTail) or global weight pointers.
__device__ __forceinline__ void actor_tail(const Params& P, const float* wl, int ldl, const float* w2, int ld2, const float* b2, const float* g2,
                                           const float* a2row, float nz, int lane, bool next, float& act_out_v, float* x2) {
  const int A = P.A;
  float x[8];
  {
    const float4 xa = *reinterpret_cast<const float4*>(a2row + lane * 8), xb = *reinterpret_cast<const float4*>(a2row + lane * 8 + 4);
    x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w; x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
  }
  float mine = 0.f, mine2 = 0.f;   // lane j keeps output j (and output j + A for the bounded-normal std)
#pragma unroll 1
  for (int j0 = 0; j0 < P.act_out; j0 += 16) {   // 16 head outputs at a time: independent FMAs, one interleaved butterfly
    float acc[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = j0 + jj < P.act_out ? j0 + jj : P.act_out - 1;
      const float4 wa = *reinterpret_cast<const float4*>(wl + (size_t)j * ldl + lane * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wl + (size_t)j * ldl + lane * 8 + 4);
      float t = x[0] * wa.x;
      t = fmaf(x[1], wa.y, t); t = fmaf(x[2], wa.z, t); t = fmaf(x[3], wa.w, t);
      t = fmaf(x[4], wb.x, t); t = fmaf(x[5], wb.y, t); t = fmaf(x[6], wb.z, t); t = fmaf(x[7], wb.w, t);
      acc[jj] = t;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) acc[jj] += __shfl_xor_sync(0xffffffffu, acc[jj], off);
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = j0 + jj;
      if (j == lane) mine = acc[jj];
      if (j == lane + A) mine2 = acc[jj];
    }
  }
  if (lane < P.act_out) mine += __ldg(P.b_last + lane);
  if (lane + A < P.act_out) mine2 += __ldg(P.b_last + lane + A);
  float act = 0.f;
  if (P.act_kind == 0) {
    if (lane < A) {
      const float sd_ = (P.max_std - P.min_std) * sigmoidf_(mine2 + 2.f) + P.min_std;
      act = tanhf(mine) + sd_ * nz;
    }
  } else {
    const bool valid = lane < A;
    const int best = sample_group<32>(mine, nz, valid, lane, A, P.act_unimix, nullptr);
    act = (valid && lane == best) ? 1.f : 0.f;
  }
  act_out_v = act;
  if (!next) return;
  const float ab = act / fmaxf(fabsf(act), 1.f);
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) x2[q] = b2[lane * 8 + q];
  for (int a = 0; a < A; ++a) {
    const float av = __shfl_sync(0xffffffffu, ab, a);
    const float4 wa = *reinterpret_cast<const float4*>(w2 + (size_t)a * ld2 + lane * 8);
    const float4 wb = *reinterpret_cast<const float4*>(w2 + (size_t)a * ld2 + lane * 8 + 4);
    x2[0] = fmaf(av, wa.x, x2[0]); x2[1] = fmaf(av, wa.y, x2[1]); x2[2] = fmaf(av, wa.z, x2[2]); x2[3] = fmaf(av, wa.w, x2[3]);
    x2[4] = fmaf(av, wb.x, x2[4]); x2[5] = fmaf(av, wb.y, x2[5]); x2[6] = fmaf(av, wb.z, x2[6]); x2[7] = fmaf(av, wb.w, x2[7]);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) ss = fmaf(x2[q], x2[q], ss);
  ss = warp_sum(ss);
  const float rs = 1.f / sqrtf(ss * (1.f / 256.f) + kRmsEps);
#pragma unroll
  for (int q = 0; q < 8; ++q) x2[q] = silu_fast((x2[q] * rs) * g2[lane * 8 + q]);
}

// constant-table layout (floats): biases of this CTA's column-split slices
enum { C_X0 = 0, C_O0 = 16, C_A0 = 32, C_X1 = 48, C_HB = 64, C_HG = 192, C_GB = 320 };

#define SD_PI_STAMP(k) do { if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) P.timing[(k)] = clock64(); } while (0)

__global__ void __launch_bounds__(THREADS, 1) imagine_persistent_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)(blockIdx.x % CL), team = (int)(blockIdx.x / CL);
  Bars b;
  b.a_full = base + kOffBar;
  b.a_empty = b.a_full + NA * 8;
  b.w_full = b.a_empty + NA * 8;
  b.w_empty = b.w_full + NW * 8;
  b.acc = b.w_empty + NW * 8;
  b.b_ready = b.acc + NACC * 8;
  const uint32_t tmem_slot = b.b_ready + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + kOffBar + kNumBar * 8);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { tc::mbar_init(b.a_full + i * 8, 1); tc::mbar_init(b.a_empty + i * 8, 1); }
    for (int i = 0; i < NW; ++i) { tc::mbar_init(b.w_full + i * 8, 1); tc::mbar_init(b.w_empty + i * 8, 1); }
    for (int i = 0; i < NACC; ++i) tc::mbar_init(b.acc + i * 8, 1);
    tc::mbar_init(b.b_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp < NEPI) {
    // per-CTA slices of the column-split biases / RMS scales and the actor-tail weights: loaded once, kept for all iterations
    float* cs = reinterpret_cast<float*>(gbase + kOffConst);
    const int t = threadIdx.x;   // 0..511
    if (t < 16) {
      const int c = rank * 16 + t;
      cs[C_X0 + t] = P.b_in0[c]; cs[C_O0 + t] = P.b_i0[c]; cs[C_A0 + t] = P.b_a0[c]; cs[C_X1 + t] = P.b_in1[c];
    }
    if (t < 128) {
      cs[C_HB + t] = P.b_hid[rank * 128 + t];
      cs[C_HG + t] = P.g_hid[rank * 128 + t];
      const int gg = rank >> 1, hf = rank & 1;
#pragma unroll
      for (int j = 0; j < 3; ++j) cs[C_GB + j * 128 + t] = P.b_gru[gg * 3 * DG + j * DG + hf * 128 + t];
    }
    if (P.tail_in_smem) {
      float* tl = reinterpret_cast<float*>(gbase + kOffTail);
      for (int i = t; i < P.act_out * 256; i += NEPI * 32) tl[i] = P.w_last[(size_t)(i >> 8) * P.ldk_last + (i & 255)];
      float* t2 = tl + P.act_out * 256;
      for (int i = t; i < P.A * 256; i += NEPI * 32) t2[i] = P.w_in2[(size_t)(i >> 8) * P.ldw_in2 + (i & 255)];
      float* t3 = t2 + P.A * 256;
      for (int i = t; i < 256; i += NEPI * 32) { t3[i] = P.b_in2[i]; t3[256 + i] = P.g_in2[i]; }
    }
    // rows 8..15 of the chain operand are never written by the prologues: clear them once (their outputs are ignored)
    for (int i = t; i < 4 * 2048 / 16; i += NEPI * 32) reinterpret_cast<uint4*>(gbase + kOffB)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp >= W_WARP && warp < A_WARP) {
    if (lane == 0) {
      WProducer v{P, b, base + kOffW, (uint32_t)(warp - W_WARP)};
      walk(P, rank, v);
    }
    __syncwarp();
  } else if (warp >= A_WARP && warp < M_WARP) {
    if (lane == 0) {
      AProducer v{P, b, base + kOffA, P.flags + (size_t)team * flags_per_team(), (uint32_t)(warp - A_WARP)};
      walk(P, rank, v);
    }
    __syncwarp();
  } else if (warp == M_WARP) {
    if (lane == 0) {
      MmaIssuer v{b, base + kOffA, base + kOffW, base + kOffB, tmem_base, blockIdx.x == 0 ? P.timing : nullptr};
      walk(P, rank, v);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int quad = warp & 3, q4 = warp >> 2;
    Epi e{P, b, gbase, tmem_base + ((uint32_t)(quad * 32) << 16), rank, quad * 32 + lane, q4, (int)threadIdx.x, warp, lane,
          P.flags + (size_t)team * flags_per_team(), P.ssq + (size_t)team * ssq_per_team()};
    const float* cs = e.consts();
    uint8_t* sB = gbase + kOffB;
    const int A = P.A, H = P.H;
    const size_t ldf = (size_t)H * F;
    // swapped 256-wide chain layers: feature tile mt = q4 / 2, row half rh = q4 % 2
    const int mt2 = q4 >> 1, rh = q4 & 1;
    const int f2 = mt2 * 128 + quad * 32 + lane;
    const float b_i1 = __ldg(P.b_i1 + f2), g_i1 = __ldg(P.g_i1 + f2), b_a1 = __ldg(P.b_a1 + f2), g_a1 = __ldg(P.g_a1 + f2),
                b_a2 = __ldg(P.b_a2 + f2), g_a2 = __ldg(P.g_a2 + f2);
    // logits: feature tile q4, all 8 rows
    const int fl = q4 * 128 + quad * 32 + lane;
    const float b_lg = __ldg(P.b_lg + fl);
    const int kcl = lane & (KC - 1);
    // tail weights
    const float* s_tl = reinterpret_cast<const float*>(gbase + kOffTail);
    for (int grp = (int)(blockIdx.x / CL); grp < P.ngroups; grp += (int)(gridDim.x / CL)) {
      const int grow = grp * BM + e.row;                   // global row of this thread in the column-split phases
      const bool rowok = grow < P.N;
      const int crow0 = grp * BM + rank * ROWS;            // first of this CTA's 8 chain rows
      const int crow = crow0 + (warp & 7);                 // global row of this WARP in the row-split prologues / the tail
      const bool crowok = crow < P.N;
      for (int i = 0; i < H; ++i) {
        const bool next = i + 1 < H;
        // ---- P7 hand-over: pre-norm x0 / o0 columns (the actor_0 partial stays in TMEM)
        e.wait_acc(ACC_P7);
        SD_PI_STAMP(32 * i + 0);
        if (q4 == 0) store_raw(e, grow, T_X0, C_X0, RX0);
        else if (q4 == 1) store_raw(e, grow, T_O0, C_O0, RO0);
        e.signal(X_P7);
        SD_PI_STAMP(32 * i + 1);
        // gumbel noise of this thread's logit feature for its CTA's 8 rows: independent of the model, done while waiting
        float gn[ROWS];
        if (i > 0) {
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            const int gr = crow0 + r;
            gn[r] = gr < P.N ? __ldg(P.u + ((size_t)gr * H + (i - 1)) * SK + fl) : 0.5f;
          }
#pragma unroll
          for (int r = 0; r < ROWS; ++r) gn[r] = -__logf(-logf(gn[r]));
        }
        // ---- img chain: x0 = SiLU(RMSNorm(.)) for the hidden layer (warps 0-7); o0 -> chain operand (warps 8-15)
        e.wait_flag(X_P7, e.xcnt_p7);
        SD_PI_STAMP(32 * i + 13);
        {
          float y[8];
          if (warp < 8) {
            norm_row(P.raw + (size_t)crow * RAW_LD + RX0, P.g_in0, lane, crowok, y);
            if (crowok) *(reinterpret_cast<uint4*>(P.act + (size_t)crow * ACT_LD + CX) + lane) = pack_bf8(y);
          } else if (i > 0) {
            norm_row(P.raw + (size_t)crow * RAW_LD + RO0, P.g_i0, lane, crowok, y);
            put_b_row(sB, warp & 7, lane, y);
          }
        }
        if (i > 0) {
          e.publish_b();
          SD_PI_STAMP(32 * i + 14);
          {
            e.wait_acc(ACC_CH);
            SD_PI_STAMP(32 * i + 15);
            float v[RH];
            chain_norm(e, T_CH + (uint32_t)(mt2 * 16), f2, rh, b_i1, g_i1, v);
            put_b_col(sB, f2, rh, v);
            e.publish_b();
          }
          SD_PI_STAMP(32 * i + 2);
          e.wait_acc(ACC_CH);
          SD_PI_STAMP(32 * i + 12);
          // logits: thread = class (16 consecutive lanes = one category), unimix + Gumbel arg-max (distributions.py:16-36);
          // the 8 rows are independent: every reduction step is issued for all of them before the next one
          {
            float lg[ROWS], ex[ROWS], red_[ROWS];
            tmem_ld8(e.tmem + T_CH + (uint32_t)(q4 * 16), lg);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) { lg[r] += b_lg; red_[r] = lg[r]; }
#pragma unroll
            for (int o = KC / 2; o > 0; o >>= 1)
#pragma unroll
              for (int r = 0; r < ROWS; ++r) red_[r] = fmaxf(red_[r], __shfl_xor_sync(0xffffffffu, red_[r], o));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) { ex[r] = __expf(lg[r] - red_[r]); red_[r] = ex[r]; }
#pragma unroll
            for (int o = KC / 2; o > 0; o >>= 1)
#pragma unroll
              for (int r = 0; r < ROWS; ++r) red_[r] += __shfl_xor_sync(0xffffffffu, red_[r], o);
            // arg max_k [ log(p_k (1-eps) + eps/K) - log(-log u_k) ] == first arg max of softmax(log p~ - lse + g): one
            // integer max over keys (order-preserving bits of the score, low 4 bits = 15 - class: ties -> smallest class;
            // scores closer than 2^-19 relative count as ties, far inside the near-tie band of the bf16 path)
            const uint32_t gmask = 0xffffu << (lane & 16);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
              const float z = __logf(fmaf(ex[r], __fdividef(1.f - P.unimix, red_[r]), P.unimix * (1.f / KC))) + gn[r];
              const uint32_t key = (fkey(z) & ~15u) | (uint32_t)(15 - kcl);
              const uint32_t best = __reduce_max_sync(gmask, key);
              const int gr = crow0 + r;
              if (gr < P.N) {
                const float one = (best == key) ? 1.f : 0.f;
                P.feats[(size_t)gr * ldf + (size_t)i * F + fl] = one;
                P.big_bf[(size_t)gr * ldf + (size_t)i * F + fl] = __float2bfloat16(one);
              }
            }
          }
        }
        e.signal(X_Z);
        SD_PI_STAMP(32 * i + 3);
        // ---- ZIN hand-over: pre-norm a0 (deter + stoch parts) / x1 columns
        e.wait_acc(ACC_ZIN);
        if (q4 == 0) store_raw(e, grow, T_A0, C_A0, RA0);
        else if (q4 == 1) store_raw(e, grow, T_X1, C_X1, RX1);
        e.signal(X_ZIN);
        SD_PI_STAMP(32 * i + 4);
        // ---- actor chain: x1 for the hidden layer (warps 0-7); a0 -> chain operand (warps 8-15) -> actor_1 -> actor_2
        const float nz = (warp < 8 && crowok && lane < A) ? __ldg(P.act_noise + ((size_t)crow * H + i) * A + lane) : 0.5f;
        e.wait_flag(X_ZIN, e.xcnt_zin);
        {
          float y[8];
          if (warp < 8) {
            if (next) {
              norm_row(P.raw + (size_t)crow * RAW_LD + RX1, P.g_in1, lane, crowok, y);
              if (crowok) *(reinterpret_cast<uint4*>(P.act + (size_t)crow * ACT_LD + CX + U) + lane) = pack_bf8(y);
            }
          } else {
            norm_row(P.raw + (size_t)crow * RAW_LD + RA0, P.g_a0, lane, crowok, y);
            put_b_row(sB, warp & 7, lane, y);
          }
        }
        e.publish_b();
        {
          e.wait_acc(ACC_CH);
          float v[RH];
          chain_norm(e, T_CH + (uint32_t)(mt2 * 16), f2, rh, b_a1, g_a1, v);
          put_b_col(sB, f2, rh, v);
          e.publish_b();
        }
        SD_PI_STAMP(32 * i + 5);
        {
          e.wait_acc(ACC_CH);
          float v[RH];
          chain_norm(e, T_CH + (uint32_t)(mt2 * 16), f2, rh, b_a2, g_a2, v);
          // a2 stays fp32 for the head: [8 rows][256] over the (now idle) chain operand
          float* a2f = reinterpret_cast<float*>(sB);
#pragma unroll
          for (int r = 0; r < RH; ++r) a2f[(rh * RH + r) * 256 + f2] = v[r];
          tc::tc_fence_before();
          epi_bar();
        }
        SD_PI_STAMP(32 * i + 6);
        // ---- actor head + action + dyn_in2 -> x2: warp w (< 8) owns chain row w
        if (warp < 8) {
          float act, x2[8];
          const float* a2row = reinterpret_cast<const float*>(sB) + warp * 256;
          if (P.tail_in_smem)
            actor_tail(P, s_tl, 256, s_tl + P.act_out * 256, 256, s_tl + (P.act_out + A) * 256, s_tl + (P.act_out + A) * 256 + 256, a2row, nz,
                       lane, next, act, x2);
          else
            actor_tail(P, P.w_last, P.ldk_last, P.w_in2, P.ldw_in2, P.b_in2, P.g_in2, a2row, nz, lane, next, act, x2);
          if (crowok && lane < A) P.actions[((size_t)crow * H + i) * A + lane] = act;
          if (next && crowok) *(reinterpret_cast<uint4*>(P.act + (size_t)crow * ACT_LD + CX + 2 * U) + lane) = pack_bf8(x2);
        }
        SD_PI_STAMP(32 * i + 7);
        if (!next) { epi_bar(); continue; }   // (the fp32 a2 tile is dead before the next iteration rewrites the operand)
        e.signal(X_X2);
        // ---- block-GRU hidden layer: RMSNorm over all 2048 columns (16 CTAs x 128) -> SiLU -> h (bf16); thread = (row, 32 columns)
        {
          e.wait_acc(ACC_HID);
          SD_PI_STAMP(32 * i + 8);
          float v[32];
          tc::tmem_ld32(e.tmem + T_HID + (uint32_t)(q4 * 32), v);
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] += cs[C_HB + q4 * 32 + j]; ss = fmaf(v[j], v[j], ss); }
          float* hss = e.red() + NEPI * RH;
          hss[q4 * BM + e.row] = ss;
          epi_bar();
          const float mine = (hss[e.row] + hss[BM + e.row]) + (hss[2 * BM + e.row] + hss[3 * BM + e.row]);
          const float tot = e.exchange(mine, q4 == 0);
          const float rs = 1.f / sqrtf(tot * (1.f / (float)D) + kRmsEps);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_fast((v[j] * rs) * cs[C_HG + q4 * 32 + j]);
          if (rowok) {
            uint4* o = reinterpret_cast<uint4*>(P.act + (size_t)grow * ACT_LD + CH + rank * 128 + q4 * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = pack_bf8(v + q * 8);
          }
          e.signal(X_H);
          SD_PI_STAMP(32 * i + 9);
        }
        // ---- gate projection -> GRU gates (rssm.py:63-75).  The 128 x 128 fp32 tile of the old / new deter is staged in
        // the (idle: the next activation load waits for X_D) activation ring so that global memory sees whole 512-byte
        // rows: chunk q (16 B) of row r lives at r * 512 + ((q ^ (r & 31)) << 4) (conflict-free for thread = row).
        {
          const size_t col = (size_t)SK + rank * 128;
          float4 pre[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // warp w loads rows w, w + 16, ...: one 512-byte row per instruction
            const int r = warp + NEPI * k, gr = grp * BM + r;
            pre[k] = gr < P.N ? *reinterpret_cast<const float4*>(P.feats + (size_t)gr * ldf + (size_t)i * F + col + lane * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          e.wait_acc(ACC_GRU);
          SD_PI_STAMP(32 * i + 10);
          uint8_t* st = gbase + kOffA;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = warp + NEPI * k;
            *reinterpret_cast<float4*>(st + r * 512 + ((lane ^ (r & 31)) << 4)) = pre[k];
          }
          epi_bar();
          const int r = e.row;
#pragma unroll 1
          for (int c16 = 0; c16 < 2; ++c16) {
            float qr[16], qc[16], qu[16];
            const uint32_t tb = e.tmem + T_GRU + (uint32_t)(q4 * 32 + c16 * 16);
            tmem_ld16(tb, qr);
            tmem_ld16(tb + 128u, qc);
            tmem_ld16(tb + 256u, qu);
            const float* gb = cs + C_GB + q4 * 32 + c16 * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4* p = reinterpret_cast<float4*>(st + r * 512 + (((q4 * 8 + c16 * 4 + q) ^ (r & 31)) << 4));
              float4 dv = *p;
              float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int u_ = q * 4 + j;
                const float reset = sigmoid_fast(qr[u_] + gb[u_]);
                const float cand = tanh_fast(reset * (qc[u_] + gb[128 + u_]));
                const float upd = sigmoid_fast((qu[u_] + gb[256 + u_]) - 1.f);
                d4[j] = fmaf(upd, cand - d4[j], d4[j]);
              }
              *p = make_float4(d4[0], d4[1], d4[2], d4[3]);
            }
          }
          tc::tc_fence_before();
          epi_bar();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = warp + NEPI * k, gr = grp * BM + rr;
            const float4 dv = *reinterpret_cast<const float4*>(st + rr * 512 + ((lane ^ (rr & 31)) << 4));
            if (gr < P.N) {
              *reinterpret_cast<float4*>(P.feats + (size_t)gr * ldf + (size_t)(i + 1) * F + col + lane * 4) = dv;
              *reinterpret_cast<uint2*>(P.big_bf + (size_t)gr * ldf + (size_t)(i + 1) * F + col + lane * 4) =
                  make_uint2(pack_bf2(dv.x, dv.y), pack_bf2(dv.z, dv.w));
            }
          }
          e.signal(X_D);
          SD_PI_STAMP(32 * i + 11);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == M_WARP) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// Re-pack the weights that several layers share an activation slab with (run once per weight refresh):
//   wp7 [768][2048]: row 48 c + 16 j + i = row 16 c + i of { dyn_in0, img_net_0, actor_0[:, 512:] }[j]
//   wz  [512][512] : row 32 c + 16 j + i = row 16 c + i of { actor_0[:, :512], dyn_in1 }[j]
__global__ void __launch_bounds__(256) pimg_pack_kernel(const __nv_bfloat16* __restrict__ in0, const __nv_bfloat16* __restrict__ img0,
                                                        const __nv_bfloat16* __restrict__ a0, const __nv_bfloat16* __restrict__ in1,
                                                        __nv_bfloat16* __restrict__ wp7, __nv_bfloat16* __restrict__ wz) {
  pdl_prologue();
  const int n7 = 768 * (D / 8), nz = 512 * (SK / 8);   // 16-byte chunks
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n7 + nz; t += gridDim.x * blockDim.x) {
    if (t < n7) {
      const int r = t / (D / 8), ch = t - r * (D / 8);
      const int c = r / 48, j = (r % 48) / 16, i = r % 16;
      const __nv_bfloat16* src = j == 0 ? in0 + (size_t)(16 * c + i) * D : j == 1 ? img0 + (size_t)(16 * c + i) * D
                                                                               : a0 + (size_t)(16 * c + i) * F + SK;
      reinterpret_cast<uint4*>(wp7 + (size_t)r * D)[ch] = reinterpret_cast<const uint4*>(src)[ch];
    } else {
      const int q = t - n7;
      const int r = q / (SK / 8), ch = q - r * (SK / 8);
      const int c = r / 32, j = (r % 32) / 16, i = r % 16;
      const __nv_bfloat16* src = j == 0 ? a0 + (size_t)(16 * c + i) * F : in1 + (size_t)(16 * c + i) * SK;
      reinterpret_cast<uint4*>(wz + (size_t)r * SK)[ch] = reinterpret_cast<const uint4*>(src)[ch];
    }
  }
}

}  // namespace pimg
}  // namespace sd
