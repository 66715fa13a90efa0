// sd_pimg.cuh -- persistent, team-resident imagination scan on tcgen05 (sm_100a).
//
// Dreamer._imagine (dreamer.py:673-692): H iterations of  feat = [stoch | deter] -> action = actor(feat).rsample()
// -> stoch, deter = img_step(stoch, deter, action)  (rssm.py:36-75, 180-195; networks.py:339-377).
//
// ONE launch runs all H iterations.  A TEAM of 16 CTAs owns a group of 128 rows for the whole rollout
// (rows-stationary; 1024 rows = 8 teams = 128 SMs, one CTA per SM); the CTAs of a team split the OUTPUT COLUMNS of every layer:
//     256-wide layers (dyn_in0/1, img_net_0/1, actor 0..2): 16 columns per CTA      logits: 2 categories per CTA
//     block-GRU hidden layer: 128 of the 2048 columns (CTA c -> block c/2, half c%2)  gate projection: the same 128 units
// Per CTA: warp 8 streams weight tiles (TMA, 5 x 16 KB ring) following a static schedule -- weights never depend on
// data, so the ring runs ahead across layer boundaries; warp 9 streams activation slabs (128 rows x 64 k, 6 x 16 KB
// ring) and is the only role that waits for other CTAs; warp 10 issues tcgen05.mma (M = 128, N = 16..128) into fixed
// TMEM column ranges; warps 0-7 are the epilogue: tcgen05.ld -> bias -> RMSNorm (row sums of squares exchanged
// between the 16 CTAs of the team) -> SiLU / GRU gates / unimix-Gumbel arg-max / actor tail -> bf16 activation slices
// written to an L2-resident exchange buffer -> release-increment of the team's counter for that exchange.
// Teams are NOT thread-block clusters: a B200 can hold only 7 co-resident 16-CTA clusters with this shared-memory
// footprint (profiles/r02_cluster_probe.txt), one short of the 8 teams of the base shape, so team-level
// synchronisation goes through L2 (red.release.gpu / ld.relaxed + fence.acq_rel.gpu on per-team counters, one 128-byte
// line each).  No kernel boundary, no grid-wide barrier, no fp32 activation round trip; `feats` (fp32 output) and its
// bf16 copy (the operand of the next iteration and of the heads) are each written once.
//
// The three single-thread roles walk the SAME schedule (`walk`), so the rings cannot get out of step.
// Every wait is bounded (trap instead of hang).
#pragma once
#include "sd_kernels.cuh"
#include "sd_tc.cuh"

namespace sd {
namespace pimg {

using tc::smem_u32;

constexpr int CL = 16;          // CTAs per team
constexpr int BM = 128, BK = 64;
constexpr int NEPI = 8;         // epilogue warps (two per TMEM lane quadrant)
// Each producer role is NPW / NPA threads in different warps that take the boxes of the schedule round-robin (a lone
// thread issues at most one 16 KB box per ~560 cycles).  Measured (profiles/r02_pimg_phase_stamps.txt): inside this kernel
// the TMA unit delivers ~25 B/clk per SM in total however many threads issue (2+2 and 3+4 give the same slab cadence).
constexpr int NPW = 2, NPA = 2;
constexpr int W_WARP = NEPI, A_WARP = W_WARP + NPW, M_WARP = A_WARP + NPA;
constexpr int THREADS = 32 * (M_WARP + 1);
constexpr int NA = 6, NW = 5;   // activation-slab ring / weight ring depth
constexpr int kSlab = BM * BK * 2;   // 16 KB: one A slab, one W ring slot

// architecture the kernel is specialised for (configs/base.yaml:117-127,252-276); the host checks it
constexpr int D = 2048, U = 256, SK = 512, KC = 16, G = 8, DG = 256, F = SK + D;
// exchange buffer columns ([rows][ACT_LD] bf16)
constexpr int CA0 = 0, CA1 = 256, CA2 = 512, CO0 = 768, CO1 = 1024, CX = 1280, CH = 2048, ACT_LD = 4096;
// TMEM columns
constexpr uint32_t T_HID = 0, T_GRU = 128, T_X0 = 128, T_O0 = 144, T_A0 = 160, T_X1 = 176, T_I1 = 192, T_LG = 208,
                   T_A1 = 240, T_A2 = 256;

enum { X_D = 0, X_P7, X_O1, X_Z, X_ZIN, X_A1, X_A2, X_X2, X_H, NX };
enum { ACC_P7 = 0, ACC_I1, ACC_LG, ACC_ZIN, ACC_A1, ACC_A2, ACC_HID, ACC_GRU, NACC };
enum { W_P7 = 0, W_Z, W_A1, W_A2, W_I1, W_LG, W_HID, W_GRU, NWMAP };
enum { A_BIG = 0, A_ACT = 1 };

constexpr int kTailFloats = 9728;   // W_last (act_out x 256) + W_in2 (A x 256) + b2 + g2 when act_out + A <= 36
constexpr int kOffA = 0;
constexpr int kOffW = kOffA + NA * kSlab;
constexpr int kOffTail = kOffW + NW * kSlab;
constexpr int kOffConst = kOffTail + kTailFloats * 4;  // per-CTA bias / gain slices
constexpr int kConstFloats = 14 * 16 + 32 + 128 + 128 + 384;
constexpr int kOffHss = kOffConst + kConstFloats * 4;  // [2][128]
constexpr int kOffBar = kOffHss + 2 * BM * 4;
constexpr int kNumBar = 2 * NA + 2 * NW + NACC;
constexpr int kSmemBytes = kOffBar + kNumBar * 8 + 16 + 1024;   // + tmem slot + alignment slack

struct Params {
  CUtensorMap ma[2];        // A_BIG: big_bf [N][H*F]; A_ACT: exchange buffer [N][ACT_LD]   (box 64 x 128)
  CUtensorMap mw[NWMAP];    // weight maps (box 64 x {48,16,16,16,16,32,128,128})
  int N, H, ngroups;
  float* feats;             // [N][H][F] fp32
  float* actions;           // [N][H][A] fp32
  __nv_bfloat16* big_bf;    // [N][H][F]
  __nv_bfloat16* act;       // [N][ACT_LD]
  const float* u;           // [N][H][SK] uniforms of the prior samples
  const float* act_noise;   // [N][H][A]
  const float *b_in0, *g_in0, *b_in1, *g_in1, *b_in2, *g_in2, *b_hid, *g_hid, *b_gru, *b_i0, *g_i0, *b_i1, *g_i1, *b_lg;
  const float *b_a0, *g_a0, *b_a1, *g_a1, *b_a2, *g_a2, *b_last;
  const float* w_last; int ldk_last;    // [act_out][ldk] fp32
  const float* w_in2; int ldw_in2;      // [A][ldw] fp32 (n contiguous)
  int A, act_out, act_kind, tail_in_smem;
  unsigned int* flags;      // [ngroups_max teams][NX + 1][32] monotonic counters (zeroed before the launch)
  float2* ssq;              // [teams][NSLOT][16 ranks][128 rows] row partial sums of squares
  float min_std, max_std, act_unimix, unimix;
  long long* timing;        // diagnostic: clock64 stamps of CTA 0 (null in production)
};

// ------------------------------------------------------------------------------------------------ primitives
constexpr int NSLOT = 6;                      // one row-statistics slot per exchange of an iteration
constexpr int kFlagStride = 32;               // one 128-byte line per counter
__host__ __device__ constexpr size_t flags_per_team() { return (size_t)(NX + 1) * kFlagStride; }
__host__ __device__ constexpr size_t ssq_per_team() { return (size_t)NSLOT * CL * BM; }

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

// ------------------------------------------------------------------------------------------------ the schedule
// One walk = everything one CTA's producer / MMA roles do, in issue order.  Visitor callbacks:
//   v.wait(x)                                  activation producer: exchange x of this iteration must have completed
//   v.a(map, col, row)                         next activation slab (128 rows x 64 k at column `col`)
//   v.w(map, k, row, n, tcol, acc, rel, bar)   weight box (n rows at `row`, k-block `k`) multiplied with the CURRENT slab
//                                              into TMEM columns [tcol, tcol+n); rel: last use of the slab;
//                                              bar >= 0: accumulator `bar` is complete after this op
template <class V>
__device__ __forceinline__ void walk(const Params& P, int rank, V& v) {
  const int g = rank >> 1, hf = rank & 1;
  for (int grp = (int)(blockIdx.x / CL); grp < P.ngroups; grp += (int)(gridDim.x / CL)) {
    const int row0 = grp * BM;
    for (int i = 0; i < P.H; ++i) {
      const bool next = i + 1 < P.H;
      // ---- layers reading deter_i (K = 2048): dyn_in0 | img_net_0 | actor_0[:, 512:] (+ this CTA's deter block -> hidden layer)
      if (i > 0) v.wait(X_D);
      for (int kb = 0; kb < D / BK; ++kb) {
        v.a(A_BIG, i * F + SK + kb * BK, row0);
        const bool hid_here = next && (kb >> 2) == g;
        v.w(W_P7, kb * BK, rank * 48, 48, T_X0, kb > 0, !hid_here, (kb == D / BK - 1 && !hid_here) ? ACC_P7 : -1);
        if (hid_here)
          v.w(W_HID, (kb & 3) * BK, g * DG + hf * 128, 128, T_HID, (kb & 3) > 0, true, kb == D / BK - 1 ? ACC_P7 : -1);
      }
      v.wait(X_P7);
      if (i > 0)   // img_net_1
        for (int kb = 0; kb < 4; ++kb) {
          v.a(A_ACT, CO0 + kb * BK, row0);
          v.w(W_I1, kb * BK, rank * 16, 16, T_I1, kb > 0, true, kb == 3 ? ACC_I1 : -1);
        }
      if (next)    // hidden layer, x0 columns
        for (int kb = 0; kb < 4; ++kb) {
          v.a(A_ACT, CX + kb * BK, row0);
          v.w(W_HID, DG + kb * BK, g * DG + hf * 128, 128, T_HID, true, true, -1);
        }
      if (i > 0) {  // logits
        v.wait(X_O1);
        for (int kb = 0; kb < 4; ++kb) {
          v.a(A_ACT, CO1 + kb * BK, row0);
          v.w(W_LG, kb * BK, rank * 32, 32, T_LG, kb > 0, true, kb == 3 ? ACC_LG : -1);
        }
        v.wait(X_Z);
      }
      // ---- layers reading stoch_i (K = 512): actor_0[:, :512] (on top of the deter part) | dyn_in1
      for (int kb = 0; kb < SK / BK; ++kb) {
        v.a(A_BIG, i * F + kb * BK, row0);
        v.w(W_Z, kb * BK, rank * 32, 16, T_A0, true, false, -1);
        v.w(W_Z, kb * BK, rank * 32 + 16, 16, T_X1, kb > 0, true, kb == SK / BK - 1 ? ACC_ZIN : -1);
      }
      v.wait(X_ZIN);
      for (int kb = 0; kb < 4; ++kb) {   // actor_1
        v.a(A_ACT, CA0 + kb * BK, row0);
        v.w(W_A1, kb * BK, rank * 16, 16, T_A1, kb > 0, true, kb == 3 ? ACC_A1 : -1);
      }
      if (next)
        for (int kb = 0; kb < 4; ++kb) {  // hidden layer, x1 columns
          v.a(A_ACT, CX + U + kb * BK, row0);
          v.w(W_HID, DG + U + kb * BK, g * DG + hf * 128, 128, T_HID, true, true, -1);
        }
      v.wait(X_A1);
      for (int kb = 0; kb < 4; ++kb) {   // actor_2
        v.a(A_ACT, CA1 + kb * BK, row0);
        v.w(W_A2, kb * BK, rank * 16, 16, T_A2, kb > 0, true, kb == 3 ? ACC_A2 : -1);
      }
      if (!next) continue;
      v.wait(X_X2);
      for (int kb = 0; kb < 4; ++kb) {   // hidden layer, x2 columns
        v.a(A_ACT, CX + 2 * U + kb * BK, row0);
        v.w(W_HID, DG + 2 * U + kb * BK, g * DG + hf * 128, 128, T_HID, true, true, kb == 3 ? ACC_HID : -1);
      }
      v.wait(X_H);
      for (int kb = 0; kb < 4; ++kb) {   // gate projection of this CTA's 128 units: reset | cand | update
        v.a(A_ACT, CH + g * DG + kb * BK, row0);
        for (int j = 0; j < 3; ++j)
          v.w(W_GRU, kb * BK, g * 3 * DG + j * DG + hf * 128, 128, T_GRU + j * 128, kb > 0, j == 2,
              (kb == 3 && j == 2) ? ACC_GRU : -1);
      }
    }
  }
}

struct Bars {
  uint32_t a_full, a_empty, w_full, w_empty, acc;
};

struct AProducer {
  const Params& P; Bars b; uint32_t sA; const unsigned int* flags; uint32_t me; uint32_t cnt = 0;
  unsigned int xcnt[NX] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  __device__ __forceinline__ void wait(int x) {
    xcnt[x] += CL;
    flag_wait(flags + x * kFlagStride, xcnt[x]);
    fence_async_global();   // the TMA (async proxy) loads below read what the peers' generic-proxy stores wrote
    if (P.timing && blockIdx.x == 0 && me == 0 && x == X_D) P.timing[16 * (xcnt[x] / CL) + 12] = clock64();
    if (P.timing && blockIdx.x == 0 && me == 0 && x == X_P7) P.timing[16 * (xcnt[x] / CL - 1) + 13] = clock64();
  }
  __device__ __forceinline__ void a(int map, int col, int row) {
    if (cnt % NPA == me) {
      const uint32_t s = cnt % NA, ph = (cnt / NA) & 1u;
      wait_local(b.a_empty + s * 8, ph ^ 1u);
      tc::mbar_expect_tx(b.a_full + s * 8, kSlab);
      tc::tma_load_2d(sA + s * kSlab, &P.ma[map], col, row, b.a_full + s * 8);
      if (P.timing && blockIdx.x == 0 && cnt >= 136 && cnt < 208) P.timing[336 + cnt - 136] = clock64();
    }
    ++cnt;
  }
  __device__ __forceinline__ void w(int, int, int, int, uint32_t, bool, bool, int) {}
};
struct WProducer {
  const Params& P; Bars b; uint32_t sW; uint32_t me; uint32_t cnt = 0;
  __device__ __forceinline__ void wait(int) {}
  __device__ __forceinline__ void a(int, int, int) {}
  __device__ __forceinline__ void w(int map, int k, int row, int n, uint32_t, bool, bool, int) {
    if (cnt % NPW == me) {
      const uint32_t s = cnt % NW, ph = (cnt / NW) & 1u;
      wait_local(b.w_empty + s * 8, ph ^ 1u);
      tc::mbar_expect_tx(b.w_full + s * 8, (uint32_t)n * BK * 2);
      tc::tma_load_2d(sW + s * kSlab, &P.mw[map], k, row, b.w_full + s * 8);
      if (P.timing && blockIdx.x == 0 && cnt >= 176 && cnt < 268) P.timing[512 + cnt - 176] = clock64();
    }
    ++cnt;
  }
};
struct MmaIssuer {
  Bars b; uint32_t sA, sW, tmem; long long* timing; uint32_t acnt = 0, wcnt = 0, cur = 0, n_p7 = 0, n_gru = 0;
  __device__ __forceinline__ void wait(int) {}
  __device__ __forceinline__ void a(int, int, int) {
    cur = acnt % NA;
    wait_local(b.a_full + cur * 8, (acnt / NA) & 1u);
    if (timing && acnt >= 136 && acnt < 208) timing[256 + acnt - 136] = clock64();
    ++acnt;
  }
  __device__ __forceinline__ void w(int, int, int, int n, uint32_t tcol, bool accum, bool rel, int bar) {
    const uint32_t s = wcnt % NW;
    wait_local(b.w_full + s * 8, (wcnt / NW) & 1u);
    if (timing && wcnt >= 176 && wcnt < 268) timing[416 + wcnt - 176] = clock64();
    ++wcnt;
    tc::tc_fence_after();
    const uint64_t da = tc::make_desc_sw128(sA + cur * kSlab), db = tc::make_desc_sw128(sW + s * kSlab);
    const uint32_t idesc = tc::make_idesc(BM, n);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k)
      tc::tc_mma_f16(tmem + tcol, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (accum || k > 0) ? 1u : 0u);
    if (timing && wcnt > 176 && wcnt <= 208) timing[640 + (wcnt - 177) * 2] = clock64();
    tc::tc_commit(b.w_empty + s * 8);
    if (rel) tc::tc_commit(b.a_empty + cur * 8);
    if (bar >= 0) tc::tc_commit(b.acc + bar * 8);
    if (timing && wcnt > 176 && wcnt <= 208) timing[641 + (wcnt - 177) * 2] = clock64();
    if (timing && bar == ACC_P7) timing[16 * (n_p7++) + 14] = clock64();
    if (timing && bar == ACC_GRU) timing[16 * (n_gru++) + 15] = clock64();
  }
};

// ------------------------------------------------------------------------------------------------ epilogue helpers
struct Epi {
  const Params& P;
  Bars b;
  uint8_t* gbase;       // generic pointer to the aligned shared-memory base
  uint32_t tmem;        // this thread's TMEM lane base (lane quadrant << 16)
  int rank, row, half, tid;   // row: 0..127 inside the group
  unsigned int* flags;  // this team's counters
  float2* ssq;          // this team's row-statistics slots
  uint32_t apar = 0;
  unsigned int scnt = 0, a2cnt = 0;

  __device__ __forceinline__ float* consts() const { return reinterpret_cast<float*>(gbase + kOffConst); }
  __device__ __forceinline__ void wait_acc(int a) {
    wait_local(b.acc + a * 8, (apar >> a) & 1u);
    apar ^= 1u << a;
    tc::tc_fence_after();
  }
  // all-to-all exchange of two per-row partial sums between the 16 CTAs of the team (through L2)
  __device__ __forceinline__ void exchange(int slot, float p0, float p1, bool sender, float& t0, float& t1) {
    float2* sl = ssq + (size_t)slot * CL * BM;
    if (sender) sl[rank * BM + row] = make_float2(p0, p1);
    epi_bar();
    scnt += CL;
    if (tid == 0) {
      flag_signal(flags + NX * kFlagStride);
      flag_wait(flags + NX * kFlagStride, scnt);
    }
    epi_bar();
    float2 q[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) q[r] = __ldcg(sl + r * BM + row);   // written by other SMs: read through L2
    t0 = 0.f; t1 = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) { t0 += q[r].x; t1 += q[r].y; }    // fixed order: deterministic
  }
  // publish this CTA's part of exchange x to the team
  __device__ __forceinline__ void signal(int x) {
    fence_async_global();       // generic-proxy global stores -> visible to the peers' TMA (async proxy) loads
    tc::tc_fence_before();      // TMEM reads of this phase are complete before anybody may overwrite the columns
    epi_bar();
    if (tid == 0) flag_signal(flags + x * kFlagStride);
  }
  __device__ __forceinline__ void wait_a2() {
    a2cnt += CL;
    if (tid == 0) flag_wait(flags + X_A2 * kFlagStride, a2cnt);
    epi_bar();
  }
};

// RMSNorm(1e-4) * gain -> SiLU of this CTA's 16 columns of one (or two) 256-wide layers; `tcol*`: TMEM columns,
// `cb*`: constant-table offsets of bias / gain, `dst*`: column in the exchange buffer.
__device__ __forceinline__ void narrow_pair(Epi& e, int slot, int grow, bool two, uint32_t tcol0, int cb0, int dst0, uint32_t tcol1,
                                            int cb1, int dst1) {
  float v0[16], v1[16];
  float s0 = 0.f, s1 = 0.f;
  const float* cs = e.consts();
  if (e.half == 0) {
    tmem_ld16(e.tmem + tcol0, v0);
#pragma unroll
    for (int j = 0; j < 16; ++j) { v0[j] += cs[cb0 + j]; s0 = fmaf(v0[j], v0[j], s0); }
    if (two) {
      tmem_ld16(e.tmem + tcol1, v1);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v1[j] += cs[cb1 + j]; s1 = fmaf(v1[j], v1[j], s1); }
    }
  }
  float t0, t1;
  e.exchange(slot, s0, s1, e.half == 0, t0, t1);
  if (e.half == 0 && grow < e.P.N) {
    const float r0 = 1.f / sqrtf(t0 * (1.f / 256.f) + kRmsEps);
    float y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = silu_fast((v0[j] * r0) * cs[cb0 + 16 + j]);
    uint4* o = reinterpret_cast<uint4*>(e.P.act + (size_t)grow * ACT_LD + dst0 + e.rank * 16);
    o[0] = pack_bf8(y); o[1] = pack_bf8(y + 8);
    if (two) {
      const float r1 = 1.f / sqrtf(t1 * (1.f / 256.f) + kRmsEps);
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = silu_fast((v1[j] * r1) * cs[cb1 + 16 + j]);
      uint4* o1 = reinterpret_cast<uint4*>(e.P.act + (size_t)grow * ACT_LD + dst1 + e.rank * 16);
      o1[0] = pack_bf8(y); o1[1] = pack_bf8(y + 8);
    }
  }
}

// constant-table layout (floats): 7 narrow layers x (16 bias + 16 gain), then logits bias, hid bias, hid gain, gru bias
enum { C_X0 = 0, C_O0 = 32, C_A0 = 64, C_X1 = 96, C_O1 = 128, C_A1 = 160, C_A2 = 192, C_LG = 224, C_HB = 256, C_HG = 384,
       C_GB = 512 };

#define SD_PI_STAMP(k) do { if (P.timing && blockIdx.x == 0 && (threadIdx.x & 31) == 0) P.timing[(k)] = clock64(); } while (0)

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
  const uint32_t tmem_slot = b.acc + NACC * 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gbase + kOffBar + kNumBar * 8);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { tc::mbar_init(b.a_full + i * 8, 1); tc::mbar_init(b.a_empty + i * 8, 1); }
    for (int i = 0; i < NW; ++i) { tc::mbar_init(b.w_full + i * 8, 1); tc::mbar_init(b.w_empty + i * 8, 1); }
    for (int i = 0; i < NACC; ++i) tc::mbar_init(b.acc + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp < NEPI) {
    // per-CTA slices of the biases / RMS scales and the actor-tail weights: loaded once, kept for all H iterations
    float* cs = reinterpret_cast<float*>(gbase + kOffConst);
    const int t = threadIdx.x;   // 0..255
    if (t < 16) {
      const int c = rank * 16 + t;
      cs[C_X0 + t] = P.b_in0[c]; cs[C_X0 + 16 + t] = P.g_in0[c];
      cs[C_O0 + t] = P.b_i0[c];  cs[C_O0 + 16 + t] = P.g_i0[c];
      cs[C_A0 + t] = P.b_a0[c];  cs[C_A0 + 16 + t] = P.g_a0[c];
      cs[C_X1 + t] = P.b_in1[c]; cs[C_X1 + 16 + t] = P.g_in1[c];
      cs[C_O1 + t] = P.b_i1[c];  cs[C_O1 + 16 + t] = P.g_i1[c];
      cs[C_A1 + t] = P.b_a1[c];  cs[C_A1 + 16 + t] = P.g_a1[c];
      cs[C_A2 + t] = P.b_a2[c];  cs[C_A2 + 16 + t] = P.g_a2[c];
    }
    if (t < 32) cs[C_LG + t] = P.b_lg[rank * 32 + t];
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
  } else if (warp >= A_WARP && warp < M_WARP) {
    if (lane == 0) {
      AProducer v{P, b, base + kOffA, P.flags + (size_t)team * flags_per_team(), (uint32_t)(warp - A_WARP)};
      walk(P, rank, v);
    }
  } else if (warp == M_WARP) {
    if (lane == 0) {
      MmaIssuer v{b, base + kOffA, base + kOffW, tmem_base, blockIdx.x == 0 ? P.timing : nullptr};
      walk(P, rank, v);
    }
  } else if (warp < NEPI) {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int quad = warp & 3, half = warp >> 2;
    Epi e{P, b, gbase, tmem_base + ((uint32_t)(quad * 32) << 16), rank, quad * 32 + lane, half, (int)threadIdx.x,
          P.flags + (size_t)team * flags_per_team(), P.ssq + (size_t)team * ssq_per_team()};
    const float* cs = e.consts();
    const int A = P.A, H = P.H;
    const size_t ldf = (size_t)H * F;
    for (int grp = (int)(blockIdx.x / CL); grp < P.ngroups; grp += (int)(gridDim.x / CL)) {
      const int grow = grp * BM + e.row;        // global row of this thread
      const bool rowok = grow < P.N;
      for (int i = 0; i < H; ++i) {
        const bool next = i + 1 < H;
        // ---- dyn_in0 -> x0, img_net_0 -> o0 (the actor_0 partial stays in TMEM)
        e.wait_acc(ACC_P7);
        SD_PI_STAMP(16 * i + 0);
        narrow_pair(e, 0, grow, true, T_X0, C_X0, CX, T_O0, C_O0, CO0);
        e.signal(X_P7);
        SD_PI_STAMP(16 * i + 1);
        if (i > 0) {
          // ---- img_net_1 -> o1
          e.wait_acc(ACC_I1);
          narrow_pair(e, 1, grow, false, T_I1, C_O1, CO1, 0, 0, 0);
          e.signal(X_O1);
          SD_PI_STAMP(16 * i + 2);
          // ---- logits -> unimix + Gumbel arg-max (distributions.py:16-36): warp half h samples category 2*rank + h
          float uu[16];
          const int cat = rank * 2 + half;
          if (rowok) {
            const float4* up = reinterpret_cast<const float4*>(P.u + ((size_t)grow * H + (i - 1)) * SK + cat * KC);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 t = __ldg(up + q); uu[4 * q] = t.x; uu[4 * q + 1] = t.y; uu[4 * q + 2] = t.z; uu[4 * q + 3] = t.w; }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) uu[j] = 0.5f;
          }
          e.wait_acc(ACC_LG);
          float lg[16];
          tmem_ld16(e.tmem + T_LG + (uint32_t)(half * 16), lg);
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) { lg[j] += cs[C_LG + half * 16 + j]; m = fmaxf(m, lg[j]); }
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { lg[j] = __expf(lg[j] - m); s += lg[j]; }
          const float inv = __fdividef(1.f - P.unimix, s), uni = P.unimix * (1.f / KC);
          // arg max_k [ log(p_k (1-eps) + eps/K) - log(-log u_k) ] == first arg max of softmax(log p~ - lse + g)
          int best = 0;
          float bv = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = __logf(fmaf(lg[j], inv, uni)) - __logf(-logf(uu[j]));
            if (z > bv) { bv = z; best = j; }
          }
          if (rowok) {
            float* fo = P.feats + (size_t)grow * ldf + (size_t)i * F + cat * KC;
            __nv_bfloat16* bo = P.big_bf + (size_t)grow * ldf + (size_t)i * F + cat * KC;
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = (j == best) ? 1.f : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(fo)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
            reinterpret_cast<uint4*>(bo)[0] = pack_bf8(y);
            reinterpret_cast<uint4*>(bo)[1] = pack_bf8(y + 8);
          }
          e.signal(X_Z);
          SD_PI_STAMP(16 * i + 3);
        }
        // ---- actor_0 -> a0, dyn_in1 -> x1
        e.wait_acc(ACC_ZIN);
        narrow_pair(e, 2, grow, true, T_A0, C_A0, CA0, T_X1, C_X1, CX + U);
        e.signal(X_ZIN);
        SD_PI_STAMP(16 * i + 4);
        e.wait_acc(ACC_A1);
        narrow_pair(e, 3, grow, false, T_A1, C_A1, CA1, 0, 0, 0);
        e.signal(X_A1);
        SD_PI_STAMP(16 * i + 5);
        e.wait_acc(ACC_A2);
        narrow_pair(e, 4, grow, false, T_A2, C_A2, CA2, 0, 0, 0);
        e.signal(X_A2);
        SD_PI_STAMP(16 * i + 6);
        // ---- actor tail (networks.py:374-377, distributions.py:217-231, rssm.py:44,48): warp w owns row 8*rank + w
        {
          const int trow = grp * BM + rank * NEPI + warp;
          const float* wl = P.tail_in_smem ? reinterpret_cast<const float*>(gbase + kOffTail) : P.w_last;
          const int ldl = P.tail_in_smem ? 256 : P.ldk_last;
          const float* w2 = P.tail_in_smem ? wl + P.act_out * 256 : P.w_in2;
          const int ld2 = P.tail_in_smem ? 256 : P.ldw_in2;
          const float* b2 = P.tail_in_smem ? w2 + A * 256 : P.b_in2;
          const float* g2 = P.tail_in_smem ? b2 + 256 : P.g_in2;
          const float nz = (trow < P.N && lane < A) ? __ldg(P.act_noise + ((size_t)trow * H + i) * A + lane) : 0.5f;
          e.wait_a2();
          if (trow < P.N) {
            // a2 row: lane owns columns [8*lane, +8) (written by 16 different CTAs: read through L2)
            const uint4 pk = __ldcg(reinterpret_cast<const uint4*>(P.act + (size_t)trow * ACT_LD + CA2) + lane);
            float x[8];
            {
              const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(p2[q]); x[2 * q] = f.x; x[2 * q + 1] = f.y; }
            }
            float mine = 0.f, mine2 = 0.f;   // lane j keeps output j (and output j + A for the bounded-normal std)
            for (int j = 0; j < P.act_out; ++j) {
              const float4 wa = *reinterpret_cast<const float4*>(wl + (size_t)j * ldl + lane * 8);
              const float4 wb = *reinterpret_cast<const float4*>(wl + (size_t)j * ldl + lane * 8 + 4);
              float acc = x[0] * wa.x;
              acc = fmaf(x[1], wa.y, acc); acc = fmaf(x[2], wa.z, acc); acc = fmaf(x[3], wa.w, acc);
              acc = fmaf(x[4], wb.x, acc); acc = fmaf(x[5], wb.y, acc); acc = fmaf(x[6], wb.z, acc); acc = fmaf(x[7], wb.w, acc);
              acc = warp_sum(acc) + __ldg(P.b_last + j);
              if (j == lane) mine = acc;
              if (j == lane + A) mine2 = acc;
            }
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
            const float ab = act / fmaxf(fabsf(act), 1.f);
            if (lane < A) P.actions[((size_t)trow * H + i) * A + lane] = act;
            if (next) {
              float vv[8], ss = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) vv[q] = b2[lane * 8 + q];
              for (int a = 0; a < A; ++a) {
                const float av = __shfl_sync(0xffffffffu, ab, a);
                const float4 wa = *reinterpret_cast<const float4*>(w2 + (size_t)a * ld2 + lane * 8);
                const float4 wb = *reinterpret_cast<const float4*>(w2 + (size_t)a * ld2 + lane * 8 + 4);
                vv[0] = fmaf(av, wa.x, vv[0]); vv[1] = fmaf(av, wa.y, vv[1]); vv[2] = fmaf(av, wa.z, vv[2]); vv[3] = fmaf(av, wa.w, vv[3]);
                vv[4] = fmaf(av, wb.x, vv[4]); vv[5] = fmaf(av, wb.y, vv[5]); vv[6] = fmaf(av, wb.z, vv[6]); vv[7] = fmaf(av, wb.w, vv[7]);
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) ss = fmaf(vv[q], vv[q], ss);
              ss = warp_sum(ss);
              const float rs = 1.f / sqrtf(ss * (1.f / 256.f) + kRmsEps);
#pragma unroll
              for (int q = 0; q < 8; ++q) vv[q] = silu_fast((vv[q] * rs) * g2[lane * 8 + q]);
              *(reinterpret_cast<uint4*>(P.act + (size_t)trow * ACT_LD + CX + 2 * U) + lane) = pack_bf8(vv);
            }
          }
          SD_PI_STAMP(16 * i + 7);
          if (!next) continue;
          e.signal(X_X2);
        }
        // ---- block-GRU hidden layer: RMSNorm over all 2048 columns (16 CTAs x 128) -> SiLU -> h (bf16)
        {
          e.wait_acc(ACC_HID);
          SD_PI_STAMP(16 * i + 8);
          float ss = 0.f;
#pragma unroll 1
          for (int c32 = 0; c32 < 2; ++c32) {
            float v[32];
            tc::tmem_ld32(e.tmem + T_HID + (uint32_t)(half * 64 + c32 * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float t = v[j] + cs[C_HB + half * 64 + c32 * 32 + j]; ss = fmaf(t, t, ss); }
          }
          float* hss = reinterpret_cast<float*>(gbase + kOffHss);
          hss[half * BM + e.row] = ss;
          epi_bar();
          const float mine = hss[e.row] + hss[BM + e.row];
          float tot, dummy;
          e.exchange(5, mine, 0.f, half == 0, tot, dummy);
          const float rs = 1.f / sqrtf(tot * (1.f / (float)D) + kRmsEps);
          uint4* o = reinterpret_cast<uint4*>(P.act + (size_t)grow * ACT_LD + CH + rank * 128 + half * 64);
#pragma unroll 1
          for (int c32 = 0; c32 < 2; ++c32) {
            float v[32];
            tc::tmem_ld32(e.tmem + T_HID + (uint32_t)(half * 64 + c32 * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = silu_fast(((v[j] + cs[C_HB + half * 64 + c32 * 32 + j]) * rs) * cs[C_HG + half * 64 + c32 * 32 + j]);
            if (rowok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) o[c32 * 4 + q] = pack_bf8(v + q * 8);
            }
          }
          e.signal(X_H);
          SD_PI_STAMP(16 * i + 9);
        }
        // ---- gate projection -> GRU gates (rssm.py:63-75): warp half h owns units [64 h, +64) of this CTA's 128
        {
          e.wait_acc(ACC_GRU);
          SD_PI_STAMP(16 * i + 10);
          const size_t col = (size_t)SK + rank * 128 + half * 64;
          const float* din = P.feats + (size_t)grow * ldf + (size_t)i * F + col;
          float* dout = P.feats + (size_t)grow * ldf + (size_t)(i + 1) * F + col;
          __nv_bfloat16* dbf = P.big_bf + (size_t)grow * ldf + (size_t)(i + 1) * F + col;
#pragma unroll 1
          for (int c16 = 0; c16 < 4; ++c16) {
            float qr[16], qc[16], qu[16], dold[16];
            if (rowok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(din + c16 * 16 + q * 4);
                dold[4 * q] = t.x; dold[4 * q + 1] = t.y; dold[4 * q + 2] = t.z; dold[4 * q + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dold[j] = 0.f;
            }
            const uint32_t tb = e.tmem + T_GRU + (uint32_t)(half * 64 + c16 * 16);
            tmem_ld16(tb, qr);
            tmem_ld16(tb + 128u, qc);
            tmem_ld16(tb + 256u, qu);
            const float* gb = cs + C_GB + half * 64 + c16 * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float reset = sigmoid_fast(qr[j] + gb[j]);
              const float cand = tanh_fast(reset * (qc[j] + gb[128 + j]));
              const float upd = sigmoid_fast((qu[j] + gb[256 + j]) - 1.f);
              dold[j] = fmaf(upd, cand - dold[j], dold[j]);
            }
            if (rowok) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(dout + c16 * 16 + q * 4) = make_float4(dold[4 * q], dold[4 * q + 1], dold[4 * q + 2], dold[4 * q + 3]);
              *reinterpret_cast<uint4*>(dbf + c16 * 16) = pack_bf8(dold);
              *reinterpret_cast<uint4*>(dbf + c16 * 16 + 8) = pack_bf8(dold + 8);
            }
          }
          e.signal(X_D);
          SD_PI_STAMP(16 * i + 11);
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
