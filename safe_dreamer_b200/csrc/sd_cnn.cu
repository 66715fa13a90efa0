// sd_cnn.cu -- host side of the CNN encoder entry points (include/safedreamer.h: sd_cnn_*).
// Replaces world_model/networks.py:192-234 (ConvEncoder) for kernel_size 5, 3-channel frames and up to 64 channels per
// stage (configs/base.yaml:300-310: depth 16, mults [2,3,4,4] -> 32/48/64/64).  Kernels: sd_cnn.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/safedreamer.h"
#include "sd_cnn.cuh"
#include "sd_cnn_bwd.cuh"
#include "sd_internal.h"

using sd::cnn::bf16;

namespace {
constexpr int kMaxLayers = 8;
inline int round16(int v) { return (v + 15) / 16 * 16; }
inline int pad_cin(int v) { return v <= 16 ? 16 : (v <= 32 ? 32 : 64); }
}  // namespace

struct sd_cnn {
  sd_cnn_config cfg;
  int L = 0;
  int H[kMaxLayers + 1], W[kMaxLayers + 1];   // H[l] x W[l]: input resolution of stage l; [L] = final feature map
  int C[kMaxLayers + 1];                      // real channels entering stage l; C[L] = final channels
  int CP[kMaxLayers];                         // output channels of stage l rounded up to 16 (TMEM columns / MMA N)
  int CS[kMaxLayers + 1];                     // channel stride of the bf16 activation entering stage l (l >= 1)
  bf16* act[kMaxLayers + 1] = {};             // act[l], l >= 1: [max_frames][H[l]][W[l]][CS[l]]
  float* pool[kMaxLayers] = {};               // tape (max_tape_frames)
  uint8_t* arg[kMaxLayers] = {};
  bf16* wpk[kMaxLayers] = {};
  bf16* wT[kMaxLayers] = {};                  // dgrad weights (only with a tape)
  bf16* dy[kMaxLayers] = {};                  // gradient of the conv output of stage l (dense, bf16 NHWC, CP[l] channels)
  float* dx[kMaxLayers + 1] = {};             // gradient of act[l], l >= 1 (fp32 NHWC, CS[l] channels)
  float* scratch = nullptr;                   // partial sums of the weight / bias / scale gradients
  size_t scratch_floats = 0;
  float* bias[kMaxLayers] = {};
  float* gain[kMaxLayers] = {};
  bool weights_set = false;
  int tape_frames = 0;
  const float* tape_obs = nullptr;
  int sms = 148;
  int device = 0;
  std::vector<void*> allocs;
};

static int cnn_validate(const sd_cnn_config& c) {
  if (c.kernel != sd::cnn::KSZ) return sd_fail(SD_ERR_INVALID, "sd_cnn: kernel_size %d unsupported (5 only)", c.kernel);
  if (c.channels != 3) return sd_fail(SD_ERR_INVALID, "sd_cnn: %d input channels unsupported (3 only)", c.channels);
  if (c.layers < 1 || c.layers > kMaxLayers) return sd_fail(SD_ERR_INVALID, "sd_cnn: layers must be 1..%d", kMaxLayers);
  if (c.max_frames < 1) return sd_fail(SD_ERR_INVALID, "sd_cnn: max_frames must be positive");
  int h = c.height, w = c.width;
  for (int l = 0; l < c.layers; ++l) {
    if (c.depths[l] < 1 || c.depths[l] > 64) return sd_fail(SD_ERR_INVALID, "sd_cnn: depth %d of stage %d outside 1..64", c.depths[l], l);
    if ((h & 1) || (w & 1) || h < 2 || w < 2)
      return sd_fail(SD_ERR_INVALID, "sd_cnn: stage %d input %dx%d is not even (MaxPool2d(2,2) would drop a row: unsupported)", l, h, w);
    h /= 2;
    w /= 2;
  }
  const int wp = c.width / 2, hp = c.height / 2;
  if (!(wp == 16 || wp == 32 || wp == 64) || (hp * wp) % 128 != 0)
    return sd_fail(SD_ERR_INVALID, "sd_cnn: frame %dx%d unsupported (width 32/64/128 and height*width/4 a multiple of 128)", c.height, c.width);
  return SD_OK;
}

extern "C" int sd_cnn_create(const sd_cnn_config* cfg, sd_cnn** out) {
  if (!cfg || !out) return sd_fail(SD_ERR_INVALID, "sd_cnn_create: null argument");
  int rc = cnn_validate(*cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return sd_fail(SD_ERR_CUDA, "sd_cnn_create: no CUDA device (this library has no CPU path)");
  sd_cnn* h = new sd_cnn();
  h->cfg = *cfg;
  h->L = cfg->layers;
  cudaGetDevice(&h->device);
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, h->device);
  h->H[0] = cfg->height; h->W[0] = cfg->width; h->C[0] = cfg->channels; h->CS[0] = cfg->channels;
  for (int l = 0; l < h->L; ++l) {
    h->H[l + 1] = h->H[l] / 2; h->W[l + 1] = h->W[l] / 2; h->C[l + 1] = cfg->depths[l];
    h->CP[l] = round16(cfg->depths[l]);
    h->CS[l + 1] = pad_cin(cfg->depths[l]);
  }
  auto alloc = [&](void** p, size_t bytes) -> bool {
    if (cudaMalloc(p, bytes) != cudaSuccess) return false;
    h->allocs.push_back(*p);
    return true;
  };
  bool ok = true;
  const size_t F = (size_t)cfg->max_frames, TF = (size_t)(cfg->max_tape_frames > 0 ? cfg->max_tape_frames : 0);
  for (int l = 0; l < h->L && ok; ++l) {
    const size_t px_out = (size_t)h->H[l + 1] * h->W[l + 1];
    if (l + 1 < h->L) ok = ok && alloc((void**)&h->act[l + 1], F * px_out * h->CS[l + 1] * sizeof(bf16));
    if (TF) {
      ok = ok && alloc((void**)&h->pool[l], TF * px_out * h->CP[l] * sizeof(float));
      ok = ok && alloc((void**)&h->arg[l], TF * px_out * h->CP[l]);
      ok = ok && alloc((void**)&h->dy[l], TF * 4 * px_out * h->CP[l] * sizeof(bf16));
      if (l >= 1) ok = ok && alloc((void**)&h->dx[l], TF * 4 * px_out * h->CS[l] * sizeof(float));
      const int cin_pad = l == 0 ? 16 : h->CS[l];
      ok = ok && alloc((void**)&h->wT[l], (size_t)25 * h->CP[l] * cin_pad * sizeof(bf16));
      const size_t wg = l == 0 ? (size_t)sd::cnn::BM * h->CP[l] : (size_t)48 * h->CS[l] * h->CP[l];   // <= 48 (group, stacked tap) slots
      if (wg * h->sms > h->scratch_floats) h->scratch_floats = wg * h->sms;
    }
    const int cinp = l == 0 ? 16 : h->CS[l];
    const size_t wel = l == 0 ? (size_t)sd::cnn::K1C * h->CP[l] * 8 : (size_t)25 * cinp * h->CP[l];
    ok = ok && alloc((void**)&h->wpk[l], wel * sizeof(bf16));
    ok = ok && alloc((void**)&h->bias[l], 64 * sizeof(float));
    ok = ok && alloc((void**)&h->gain[l], 64 * sizeof(float));
  }
  if (TF && ok) {
    if ((size_t)h->sms * 8 * 128 > h->scratch_floats) h->scratch_floats = (size_t)h->sms * 8 * 128;
    ok = alloc((void**)&h->scratch, h->scratch_floats * sizeof(float));
  }
  if (!ok) {
    for (void* p : h->allocs) cudaFree(p);
    delete h;
    return sd_fail(SD_ERR_CUDA, "sd_cnn_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = h;
  return SD_OK;
}

extern "C" int sd_cnn_destroy(sd_cnn* h) {
  if (!h) return SD_OK;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return SD_OK;
}

extern "C" int64_t sd_cnn_embed_size(const sd_cnn* h) {
  return h ? (int64_t)h->C[h->L] * h->H[h->L] * h->W[h->L] : 0;
}

extern "C" int sd_cnn_set_weights(sd_cnn* h, const float* const* tensors, int count, void* stream) {
  if (!h || !tensors) return sd_fail(SD_ERR_INVALID, "sd_cnn_set_weights: null argument");
  if (count != 3 * h->L) return sd_fail(SD_ERR_INVALID, "sd_cnn_set_weights: expected %d tensors (conv weight, conv bias, RMS scale per stage), got %d", 3 * h->L, count);
  cudaStream_t st = (cudaStream_t)stream;
  for (int l = 0; l < h->L; ++l) {
    const float *w = tensors[3 * l], *b = tensors[3 * l + 1], *g = tensors[3 * l + 2];
    if (!w || !b || !g) return sd_fail(SD_ERR_INVALID, "sd_cnn_set_weights: null tensor for stage %d", l);
    if (l == 0)
      sd::cnn::pack_conv1_kernel<<<8, 256, 0, st>>>(w, b, g, h->C[1], h->CP[0], h->wpk[0], h->bias[0], h->gain[0]);
    else
      sd::cnn::pack_conv_kernel<<<64, 256, 0, st>>>(w, b, g, h->C[l + 1], h->C[l], h->CP[l], h->CS[l], h->wpk[l], h->bias[l], h->gain[l]);
    if (h->wT[l]) sd::cnn::pack_dgrad_kernel<<<64, 256, 0, st>>>(w, h->C[l + 1], h->C[l], h->CP[l], l == 0 ? 16 : h->CS[l], h->wT[l]);
  }
  sd_count_launches(h->L * (h->wT[0] ? 2 : 1));
  SD_CUDA_TRY(cudaGetLastError());
  h->weights_set = true;
  return SD_OK;
}

template <typename K>
static cudaError_t ensure_smem(K kernel, int bytes, unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return cudaSuccess;
  mask |= bit;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

static sd::cnn::StageOut stage_out(sd_cnn* h, int l, bool tape, float* embed) {
  sd::cnn::StageOut o;
  const bool last = l + 1 == h->L;
  o.y = last ? nullptr : h->act[l + 1];
  o.pool = tape ? h->pool[l] : nullptr;
  o.arg = tape ? h->arg[l] : nullptr;
  o.embed = last ? embed : nullptr;
  o.bias = h->bias[l];
  o.gain = h->gain[l];
  o.cout = h->C[l + 1];
  o.cp = h->CP[l];
  o.cnext = last ? h->CP[l] : h->CS[l + 1];
  o.HpWp = h->H[l + 1] * h->W[l + 1];
  o.total = 0;
  return o;
}

template <int CIN, bool RES>
static int launch_conv_t(sd_cnn* h, int l, int frames, const sd::cnn::StageOut& o, cudaStream_t st) {
  static unsigned long long mask = 0;
  using L = sd::cnn::ConvSmem<CIN, RES>;
  SD_CUDA_TRY(ensure_smem(sd::cnn::conv_pool_kernel<CIN, RES>, sd::cnn::kConvSmemBudget, mask));
  sd::cnn::ConvParams p;
  p.x = h->act[l];
  p.wpk = h->wpk[l];
  p.out = o;
  p.Hin = h->H[l]; p.Win = h->W[l]; p.Hp = h->H[l + 1]; p.Wp = h->W[l + 1];
  p.out.total = frames * p.Hp * p.Wp;
  p.tiles = (p.out.total + sd::cnn::BM - 1) / sd::cnn::BM;
  p.stages = L::stages(o.cp);
  const int grid = p.tiles < h->sms ? p.tiles : h->sms;
  static const bool trace2 = getenv("SD_TRACE_CNN") && atoi(getenv("SD_TRACE_CNN")) >= 2;
  static long long* dbg = nullptr;
  p.dbg = nullptr;
  p.dbg_skip = getenv("SD_CNN_SKIP") ? atoi(getenv("SD_CNN_SKIP")) : 0;
  if (trace2) {
    if (!dbg) cudaMalloc(&dbg, 16 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
    p.dbg = dbg;
  }
  sd::cnn::conv_pool_kernel<CIN, RES><<<grid, sd::cnn::THREADS, L::total(o.cp), st>>>(p);
  if (trace2) {
    long long v[16];
    cudaMemcpyAsync(v, dbg, sizeof(v), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[SD_TRACE_CNN] stage %d (%s weights, %d ring stages) CTA0: producer total %lld empty-wait %lld group-wait %lld views %lld | mma total %lld full-wait %lld free-wait %lld tiles %lld | epilogue wait %lld work %lld | mma fence %lld issue %lld commit %lld\n",
            l + 1, RES ? "resident" : "streamed", p.stages, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12]);
  }
  return SD_OK;
}
// weights stay resident in shared memory when all 25 tiles leave room for a useful ring (<= 100 KB)
template <int CIN>
static int launch_conv(sd_cnn* h, int l, int frames, const sd::cnn::StageOut& o, cudaStream_t st) {
  if (25 * (CIN / 8) * o.cp * 16 <= 100 * 1024) return launch_conv_t<CIN, true>(h, l, frames, o, st);
  return launch_conv_t<CIN, false>(h, l, frames, o, st);
}

extern "C" int sd_cnn_forward(sd_cnn* h, int frames, const float* obs, float* embed, uint32_t flags, void* stream) {
  if (!h || !obs || !embed) return sd_fail(SD_ERR_INVALID, "sd_cnn_forward: null argument");
  if (!h->weights_set) return sd_fail(SD_ERR_WEIGHTS, "sd_cnn_forward: sd_cnn_set_weights was never called");
  if (frames < 1 || frames > h->cfg.max_frames) return sd_fail(SD_ERR_WORKSPACE, "sd_cnn_forward: %d frames exceed max_frames %d", frames, h->cfg.max_frames);
  const bool tape = (flags & SD_FLAG_SAVE_TAPE) != 0;
  if (tape && frames > h->cfg.max_tape_frames)
    return sd_fail(SD_ERR_WORKSPACE, "sd_cnn_forward: %d frames exceed max_tape_frames %d", frames, h->cfg.max_tape_frames);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool trace = getenv("SD_TRACE_CNN") != nullptr;   // diagnostic: per-stage times (synchronises)
  cudaEvent_t ev[kMaxLayers + 1];
  if (trace) {
    for (int l = 0; l <= h->L; ++l) cudaEventCreate(&ev[l]);
    cudaEventRecord(ev[0], st);
  }
  {
    static unsigned long long mask = 0;
    SD_CUDA_TRY(ensure_smem(sd::cnn::conv1_pool_kernel, sd::cnn::Conv1Smem::kTotal, mask));
    sd::cnn::Conv1Params p;
    p.obs = obs;
    p.wpk = h->wpk[0];
    p.out = stage_out(h, 0, tape, embed);
    p.Hin = h->H[0]; p.Win = h->W[0]; p.Hp = h->H[1]; p.Wp = h->W[1];
    p.out.total = frames * p.Hp * p.Wp;
    p.tiles = p.out.total / sd::cnn::BM;
    const int grid = p.tiles < h->sms ? p.tiles : h->sms;
    static const bool trace2 = getenv("SD_TRACE_CNN") && atoi(getenv("SD_TRACE_CNN")) >= 2;
    static long long* dbg = nullptr;
    p.dbg = nullptr;
    if (trace2) {
      if (!dbg) cudaMalloc(&dbg, 16 * sizeof(long long));
      cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
      p.dbg = dbg;
    }
    sd::cnn::conv1_pool_kernel<<<grid, sd::cnn::THREADS1, sd::cnn::Conv1Smem::kTotal, st>>>(p);
    if (trace2) {
      long long v[16];
      cudaMemcpyAsync(v, dbg, sizeof(v), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      fprintf(stderr, "[SD_TRACE_CNN] stage 1 CTA0: producer total %lld patch %lld empty-wait %lld build %lld | mma total %lld full-wait %lld free-wait %lld tiles %lld | epilogue wait %lld work %lld\n",
              v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9]);
    }
    if (trace) cudaEventRecord(ev[1], st);
  }
  for (int l = 1; l < h->L; ++l) {
    const sd::cnn::StageOut o = stage_out(h, l, tape, embed);
    int rc;
    switch (h->CS[l]) {
      case 16: rc = launch_conv<16>(h, l, frames, o, st); break;
      case 32: rc = launch_conv<32>(h, l, frames, o, st); break;
      default: rc = launch_conv<64>(h, l, frames, o, st); break;
    }
    if (rc) return rc;
    if (trace) cudaEventRecord(ev[l + 1], st);
  }
  if (trace) {
    cudaEventSynchronize(ev[h->L]);
    fprintf(stderr, "[SD_TRACE_CNN] forward %d frames:", frames);
    for (int l = 0; l < h->L; ++l) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[l], ev[l + 1]);
      fprintf(stderr, " stage%d %.1f us", l + 1, ms * 1e3f);
    }
    fprintf(stderr, "\n");
    for (int l = 0; l <= h->L; ++l) cudaEventDestroy(ev[l]);
  }
  sd_count_launches(h->L);
  SD_CUDA_TRY(cudaGetLastError());
  h->tape_frames = tape ? frames : 0;
  h->tape_obs = tape ? obs : nullptr;
  return SD_OK;
}

// ------------------------------------------------------------------------------------------------ backward
template <int CK, bool RES>
static int launch_dgrad_t(sd_cnn* h, const sd::cnn::DgradParams& p0, cudaStream_t st) {
  static unsigned long long mask = 0;
  using L = sd::cnn::DgradSmem<CK, RES>;
  SD_CUDA_TRY(ensure_smem(sd::cnn::conv_dgrad_kernel<CK, RES>, sd::cnn::kConvSmemBudget, mask));
  sd::cnn::DgradParams p = p0;
  p.stages = L::stages(p.cinp);
  const int grid = p.tiles < h->sms ? p.tiles : h->sms;
  sd::cnn::conv_dgrad_kernel<CK, RES><<<grid, sd::cnn::DG_THREADS, L::total(p.cinp), st>>>(p);
  return SD_OK;
}
template <int CK, bool RES>
static int launch_dgrad_patch_t(sd_cnn* h, const sd::cnn::DgradParams& p0, cudaStream_t st) {
  static unsigned long long mask = 0;
  using L = sd::cnn::DgradPatchSmem<CK, RES>;
  SD_CUDA_TRY(ensure_smem(sd::cnn::conv_dgrad_patch_kernel<CK, RES>, sd::cnn::kConvSmemBudget, mask));
  sd::cnn::DgradParams p = p0;
  p.tiles = p.total / sd::cnn::BM;                 // 16 x 8 blocks: H % 16 == 0 and W % 8 == 0
  const int grid = p.tiles < h->sms ? p.tiles : h->sms;
  sd::cnn::conv_dgrad_patch_kernel<CK, RES><<<grid, sd::cnn::DG_THREADS, L::total(p.cinp), st>>>(p);
  return SD_OK;
}
template <int CK>
static int launch_dgrad(sd_cnn* h, const sd::cnn::DgradParams& p, cudaStream_t st) {
  const bool res = 25 * (CK / 8) * p.cinp * 16 <= 100 * 1024;
  static const bool no_patch = getenv("SD_CNN_NO_PATCH") != nullptr;   // A/B switch: stage every view separately
  if (!no_patch && p.H % sd::cnn::PT_H == 0 && p.W % sd::cnn::PT_W == 0)
    return res ? launch_dgrad_patch_t<CK, true>(h, p, st) : launch_dgrad_patch_t<CK, false>(h, p, st);
  return res ? launch_dgrad_t<CK, true>(h, p, st) : launch_dgrad_t<CK, false>(h, p, st);
}
template <int CX>
static int launch_wgrad(sd_cnn* h, const sd::cnn::WgradParams& p0, int* nblocks, cudaStream_t st) {
  static unsigned long long mask = 0;
  SD_CUDA_TRY(ensure_smem(sd::cnn::conv_wgrad_kernel<CX>, sd::cnn::kWgSmem, mask));
  sd::cnn::WgradParams p = p0;
  const int tpg = 128 / CX, ngrp = (25 + tpg - 1) / tpg;
  int gph = 512 / p.cp;
  if (gph > ngrp) gph = ngrp;
  const int halves = (ngrp + gph - 1) / gph;
  gph = (ngrp + halves - 1) / halves;
  p.gph = gph;
  p.taps_padded = halves * gph * tpg;
  int gx = h->sms / halves;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  *nblocks = gx;
  static const bool trace2 = getenv("SD_TRACE_CNN") && atoi(getenv("SD_TRACE_CNN")) >= 2;
  static long long* dbg = nullptr;
  p.dbg = nullptr;
  if (trace2) {
    if (!dbg) cudaMalloc(&dbg, 16 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
    p.dbg = dbg;
  }
  sd::cnn::conv_wgrad_kernel<CX><<<dim3(gx, halves), sd::cnn::WG_THREADS, sd::cnn::kWgSmem, st>>>(p);
  if (trace2) {
    long long v[16];
    cudaMemcpyAsync(v, dbg, sizeof(v), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[SD_TRACE_CNN] wgrad CX=%d cp=%d grid %dx%d gph %d: producer total %lld empty-wait %lld dy-empty-wait %lld tiles %lld | mma total %lld full-wait %lld dy-full-wait %lld | epilogue waited %lld stored %lld\n",
            CX, p.cp, gx, halves, gph, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[8], v[9]);
  }
  return SD_OK;
}

// patch-resident wgrad (CX = 32 / 64; maps W % 8 == 0 and H % 16 == 0 or H == 8 with an even frame count)
template <int CX>
static int launch_wgrad_patch(sd_cnn* h, const bf16* x, const bf16* dy, int frames, int H, int W, int cp, int cout, int cin, float* g_w,
                              cudaStream_t st) {
  using Cfg = sd::cnn::WgradPatchCfg<CX>;
  static unsigned long long mask = 0;
  SD_CUDA_TRY(ensure_smem(sd::cnn::conv_wgrad_patch_kernel<CX>, sd::cnn::kConvSmemBudget, mask));
  sd::cnn::WgradPatchParams p;
  p.x = x; p.dy = dy; p.partial = h->scratch;
  p.H = H; p.W = W; p.cp = cp;
  p.sh = H % 16 == 0 ? 16 : 8;
  p.nsub = H % 16 == 0 ? 1 : 2;
  p.patch_bytes = Cfg::patch_bytes(p.sh, p.nsub);
  const int smem = Cfg::smem(p.sh, p.nsub);
  if (smem > sd::cnn::kConvSmemBudget) return 1;    // caller falls back to the view-staging kernel
  p.tiles = frames * H * W / sd::cnn::BM;
  int gph = 512 / cp;
  if (gph > Cfg::NGRP) gph = Cfg::NGRP;
  const int halves = (Cfg::NGRP + gph - 1) / gph;
  gph = (Cfg::NGRP + halves - 1) / halves;
  p.gph = gph;
  p.ngroups_padded = halves * gph;
  int gx = h->sms / halves;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  static const bool trace2 = getenv("SD_TRACE_CNN") && atoi(getenv("SD_TRACE_CNN")) >= 2;
  static long long* dbg = nullptr;
  p.dbg = nullptr;
  if (trace2) {
    if (!dbg) cudaMalloc(&dbg, 16 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
    p.dbg = dbg;
  }
  sd::cnn::conv_wgrad_patch_kernel<CX><<<dim3(gx, halves), sd::cnn::WGP_THREADS, smem, st>>>(p);
  if (trace2) {
    long long v[16];
    cudaMemcpyAsync(v, dbg, sizeof(v), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[SD_TRACE_CNN] wgrad-patch CX=%d cp=%d grid %dx%d gph %d smem %d: producer total %lld patch-empty-wait %lld tiles %lld | issuer0 total %lld patch-full-wait %lld\n",
            CX, cp, gx, halves, gph, smem, v[0], v[1], v[3], v[4], v[5]);
  }
  sd::cnn::wgrad_patch_reduce_kernel<<<148, 256, 0, st>>>(h->scratch, gx, p.ngroups_padded * Cfg::TPG, Cfg::TPG, Cfg::NKG, CX, cp, cout, cin, g_w);
  return SD_OK;
}

extern "C" int sd_cnn_backward(sd_cnn* h, int frames, const float* d_embed, const float* obs, float* d_obs, float* const* weight_grads,
                               void* stream) {
  if (!h || !d_embed) return sd_fail(SD_ERR_INVALID, "sd_cnn_backward: null argument");
  if (h->tape_frames != frames || frames < 1)
    return sd_fail(SD_ERR_NO_TAPE, "sd_cnn_backward: no SD_FLAG_SAVE_TAPE forward with %d frames on this handle (last taped forward: %d frames)", frames, h->tape_frames);
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t launches = 0;
  static const bool trace = getenv("SD_TRACE_CNN") != nullptr;   // diagnostic: per-kernel times (synchronises)
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  auto mark = [&](const char* what, int l) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks.push_back({std::string(what) + std::to_string(l + 1), e});
  };
  mark("start", -1);
  for (int l = h->L - 1; l >= 0; --l) {
    const bool last = l + 1 == h->L;
    const int Hp = h->H[l + 1], Wp = h->W[l + 1], cp = h->CP[l];
    float* g_w = weight_grads ? weight_grads[3 * l] : nullptr;
    float* g_b = weight_grads ? weight_grads[3 * l + 1] : nullptr;
    float* g_g = weight_grads ? weight_grads[3 * l + 2] : nullptr;
    {  // 1. SiLU / RMSNorm / max-pool backward
      sd::cnn::NormBwdParams p;
      p.pool = h->pool[l]; p.arg = h->arg[l]; p.gain = h->gain[l];
      p.dout = last ? d_embed : h->dx[l + 1];
      p.dy = h->dy[l];
      p.partial = h->scratch;
      p.total = frames * Hp * Wp; p.cp = cp; p.cout = h->C[l + 1];
      p.ldo = last ? 0 : h->CS[l + 1]; p.embed = last ? 1 : 0; p.HpWp = Hp * Wp; p.Wp = Wp;
      int blocks;
      static const bool lanes = getenv("SD_CNN_NORM_LANES") != nullptr;   // A/B switch: the lane-per-channel kernel
      const bool aligned = last || (h->CS[l + 1] % 4 == 0);
      static const bool perpx = getenv("SD_CNN_NORM_PX") != nullptr;      // A/B switch: the thread-per-pixel kernel
      if (!lanes && !perpx && aligned && (cp == 32 || cp == 48 || cp == 64)) {
        const int ppb = cp == 32 ? 128 : 64;       // pixels per 256-thread block (2 or 4 lanes per pixel)
        blocks = (p.total + ppb - 1) / ppb;
        if (blocks > h->sms * 8) blocks = h->sms * 8;
        if (cp == 32) sd::cnn::norm_pool_bwd_split_kernel<32, 2><<<blocks, 256, 0, st>>>(p);
        else if (cp == 48) sd::cnn::norm_pool_bwd_split_kernel<48, 4><<<blocks, 256, 0, st>>>(p);
        else sd::cnn::norm_pool_bwd_split_kernel<64, 4><<<blocks, 256, 0, st>>>(p);
      } else if (!lanes && aligned && (cp == 32 || cp == 48 || cp == 64)) {
        blocks = (p.total + 127) / 128;
        if (blocks > h->sms * 4) blocks = h->sms * 4;
        if (cp == 32) sd::cnn::norm_pool_bwd_px_kernel<32><<<blocks, 128, 0, st>>>(p);
        else if (cp == 48) sd::cnn::norm_pool_bwd_px_kernel<48><<<blocks, 128, 0, st>>>(p);
        else sd::cnn::norm_pool_bwd_px_kernel<64><<<blocks, 128, 0, st>>>(p);
      } else {
        blocks = (p.total + 31) / 32;
        if (blocks > h->sms * 8) blocks = h->sms * 8;
        if (cp <= 32) sd::cnn::norm_pool_bwd_kernel<1><<<blocks, 256, 0, st>>>(p);
        else sd::cnn::norm_pool_bwd_kernel<2><<<blocks, 256, 0, st>>>(p);
      }
      ++launches;
      if (g_b || g_g) {
        sd::cnn::norm_bwd_reduce_kernel<<<1, 128, 0, st>>>(h->scratch, blocks, h->C[l + 1], g_g, g_b);
        ++launches;
      }
      mark("norm_pool_bwd", l);
    }
    const int Hc = h->H[l], Wc = h->W[l], total = frames * Hc * Wc;
    if (g_w) {  // 3. weight gradient
      int nblocks = 0;
      if (l == 0) {
        static unsigned long long mask = 0;
        SD_CUDA_TRY(ensure_smem(sd::cnn::conv1_wgrad_kernel, sd::cnn::kW1Smem, mask));
        sd::cnn::Wgrad1Params p;
        p.obs = obs ? obs : h->tape_obs; p.dy = h->dy[0]; p.partial = h->scratch;
        p.H = Hc; p.W = Wc; p.total = total; p.tiles = total / sd::cnn::BM; p.cp = cp;
        nblocks = p.tiles < h->sms ? p.tiles : h->sms;
        sd::cnn::conv1_wgrad_kernel<<<nblocks, sd::cnn::W1_THREADS, sd::cnn::kW1Smem, st>>>(p);
        sd::cnn::wgrad1_reduce_kernel<<<20, 256, 0, st>>>(h->scratch, nblocks, cp, h->C[1], g_w);
      } else {
        int rc = 1;
        if (!getenv("SD_CNN_NO_PATCH") && (h->CS[l] == 32 || h->CS[l] == 64) && Wc % 8 == 0 && (Hc % 16 == 0 || (Hc == 8 && frames % 2 == 0)))
          rc = h->CS[l] == 64 ? launch_wgrad_patch<64>(h, h->act[l], h->dy[l], frames, Hc, Wc, cp, h->C[l + 1], h->C[l], g_w, st)
                              : launch_wgrad_patch<32>(h, h->act[l], h->dy[l], frames, Hc, Wc, cp, h->C[l + 1], h->C[l], g_w, st);
        if (rc < 0) return rc;
        if (rc == 1) {   // view-staging kernel (any map size)
        sd::cnn::WgradParams p;
        p.x = h->act[l]; p.dy = h->dy[l]; p.partial = h->scratch;
        p.H = Hc; p.W = Wc; p.total = total; p.tiles = (total + sd::cnn::BM - 1) / sd::cnn::BM; p.cp = cp;
        switch (h->CS[l]) {
          case 16: rc = launch_wgrad<16>(h, p, &nblocks, st); break;
          case 32: rc = launch_wgrad<32>(h, p, &nblocks, st); break;
          default: rc = launch_wgrad<64>(h, p, &nblocks, st); break;
        }
        if (rc) return rc;
        const int tpg = 128 / h->CS[l], ngrp = (25 + tpg - 1) / tpg;
        int gph = 512 / cp;
        if (gph > ngrp) gph = ngrp;
        const int halves = (ngrp + gph - 1) / gph;
        gph = (ngrp + halves - 1) / halves;
        sd::cnn::wgrad_reduce_kernel<<<400, 256, 0, st>>>(h->scratch, nblocks, halves * gph * tpg, h->CS[l], cp, h->C[l + 1], h->C[l], g_w);
        }
      }
      launches += 2;
      mark("wgrad", l);
    }
    if (l >= 1 || d_obs) {  // 2. input gradient
      sd::cnn::DgradParams p;
      p.dy = h->dy[l]; p.wT = h->wT[l];
      p.dx = l >= 1 ? h->dx[l] : d_obs;
      p.H = Hc; p.W = Wc; p.total = total; p.tiles = (total + sd::cnn::BM - 1) / sd::cnn::BM;
      p.cinp = l >= 1 ? h->CS[l] : 16;
      p.ldx = l >= 1 ? h->CS[l] : 3;
      p.nwrite = l >= 1 ? h->CS[l] : 3;
      p.stages = 0;
      int rc;
      switch (cp) {
        case 16: rc = launch_dgrad<16>(h, p, st); break;
        case 32: rc = launch_dgrad<32>(h, p, st); break;
        case 48: rc = launch_dgrad<48>(h, p, st); break;
        default: rc = launch_dgrad<64>(h, p, st); break;
      }
      if (rc) return rc;
      ++launches;
      mark("dgrad", l);
    }
  }
  if (trace) {
    cudaEventSynchronize(marks.back().second);
    fprintf(stderr, "[SD_TRACE_CNN] backward %d frames:", frames);
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      fprintf(stderr, " %s %.1f us", marks[i].first.c_str(), ms * 1e3f);
    }
    fprintf(stderr, "\n");
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
  sd_count_launches(launches);
  SD_CUDA_TRY(cudaGetLastError());
  return SD_OK;
}
