/*
 * safedreamer.h -- C ABI of the B200-native RSSM latent-dynamics library
 * (libsafedreamer.so, built from safe_dreamer_b200/csrc/, sm_100a only).
 *
 * The reference (sharmaabhijith/safe-dreamer) has no FFI; its boundary for
 * this path is the Python method surface of world_model/rssm.py and the
 * imagination / lambda-return helpers of world_model/dreamer.py.  Every entry
 * point below names the reference method it replaces (file:line).  Tensors
 * cross the ABI as raw device pointers in the reference's own layouts
 * (row-major, fp32 unless noted); weights are passed in state_dict layout and
 * repacked internally (sd_set_weights).  No torch types, no exceptions, no
 * host synchronisation and no allocation inside the compute calls: the
 * workspace is sized by sd_workspace_bytes() and owned by the handle.
 *
 * All compute calls enqueue on `stream` (a cudaStream_t passed as void*) and
 * return 0 on success or a negative sd_status; sd_last_error_string() gives
 * the reason.  There is no CPU fallback: without a CUDA device sd_create()
 * fails.
 */
#ifndef SAFEDREAMER_H_
#define SAFEDREAMER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 1

typedef enum sd_status {
  SD_OK = 0,
  SD_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  SD_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  SD_ERR_WORKSPACE = -3,   /* rows/steps exceed what the handle was created for */
  SD_ERR_WEIGHTS = -4,     /* weights for a needed module were never set */
  SD_ERR_NO_TAPE = -5      /* backward called without a matching SD_FLAG_SAVE_TAPE forward */
} sd_status;

/* Weight modules (state_dict owners in world_model/dreamer.py:65-160). */
typedef enum sd_module {
  SD_MOD_RSSM = 0,        /* Dreamer.rssm / _frozen_rssm   (rssm.py:78-131)   */
  SD_MOD_ACTOR = 1,       /* Dreamer.actor                 (networks.py:339)  */
  SD_MOD_REWARD = 2,      /* Dreamer.reward                                   */
  SD_MOD_CONT = 3,        /* Dreamer.cont                                     */
  SD_MOD_VALUE = 4,       /* Dreamer.value                                    */
  SD_MOD_SLOW_VALUE = 5,  /* Dreamer._slow_value                              */
  SD_MOD_COUNT = 6
} sd_module;

/* Call flags. */
#define SD_FLAG_BF16       1u  /* dense layers on tcgen05 (bf16 operands, fp32 accumulate) when rows >= 128;
                                  default is the fp32 SIMT path (bit-faithful parity mode, small batches) */
#define SD_FLAG_SAVE_TAPE  2u  /* keep per-step activations for the matching *_bwd call */
#define SD_FLAG_GRAPH      4u  /* replay the step sequence as a cached CUDA graph (ignored while the stream is
                                  already being captured, e.g. under torch.compile reduce-overhead) */
#define SD_FLAG_FEATS_FROM_IMAGINE 8u /* sd_heads_lambda_fwd only: `feats` is the UNMODIFIED feats output of the preceding
                                  SD_FLAG_BF16 sd_imagine_fwd on this handle (same N, H), so its bf16 copy, written step by
                                  step during the rollout, is reused instead of re-casting 168 MB; SD_ERR_INVALID when the
                                  pointer / sizes do not match that call */
#define SD_FLAG_BACKGROUND 16u /* with SD_FLAG_GRAPH: this call is off the caller's critical path; its kernel nodes are
                                  captured WITHOUT programmatic dependent launch, so only one of its kernels holds SM
                                  resources at a time and concurrent latency-critical work on another stream finds
                                  room (two-stream schedule). */
#define SD_FLAG_PERSISTENT 32u /* sd_imagine_fwd only: force the persistent team-resident kernel (csrc/sd_pimg.cuh: the whole
                                  rollout is ONE launch).  It is already the default for SD_FLAG_BF16 calls without a tape on the
                                  base.yaml architecture (D=2048, U=units=256, 32x16 latents, G=8, 2 img / 3 actor layers);
                                  SD_PIMG=0 in the environment makes the launch sequence the default again. */
#define SD_FLAG_LAYERWISE  64u /* sd_imagine_fwd only: force the layer-by-layer launch sequence (13 launches per step); for A/B
                                  measurements and cross-checks.  Both paths agree up to bf16 rounding order. */

/* Sizes of the path: configs/base.yaml:117-127,252-276,340-420. */
typedef struct sd_config {
  int32_t D;            /* deter (2048) */
  int32_t U;            /* rssm hidden (256) */
  int32_t S;            /* stoch categories per row (32) */
  int32_t K;            /* classes per category, `discrete` (16); <= 32 */
  int32_t G;            /* blocks (8) */
  int32_t E;            /* embed size (1024 vision / 256 proprio) */
  int32_t A;            /* action dim */
  int32_t obs_layers;   /* 1 */
  int32_t img_layers;   /* 2 */
  int32_t act_kind;     /* 0 = bounded_normal (continuous), 1 = onehot (discrete, A <= 32) */
  int32_t units;        /* head hidden (256) */
  int32_t actor_layers, value_layers, reward_layers, cont_layers; /* 3,3,1,1 */
  int32_t bins;         /* two-hot bins (255) */
  float unimix;         /* rssm.unimix_ratio (0.01) */
  float act_unimix;     /* actor.dist.disc.unimix_ratio (0.01) */
  float min_std, max_std; /* actor.dist.cont (0.1, 1.0) */
  int32_t max_rows;     /* largest row count of any call (B for observe, N for imagine) */
  int32_t max_steps;    /* largest T / H of any call */
  int32_t max_tape_rows;/* largest B of a SD_FLAG_SAVE_TAPE call (0 = no backward) */
} sd_config;

typedef struct sd_handle sd_handle;

/* ---- lifetime -------------------------------------------------------------------------------- */
int sd_abi_version(void);
const char* sd_last_error_string(void);
/* Bytes of device workspace sd_create() will allocate for this config. */
size_t sd_workspace_bytes(const sd_config* cfg);
int sd_create(const sd_config* cfg, sd_handle** out);
int sd_destroy(sd_handle* h);

/* Number / state_dict name / element count of the i-th weight tensor of a module, in the order
 * sd_set_weights() expects (names as in SURVEY.md section 8b, e.g. "_deter_net._dyn_gru.weight"). */
int sd_weight_count(const sd_handle* h, int module);
const char* sd_weight_name(const sd_handle* h, int module, int i);
int64_t sd_weight_numel(const sd_handle* h, int module, int i);

/* Repack one module's weights (fp32, reference layouts, device pointers) into the internal layouts.
 * Must be called again whenever the parameters changed (optimizer step; laprop.py:116 updates
 * storage in place) -- the host mirror keys this on (data_ptr, _version). */
int sd_set_weights(sd_handle* h, int module, const float* const* tensors, int count, void* stream);

/* ---- posterior scan ------------------------------------------------------------------------- */
/* RSSM.observe (rssm.py:140-156); T == 1 is RSSM.obs_step (rssm.py:158-178).
 *   embed (B,T,E)  action (B,T,A)  init_stoch (B,S,K)  init_deter (B,D)
 *   is_first (B,T) uint8   u (B,T,S,K) uniforms in (0,1) that drive the Gumbel draws
 *   out: stochs (B,T,S,K) exact one-hot, deters (B,T,D), logits (B,T,S,K)            */
int sd_observe_fwd(sd_handle* h, int B, int T, const float* embed, const float* action,
                   const float* init_stoch, const float* init_deter, const uint8_t* is_first,
                   const float* u, float* stochs, float* deters, float* logits,
                   uint32_t flags, void* stream);

/* Reverse-time backward of the last SD_FLAG_SAVE_TAPE sd_observe_fwd (autograd of rssm.py:140-178).
 *   d_stochs/d_deters/d_logits: upstream grads (nullable = zero)
 *   out: d_embed (B,T,E), d_init_stoch (B,S,K), d_init_deter (B,D) (each nullable),
 *        weight_grads: sd_weight_count(SD_MOD_RSSM) fp32 tensors in reference layouts, ACCUMULATED into
 *        (nullable entries / nullable array = skip: dgrad-only, the frozen-weights attack shape). */
int sd_observe_bwd(sd_handle* h, int B, int T, const float* d_stochs, const float* d_deters,
                   const float* d_logits, float* d_embed, float* d_init_stoch, float* d_init_deter,
                   float* const* weight_grads, uint32_t flags, void* stream);

/* ---- prior steps ---------------------------------------------------------------------------- */
/* RSSM.prior (rssm.py:189-195) on R rows: logit = _img_net(deter); stoch = rsample. */
int sd_prior(sd_handle* h, int R, const float* deter, const float* u, float* stoch, float* logit,
             uint32_t flags, void* stream);
/* Backward of the last SD_FLAG_SAVE_TAPE sd_prior (autograd of rssm.py:189-195 as used at dreamer.py:485-486):
 *   d_stoch (R,S,K) nullable, d_logit (R,S,K) nullable (not both) -> d_deter (R,D) nullable; weight_grads as in
 *   sd_observe_bwd (only the _img_net slots are touched; accumulated). */
int sd_prior_bwd(sd_handle* h, int R, const float* d_stoch, const float* d_logit, float* d_deter,
                 float* const* weight_grads, uint32_t flags, void* stream);
/* RSSM.imagine_with_action (rssm.py:197-209); T == 1 is RSSM.img_step (rssm.py:180-187).
 *   stoch (R,S,K) deter (R,D) actions (R,T,A) u (R,T,S,K) -> stochs (R,T,S,K) deters (R,T,D) */
int sd_imagine_with_action(sd_handle* h, int R, int T, const float* stoch, const float* deter,
                           const float* actions, const float* u, float* stochs, float* deters,
                           uint32_t flags, void* stream);

/* ---- imagination rollout with in-loop actor -------------------------------------------------- */
/* Dreamer._imagine (dreamer.py:673-692): H iterations of get_feat -> actor.rsample -> img_step.
 *   stoch0 (N,S,K) deter0 (N,D)  u (N,H,S,K) uniforms  act_noise (N,H,A): N(0,1) eps for the
 *   bounded-normal actor, uniforms for the one-hot actor
 *   out: feats (N,H,S*K+D) [stoch first, rssm.py:211-217], actions (N,H,A)          */
int sd_imagine_fwd(sd_handle* h, int N, int H, const float* stoch0, const float* deter0,
                   const float* u, const float* act_noise, float* feats, float* actions,
                   uint32_t flags, void* stream);
/* dgrad-only backward of the last SD_FLAG_SAVE_TAPE sd_imagine_fwd (frozen weights; the patch-attack
 * shape, README.md:68-116): d_feats (N,H,F), d_actions (N,H,A) -> d_stoch0 (N,S,K), d_deter0 (N,D). */
int sd_imagine_bwd(sd_handle* h, int N, int H, const float* d_feats, const float* d_actions,
                   float* d_stoch0, float* d_deter0, uint32_t flags, void* stream);

/* ---- heads + lambda-return on imagined trajectories ------------------------------------------ */
/* dreamer.py:589-602: frozen reward/cont/value/slow-value heads on feats (N,H,F), TwoHot.mode with the
 * reference pairing (distributions.py:78-98), weight = cumprod(cont*disc), ret = _lambda_return(...).
 *   out (each nullable): reward, cont, value, slow_value, weight (N,H,1); ret (N,H-1,1). */
int sd_heads_lambda_fwd(sd_handle* h, int N, int H, const float* feats, float disc, float lamb,
                        float* reward, float* cont, float* value, float* slow_value, float* weight,
                        float* ret, uint32_t flags, void* stream);
/* Backward of sd_heads_lambda_fwd w.r.t. feats with frozen head weights (the adversarial-patch attack's
 * d(imagined return)/d(feats), README.md:68-116; dreamer.py:589-602 with _lambda_return differentiated).
 *   in : feats (N,H,F); d_ret (N,H-1,1) and optional direct cotangents d_reward / d_cont / d_value (N,H,1), each nullable
 *   out: d_feats (N,H,F).  Needs a handle created with max_tape_rows >= N.  H <= 64. */
int sd_heads_lambda_bwd(sd_handle* h, int N, int H, const float* feats, float disc, float lamb, const float* d_ret,
                        const float* d_reward, const float* d_cont, const float* d_value, float* d_feats,
                        uint32_t flags, void* stream);
/* Dreamer._lambda_return (dreamer.py:694-707) on (N,T,1) inputs -> out (N,T-1,1). */
int sd_lambda_return(int N, int T, const float* last, const float* term, const float* reward,
                     const float* value, const float* boot, float disc, float lamb, float* out,
                     void* stream);
/* RSSM.kl_loss (rssm.py:222-230, distributions.py:266-271) on R rows of (S,K) raw logits:
 * dyn = rep = max(sum_s KL(post||prior), free); also the unimix entropies logged at dreamer.py:575-576
 * (each output nullable). */
int sd_kl_loss(sd_handle* h, int R, const float* post_logit, const float* prior_logit, float free_nats,
               float* dyn_loss, float* rep_loss, float* post_entropy, float* prior_entropy, void* stream);

/* Backward of sd_kl_loss w.r.t. the raw logits (autograd of rssm.py:222-230 with its detach pattern): g_dyn / g_rep are the
 * upstream gradients of the per-row dyn / rep losses (nullable = ones); d_post receives the rep term, d_prior the dyn term
 * (each nullable); rows whose summed KL is below free_nats get zero gradient (the clip). */
int sd_kl_loss_bwd(sd_handle* h, int R, const float* post_logit, const float* prior_logit, float free_nats,
                   const float* g_dyn, const float* g_rep, float* d_post_logit, float* d_prior_logit, void* stream);

/* TwoHot.log_prob (distributions.py:100-129) of R rows of `n` logits (row stride ld) against scalar targets and the bin
 * positions `bins[n]` (ascending; symexp_twohot: distributions.py:242-251): out[r] = sum(two_hot(target_r) * log_softmax). */
int sd_twohot_logprob(const float* logits, int ld, const float* bins, int n, const float* target, int R, float* out,
                      void* stream);
/* Its gradient w.r.t. the logits: d_logits[r][j] = g[r] * (two_hot_j - softmax_j)  (g nullable = ones). */
int sd_twohot_logprob_bwd(const float* logits, int ld, const float* bins, int n, const float* target, const float* g, int R,
                          float* d_logits, int ld_d, void* stream);

/* ReturnEMA.__call__ (networks.py:416-422): q05/q95 = torch.quantile(ret.flatten(), [0.05, 0.95]) (linear interpolation),
 * ema_vals[2] (device, in/out) <- alpha * q + (1 - alpha) * ema_vals, offset = ema_vals[0],
 * scale = max(ema_vals[1] - ema_vals[0], 1).  `ret` holds n device floats; offset / scale are device scalars (nullable). */
int sd_return_ema(const float* ret, int64_t n, double alpha, float* ema_vals, float* offset, float* scale, void* stream);

/* Replay latent write-back = Buffer.update (utils/buffer.py:44-53; call site dreamer.py:450): row i of the freshly
 * inferred posterior (stoch (R, S, K) one-hot, deter (R, D)) is written to storage slot (time_idx[i], env_idx[i]) of a
 * (n_time, n_env, ...) storage (the reference's LazyTensorStorage(ndim=2): length first, environments second).
 *   store_idx   (n_time, n_env, S) uint8 : class index per categorical (first arg-max over K; K <= 256) -- 32 B per row
 *                                          instead of the reference's 2 KB one-hot
 *   store_stoch (n_time, n_env, S, K) f32: optional verbatim mirror in the reference's layout (nullable)
 *   store_deter (n_time, n_env, D) f32
 * When several rows of one call address the same slot (overlapping sampled slices) the row with the largest i wins, as
 * sequential assignment would (deterministic).  Rows whose index is out of range are skipped and added to *n_bad (device
 * int, nullable; the reference raises IndexError -- the host mirror checks it).  All pointers are device pointers. */
int sd_latent_writeback(const int64_t* env_idx, const int64_t* time_idx, int R, const float* stoch, const float* deter,
                        int S, int K, int D, int64_t n_time, int64_t n_env, uint8_t* store_idx, float* store_stoch,
                        float* store_deter, int* n_bad, void* stream);
/* Read side, Buffer.sample's `initial` (utils/buffer.py:40): gathers R rows from the storage written above; stoch comes
 * back as exact one-hots decoded from the class indices.  Out-of-range rows are zero-filled and counted in *n_bad. */
int sd_latent_gather(const int64_t* env_idx, const int64_t* time_idx, int R, int S, int K, int D, int64_t n_time,
                     int64_t n_env, const uint8_t* store_idx, const float* store_deter, float* stoch, float* deter,
                     int* n_bad, void* stream);

/* Barlow-twins redundancy loss of dreamer.py:525-532 between projected latents x1 (N, E) and (detached) embeddings x2 (N, E):
 * columns standardised with the unbiased std (+1e-8), c = x1n^T x2n / N, loss = sum_i (c_ii - 1)^2 + lambd * sum_{i!=j} c_ij^2.
 * Writes the scalar loss and (nullable) d(loss)/d(x1) (N, E).  E must be a multiple of 16, N >= 2.  `scratch`: device buffer
 * of sd_barlow_scratch_bytes(N, E) bytes.  fp32 throughout (3xTF32 tensor tiles for the two contractions). */
size_t sd_barlow_scratch_bytes(int N, int E);
int sd_barlow_loss(const float* x1, const float* x2, int N, int E, float lambd, float* loss, float* d_x1, void* scratch,
                   void* stream);

/* Fused multi-tensor optimiser step = clip_grad_agc_ (utils/optim/agc.py:15-60) followed by LaProp.step
 * (utils/optim/laprop.py:46-118, amsgrad = centered = False) for `count` fp32 tensors in three launches.
 *   unscale: g' = g * inv_scale (GradScaler.unscale_, dreamer.py:422; 1 when unused) -- BEFORE the clip, as in the reference;
 *   AGC    : per tensor scale = 1 / max(||g'||_2 / (clip * max(||p||_2, pmin)), 1), g' *= scale (clip <= 0: no clipping);
 *            p.grad is left holding the unscaled, clipped gradient;
 *   LaProp : v = beta2 v + one_minus_beta2 g'^2;
 *            m = beta1 m + lr_term * g' / (sqrt(v / bias_correction2) + eps), lr_term = (1 - beta1) * lr;
 *            p -= step_size * m (step_size = 1 / bias_correction1); p -= weight_decay * p.
 * The scalar state (exp_avg_lr_1/2 -> step_size, bias_correction2) stays on the host as in the reference.
 * `table_dev` / `scratch_dev`: device buffers of sd_opt_table_bytes(count) / sd_opt_scratch_bytes(tensors, count) bytes.
 * found_inf (device int, nullable): set to 1 and the update skipped when a gradient norm is not finite.
 * mode 0 = AGC + LaProp, 1 = AGC only (scales the gradients in place, nothing else), 2 = finite check only (raises
 * *found_inf when a gradient norm is not finite; nothing is written: run it over ALL tensors before a step that is
 * split into several calls, so that an overflow skips every one of them). */
typedef struct sd_opt_tensor {
  float* param; float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t numel;
} sd_opt_tensor;
size_t sd_opt_table_bytes(int count);
size_t sd_opt_scratch_bytes(const sd_opt_tensor* tensors, int count);
int sd_agc_laprop_step(const sd_opt_tensor* tensors, int count, int mode, float clip, float pmin, float inv_scale,
                       float beta1, float beta2, float one_minus_beta2, float lr_term, float step_size,
                       float bias_correction2, float eps, float weight_decay, void* table_dev, void* scratch_dev,
                       int* found_inf, void* stream);

/* ---- CNN encoder ------------------------------------------------------ */
/* ConvEncoder (networks.py:192-234): obs - 0.5 -> `layers` x [Conv2dSamePad k=5 stride 1 (networks.py:59-86) ->
 * MaxPool2d(2,2) -> RMSNorm2D over channels, eps 1e-4 (networks.py:89-98,212) -> SiLU] -> flatten in (C,H,W) order.
 * Implicit-GEMM convolutions on tcgen05 (bf16 operands, fp32 accumulate) with the pool / norm / activation in the
 * epilogue; activations between stages are bf16 NHWC and owned by the handle.
 * Supported: kernel 5, 3 input channels, even sizes at every stage, frame width 32 / 64 / 128, depths <= 64. */
typedef struct sd_cnn_config {
  int32_t height, width, channels; /* frame (64, 64, 3) */
  int32_t layers;                  /* stages (4) */
  int32_t kernel;                  /* 5 */
  int32_t depths[8];               /* output channels per stage: depth * mults (32, 48, 64, 64) */
  int32_t max_frames;              /* largest B*T of a call */
  int32_t max_tape_frames;         /* largest B*T of a SD_FLAG_SAVE_TAPE call (0 = forward only) */
} sd_cnn_config;
typedef struct sd_cnn sd_cnn;
int sd_cnn_create(const sd_cnn_config* cfg, sd_cnn** out);
int sd_cnn_destroy(sd_cnn* h);
int64_t sd_cnn_embed_size(const sd_cnn* h);   /* depths[layers-1] * (height >> layers) * (width >> layers) */
/* tensors: per stage (conv weight (Cout,Cin,5,5), conv bias (Cout), RMS scale (Cout)) = state_dict entries
 * layers.{4i}.weight, layers.{4i}.bias, layers.{4i+2}.weight; fp32 device pointers, repacked to bf16 internally. */
int sd_cnn_set_weights(sd_cnn* h, const float* const* tensors, int count, void* stream);
/* ConvEncoder.forward: obs (frames, H, W, 3) fp32 in [0,1] -> embed (frames, sd_cnn_embed_size) fp32.
 * SD_FLAG_SAVE_TAPE keeps the pooled pre-norm maps and arg-max positions for sd_cnn_backward (obs must stay alive). */
int sd_cnn_forward(sd_cnn* h, int frames, const float* obs, float* embed, uint32_t flags, void* stream);
/* Backward of the last SD_FLAG_SAVE_TAPE sd_cnn_forward (autograd of networks.py:192-234).
 *   d_embed (frames, embed_size) fp32 -> d_obs (frames, H, W, 3) fp32 (nullable: skipped; only the attack needs it),
 *   obs: the frames of that forward (the stage-1 weight gradient re-reads them); NULL = the pointer the forward was given,
 *   weight_grads: 3 * layers fp32 tensors in the order / layouts of sd_cnn_set_weights, ACCUMULATED into
 *   (nullable array or entries = skip: frozen encoder).  Deterministic: partial sums are reduced in a fixed order. */
int sd_cnn_backward(sd_cnn* h, int frames, const float* d_embed, const float* obs, float* d_obs, float* const* weight_grads,
                    void* stream);

/* Kernels launched by this library since process start (all handles): bench.py's gpu_launches. */
uint64_t sd_launch_count(void);

/* Hand-off mode of the persistent posterior scan (sd_observe_fwd at B <= 16; csrc/sd_scan.cuh) on the CURRENT device:
 *   0 = five grid barriers per step, 1 = two grid barriers + flagged (value + tag) hand-offs, 2 = 1 + helper CTAs.
 * All modes compute bit-identical results.  By default the first full-length (T >= 16) direct call on a device times the
 * three modes once on its own inputs (about 10 ms, one stream synchronisation, never inside a stream capture) and keeps the
 * fastest; the environment variable SD_SCAN_LL pins the mode for the process.
 *   set = -1: query only; set = 0 | 1 | 2: pin the mode for this device; set = -2: forget it (the next call tunes again).
 * Returns the mode in force BEFORE the call (-1 = not tuned yet).  No reference counterpart (the reference has no such kernel). */
int sd_scan_mode(int set);

#ifdef __cplusplus
}
#endif
#endif /* SAFEDREAMER_H_ */
