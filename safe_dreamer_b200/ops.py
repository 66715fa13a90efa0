"""torch.library custom operators over the C ABI (include/safedreamer.h): the boundary BASELINE.json's north_star names.

Why custom ops and not only `torch.autograd.Function`: the reference wraps `_cal_grad` in
`torch.compile(mode="reduce-overhead")` (dreamer.py:231-233, configs/base.yaml:172).  Dynamo cannot trace through ctypes
calls, but it treats a registered operator as one opaque graph node: every entry point below has
  * a CUDA implementation that calls the library on the current stream (inside torch's CUDA-graph capture the library
    just enqueues its kernels: csrc/sd_api.cu `run`),
  * a fake (meta) implementation, so shapes propagate without running anything,
  * where the reference differentiates through it, an autograd formula that calls the matching `*_bwd` operator.
The module that owns the weights is passed as an integer key into `_MODULES` (operators only take tensors and scalars);
its parameters are ALSO passed as a tensor list so autograd sees them as inputs and routes their gradients.

    safedreamer::observe / observe_bwd          RSSM.observe            rssm.py:140-178 (+ autograd)
    safedreamer::prior / prior_bwd              RSSM.prior (batched)    rssm.py:189-195, dreamer.py:485
    safedreamer::kl_loss / kl_loss_bwd          RSSM.kl_loss            rssm.py:222-230
    safedreamer::imagine                        Dreamer._imagine        dreamer.py:673-692 (no_grad in training)
    safedreamer::heads_lambda                   frozen heads + weights + _lambda_return   dreamer.py:589-602
    safedreamer::lambda_return                  Dreamer._lambda_return  dreamer.py:694-707
    safedreamer::cnn_encoder / cnn_encoder_bwd  ConvEncoder.forward     networks.py:192-234 (+ autograd)
"""
from __future__ import annotations

import weakref
from typing import List, Tuple

import torch
from torch import Tensor

from .engine import SD_FLAG_SAVE_TAPE

_MODULES = weakref.WeakValueDictionary()


def module_key(m) -> int:
    """Register `m` (an RSSM mirror) and return the integer the operators take."""
    k = id(m)
    _MODULES[k] = m
    return k


def _mod(key: int):
    m = _MODULES.get(key)
    if m is None:
        raise RuntimeError("safedreamer op: the RSSM module this graph was traced with no longer exists")
    return m


def _capturing(t: Tensor) -> bool:
    return t.is_cuda and torch.cuda.is_current_stream_capturing()


def _engine(rssm, rows, steps, ref: Tensor, **kw):
    """Engine of `rssm`; under CUDA-graph capture the weight re-pack launches must be part of the captured graph (the host
    code that decides "unchanged, skip" does not run at replay), so it is forced."""
    eng = rssm._get_engine(rows, steps, **kw)
    if _capturing(ref):
        rssm.refresh_weights(force=True)
    return eng


# ------------------------------------------------------------------------------------------------ observe
@torch.library.custom_op("safedreamer::observe", mutates_args=())
def observe(embed: Tensor, action: Tensor, init_stoch: Tensor, init_deter: Tensor, reset: Tensor, u: Tensor,
            params: List[Tensor], mod: int, taped: bool) -> Tuple[Tensor, Tensor, Tensor]:
    rssm = _mod(mod)
    B, T = action.shape[0], action.shape[1]
    eng = _engine(rssm, B, T, embed, tape=taped)
    flags = rssm._flags() | (SD_FLAG_SAVE_TAPE if taped else 0)
    c = eng.cfg
    out = (torch.empty(B, T, c.S, c.K, device=embed.device), torch.empty(B, T, c.D, device=embed.device),
           torch.empty(B, T, c.S, c.K, device=embed.device))
    eng.observe(embed, action, init_stoch, init_deter, reset, u, flags=flags, out=out)
    return out


@observe.register_fake
def _(embed, action, init_stoch, init_deter, reset, u, params, mod, taped):
    B, T = action.shape[0], action.shape[1]
    S, K = init_stoch.shape[-2], init_stoch.shape[-1]
    D = init_deter.shape[-1]
    return (embed.new_empty(B, T, S, K, dtype=torch.float32), embed.new_empty(B, T, D, dtype=torch.float32),
            embed.new_empty(B, T, S, K, dtype=torch.float32))


@torch.library.custom_op("safedreamer::observe_bwd", mutates_args=())
def observe_bwd(d_stochs: Tensor, d_deters: Tensor, d_logits: Tensor, params: List[Tensor], mod: int, need_embed: bool,
                need_init: bool, need_w: bool) -> Tuple[Tensor, Tensor, Tensor, List[Tensor]]:
    rssm = _mod(mod)
    eng = rssm._rt.engine
    if eng is None:
        raise RuntimeError("safedreamer::observe_bwd: no engine (forward not run?)")
    B, T = d_deters.shape[0], d_deters.shape[1]
    c = eng.cfg
    dev = d_deters.device
    names = eng.weight_names(0)
    pnames = [n for n, _ in rssm._param_dicts()[0][1].items()]
    wg = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in zip(pnames, params)} if need_w else None
    so, eng.static_outputs = eng.static_outputs, False      # fresh outputs: operators must not alias engine-owned buffers
    try:
        d_embed, d_is, d_id = eng.observe_bwd(B, T, d_stochs, d_deters, d_logits, need_embed, need_init, wg, rssm._flags())
    finally:
        eng.static_outputs = so
    empty = lambda: torch.empty(0, device=dev)      # (every returned tensor must be its own object: no aliasing between outputs)
    assert set(names) == set(pnames)
    return (d_embed if d_embed is not None else empty(), d_is if d_is is not None else empty(),
            d_id if d_id is not None else empty(), [wg[n] for n in pnames] if need_w else [empty() for _ in pnames])


@observe_bwd.register_fake
def _(d_stochs, d_deters, d_logits, params, mod, need_embed, need_init, need_w):
    rssm = _mod(mod)
    B, T = d_deters.shape[0], d_deters.shape[1]
    e = lambda: d_deters.new_empty(0)
    return (d_deters.new_empty(B, T, rssm._embed_size) if need_embed else e(),
            d_deters.new_empty(B, rssm._stoch, rssm._discrete) if need_init else e(),
            d_deters.new_empty(B, rssm._deter) if need_init else e(),
            [torch.empty_like(p, dtype=torch.float32) if need_w else e() for p in params])


def _observe_setup(ctx, inputs, output):
    embed, action, init_stoch, init_deter, reset, u, params, mod, taped = inputs
    ctx.mod = mod
    ctx.params = params
    ctx.need = (embed.requires_grad, init_stoch.requires_grad or init_deter.requires_grad, any(p.requires_grad for p in params))
    if not taped and any(ctx.need):
        raise RuntimeError("safedreamer::observe: gradients requested from an untaped forward")


def _observe_backward(ctx, d_st, d_dt, d_lg):
    need_embed, need_init, need_w = ctx.need
    de, dis, did, wg = torch.ops.safedreamer.observe_bwd(d_st.contiguous(), d_dt.contiguous(), d_lg.contiguous(), ctx.params,
                                                         ctx.mod, need_embed, need_init, need_w)
    return (de if need_embed else None, None, dis if need_init else None, did if need_init else None, None, None,
            list(wg) if need_w else [None] * len(ctx.params), None, None)


observe.register_autograd(_observe_backward, setup_context=_observe_setup)


# ------------------------------------------------------------------------------------------------ batched prior
@torch.library.custom_op("safedreamer::prior", mutates_args=())
def prior(deter: Tensor, u: Tensor, params: List[Tensor], mod: int, taped: bool) -> Tuple[Tensor, Tensor]:
    rssm = _mod(mod)
    lead = deter.shape[:-1]
    rows = 1
    for d in lead:
        rows *= int(d)
    eng = _engine(rssm, min(rows, rssm.max_rows), 1, deter, tape=taped, tape_rows=1)
    so, eng.static_outputs = eng.static_outputs, False
    try:
        stoch, logit = eng.prior(deter, u, flags=rssm._flags() | (SD_FLAG_SAVE_TAPE if taped else 0))
    finally:
        eng.static_outputs = so
    return stoch, logit


@prior.register_fake
def _(deter, u, params, mod, taped):
    rssm = _mod(mod)
    lead = deter.shape[:-1]
    return (deter.new_empty(*lead, rssm._stoch, rssm._discrete, dtype=torch.float32),
            deter.new_empty(*lead, rssm._stoch, rssm._discrete, dtype=torch.float32))


@torch.library.custom_op("safedreamer::prior_bwd", mutates_args=())
def prior_bwd(d_stoch: Tensor, d_logit: Tensor, params: List[Tensor], mod: int, need_deter: bool,
              need_w: bool) -> Tuple[Tensor, List[Tensor]]:
    rssm = _mod(mod)
    eng = rssm._rt.engine
    lead = d_logit.shape[:-2]
    rows = 1
    for d in lead:
        rows *= int(d)
    pnames = [n for n, _ in rssm.named_parameters()]
    wg = None
    if need_w:
        wg = {n: (torch.zeros_like(p, dtype=torch.float32) if n.startswith("_img_net") else None) for n, p in zip(pnames, params)}
    so, eng.static_outputs = eng.static_outputs, False
    try:
        d_deter = eng.prior_bwd(rows, d_stoch, d_logit, need_deter, wg, rssm._flags())
    finally:
        eng.static_outputs = so
    empty = lambda: torch.empty(0, device=d_logit.device)
    return (d_deter.reshape(*lead, -1).clone() if d_deter is not None else empty(),
            [(wg[n] if (need_w and wg[n] is not None) else empty()) for n in pnames])


@prior_bwd.register_fake
def _(d_stoch, d_logit, params, mod, need_deter, need_w):
    rssm = _mod(mod)
    lead = d_logit.shape[:-2]
    e = lambda: d_logit.new_empty(0)
    names = [n for n, _ in rssm.named_parameters()]
    return (d_logit.new_empty(*lead, rssm._deter) if need_deter else e(),
            [torch.empty_like(p, dtype=torch.float32) if (need_w and n.startswith("_img_net")) else e() for n, p in zip(names, params)])


def _prior_setup(ctx, inputs, output):
    deter, u, params, mod, taped = inputs
    ctx.mod, ctx.params = mod, params
    ctx.need = (deter.requires_grad, any(p.requires_grad for p in params))
    if not taped and any(ctx.need):
        raise RuntimeError("safedreamer::prior: gradients requested from an untaped forward")


def _prior_backward(ctx, d_stoch, d_logit):
    need_deter, need_w = ctx.need
    dd, wg = torch.ops.safedreamer.prior_bwd(d_stoch.contiguous(), d_logit.contiguous(), ctx.params, ctx.mod, need_deter, need_w)
    grads = [(g if (need_w and g.numel() > 0) else None) for g in wg]
    return dd if need_deter else None, None, grads, None, None


prior.register_autograd(_prior_backward, setup_context=_prior_setup)


# ------------------------------------------------------------------------------------------------ KL
@torch.library.custom_op("safedreamer::kl_loss", mutates_args=())
def kl_loss(post_logit: Tensor, prior_logit: Tensor, free: float, mod: int) -> Tuple[Tensor, Tensor]:
    rssm = _mod(mod)
    lead = post_logit.shape[:-2]
    rows = 1
    for d in lead:
        rows *= int(d)
    eng = _engine(rssm, min(max(rows, 1), rssm.max_rows), 1, post_logit)
    dyn, rep = eng.kl_loss(post_logit, prior_logit, free)
    return dyn.reshape(lead).clone(), rep.reshape(lead).clone()


@kl_loss.register_fake
def _(post_logit, prior_logit, free, mod):
    lead = post_logit.shape[:-2]
    return post_logit.new_empty(lead, dtype=torch.float32), post_logit.new_empty(lead, dtype=torch.float32)


@torch.library.custom_op("safedreamer::kl_loss_bwd", mutates_args=())
def kl_loss_bwd(post_logit: Tensor, prior_logit: Tensor, free: float, g_dyn: Tensor, g_rep: Tensor, mod: int) -> Tuple[Tensor, Tensor]:
    rssm = _mod(mod)
    eng = rssm._rt.engine
    d_post, d_prior = eng.kl_loss_bwd(post_logit, prior_logit, free, g_dyn.contiguous(), g_rep.contiguous(), True, True)
    return d_post.clone(), d_prior.clone()


@kl_loss_bwd.register_fake
def _(post_logit, prior_logit, free, g_dyn, g_rep, mod):
    return torch.empty_like(post_logit, dtype=torch.float32), torch.empty_like(prior_logit, dtype=torch.float32)


def _kl_setup(ctx, inputs, output):
    post_logit, prior_logit, free, mod = inputs
    ctx.save_for_backward(post_logit, prior_logit)
    ctx.free, ctx.mod = free, mod
    ctx.need = (post_logit.requires_grad, prior_logit.requires_grad)


def _kl_backward(ctx, g_dyn, g_rep):
    post_logit, prior_logit = ctx.saved_tensors
    d_post, d_prior = torch.ops.safedreamer.kl_loss_bwd(post_logit, prior_logit, ctx.free, g_dyn, g_rep, ctx.mod)
    return d_post if ctx.need[0] else None, d_prior if ctx.need[1] else None, None, None


kl_loss.register_autograd(_kl_backward, setup_context=_kl_setup)


# ------------------------------------------------------------------------------------------------ imagination side (no grad)
@torch.library.custom_op("safedreamer::imagine", mutates_args=())
def imagine(stoch: Tensor, deter: Tensor, u: Tensor, act_noise: Tensor, horizon: int, mod: int) -> Tuple[Tensor, Tensor]:
    rssm = _mod(mod)
    N = deter.shape[0]
    eng = _engine(rssm, N, horizon, deter)
    so, eng.static_outputs = eng.static_outputs, False
    try:
        feats, actions = eng.imagine(stoch, deter, u, act_noise, horizon, flags=rssm._flags())
    finally:
        eng.static_outputs = so
    return feats, actions


@imagine.register_fake
def _(stoch, deter, u, act_noise, horizon, mod):
    rssm = _mod(mod)
    N = deter.shape[0]
    return (deter.new_empty(N, horizon, rssm.feat_size, dtype=torch.float32),
            deter.new_empty(N, horizon, rssm._act_dim, dtype=torch.float32))


@torch.library.custom_op("safedreamer::heads_lambda", mutates_args=())
def heads_lambda(feats: Tensor, disc: float, lamb: float, mod: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    rssm = _mod(mod)
    N, H = feats.shape[0], feats.shape[1]
    eng = _engine(rssm, N, H, feats)
    so, eng.static_outputs = eng.static_outputs, False
    try:
        out = eng.heads_lambda(feats, disc, lamb, flags=rssm._flags())
    finally:
        eng.static_outputs = so
    return tuple(out)


@heads_lambda.register_fake
def _(feats, disc, lamb, mod):
    N, H = feats.shape[0], feats.shape[1]
    e = lambda h: feats.new_empty(N, h, 1, dtype=torch.float32)
    return e(H), e(H), e(H), e(H), e(H), e(H - 1)


@torch.library.custom_op("safedreamer::lambda_return", mutates_args=())
def lambda_return(last: Tensor, term: Tensor, reward: Tensor, value: Tensor, boot: Tensor, disc: float, lamb: float,
                  mod: int) -> Tensor:
    rssm = _mod(mod)
    eng = _engine(rssm, 1, 1, reward)
    so, eng.static_outputs = eng.static_outputs, False
    try:
        out = eng.lambda_return(last.float(), term.float(), reward, value, boot, disc, lamb)
    finally:
        eng.static_outputs = so
    return out


@lambda_return.register_fake
def _(last, term, reward, value, boot, disc, lamb, mod):
    N, T = reward.shape[0], reward.shape[1]
    return reward.new_empty(N, T - 1, 1, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ CNN encoder
@torch.library.custom_op("safedreamer::cnn_encoder", mutates_args=())
def cnn_encoder(obs: Tensor, params: List[Tensor], mod: int, taped: bool) -> Tensor:
    enc = _mod(mod)
    frames = int(obs.numel() // (enc._input_shape[0] * enc._input_shape[1] * enc._input_shape[2]))
    if _capturing(obs):
        enc._wkey = None       # the weight re-pack must be part of the captured graph (see _engine above)
    eng = enc._engine(frames, taped)
    return eng.forward(obs, tape=taped, keep_obs=False)


@cnn_encoder.register_fake
def _(obs, params, mod, taped):
    enc = _mod(mod)
    return obs.new_empty(*obs.shape[:-3], enc.out_dim, dtype=torch.float32)


@torch.library.custom_op("safedreamer::cnn_encoder_bwd", mutates_args=())
def cnn_encoder_bwd(d_embed: Tensor, obs: Tensor, params: List[Tensor], mod: int, need_obs: bool,
                    need_w: bool) -> Tuple[Tensor, List[Tensor]]:
    enc = _mod(mod)
    eng = enc._eng
    if eng is None:
        raise RuntimeError("safedreamer::cnn_encoder_bwd: no engine (forward not run?)")
    wg = [torch.zeros_like(p, dtype=torch.float32) for p in params] if need_w else None
    d_obs = eng.backward(d_embed, want_obs_grad=need_obs, weight_grads=wg, obs=obs)
    empty = lambda: torch.empty(0, device=d_embed.device)
    return (d_obs if d_obs is not None else empty(), wg if need_w else [empty() for _ in params])


@cnn_encoder_bwd.register_fake
def _(d_embed, obs, params, mod, need_obs, need_w):
    enc = _mod(mod)
    frames = 1
    for d in d_embed.shape[:-1]:
        frames *= int(d)
    e = lambda: d_embed.new_empty(0)
    return (d_embed.new_empty(frames, *enc._input_shape, dtype=torch.float32) if need_obs else e(),
            [torch.empty_like(p, dtype=torch.float32) if need_w else e() for p in params])


def _cnn_setup(ctx, inputs, output):
    obs, params, mod, taped = inputs
    ctx.mod, ctx.params, ctx.obs_shape = mod, params, obs.shape
    ctx.save_for_backward(obs)
    ctx.need = (obs.requires_grad, any(p.requires_grad for p in params))
    if not taped and any(ctx.need):
        raise RuntimeError("safedreamer::cnn_encoder: gradients requested from an untaped forward")


def _cnn_backward(ctx, g):
    need_obs, need_w = ctx.need
    (obs,) = ctx.saved_tensors
    d_obs, wg = torch.ops.safedreamer.cnn_encoder_bwd(g.contiguous(), obs, ctx.params, ctx.mod, need_obs, need_w)
    return (d_obs.reshape(ctx.obs_shape) if need_obs else None, list(wg) if need_w else [None] * len(ctx.params), None, None)


cnn_encoder.register_autograd(_cnn_backward, setup_context=_cnn_setup)
