"""CUDA replacements for the imagination part of world_model/dreamer.py:
`imagine`  == Dreamer._imagine (dreamer.py:673-692),
`heads_lambda` == frozen heads + weights + `_lambda_return` on imagined feats (dreamer.py:589-602),
`lambda_return` == Dreamer._lambda_return (dreamer.py:694-707).
Heads are passed as modules whose state_dict follows networks.MLPHead; they are registered on the
RSSM so their weights are repacked together with it."""
import torch

from .engine import MOD_ACTOR, MOD_CONT, MOD_REWARD, MOD_SLOW_VALUE, MOD_VALUE

_U_LO = 2.0 ** -24


def _layers(m):
    return sum(1 for n, _ in m.named_parameters() if n.endswith(".weight") and "_linear" in n)


def attach_heads(rssm, actor=None, reward=None, cont=None, value=None, slow_value=None, act_kind="cont",
                 min_std=0.1, max_std=1.0, act_unimix=0.01, bins=255):
    """Register head modules on `rssm` (fixes the engine's head dimensions)."""
    mods = {MOD_ACTOR: actor, MOD_REWARD: reward, MOD_CONT: cont, MOD_VALUE: value, MOD_SLOW_VALUE: slow_value}
    rssm.head_modules = {k: v for k, v in mods.items() if v is not None}
    dims = dict(act_kind=0 if act_kind == "cont" else 1, min_std=min_std, max_std=max_std, act_unimix=act_unimix,
                bins=bins)
    some = next(iter(rssm.head_modules.values()))
    dims["units"] = some.last.in_features
    if actor is not None:
        dims["actor_layers"] = _layers(actor)
    if value is not None:
        dims["value_layers"] = _layers(value)
    if reward is not None:
        dims["reward_layers"] = _layers(reward)
    if cont is not None:
        dims["cont_layers"] = _layers(cont)
    rssm._head_dims = dims
    rssm._rt.engine = None  # dimensions changed: rebuild lazily
    rssm._rt.limits = (0, 0, 0)


@torch.no_grad()
def imagine(rssm, start, imag_horizon, act_noise=None, u=None):
    """Dreamer._imagine: returns feats (N,H,F), actions (N,H,A)."""
    stoch, deter = start
    N = deter.shape[0]
    eng = rssm._get_engine(N, imag_horizon)
    dev = deter.device
    if u is None:
        u = rssm._uniform(N, imag_horizon, rssm._stoch, rssm._discrete)
    if act_noise is None:
        shape = (N, imag_horizon, rssm._act_dim)
        if rssm.static_outputs and not rssm.stage_inputs:   # pointer-stable noise buffer, regenerated in place
            buf = rssm._rt.ubuf.get(("act",) + shape)
            if buf is None or buf.device != dev:
                buf = rssm._rt.ubuf[("act",) + shape] = torch.empty(*shape, device=dev)
            act_noise = buf.normal_() if eng.cfg.act_kind == 0 else buf.uniform_().clamp_(_U_LO, 1 - _U_LO)
        elif eng.cfg.act_kind == 0:
            act_noise = torch.randn(*shape, device=dev)
        else:
            act_noise = torch.rand(*shape, device=dev).clamp_(_U_LO, 1 - _U_LO)
    return eng.imagine(stoch, deter, u, act_noise, imag_horizon, flags=rssm._flags())


@torch.no_grad()
def heads_lambda(rssm, imag_feat, horizon=333, lamb=0.95, slow=True):
    """dreamer.py:589-602 -> reward, cont, value, slow_value, weight, ret."""
    N, H = imag_feat.shape[:2]
    eng = rssm._get_engine(N, H)
    return eng.heads_lambda(imag_feat, 1 - 1 / horizon, lamb, flags=rssm._flags(), slow=slow)


class _ImagineFn(torch.autograd.Function):
    """Grad-enabled rollout: sd_imagine_fwd with a tape / sd_imagine_bwd (dgrad only: frozen weights)."""

    @staticmethod
    def forward(ctx, rssm, stoch, deter, u, act_noise, H):
        N = deter.shape[0]
        eng = rssm._get_engine(N, H, tape=True, tape_rows=N)
        from .engine import SD_FLAG_SAVE_TAPE
        feats, actions = eng.imagine(stoch, deter, u, act_noise, H, flags=rssm._flags() | SD_FLAG_SAVE_TAPE)
        ctx.rssm, ctx.eng, ctx.gen, ctx.N, ctx.H = rssm, eng, eng.tape_gen["scan"], N, H
        ctx.flags = rssm._flags()
        return feats.clone(), actions.clone()

    @staticmethod
    def backward(ctx, d_feats, d_actions):
        eng = ctx.rssm._rt.engine
        if eng is not ctx.eng:
            raise RuntimeError("imagine_grad backward: the engine holding this rollout's tape was rebuilt; raise rssm.max_rows / "
                               "max_steps before the forward")
        eng.check_tape("scan", ctx.gen, "imagine_grad backward")
        d_s, d_d = eng.imagine_bwd(ctx.N, ctx.H, d_feats.contiguous(), d_actions.contiguous(), flags=ctx.flags)
        return None, d_s, d_d, None, None, None


def imagine_grad(rssm, start, imag_horizon, act_noise=None, u=None):
    """Differentiable Dreamer._imagine (the adversarial-patch attack shape, README.md:68-116): gradients flow from
    feats / actions back to the start state through the frozen actor and img_step (straight-through samples)."""
    stoch, deter = start
    N = deter.shape[0]
    dev = deter.device
    if u is None:
        u = rssm._uniform(N, imag_horizon, rssm._stoch, rssm._discrete)
    if act_noise is None:
        eng = rssm._get_engine(N, imag_horizon, tape=True, tape_rows=N)
        shape = (N, imag_horizon, rssm._act_dim)
        act_noise = torch.randn(*shape, device=dev) if eng.cfg.act_kind == 0 else torch.rand(*shape, device=dev).clamp_(_U_LO, 1 - _U_LO)
    return _ImagineFn.apply(rssm, stoch.float(), deter.float(), u, act_noise, int(imag_horizon))


class _HeadsLambdaFn(torch.autograd.Function):
    """Differentiable heads + lambda-return on imagined feats: sd_heads_lambda_fwd / sd_heads_lambda_bwd."""

    @staticmethod
    def forward(ctx, rssm, feats, horizon, lamb):
        N, H = feats.shape[:2]
        eng = rssm._get_engine(N, H, tape=True, tape_rows=N)
        disc = 1 - 1 / horizon
        out = eng.heads_lambda(feats, disc, lamb, flags=rssm._flags())
        ctx.save_for_backward(feats)
        ctx.rssm, ctx.disc, ctx.lamb, ctx.flags = rssm, disc, lamb, rssm._flags()
        ctx.mark_non_differentiable(out[3], out[4])     # slow value and the cumprod weights carry no gradient here
        return tuple(o.clone() for o in out)

    @staticmethod
    def backward(ctx, d_rew, d_cont, d_val, d_sval, d_wgt, d_ret):
        (feats,) = ctx.saved_tensors
        eng = ctx.rssm._rt.engine
        d_feats = eng.heads_lambda_bwd(feats, ctx.disc, ctx.lamb, d_ret, d_rew, d_cont, d_val, flags=ctx.flags)
        return None, d_feats, None, None


def heads_lambda_grad(rssm, imag_feat, horizon=333, lamb=0.95):
    """heads_lambda with gradients w.r.t. imag_feat (frozen heads; the lambda-return recursion differentiated)."""
    return _HeadsLambdaFn.apply(rssm, imag_feat.float(), horizon, lamb)


@torch.no_grad()
def lambda_return(rssm, last, term, reward, value, boot, disc, lamb):
    eng = rssm._get_engine(1, 1)
    return eng.lambda_return(last, term, reward, value, boot, disc, lamb)


@torch.no_grad()
def act(rssm, embed, state, is_first, eval=False, u=None, act_noise=None):
    """Dreamer.act after the encoder (dreamer.py:345-357): obs_step on the previous latent, then the frozen actor on the new
    feat.  state = (stoch (B,S,K), deter (B,D), prev_action (B,A)); returns action (B,A) and the new (stoch, deter, action).
    Two library calls (the single-step persistent posterior kernel and the actor), each one cached CUDA graph when
    `rssm.use_graph`.  eval=True takes the distribution's mode: zero normal noise (tanh(mean)) for the bounded-normal actor,
    constant uniforms (argmax of the logits) for the one-hot actor."""
    prev_stoch, prev_deter, prev_action = state
    B = prev_deter.shape[0]
    dev = prev_deter.device
    eng = rssm._get_engine(B, 1)
    if u is None:
        u = rssm._uniform(B, 1, rssm._stoch, rssm._discrete)
    stoch, deter, _ = eng.observe(embed.reshape(B, 1, -1), prev_action.reshape(B, 1, -1), prev_stoch, prev_deter,
                                  is_first.reshape(B, 1), u.reshape(B, 1, rssm._stoch, rssm._discrete), flags=rssm._flags() & ~1)
    stoch, deter = stoch[:, 0], deter[:, 0]
    if act_noise is None:
        if eng.cfg.act_kind == 0:
            act_noise = torch.zeros(B, 1, rssm._act_dim, device=dev) if eval else torch.randn(B, 1, rssm._act_dim, device=dev)
        else:
            act_noise = (torch.full((B, 1, rssm._act_dim), 0.5, device=dev) if eval
                         else torch.rand(B, 1, rssm._act_dim, device=dev).clamp_(_U_LO, 1 - _U_LO))
    iu = torch.full((B, 1, rssm._stoch, rssm._discrete), 0.5, device=dev)   # the prior sample of a 1-step rollout is never drawn
    _, actions = eng.imagine(stoch, deter, iu, act_noise.reshape(B, 1, -1), 1, flags=rssm._flags() & ~1)
    action = actions[:, 0]
    return action, (stoch, deter, action)


class _BarlowFn(torch.autograd.Function):
    """sd_barlow_loss: loss and d(loss)/d(x1) in one pass (x2 is detached, as at dreamer.py:522)."""

    @staticmethod
    def forward(ctx, x1, x2, lambd):
        from . import _lib
        lib = _lib.load()
        N, E = x1.shape
        a, b = x1.float().contiguous(), x2.detach().float().contiguous()
        scratch = torch.empty(int(lib.sd_barlow_scratch_bytes(N, E)), dtype=torch.uint8, device=a.device)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        need = ctx.needs_input_grad[0]
        d_x1 = torch.empty_like(a) if need else None
        stream = torch.cuda.current_stream(a.device).cuda_stream
        _lib.check(lib.sd_barlow_loss(a.data_ptr(), b.data_ptr(), N, E, float(lambd), loss.data_ptr(),
                                      d_x1.data_ptr() if need else None, scratch.data_ptr(), stream), "sd_barlow_loss")
        if need:
            ctx.save_for_backward(d_x1)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d_x1,) = ctx.saved_tensors
        return d_x1 * g, None, None


def barlow_loss(x1, x2, lambd):
    """dreamer.py:525-532: Barlow-twins loss between projected latents x1 (B*T, E) and detached embeddings x2 (B*T, E)."""
    if not x1.is_cuda:
        raise RuntimeError("barlow_loss: expected CUDA tensors (no CPU implementation)")
    return _BarlowFn.apply(x1, x2, lambd)
