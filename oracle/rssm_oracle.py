"""CPU oracle for the RSSM latent-dynamics hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm.  It is the
*checker* used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``safe_dreamer_b200/`` (the product) imports it; the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pins are outputs of the unmodified reference
modules executed in the build container by ``tests/golden/make_golden.py``
and committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function here against them.

Reference lines followed (all relative to the reference checkout):
  world_model/rssm.py:36-75        Deter.forward           -> deter_step
  world_model/rssm.py:158-178      RSSM.obs_step           -> obs_step
  world_model/rssm.py:140-156      RSSM.observe            -> observe
  world_model/rssm.py:180-195      RSSM.img_step / prior   -> img_step / prior
  world_model/rssm.py:197-209      imagine_with_action     -> imagine_with_action
  world_model/rssm.py:211-217      get_feat                -> get_feat
  world_model/rssm.py:222-230      kl_loss                 -> kl_loss
  world_model/networks.py:43-56    BlockLinear.forward     -> block_linear
  world_model/networks.py:331-336  MLP.forward             -> mlp
  world_model/networks.py:374-377  MLPHead.forward         -> head_*
  world_model/distributions.py:16-36   OneHotDist (unimix + ST gumbel) -> sample_onehot
  world_model/distributions.py:78-98   TwoHot.mode             -> twohot_mode
  world_model/distributions.py:217-222 bounded_normal          -> actor_sample
  world_model/distributions.py:238-251 binary / symexp_twohot  -> twohot_bins
  world_model/distributions.py:266-271 kl                      -> kl
  world_model/dreamer.py:673-692   Dreamer._imagine        -> imagine
  world_model/dreamer.py:589-602   heads + weights + return-> heads_lambda
  world_model/dreamer.py:694-707   Dreamer._lambda_return  -> lambda_return
Third-party arithmetic restated from its published definition: torch
(``requirements.txt:1`` pins 2.8.0; 2.11.0 is installed here) --
``F.gumbel_softmax`` (torch/nn/functional.py), ``nn.RMSNorm``,
``OneHotCategorical`` logit normalisation, ``torch.linspace``.

Noise is *injected*: categorical draws take uniforms ``u`` in (0,1) and use
``g = -log(-log(u))`` (the Gumbel that ``-log(Exponential(1))`` produces);
the continuous actor takes standard normals ``eps``.
"""
from __future__ import annotations

import numpy as np

RMS_EPS = 1e-4


from safe_dreamer_b200.synth import (Cfg, all_param_shapes, cast_params, clamp_u, head_param_shapes,  # noqa: E402,F401
                                     init_params, rssm_param_shapes, synth_imagine_inputs, synth_observe_inputs)


# --------------------------------------------------------------------------- primitives
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def silu(x):
    return x * sigmoid(x)


def rms_norm(x, w):
    """nn.RMSNorm(eps=1e-4): x * rsqrt(mean(x^2) + eps) * w (rssm.py:17)."""
    ms = np.mean(x * x, axis=-1, keepdims=True)
    return x * (1.0 / np.sqrt(ms + x.dtype.type(RMS_EPS))) * w


def linear(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def block_linear(x, w, b, G):
    """BlockLinear.forward (networks.py:43-56); weight is (O/G, I/G, G)."""
    R = x.shape[0]
    xg = x.reshape(R, G, -1)
    ys = [xg[:, g, :] @ w[:, :, g].T for g in range(G)]
    return np.stack(ys, axis=1).reshape(R, -1) + b


def softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(-1, keepdims=True)


def logsumexp(x):
    m = x.max(-1, keepdims=True)
    return m + np.log(np.exp(x - m).sum(-1, keepdims=True))


def log_softmax(x):
    return x - logsumexp(x)


def mlp(x, p, prefix_lin, prefix_norm, layers):
    """[Linear -> RMSNorm(1e-4) -> SiLU] x layers (networks.py:313-336, rssm.py:106-130)."""
    for i in range(layers):
        x = silu(rms_norm(linear(x, p[f"{prefix_lin}{i}.weight"], p[f"{prefix_lin}{i}.bias"]),
                          p[f"{prefix_norm}{i}.weight"]))
    return x


# --------------------------------------------------------------------------- distributions
def unimix_logits(logit, unimix):
    """OneHotDist.__init__ (distributions.py:17-23) incl. OneHotCategorical renormalisation."""
    one = logit.dtype.type(1.0)
    um = logit.dtype.type(unimix)
    p = softmax(logit)
    p = p * (one - um) + um / logit.dtype.type(logit.shape[-1])
    l = np.log(p)
    return l - logsumexp(l)


def sample_onehot(logit, u, unimix):
    """OneHotDist.rsample = F.gumbel_softmax(hard=True) with injected uniforms.

    Returns (straight-through value, index, y_soft, normalised logits).
    """
    l = unimix_logits(logit, unimix)
    g = -np.log(-np.log(u.astype(logit.dtype)))
    y = softmax(l + g)
    idx = np.argmax(y, axis=-1)  # first max, like torch.max
    hard = np.zeros_like(y)
    np.put_along_axis(hard, idx[..., None], 1.0, axis=-1)
    return (hard - y) + y, idx, y, l


def onehot_entropy(logit, unimix):
    l = unimix_logits(logit, unimix)
    return -(np.exp(l) * l).sum(-1)


def kl(l_left, l_right):
    """distributions.kl (distributions.py:266-271) on raw logits."""
    a, b = log_softmax(l_left), log_softmax(l_right)
    return (softmax(l_left) * (a - b)).sum(-1)


def kl_loss(post_logit, prior_logit, free):
    """RSSM.kl_loss (rssm.py:222-230): values only (detach only matters for grads)."""
    k = kl(post_logit, prior_logit).sum(-1)
    v = np.maximum(k, post_logit.dtype.type(free))
    return v.copy(), v.copy()  # dyn_loss, rep_loss


def kl_loss_bwd(post_logit, prior_logit, free, g_dyn, g_rep):
    """Autograd of RSSM.kl_loss (rssm.py:222-230): returns (d_post_logit, d_prior_logit).

    rep = clip(KL(post || sg(prior)).sum(-1), free) -> posterior logits; dyn = clip(KL(sg(post) || prior).sum(-1), free) ->
    prior logits; torch.clip passes gradient where the clipped value >= free."""
    lp, lq = log_softmax(post_logit), log_softmax(prior_logit)
    p, q = np.exp(lp), np.exp(lq)
    kls = (p * (lp - lq)).sum(-1, keepdims=True)                      # (..., S, 1)
    act = (kls.sum(-2, keepdims=True) >= post_logit.dtype.type(free)).astype(post_logit.dtype)
    gd = np.asarray(g_dyn, post_logit.dtype)[..., None, None] * act
    gr = np.asarray(g_rep, post_logit.dtype)[..., None, None] * act
    return gr * p * ((lp - lq) - kls), gd * (q - p)


def torch_linspace_f32(start, end, steps):
    """torch.linspace's symmetric fp32 algorithm (front half from start, back half from end)."""
    start, end = np.float32(start), np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    front = (start + step * i.astype(np.float32)).astype(np.float32)
    back = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, front, back).astype(np.float32)


def symexp(x):
    return np.sign(x) * np.expm1(np.abs(x))


def twohot_bins(n=255):
    """symexp_twohot bins (distributions.py:242-251)."""
    if n % 2 == 1:
        half = symexp(torch_linspace_f32(-20.0, 0.0, (n - 1) // 2 + 1)).astype(np.float32)
        return np.concatenate([half, -half[:-1][::-1]]).astype(np.float32)
    half = symexp(torch_linspace_f32(-20.0, 0.0, n // 2)).astype(np.float32)
    return np.concatenate([half, -half[::-1]]).astype(np.float32)


def twohot_mode(logit, bins):
    """TwoHot.mode with the reference's symmetric pairing (distributions.py:78-98)."""
    p = softmax(logit)
    n = logit.shape[-1]
    b = bins.astype(logit.dtype)
    if n % 2 == 1:
        m = (n - 1) // 2
        pair = (p[..., :m] * b[:m])[..., ::-1] + p[..., m + 1:] * b[m + 1:]
        return (p[..., m:m + 1] * b[m:m + 1]).sum(-1, keepdims=True) + pair.sum(-1, keepdims=True)
    h = n // 2
    pair = (p[..., :h] * b[:h])[..., ::-1] + p[..., h:] * b[h:]
    return pair.sum(-1, keepdims=True)


def twohot_logprob(logit, bins, target):
    """TwoHot.log_prob (distributions.py:100-129) with squash = identity: returns (log_prob, mixed_target)."""
    f = logit.dtype.type
    b = bins.astype(logit.dtype)
    n = b.size
    t = np.asarray(target, logit.dtype).reshape(logit.shape[:-1])
    below = np.clip((b <= t[..., None]).sum(-1) - 1, 0, n - 1)
    above = np.clip(n - (b > t[..., None]).sum(-1), 0, n - 1)
    equal = below == above
    db = np.where(equal, f(1.0), np.abs(b[below] - t)).astype(logit.dtype)
    da = np.where(equal, f(1.0), np.abs(b[above] - t)).astype(logit.dtype)
    total = db + da
    wb, wa = da / total, db / total
    mixed = np.zeros_like(logit)
    np.add.at(mixed, tuple(np.indices(below.shape)) + (below,), wb)
    np.add.at(mixed, tuple(np.indices(above.shape)) + (above,), wa)
    log_pred = log_softmax(logit)
    return (mixed * log_pred).sum(-1), mixed


def twohot_logprob_bwd(logit, mixed, g):
    """d(sum_r g_r * log_prob_r)/d(logit) = g * (mixed - softmax) (mixed sums to one)."""
    return np.asarray(g, logit.dtype)[..., None] * (mixed - softmax(logit))


# --------------------------------------------------------------------------- RSSM
def get_feat(stoch, deter):
    """RSSM.get_feat (rssm.py:211-217): [stoch.flat | deter]."""
    return np.concatenate([stoch.reshape(*stoch.shape[:-2], -1), deter], -1)


def deter_step(c: Cfg, P, stoch, deter, action, tape=None):
    """Deter.forward (rssm.py:36-75).  stoch (R,S,K) or (R,SK); deter (R,D); action (R,A)."""
    R = deter.shape[0]
    z = stoch.reshape(R, -1)
    a = action / np.maximum(np.abs(action), action.dtype.type(1.0))
    pre = "_deter_net."
    v0 = linear(deter, P[pre + "_dyn_in0.0.weight"], P[pre + "_dyn_in0.0.bias"])
    v1 = linear(z, P[pre + "_dyn_in1.0.weight"], P[pre + "_dyn_in1.0.bias"])
    v2 = linear(a, P[pre + "_dyn_in2.0.weight"], P[pre + "_dyn_in2.0.bias"])
    x0 = silu(rms_norm(v0, P[pre + "_dyn_in0.1.weight"]))
    x1 = silu(rms_norm(v1, P[pre + "_dyn_in1.1.weight"]))
    x2 = silu(rms_norm(v2, P[pre + "_dyn_in2.1.weight"]))
    x = np.concatenate([x0, x1, x2], -1)
    dg = deter.reshape(R, c.G, -1)
    xin = np.concatenate([dg, np.broadcast_to(x[:, None, :], (R, c.G, x.shape[-1]))], -1).reshape(R, -1)
    hpre = block_linear(xin, P[pre + "_dyn_hid.dyn_hid_0.weight"], P[pre + "_dyn_hid.dyn_hid_0.bias"], c.G)
    h = silu(rms_norm(hpre, P[pre + "_dyn_hid.norm_0.weight"]))
    q = block_linear(h, P[pre + "_dyn_gru.weight"], P[pre + "_dyn_gru.bias"], c.G)
    qg = q.reshape(R, c.G, 3, -1)
    r_ = sigmoid(qg[:, :, 0, :].reshape(R, -1))
    cand = np.tanh(r_ * qg[:, :, 1, :].reshape(R, -1))
    upd = sigmoid(qg[:, :, 2, :].reshape(R, -1) - q.dtype.type(1.0))
    new = upd * cand + (q.dtype.type(1.0) - upd) * deter
    if tape is not None:
        tape.update(z=z, a=a, act_raw=action, deter_in=deter, v0=v0, v1=v1, v2=v2, x=x, hpre=hpre, h=h, q=q)
    return new


def obs_logit(c: Cfg, P, deter, embed, tape=None):
    """_obs_net (rssm.py:106-117) on [deter | embed]."""
    x = np.concatenate([deter, embed], -1)
    acts = []
    for i in range(c.obs_layers):
        v = linear(x, P[f"_obs_net.obs_net_{i}.weight"], P[f"_obs_net.obs_net_{i}.bias"])
        acts.append((x, v))
        x = silu(rms_norm(v, P[f"_obs_net.obs_net_n_{i}.weight"]))
    lg = linear(x, P["_obs_net.obs_net_logit.weight"], P["_obs_net.obs_net_logit.bias"])
    if tape is not None:
        tape.update(obs_acts=acts, obs_last_in=x)
    return lg.reshape(-1, c.S, c.K)


def img_logit(c: Cfg, P, deter, tape=None):
    """_img_net (rssm.py:119-130)."""
    x = deter
    acts = []
    for i in range(c.img_layers):
        v = linear(x, P[f"_img_net.img_net_{i}.weight"], P[f"_img_net.img_net_{i}.bias"])
        acts.append((x, v))
        x = silu(rms_norm(v, P[f"_img_net.img_net_n_{i}.weight"]))
    lg = linear(x, P["_img_net.img_net_logit.weight"], P["_img_net.img_net_logit.bias"])
    if tape is not None:
        tape.update(img_acts=acts, img_last_in=x)
    return lg.reshape(*deter.shape[:-1], c.S, c.K)


def obs_step(c: Cfg, P, stoch, deter, prev_action, embed, reset, u, tape=None):
    """RSSM.obs_step (rssm.py:158-178).  reset: (R,) or (R,1) bool."""
    rs = np.asarray(reset).reshape(-1).astype(bool)
    stoch = np.where(rs[:, None, None], 0, stoch).astype(deter.dtype)
    deter = np.where(rs[:, None], 0, deter).astype(deter.dtype)
    prev_action = np.where(rs[:, None], 0, prev_action).astype(deter.dtype)
    deter = deter_step(c, P, stoch, deter, prev_action, tape)
    logit = obs_logit(c, P, deter, embed, tape)
    st, idx, y, l = sample_onehot(logit, u, c.unimix)
    if tape is not None:
        tape.update(reset=rs, logit=logit, y=y, idx=idx, u=u, deter_out=deter)
    return st, deter, logit, idx


def observe(c: Cfg, P, embed, action, initial, reset, u, tapes=None):
    """RSSM.observe (rssm.py:140-156).  embed (B,T,E) action (B,T,A) reset (B,T[,1]) u (B,T,S,K)."""
    B, T = action.shape[:2]
    stoch, deter = initial
    reset = np.asarray(reset).reshape(B, T)
    S_, D_, L_, I_ = [], [], [], []
    for t in range(T):
        tp = {} if tapes is not None else None
        stoch, deter, logit, idx = obs_step(c, P, stoch, deter, action[:, t], embed[:, t], reset[:, t], u[:, t], tp)
        if tapes is not None:
            tapes.append(tp)
        S_.append(stoch); D_.append(deter); L_.append(logit); I_.append(idx)
    return np.stack(S_, 1), np.stack(D_, 1), np.stack(L_, 1), np.stack(I_, 1)


def prior(c: Cfg, P, deter, u):
    """RSSM.prior (rssm.py:189-195)."""
    logit = img_logit(c, P, deter)
    st, idx, _, _ = sample_onehot(logit, u, c.unimix)
    return st, logit, idx


def img_step(c: Cfg, P, stoch, deter, prev_action, u, tape=None):
    """RSSM.img_step (rssm.py:180-187)."""
    deter = deter_step(c, P, stoch, deter, prev_action, tape)
    logit = img_logit(c, P, deter, tape)
    st, idx, y, l = sample_onehot(logit, u, c.unimix)
    if tape is not None:
        tape.update(logit=logit, y=y, idx=idx, u=u, deter_out=deter)
    return st, deter, logit, idx


def imagine_with_action(c: Cfg, P, stoch, deter, actions, u):
    """RSSM.imagine_with_action (rssm.py:197-209)."""
    S_, D_ = [], []
    for t in range(actions.shape[1]):
        stoch, deter, _, _ = img_step(c, P, stoch, deter, actions[:, t], u[:, t])
        S_.append(stoch); D_.append(deter)
    return np.stack(S_, 1), np.stack(D_, 1)


def return_ema(x, ema_vals, alpha=1e-2):
    """ReturnEMA.__call__ (networks.py:416-422): returns (new ema_vals, offset, scale), all float32.

    torch.quantile(flat, [0.05, 0.95]) with linear interpolation computes, in float32: rank = q * (n - 1),
    below = floor(rank), above = ceil(rank), w = rank - below, lerp(v[below], v[above], w) with torch's lerp
    (w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)); then alpha * q + (1 - alpha) * ema with the two Python-float
    scalars rounded to float32 and no fused multiply-add."""
    f = np.float32
    v = np.sort(np.asarray(x, np.float32).reshape(-1))
    n = v.size
    qv = []
    for q in (f(0.05), f(0.95)):
        rank = f(q * f(n - 1))
        lo, hi = int(np.floor(rank)), int(np.ceil(rank))
        w = f(rank - f(np.floor(rank)))
        a, b = v[lo], v[hi]
        d = f(b - a)
        qv.append(f(a + f(w * d)) if w < f(0.5) else f(b - f(d * f(f(1.0) - w))))
    a32, b32 = f(alpha), f(1.0 - alpha)
    ema = np.asarray(ema_vals, np.float32)
    new = np.array([f(f(a32 * qv[0]) + f(b32 * ema[0])), f(f(a32 * qv[1]) + f(b32 * ema[1]))], np.float32)
    return new, new[0], np.maximum(f(new[1] - new[0]), f(1.0))


# --------------------------------------------------------------------------- heads
def head_logits(p, name, layers, feat):
    """MLPHead.forward up to the distribution factory (networks.py:339-377)."""
    x = mlp(feat, p, f"mlp.layers.{name}_linear", f"mlp.layers.{name}_norm", layers)
    return linear(x, p["last.weight"], p["last.bias"])


def actor_sample(c: Cfg, PA, feat, noise, tape=None):
    """_frozen_actor(feat).rsample() (dreamer.py:684; distributions.py:217-222,230-231).

    cont: noise = eps ~ N(0,1) (R,A) -> tanh(mean) + std*eps (unclipped).
    onehot: noise = uniforms (R,A) -> straight-through gumbel one-hot.
    """
    out = head_logits(PA, "actor", c.actor_layers, feat)
    if c.act_kind == "cont":
        mean, sraw = out[..., :c.A], out[..., c.A:]
        std = feat.dtype.type(c.max_std - c.min_std) * sigmoid(sraw + feat.dtype.type(2.0)) + feat.dtype.type(c.min_std)
        tm = np.tanh(mean)
        if tape is not None:
            tape.update(act_out=out, act_std=std, act_tm=tm, act_noise=noise)
        return tm + std * noise.astype(feat.dtype)
    st, idx, y, l = sample_onehot(out, noise, c.act_unimix)
    if tape is not None:
        tape.update(act_out=out, act_y=y, act_idx=idx, act_noise=noise)
    return st


def imagine(c: Cfg, P, PA, start, H, u, act_noise, tapes=None):
    """Dreamer._imagine (dreamer.py:673-692): returns feats (N,H,F), actions (N,H,A)."""
    stoch, deter = start
    feats, actions = [], []
    for t in range(H):
        tp = {} if tapes is not None else None
        feat = get_feat(stoch, deter)
        action = actor_sample(c, PA, feat, act_noise[:, t], tp)
        feats.append(feat); actions.append(action)
        stoch, deter, _, _ = img_step(c, P, stoch, deter, action, u[:, t], tp)
        if tapes is not None:
            tp["feat"] = feat
            tapes.append(tp)
    return np.stack(feats, 1), np.stack(actions, 1)


def lambda_return(last, term, reward, value, boot, disc, lamb):
    """Dreamer._lambda_return (dreamer.py:694-707).  All inputs (N,T,1)."""
    dt = reward.dtype.type
    live = (dt(1.0) - term)[:, 1:] * dt(disc)
    cont = (dt(1.0) - last)[:, 1:] * dt(lamb)
    interm = reward[:, 1:] + (dt(1.0) - cont) * live * boot[:, 1:]
    out = [boot[:, -1]]
    for i in reversed(range(live.shape[1])):
        out.append(interm[:, i] + live[:, i] * cont[:, i] * out[-1])
    return np.stack(list(reversed(out))[:-1], 1)


def heads_lambda(c: Cfg, PR, PC, PV, PSV, feats):
    """dreamer.py:589-602: frozen reward/cont/value/slow-value on imagined feats,
    discount weights and the lambda-return."""
    bins = twohot_bins(c.bins)
    dt = feats.dtype.type
    rew = twohot_mode(head_logits(PR, "reward", c.reward_layers, feats), bins)
    cont = sigmoid(head_logits(PC, "cont", c.cont_layers, feats))
    val = twohot_mode(head_logits(PV, "value", c.value_layers, feats), bins)
    sval = twohot_mode(head_logits(PSV, "value", c.value_layers, feats), bins)
    disc = dt(1.0 - 1.0 / c.horizon)
    weight = np.cumprod(cont * disc, axis=1)
    ret = lambda_return(np.zeros_like(cont), dt(1.0) - cont, rew, val, val, disc, c.lamb)
    return rew, cont, val, sval, weight, ret


# --------------------------------------------------------------------------- backward (manual autograd)
# Restates what torch.autograd computes for the reference modules; pinned against
# autograd of the real reference by tests/golden (keys bwd_* / imag_bwd_*).
def normact_bwd(dout, v, w):
    """out = silu(rms_norm(v, w)).  Returns dv, dw."""
    dt = v.dtype.type
    rho = dt(1.0) / np.sqrt(np.mean(v * v, axis=-1, keepdims=True) + dt(RMS_EPS))
    n = v * rho
    m = n * w
    sig = sigmoid(m)
    dm = dout * (sig * (dt(1.0) + m * (dt(1.0) - sig)))
    dw = (dm * n).sum(0)
    dn = dm * w
    dv = rho * (dn - n * np.mean(dn * n, axis=-1, keepdims=True))
    return dv, dw


def linear_bwd(dy, x, w):
    return dy @ w, dy.T @ x, dy.sum(0)


def block_linear_bwd(dy, x, w, G):
    R = x.shape[0]
    xg, dyg = x.reshape(R, G, -1), dy.reshape(R, G, -1)
    dx = np.stack([dyg[:, g] @ w[:, :, g] for g in range(G)], 1).reshape(R, -1)
    dw = np.stack([dyg[:, g].T @ xg[:, g] for g in range(G)], -1)
    return dx, dw, dy.sum(0)


def sample_bwd(dz, logit, y, unimix):
    """Gradient of the straight-through sample w.r.t. the raw logits (only the +y term carries grad)."""
    dt = logit.dtype.type
    K = logit.shape[-1]
    dl = y * (dz - (dz * y).sum(-1, keepdims=True))
    p = softmax(logit)
    pt = p * (dt(1.0) - dt(unimix)) + dt(unimix) / dt(K)
    dlp = dl - pt * dl.sum(-1, keepdims=True)
    dp = dlp / pt * (dt(1.0) - dt(unimix))
    return p * (dp - (dp * p).sum(-1, keepdims=True))


def _acc(G, k, v):
    G[k] = v if k not in G else G[k] + v


def mlp_logits_bwd(dlogit, P, G, acts, last_in, lin, norm, last):
    """Backward of [Linear->RMS->SiLU]*n -> Linear.  acts = [(x_in, v)], returns d(input of layer 0)."""
    dx, dw, db = linear_bwd(dlogit, last_in, P[last + ".weight"])
    if G is not None:
        _acc(G, last + ".weight", dw); _acc(G, last + ".bias", db)
    for i in reversed(range(len(acts))):
        x_in, v = acts[i]
        dv, dg = normact_bwd(dx, v, P[f"{norm}{i}.weight"])
        dx, dw, db = linear_bwd(dv, x_in, P[f"{lin}{i}.weight"])
        if G is not None:
            _acc(G, f"{norm}{i}.weight", dg); _acc(G, f"{lin}{i}.weight", dw); _acc(G, f"{lin}{i}.bias", db)
    return dx


def deter_step_bwd(c: Cfg, P, G, tp, g_d):
    """Backward of deter_step given grad of its output.  Returns d_stoch(z flat), d_deter_in, d_action_raw."""
    pre = "_deter_net."
    dt = g_d.dtype.type
    R = g_d.shape[0]
    Dg = c.D // c.G
    q, d_in = tp["q"], tp["deter_in"]
    qg = q.reshape(R, c.G, 3, Dg)
    r_, c_, u_ = (qg[:, :, j, :].reshape(R, -1) for j in range(3))
    Rg = sigmoid(r_); C = np.tanh(Rg * c_); Uu = sigmoid(u_ - dt(1.0))
    dUu = g_d * (C - d_in); dC = g_d * Uu; dd = g_d * (dt(1.0) - Uu)
    du = dUu * Uu * (dt(1.0) - Uu)
    dtn = dC * (dt(1.0) - C * C)
    dc = dtn * Rg; dR = dtn * c_
    dr = dR * Rg * (dt(1.0) - Rg)
    dq = np.stack([dr.reshape(R, c.G, Dg), dc.reshape(R, c.G, Dg), du.reshape(R, c.G, Dg)], 2).reshape(R, -1)
    dh, dw, db = block_linear_bwd(dq, tp["h"], P[pre + "_dyn_gru.weight"], c.G)
    if G is not None:
        _acc(G, pre + "_dyn_gru.weight", dw); _acc(G, pre + "_dyn_gru.bias", db)
    dhpre, dg = normact_bwd(dh, tp["hpre"], P[pre + "_dyn_hid.norm_0.weight"])
    x = tp["x"]
    xin = np.concatenate([d_in.reshape(R, c.G, Dg), np.broadcast_to(x[:, None, :], (R, c.G, x.shape[-1]))], -1).reshape(R, -1)
    dxin, dw, db = block_linear_bwd(dhpre, xin, P[pre + "_dyn_hid.dyn_hid_0.weight"], c.G)
    if G is not None:
        _acc(G, pre + "_dyn_hid.norm_0.weight", dg)
        _acc(G, pre + "_dyn_hid.dyn_hid_0.weight", dw); _acc(G, pre + "_dyn_hid.dyn_hid_0.bias", db)
    dxin = dxin.reshape(R, c.G, -1)
    dd = dd + dxin[:, :, :Dg].reshape(R, -1)
    dx = dxin[:, :, Dg:].sum(1)
    U = c.U
    outs = []
    for j, (name, inp) in enumerate((("_dyn_in0", d_in), ("_dyn_in1", tp["z"]), ("_dyn_in2", tp["a"]))):
        dv, dgj = normact_bwd(dx[:, j * U:(j + 1) * U], tp[f"v{j}"], P[pre + name + ".1.weight"])
        dxi, dw, db = linear_bwd(dv, inp, P[pre + name + ".0.weight"])
        if G is not None:
            _acc(G, pre + name + ".1.weight", dgj); _acc(G, pre + name + ".0.weight", dw); _acc(G, pre + name + ".0.bias", db)
        outs.append(dxi)
    dd = dd + outs[0]
    d_act = outs[2] / np.maximum(np.abs(tp["act_raw"]), dt(1.0))
    return outs[1], dd, d_act


def observe_bwd(c: Cfg, P, tapes, d_stochs, d_deters, d_logits):
    """Reverse-time backward of observe().  Returns (param grads dict, d_embed, d_init_stoch, d_init_deter)."""
    T = len(tapes)
    B = tapes[0]["deter_in"].shape[0]
    G = {}
    d_embed = np.zeros((B, T, c.E), d_deters.dtype)
    g_z = np.zeros((B, c.S, c.K), d_deters.dtype)
    g_d = np.zeros((B, c.D), d_deters.dtype)
    for t in reversed(range(T)):
        tp = tapes[t]
        g_z = g_z + d_stochs[:, t]
        g_d = g_d + d_deters[:, t]
        g_l = d_logits[:, t] + sample_bwd(g_z, tp["logit"], tp["y"], c.unimix)
        dxe = mlp_logits_bwd(g_l.reshape(B, -1), P, G, tp["obs_acts"], tp["obs_last_in"], "_obs_net.obs_net_",
                             "_obs_net.obs_net_n_", "_obs_net.obs_net_logit")
        g_d = g_d + dxe[:, :c.D]
        d_embed[:, t] = dxe[:, c.D:]
        dz, dd, _ = deter_step_bwd(c, P, G, tp, g_d)
        keep = (~tp["reset"])[:, None].astype(dd.dtype)
        g_z = (dz * keep).reshape(B, c.S, c.K)
        g_d = dd * keep
    for k, shp in rssm_param_shapes(c).items():
        if k not in G:
            G[k] = np.zeros(shp, d_deters.dtype)
    return G, d_embed, g_z, g_d


def actor_bwd(c: Cfg, PA, tp, d_action):
    """dgrad-only backward of actor_sample w.r.t. feat."""
    dt = d_action.dtype.type
    if c.act_kind == "cont":
        std, tm, eps = tp["act_std"], tp["act_tm"], tp["act_noise"].astype(d_action.dtype)
        dmean = d_action * (dt(1.0) - tm * tm)
        dstd = d_action * eps
        sg = (std - dt(c.min_std)) / dt(c.max_std - c.min_std)
        dsraw = dstd * dt(c.max_std - c.min_std) * sg * (dt(1.0) - sg)
        dout = np.concatenate([dmean, dsraw], -1)
    else:
        dout = sample_bwd(d_action, tp["act_out"], tp["act_y"], c.act_unimix)
    feat = tp["feat"]
    # recompute trunk activations
    acts, x = [], feat
    for i in range(c.actor_layers):
        v = linear(x, PA[f"mlp.layers.actor_linear{i}.weight"], PA[f"mlp.layers.actor_linear{i}.bias"])
        acts.append((x, v))
        x = silu(rms_norm(v, PA[f"mlp.layers.actor_norm{i}.weight"]))
    return mlp_logits_bwd(dout, {**PA}, None, acts, x, "mlp.layers.actor_linear", "mlp.layers.actor_norm", "last")


def imagine_bwd(c: Cfg, P, PA, tapes, d_feats, d_actions):
    """dgrad-only backward of imagine() (frozen weights): returns d_stoch0 (N,S,K), d_deter0 (N,D)."""
    H = len(tapes)
    N = d_feats.shape[0]
    SK = c.SK
    g_z = np.zeros((N, SK), d_feats.dtype)   # grad w.r.t. stoch entering step t+1 (output of img_step t)
    g_d = np.zeros((N, c.D), d_feats.dtype)
    for t in reversed(range(H)):
        tp = tapes[t]
        d_act = d_actions[:, t].copy()
        if t < H - 1:
            # g_z/g_d currently hold grads of (stoch_{t+1}, deter_{t+1}) = outputs of img_step at step t
            g_l = sample_bwd(g_z.reshape(N, c.S, c.K), tp["logit"], tp["y"], c.unimix)
            dd_img = mlp_logits_bwd(g_l.reshape(N, -1), P, None, tp["img_acts"], tp["img_last_in"], "_img_net.img_net_",
                                    "_img_net.img_net_n_", "_img_net.img_net_logit")
            dz, dd, da = deter_step_bwd(c, P, None, tp, g_d + dd_img)
            d_act = d_act + da
        else:
            dz = np.zeros((N, SK), d_feats.dtype); dd = np.zeros((N, c.D), d_feats.dtype)
        dfeat = actor_bwd(c, PA, tp, d_act) + d_feats[:, t]
        g_z = dz + dfeat[:, :SK]
        g_d = dd + dfeat[:, SK:]
    return g_z.reshape(N, c.S, c.K), g_d


# --------------------------------------------------------------------------- optimiser (section 8f rank 3)
def agc_scale(p, g, clip, pmin):
    """clip_grad_agc_ (utils/optim/agc.py:15-60): per-tensor factor 1 / max(||g|| / (clip * max(||p||, pmin)), 1)."""
    f = np.float32
    pn = np.sqrt(np.sum(p.astype(np.float32) ** 2, dtype=np.float32))
    gn = np.sqrt(np.sum(g.astype(np.float32) ** 2, dtype=np.float32))
    upper = f(np.maximum(pn, f(pmin)) * f(clip))
    return f(1.0) / np.maximum(f(gn / upper), f(1.0))


def laprop_step(p, g, m, v, st, lr, beta1=0.9, beta2=0.999, eps=1e-15, wd=0.0):
    """LaProp.step for one tensor (utils/optim/laprop.py:83-118, amsgrad = centered = False); st = dict(step, lr1, lr2).
    Returns the new (p, m, v); the Python-float state is updated in place."""
    f = np.float32
    st["step"] += 1
    st["lr1"] = st["lr1"] * beta1 + (1 - beta1) * lr
    st["lr2"] = st["lr2"] * beta2 + (1 - beta2)
    bc1 = st["lr1"] / lr if lr != 0.0 else 1.0
    step_size, bc2 = 1 / bc1, st["lr2"]
    v = (v * f(beta2) + (f(1 - beta2) * g) * g).astype(np.float32)
    denom = (np.sqrt(v / f(bc2)) + f(eps)).astype(np.float32)
    m = (m * f(beta1) + f((1 - beta1) * lr) * (g / denom).astype(np.float32)).astype(np.float32)
    p = (p + f(-step_size) * m).astype(np.float32)
    if wd != 0:
        p = (p + f(-wd) * p).astype(np.float32)
    return p, m, v


# --------------------------------------------------------------------------- Barlow loss (section 8f rank 2)
def barlow_loss(x1, x2, lambd):
    """dreamer.py:525-532: returns (loss, d loss / d x1); x2 carries no gradient (detached at :522)."""
    f = x1.dtype.type
    n, e = x1.shape
    mu1, mu2 = x1.mean(0), x2.mean(0)
    s1 = x1.std(0, ddof=1) + f(1e-8)
    s2 = x2.std(0, ddof=1) + f(1e-8)
    x1n, x2n = (x1 - mu1) / s1, (x2 - mu2) / s2
    c = (x1n.T @ x2n) / f(n)
    eye = np.eye(e, dtype=bool)
    loss = ((np.diag(c) - f(1.0)) ** 2).sum() + f(lambd) * (c[~eye] ** 2).sum()
    dc = np.where(eye, f(2.0) * (c - f(1.0)), f(2.0) * f(lambd) * c).astype(x1.dtype)
    g = (x2n @ dc.T) / f(n)                                            # d loss / d x1n
    sigma = x1.std(0, ddof=1)
    dx = (g - g.mean(0)) / s1 - x1n * ((g * x1n).sum(0) / (f(n - 1) * sigma))
    return loss, dx.astype(x1.dtype)


# --------------------------------------------------------------------------- replay latent write-back (section 8f rank 4)
def latent_writeback(index, stoch, deter, store_deter, store_idx=None, store_stoch=None):
    """Buffer.update (utils/buffer.py:44-53; call site dreamer.py:450), in place: flatten (B, T, ...) -> (B*T, ...), then
    storage[index[1], index[0]] <- row (storage is (length, envs, ...): index[1] addresses dim 0, index[0] dim 1).
    Rows are assigned in order, so for a repeated slot the LAST row wins.  `store_stoch` is the reference's one-hot
    storage; `store_idx` (length, envs, S) uint8 is the class-index form the B200 path keeps (first arg-max over K).
    The storage itself lives in torchrl 's LazyTensorStorage (third-party, absent from the image): its published
    behaviour for this call is plain index assignment, which is what is restated here."""
    i0 = np.asarray(index[1]).reshape(-1)
    i1 = np.asarray(index[0]).reshape(-1)
    st = stoch.reshape(-1, *stoch.shape[2:])
    dt = deter.reshape(-1, *deter.shape[2:])
    for r in range(i0.shape[0]):
        store_deter[i0[r], i1[r]] = dt[r]
        if store_stoch is not None:
            store_stoch[i0[r], i1[r]] = st[r]
        if store_idx is not None:
            store_idx[i0[r], i1[r]] = st[r].argmax(-1).astype(np.uint8)


def latent_initial(index, store_idx, store_deter, K):
    """Read side, `initial` of Buffer.sample (utils/buffer.py:40): (stoch one-hot (R, S, K), deter (R, D)) of the rows at
    storage[index[1], index[0]], decoded from the class-index storage."""
    i0 = np.asarray(index[1]).reshape(-1)
    i1 = np.asarray(index[0]).reshape(-1)
    idx = store_idx[i0, i1]
    return np.eye(K, dtype=np.float32)[idx], store_deter[i0, i1].copy()
