"""Generate golden vectors by executing the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box never runs this; tests read the
committed .npz files.

How the reference is imported: ``world_model``,
``utils`` and ``ablations`` are pre-registered as empty namespace packages so
their ``__init__`` (which pull tensordict / torchrl) are skipped, and
``tensordict.TensorDict`` is stubbed.  Noise is injected by replacing the two
RNG draws the path makes: ``OneHotDist.rsample`` (same formula as
``F.gumbel_softmax(hard=True)`` with g=-log(-log(u)) from a queue) and
``torch.distributions.utils._standard_normal`` (queue of eps).
Weights/inputs come from ``oracle.rssm_oracle.init_params`` / ``synth_*`` (seeded
numpy Philox), so only *outputs* need to be stored for the full-size config.
"""
import os
import sys
import types
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SD_REFERENCE", "/root/reference")

from oracle import rssm_oracle as O  # noqa: E402


def import_reference():
    for pkg in ("world_model", "utils", "ablations"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, pkg)]
        sys.modules[pkg] = m
    td = types.ModuleType("tensordict")

    class TensorDict(dict):
        pass

    td.TensorDict = TensorDict
    sys.modules["tensordict"] = td
    sys.path.insert(0, REF)
    import world_model.rssm as rssm
    import world_model.distributions as dists
    import world_model.networks as networks
    try:
        import world_model.dreamer as dreamer
    except Exception as e:  # pragma: no cover - depends on optional deps
        print("dreamer.py not importable:", repr(e))
        dreamer = None
    return rssm, dists, networks, dreamer


U_QUEUE, EPS_QUEUE = [], []


def patch_noise(dists):
    def rsample(self, sample_shape=(), temperature=1.0):
        u = U_QUEUE.pop(0)
        assert u.shape == self.logits.shape, (u.shape, self.logits.shape)
        g = -torch.log(-torch.log(u))
        y = ((self.logits + g) / temperature).softmax(-1)
        index = y.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(self.logits, memory_format=torch.legacy_contiguous_format).scatter_(-1, index, 1.0)
        return y_hard - y.detach() + y

    dists.OneHotDist.rsample = rsample
    import torch.distributions.normal as tdn

    def std_normal(shape, dtype, device):
        e = EPS_QUEUE.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype)

    tdn._standard_normal = std_normal


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def rssm_cfg(c):
    return NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix,
              initial="learned", device="cpu", obs_layers=c.obs_layers, img_layers=c.img_layers,
              dyn_layers=1, blocks=c.G, norm=True)


def head_cfg(c, name, layers, out, dist):
    return NS(act="SiLU", symlog_inputs=False, device="cpu", layers=layers, units=c.units, name=name,
              dist=dist, outscale=1.0, shape=[out], norm=True)


def build_reference(c, P, rssm_mod, networks):
    R = rssm_mod.RSSM(rssm_cfg(c), c.E, c.A)
    R.load_state_dict({k: t(v) for k, v in P["rssm"].items()}, strict=True)
    if c.act_kind == "cont":
        adist = NS(name="bounded_normal", min_std=c.min_std, max_std=c.max_std)
    else:
        adist = NS(name="onehot", unimix_ratio=c.act_unimix)
    heads = {}
    spec = {
        "actor": ("actor", c.actor_layers, c.A, adist),
        "reward": ("reward", c.reward_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
        "cont": ("cont", c.cont_layers, 1, NS(name="binary")),
        "value": ("value", c.value_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
        "slow_value": ("value", c.value_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
    }
    for key, (name, layers, out, dist) in spec.items():
        h = networks.MLPHead(head_cfg(c, name, layers, out, dist), c.F)
        h.load_state_dict({k: t(v) for k, v in P[key].items()}, strict=True)
        heads[key] = h
    return R, heads


def run_case(tag, c, B, T, N, H, rssm_mod, dists, networks, dreamer, store_inputs):
    torch.manual_seed(0)
    P = O.init_params(c, seed=0)
    R, heads = build_reference(c, P, rssm_mod, networks)
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    out = {"cfg_keys": np.array(sorted(c.as_dict().keys())),
           "cfg_vals": np.array([str(c.as_dict()[k]) for k in sorted(c.as_dict().keys())]),
           "B": B, "T": T, "N": N, "H": H}

    # ---- observe (rssm.py:140-156) with a non-trivial initial state
    rng = np.random.Generator(np.random.Philox(7))
    init_idx = rng.integers(0, c.K, size=(B, c.S))
    init_stoch = np.eye(c.K, dtype=np.float32)[init_idx]
    init_deter = np.tanh(rng.standard_normal((B, c.D), dtype=np.float32)).astype(np.float32)
    reset2 = reset.copy()
    reset2[0, 0] = False  # row 0 keeps its initial state: exercises the carried-in path
    emb_t = t(embed).requires_grad_(True)
    is_t, id_t = t(init_stoch).requires_grad_(True), t(init_deter).requires_grad_(True)
    U_QUEUE[:] = [t(u[:, i]) for i in range(T)]
    stochs, deters, logits = R.observe(emb_t, t(action), (is_t, id_t), t(reset2)[..., None])
    assert not U_QUEUE
    out.update(obs_stoch_idx=stochs.detach().argmax(-1).numpy().astype(np.int8),
               obs_stoch_maxdev=np.float32((stochs.detach() - torch.nn.functional.one_hot(stochs.detach().argmax(-1), c.K)).abs().max()),
               obs_deter=deters.detach().numpy(), obs_logit=logits.detach().numpy())
    # batched prior + kl (dreamer.py:485-486) + entropy metrics (dreamer.py:575-576)
    up = O.clamp_u(np.random.Generator(np.random.Philox(11)).random((B, T, c.S, c.K), dtype=np.float32))
    U_QUEUE[:] = [t(up)]
    pst, plog = R.prior(deters)
    dyn, rep = R.kl_loss(logits, plog, 1.0)
    out.update(prior_logit=plog.detach().numpy(), prior_idx=pst.detach().argmax(-1).numpy().astype(np.int8),
               kl_dyn=dyn.detach().numpy(), kl_rep=rep.detach().numpy(),
               ent_post=R.get_dist(logits).entropy().detach().numpy(),
               ent_prior=R.get_dist(plog).entropy().detach().numpy())
    # backward through observe + prior + kl with seeded cotangents (autograd == ground truth for K2)
    g = np.random.Generator(np.random.Philox(13))
    c_st = g.standard_normal(stochs.shape, dtype=np.float32)
    c_dt = g.standard_normal(deters.shape, dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal(logits.shape, dtype=np.float32) * np.float32(0.1)
    loss = (stochs * t(c_st)).sum() + (deters * t(c_dt)).sum() + (logits * t(c_lg)).sum()
    params = dict(R.named_parameters())
    grads = torch.autograd.grad(loss, [emb_t, is_t, id_t] + list(params.values()), allow_unused=True)
    out.update(bwd_d_embed=grads[0].numpy(), bwd_d_init_stoch=grads[1].numpy(), bwd_d_init_deter=grads[2].numpy())
    for (name, _), gr in zip(params.items(), grads[3:]):
        gnp = np.zeros(P["rssm"][name].shape, np.float32) if gr is None else gr.numpy()
        if store_inputs or gnp.size <= 4096:
            out["bwd_g/" + name] = gnp
        else:  # full-size config: keep a strided slice + norms to stay small
            out["bwd_gs/" + name] = gnp.reshape(-1)[:: max(1, gnp.size // 2048)][:2048].copy()
        out["bwd_gn/" + name] = np.float64(np.sqrt((gnp.astype(np.float64) ** 2).sum()))

    # ---- single obs_step / img_step (rssm.py:158-187): the act() shape, reset as (B,1)
    U_QUEUE[:] = [t(u[:, 0])]
    s1, d1, l1 = R.obs_step(t(init_stoch), t(init_deter), t(action[:, 0]), t(embed[:, 0]), t(reset2[:, :1]))
    out.update(step_obs_idx=s1.argmax(-1).numpy().astype(np.int8), step_obs_deter=d1.detach().numpy(),
               step_obs_logit=l1.detach().numpy())
    U_QUEUE[:] = [t(u[:, 1])]
    s2, d2 = R.img_step(t(init_stoch), t(init_deter), t(action[:, 1]))
    out.update(step_img_idx=s2.argmax(-1).numpy().astype(np.int8), step_img_deter=d2.detach().numpy())
    # imagine_with_action (rssm.py:197-209)
    U_QUEUE[:] = [t(u[:, i]) for i in range(T)]
    ws, wd = R.imagine_with_action(t(init_stoch), t(init_deter), t(action))
    out.update(iwa_idx=ws.argmax(-1).numpy().astype(np.int8), iwa_deter=wd.detach().numpy())

    # ---- imagination (dreamer.py:673-692) with grad enabled (attack shape: dgrad-only)
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    st_t, dt_t = t(st0).requires_grad_(True), t(dt0).requires_grad_(True)
    fake = NS(_frozen_rssm=R, _frozen_actor=heads["actor"])
    for p_ in list(R.parameters()) + [q for h in heads.values() for q in h.parameters()]:
        p_.requires_grad_(False)
    imagine = dreamer.Dreamer._imagine.__wrapped__ if dreamer is not None else None
    U_QUEUE[:] = []
    EPS_QUEUE[:] = []
    for i in range(H):
        if c.act_kind == "cont":
            EPS_QUEUE.append(t(noise[:, i]))
        else:
            U_QUEUE.append(t(noise[:, i]))
        U_QUEUE.append(t(ui[:, i]))
    if imagine is not None:
        feats, acts = imagine(fake, (st_t, dt_t), H)
    else:  # literal restatement of dreamer.py:680-688 on the reference modules
        fl, al = [], []
        s_, d_ = st_t, dt_t
        for _ in range(H):
            f_ = R.get_feat(s_, d_)
            a_ = heads["actor"](f_).rsample()
            fl.append(f_); al.append(a_)
            s_, d_ = R.img_step(s_, d_, a_)
        feats, acts = torch.stack(fl, 1), torch.stack(al, 1)
    assert not U_QUEUE and not EPS_QUEUE
    out.update(imag_feat_idx=feats.detach()[..., :c.SK].reshape(N, H, c.S, c.K).argmax(-1).numpy().astype(np.int8),
               imag_deter=feats.detach()[..., c.SK:].numpy(), imag_action=acts.detach().numpy())
    # heads + lambda-return (dreamer.py:589-602)
    rew = heads["reward"](feats).mode()
    cont = heads["cont"](feats).mean
    val = heads["value"](feats).mode()
    sval = heads["slow_value"](feats).mode()
    disc = 1 - 1 / c.horizon
    weight = torch.cumprod(cont * disc, dim=1)
    if dreamer is not None:
        lam = dreamer.Dreamer._lambda_return.__wrapped__
        ret = lam(None, torch.zeros_like(cont), 1 - cont, rew, val, val, disc, c.lamb)
    else:
        ret = t(O.lambda_return(np.zeros_like(cont.detach().numpy()), 1 - cont.detach().numpy(),
                                rew.detach().numpy(), val.detach().numpy(), val.detach().numpy(), disc, c.lamb))
    out.update(imag_reward=rew.detach().numpy(), imag_cont=cont.detach().numpy(), imag_value=val.detach().numpy(),
               imag_slow_value=sval.detach().numpy(), imag_weight=weight.detach().numpy(), imag_ret=ret.detach().numpy(),
               used_dreamer_py=np.bool_(dreamer is not None))
    # dgrad-only backward through imagination + heads (attack shape, README.md:70)
    g2 = np.random.Generator(np.random.Philox(17))
    c_f = g2.standard_normal(feats.shape, dtype=np.float32) * np.float32(0.1)
    c_a = g2.standard_normal(acts.shape, dtype=np.float32)
    loss2 = (feats * t(c_f)).sum() + (acts * t(c_a)).sum()
    gs, gd = torch.autograd.grad(loss2, [st_t, dt_t])
    out.update(imag_bwd_d_stoch=gs.numpy(), imag_bwd_d_deter=gd.numpy())

    if store_inputs:
        for mod, d in P.items():
            for k, v in d.items():
                out[f"P/{mod}/{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", f"{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "dreamer.py used:", dreamer is not None)


def return_ema_inputs(case, call):
    """Seeded return tensors for the ReturnEMA goldens (shape as dreamer.py:600-602: (N, H-1, 1)); shared with the tests."""
    n_rows, scale, shift, ties = [(1024, 3.0, 1.0, False), (7, 1.0, 0.0, False), (1, 1.0, 2.0, False), (333, 50.0, -20.0, True),
                                  (9001, 0.01, 0.0, False)][case]
    rng = np.random.Generator(np.random.Philox(900 + 17 * case + call))
    x = (rng.standard_normal((n_rows, 15, 1), dtype=np.float32) * np.float32(scale) + np.float32(shift + 0.5 * call)).astype(np.float32)
    if ties:
        x = np.round(x)          # many equal values
    return x


def kl_grad_inputs():
    """Seeded (R, S, K) posterior / prior logits and upstream row gradients for the kl_loss backward goldens; the scale
    makes some rows fall under free = 1 (clipped: zero gradient) and some above.  Shared with the tests."""
    rng = np.random.Generator(np.random.Philox(4242))
    R, S, K = 12, 32, 16
    post = rng.standard_normal((R, S, K), dtype=np.float32) * np.linspace(0.05, 1.5, R, dtype=np.float32)[:, None, None]
    prior = rng.standard_normal((R, S, K), dtype=np.float32) * np.linspace(0.05, 1.5, R, dtype=np.float32)[:, None, None]
    g_dyn = rng.standard_normal(R, dtype=np.float32)
    g_rep = rng.standard_normal(R, dtype=np.float32)
    return post.astype(np.float32), prior.astype(np.float32), g_dyn, g_rep


def run_kl_grad(rssm_mod):
    """Autograd of the reference RSSM.kl_loss (rssm.py:222-230) on seeded logits."""
    c = O.Cfg()
    R = rssm_mod.RSSM(rssm_cfg(c), c.E, c.A)
    post, prior, g_dyn, g_rep = kl_grad_inputs()
    a, b = t(post).requires_grad_(True), t(prior).requires_grad_(True)
    dyn, rep = R.kl_loss(a, b, 1.0)
    loss = (dyn * t(g_dyn)).sum() + (rep * t(g_rep)).sum()
    da, db = torch.autograd.grad(loss, [a, b])
    path = os.path.join(ROOT, "tests", "golden", "kl_grad.npz")
    np.savez_compressed(path, dyn=dyn.detach().numpy(), rep=rep.detach().numpy(), d_post=da.numpy(), d_prior=db.numpy())
    print("kl_grad ->", path, "rows clipped:", int((dyn.detach() <= 1.0).sum().item()), "of", post.shape[0])


def twohot_inputs(bins):
    """Seeded logits (R, 255), targets and upstream gradients for the TwoHot.log_prob goldens: generic targets, exact bin
    hits, values beyond both ends of the bin range.  Shared with the tests (restated there)."""
    rng = np.random.Generator(np.random.Philox(777))
    R, n = 96, len(bins)
    logits = (rng.standard_normal((R, n), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    target = (rng.standard_normal(R, dtype=np.float32) * np.float32(30.0)).astype(np.float32)
    target[:8] = bins[[0, 1, n // 2, n // 2 + 1, n - 2, n - 1, 17, 200]]          # exact hits
    target[8:12] = np.array([-1e9, 1e9, bins[0] * 2, bins[-1] * 2], np.float32)   # out of range
    target[12:16] = np.array([0.0, 1e-6, -1e-6, 0.5], np.float32)
    g = rng.standard_normal(R, dtype=np.float32)
    return logits, target.astype(np.float32), g


def run_twohot(dists):
    """TwoHot.log_prob (distributions.py:100-129) built by symexp_twohot (distributions.py:242-251), with autograd."""
    bins = dists.symexp_twohot(torch.zeros(1, 255), 255).bins.numpy().copy()   # the reference's own bin positions (torch
    np.testing.assert_allclose(O.twohot_bins(255), bins, rtol=4e-6)             # expm1; numpy's differs by ~1 ulp)
    logits, target, g = twohot_inputs(bins)
    lg = t(logits).requires_grad_(True)
    dist = dists.symexp_twohot(lg, 255)
    lp = dist.log_prob(t(target)[..., None])
    (dlg,) = torch.autograd.grad((lp * t(g)).sum(), [lg])
    path = os.path.join(ROOT, "tests", "golden", "twohot_logprob.npz")
    np.savez_compressed(path, bins=bins, log_prob=lp.detach().numpy(), d_logits=dlg.numpy(), mode=dist.mode().detach().numpy())
    print("twohot_logprob ->", path)


OPT_SHAPES = [(48, 200), (256,), (8, 12, 8), (12, 64), (1,), (255, 16), (3, 5, 7)]


def optim_inputs(step):
    """Seeded parameters (step < 0) / gradients of step `step` for the optimiser goldens; tensor 1's gradients are large
    (AGC clips), tensor 4 is a single element, tensor 6 has a tiny parameter norm (pmin floor).  Shared with the tests."""
    rng = np.random.Generator(np.random.Philox(31337 + (step if step >= 0 else 1000)))
    out = []
    for i, shp in enumerate(OPT_SHAPES):
        x = rng.standard_normal(shp, dtype=np.float32)
        if step < 0:
            x = x * np.float32(1e-5 if i == 6 else 0.05)
        else:
            x = x * np.float32([1e-3, 5.0, 1e-2, 1e-4, 0.3, 2e-2, 1e-3][i])
        out.append(x.astype(np.float32))
    return out


def run_optim():
    """clip_grad_agc_ (utils/optim/agc.py) + LaProp.step (utils/optim/laprop.py) for three steps on seeded tensors."""
    import utils.optim.laprop as laprop
    import utils.optim.agc as agc
    params = [torch.nn.Parameter(t(x.copy())) for x in optim_inputs(-1)]
    opt = laprop.LaProp(params, lr=4e-5, betas=(0.9, 0.999), eps=1e-20)
    out = {}
    for step in range(3):
        for p_, g_ in zip(params, optim_inputs(step)):
            p_.grad = t(g_.copy())
        agc.clip_grad_agc_(params, 0.3, 1e-3, foreach=True)
        for i, p_ in enumerate(params):
            out[f"s{step}_g{i}"] = p_.grad.detach().numpy().copy()
        opt.step()
        for i, p_ in enumerate(params):
            out[f"s{step}_p{i}"] = p_.detach().numpy().copy()
            out[f"s{step}_m{i}"] = opt.state[p_]["exp_avg"].numpy().copy()
            out[f"s{step}_v{i}"] = opt.state[p_]["exp_avg_sq"].numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "optim.npz")
    np.savez_compressed(path, **out)
    print("optim ->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


def barlow_inputs():
    """Seeded projected latents x1 / embeddings x2 (N, E) for the Barlow goldens (correlated, so diag(c) is not ~0)."""
    rng = np.random.Generator(np.random.Philox(2718))
    N, E = 64, 48
    z = rng.standard_normal((N, E), dtype=np.float32)
    x1 = (z * np.float32(1.5) + np.float32(0.3) * rng.standard_normal((N, E), dtype=np.float32) + np.float32(0.7)).astype(np.float32)
    x2 = (z @ (np.eye(E, dtype=np.float32) + np.float32(0.1) * rng.standard_normal((E, E), dtype=np.float32))
          + np.float32(0.5) * rng.standard_normal((N, E), dtype=np.float32)).astype(np.float32)
    return x1, x2


def run_barlow():
    """The Barlow block of Dreamer._cal_grad is inline code (dreamer.py:525-532), not a callable: those source lines are read
    from the reference checkout and executed as they are (dedented) on seeded tensors under torch autograd, with x2 detached
    as at :522 and `self.barlow_lambd` = configs/base.yaml:214."""
    import textwrap
    x1_np, x2_np = barlow_inputs()
    x1 = t(x1_np).requires_grad_(True)
    x2 = t(x2_np).detach()
    src = open(os.path.join(REF, "world_model", "dreamer.py")).read().splitlines()[524:532]
    assert src[0].strip().startswith("x1_norm") and src[-1].strip().startswith('losses["barlow"]'), src
    ns = {"torch": torch, "x1": x1, "x2": x2, "B": x1.shape[0], "T": 1, "self": NS(barlow_lambd=5e-4), "losses": {}}
    exec(textwrap.dedent("\n".join(src)), ns)
    loss = ns["losses"]["barlow"]
    (dx1,) = torch.autograd.grad(loss, [x1])
    path = os.path.join(ROOT, "tests", "golden", "barlow.npz")
    np.savez_compressed(path, loss=np.float32(loss.item()), d_x1=dx1.numpy(), lambd=np.float32(5e-4))
    print("barlow ->", path, "loss", float(loss))


def latent_store_inputs():
    """Seeded inputs for the replay latent write-back (restated in tests/test_latent_store.py)."""
    rng = np.random.Generator(np.random.Philox(4242))
    B, T, S, K, D, n_len, n_env = 3, 5, 4, 8, 12, 20, 4
    # B slices of T consecutive steps, distinct (time, env) slots (unique => index_put_ is deterministic)
    env = np.array([0, 2, 3], np.int64)[:, None].repeat(T, 1)
    t0 = np.array([1, 7, 13], np.int64)[:, None]
    time = t0 + np.arange(T, dtype=np.int64)[None]
    cls = rng.integers(0, K, size=(B, T, S))
    stoch = np.eye(K, dtype=np.float32)[cls]
    deter = rng.standard_normal((B, T, D), dtype=np.float32)
    store_stoch = rng.standard_normal((n_len, n_env, S, K), dtype=np.float32)
    store_deter = rng.standard_normal((n_len, n_env, D), dtype=np.float32)
    return env, time, stoch, deter, store_stoch, store_deter


def run_latent_store():
    """Buffer.update (utils/buffer.py:44-53) executed from the reference checkout.  torchrl / tensordict (the storage the
    reference delegates to; requirements pin torchrl, absent from this image) are replaced by a minimal stand-in whose
    `storage[i0, i1].set_(key, value)` is torch's index assignment on plain tensors -- so the reference's own lines
    (flattening, index ORDER index[1] -> dim 0, index[0] -> dim 1) are what runs."""
    import types
    for name in ("torchrl", "torchrl.data", "torchrl.data.replay_buffers", "torchrl.data.replay_buffers.samplers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchrl.data.replay_buffers"].LazyTensorStorage = object
    sys.modules["torchrl.data.replay_buffers"].ReplayBuffer = object
    sys.modules["torchrl.data.replay_buffers.samplers"].SliceSampler = object
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_buffer", os.path.join(REF, "utils", "buffer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class View:
        def __init__(self, store, key):
            self.store, self.key = store, key

        def set_(self, name, value):
            self.store.t[name][self.key] = value

    class Store:
        def __init__(self, **t):
            self.t = t

        def __getitem__(self, key):
            return View(self, key)

    env, time, stoch, deter, store_stoch, store_deter = latent_store_inputs()
    st = Store(stoch=t(store_stoch).clone(), deter=t(store_deter).clone())
    holder = types.SimpleNamespace(_buffer=st)
    mod.Buffer.update(holder, [t(env), t(time)], t(stoch), t(deter))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "latent_store.npz"), store_stoch=st.t["stoch"].numpy(), store_deter=st.t["deter"].numpy())
    print("wrote latent_store.npz")


def cnn_encoder_inputs(hw, n):
    """Seeded image batch (B, T, H, W, 3) in [0, 1] and an upstream gradient for the embedding (restated in tests/test_cnn_oracle.py)."""
    rng = np.random.Generator(np.random.Philox(5150 + hw))
    obs = rng.random((n, 2, hw, hw, 3), dtype=np.float32)
    return obs, rng


def run_cnn_encoder(networks):
    """ConvEncoder (networks.py:192-234) of the reference, fp32 on CPU: forward embedding and autograd gradients for a tiny
    (32x32, depth 4) and the base (64x64, depth 16, mults 2/3/4/4 -> E = 1024) configuration."""
    from types import SimpleNamespace as NS
    from oracle import cnn_oracle as CO
    out = {}
    for tag, hw, depth, n in (("tiny", 32, 4, 2), ("base", 64, 16, 1)):
        mults = (2, 3, 4, 4)
        cfg = NS(act="SiLU", norm=True, kernel_size=5, minres=4, depth=depth, mults=list(mults))
        enc = networks.ConvEncoder(cfg, (hw, hw, 3))
        P = CO.encoder_params([depth * m for m in mults], 3, 5, seed=77 + hw)
        enc.load_state_dict({k: t(v) for k, v in P.items()})
        obs, rng = cnn_encoder_inputs(hw, n)
        x = t(obs).requires_grad_(True)
        emb = enc(x)
        g = rng.standard_normal(tuple(emb.shape), dtype=np.float32)
        emb.backward(t(g))
        out[f"{tag}/emb"] = emb.detach().numpy()
        out[f"{tag}/g"] = g
        out[f"{tag}/d_obs"] = x.grad.numpy()
        for k, p_ in enc.named_parameters():
            out[f"{tag}/grad/{k}"] = p_.grad.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cnn_encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("emb")})


def run_return_ema(networks):
    """ReturnEMA (networks.py:405-422), four consecutive calls per case (the buffer carries over)."""
    out = {}
    for case in range(5):
        ema = networks.ReturnEMA(device="cpu")
        for call in range(4):
            x = return_ema_inputs(case, call)
            off, scl = ema(t(x))
            out[f"c{case}_{call}_ema"] = ema.ema_vals.numpy().copy()
            out[f"c{case}_{call}_offset"] = np.float32(off.item())
            out[f"c{case}_{call}_scale"] = np.float32(scl.item())
    path = os.path.join(ROOT, "tests", "golden", "return_ema.npz")
    np.savez_compressed(path, **out)
    print("return_ema ->", path)


def run_attack(tag, c, N, H, rssm_mod, dists, networks, dreamer):
    """The adversarial-patch attack's gradient path after the posterior (README.md:68-116): grad-enabled
    Dreamer._imagine body -> frozen reward / cont / value heads -> Dreamer._lambda_return body (differentiated) -> a seeded
    linear objective over ret; autograd gives d(objective)/d(feats) and d/d(start state)."""
    torch.manual_seed(0)
    P = O.init_params(c, seed=0)
    R, heads = build_reference(c, P, rssm_mod, networks)
    for p_ in list(R.parameters()) + [q for h in heads.values() for q in h.parameters()]:
        p_.requires_grad_(False)
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    st_t, dt_t = t(st0).requires_grad_(True), t(dt0).requires_grad_(True)
    fake = NS(_frozen_rssm=R, _frozen_actor=heads["actor"])
    U_QUEUE[:] = []
    EPS_QUEUE[:] = []
    for i in range(H):
        if c.act_kind == "cont":
            EPS_QUEUE.append(t(noise[:, i]))
        else:
            U_QUEUE.append(t(noise[:, i]))
        U_QUEUE.append(t(ui[:, i]))
    feats, acts = dreamer.Dreamer._imagine.__wrapped__(fake, (st_t, dt_t), H)
    assert not U_QUEUE and not EPS_QUEUE
    feats.retain_grad()
    rew = heads["reward"](feats).mode()
    cont = heads["cont"](feats).mean
    val = heads["value"](feats).mode()
    disc = 1 - 1 / c.horizon
    ret = dreamer.Dreamer._lambda_return.__wrapped__(None, torch.zeros_like(cont), 1 - cont, rew, val, val, disc, c.lamb)
    g = np.random.Generator(np.random.Philox(19))
    c_ret = g.standard_normal(tuple(ret.shape), dtype=np.float32)
    c_rew = g.standard_normal(tuple(rew.shape), dtype=np.float32) * np.float32(0.1)
    loss = (ret * t(c_ret)).sum() + (rew * t(c_rew)).sum()
    loss.backward()
    out = dict(N=N, H=H, feats=feats.detach().numpy(), ret=ret.detach().numpy(), c_ret=c_ret, c_rew=c_rew,
               d_feats=feats.grad.numpy(), d_stoch=st_t.grad.numpy(), d_deter=dt_t.grad.numpy(), loss=np.float64(loss.item()))
    path = os.path.join(ROOT, "tests", "golden", f"attack_{tag}.npz")
    np.savez_compressed(path, **out)
    print("attack", tag, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB  loss {loss.item():.6f}")


def run_attacks(rssm_mod, dists, networks, dreamer):
    patch_noise(dists)
    tiny = dict(D=256, U=64, S=8, K=8, G=4, E=48, units=64)
    run_attack("tiny_cont", O.Cfg(A=3, **tiny), 6, 5, rssm_mod, dists, networks, dreamer)
    run_attack("base_cont", O.Cfg(), 4, 16, rssm_mod, dists, networks, dreamer)


def main():
    rssm_mod, dists, networks, dreamer = import_reference()
    if "--c1-k32-only" in sys.argv:
        patch_noise(dists)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        run_case("base_e256", O.Cfg(E=256), 2, 5, 4, 3, rssm_mod, dists, networks, dreamer, False)
        run_case("base_k32", O.Cfg(K=32), 2, 4, 3, 3, rssm_mod, dists, networks, dreamer, False)
        return
    if "--attack-only" in sys.argv:
        run_attacks(rssm_mod, dists, networks, dreamer)
        return
    if "--return-ema-only" in sys.argv:
        run_return_ema(networks)
        return
    if "--barlow-only" in sys.argv:
        run_barlow()
        return
    if "--latent-store-only" in sys.argv:
        run_latent_store()
        return
    if "--cnn-only" in sys.argv:
        run_cnn_encoder(networks)
        return
    if "--optim-only" in sys.argv:
        run_optim()
        return
    if "--twohot-only" in sys.argv:
        run_twohot(dists)
        return
    if "--kl-grad-only" in sys.argv:
        run_kl_grad(rssm_mod)
        return
    patch_noise(dists)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    tiny = dict(D=256, U=64, S=8, K=8, G=4, E=48, units=64)
    run_case("tiny_cont", O.Cfg(A=3, **tiny), 3, 6, 5, 4, rssm_mod, dists, networks, dreamer, True)
    run_case("tiny_onehot", O.Cfg(A=5, act_kind="onehot", **tiny), 3, 6, 5, 4, rssm_mod, dists, networks, dreamer, True)
    run_case("base_cont", O.Cfg(), 2, 5, 4, 3, rssm_mod, dists, networks, dreamer, False)
    run_case("base_onehot18", O.Cfg(A=18, act_kind="onehot"), 2, 3, 3, 3, rssm_mod, dists, networks, dreamer, False)
    run_case("base_e256", O.Cfg(E=256), 2, 5, 4, 3, rssm_mod, dists, networks, dreamer, False)   # config C1: proprio embed size
    run_case("base_k32", O.Cfg(K=32), 2, 4, 3, 3, rssm_mod, dists, networks, dreamer, False)     # 32 classes at base widths
    run_return_ema(networks)
    run_kl_grad(rssm_mod)
    run_twohot(dists)
    run_optim()
    run_barlow()
    run_latent_store()
    run_cnn_encoder(networks)
    run_attacks(rssm_mod, dists, networks, dreamer)


if __name__ == "__main__":
    main()
