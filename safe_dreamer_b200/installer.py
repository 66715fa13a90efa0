"""install(agent): swap the RSSM hot path of a *reference* `Dreamer` instance (world_model/dreamer.py) for this library,
in place, without editing dreamer.py.

What is replaced (everything else -- encoder, decoder / projector, heads with gradients, losses, logging -- stays the
reference's own code):
  agent.rssm / agent._frozen_rssm   -> safe_dreamer_b200.rssm.RSSM adopting the SAME Parameter objects (the reference's
                                       optimizer, `_named_params`, checkpoints and clone_and_freeze keep working)
                                       observe / obs_step / prior / kl_loss / get_feat / get_dist / initial   dreamer.py:348,361,383,483-486,575
  agent._imagine                    -> the CUDA rollout with the frozen actor in the loop                      dreamer.py:585,673-692
  agent._lambda_return              -> sd_lambda_return (imagined H = 16 and replay T = 64 calls)              dreamer.py:600,651,694-707
  agent._frozen_reward/_cont/_value/_slow_value (eager mode only) -> proxies that evaluate all four heads, the
                                       cumprod weights and the imagined lambda-return in one fused call          dreamer.py:589-602
  agent.return_ema                  -> sd_return_ema (exact quantiles + EMA in one launch)                     networks.py:405-422
  agent._optimizer / agent._agc     -> fused multi-tensor AGC + LaProp (optional)                              dreamer.py:219-224,432-433
  agent.encoder.encoders[i] (ConvEncoder; optional, `encoder=True`, eager mode) -> safe_dreamer_b200.encoder.ConvEncoder
                                       adopting the same Parameter objects: tcgen05 implicit-GEMM forward / backward  networks.py:192-234
With `custom_ops=True` (default when the agent was built with config.compile) the RSSM calls go through the
torch.library operators of `safe_dreamer_b200.ops`, which torch.compile(mode="reduce-overhead") traces as opaque nodes.
"""
from __future__ import annotations

import types
from types import SimpleNamespace as NS

import torch

from . import dreamer_ops
from .rssm import RSSM


def _adopt_parameters(dst, src):
    """Make `dst` hold the very Parameter objects of `src` (same names / shapes: tests/test_abi_symbols.py)."""
    sp = dict(src.named_parameters())
    dp = dict(dst.named_parameters())
    if set(sp) != set(dp):
        raise RuntimeError(f"state_dict mismatch: {sorted(set(sp) ^ set(dp))}")
    for name, p in sp.items():
        if tuple(p.shape) != tuple(dp[name].shape):
            raise RuntimeError(f"shape mismatch for {name}: {tuple(p.shape)} vs {tuple(dp[name].shape)}")
        mod = dst
        parts = name.split(".")
        for part in parts[:-1]:
            mod = getattr(mod, part)
        mod._parameters[parts[-1]] = p


def mirror_conv_encoder(ref, input_shape):
    """A safe_dreamer_b200 ConvEncoder sharing the parameters of the reference ConvEncoder `ref` (networks.py:192-234)."""
    from .encoder import ConvEncoder
    n = len(ref.depths)
    if len(ref.layers) != 4 * n or not isinstance(ref.layers[3], torch.nn.SiLU):
        raise NotImplementedError("install(encoder=True): only conv -> maxpool -> RMSNorm2D -> SiLU stages (configs/base.yaml) have kernels")
    cfg = NS(act="SiLU", norm=True, depth=1, mults=list(ref.depths), kernel_size=ref.kernel_size)
    new = ConvEncoder(cfg, tuple(int(v) for v in input_shape)).to(next(ref.parameters()).device)
    _adopt_parameters(new, ref)
    return new


def mirror_rssm(ref, embed_size):
    """A safe_dreamer_b200 RSSM sharing the parameters of the reference RSSM `ref`."""
    dev = next(ref.parameters()).device
    cfg = NS(stoch=ref._stoch, deter=ref._deter, hidden=ref._hidden, discrete=ref._discrete, act="SiLU",
             unimix_ratio=ref._unimix_ratio, initial=ref._initial, device=str(dev), obs_layers=ref._obs_layers,
             img_layers=ref._img_layers, dyn_layers=ref._dyn_layers, blocks=ref._blocks)
    new = RSSM(cfg, embed_size, ref._act_dim).to(dev)
    _adopt_parameters(new, ref)
    return new


class _HeadProxy:
    """Stands in for one frozen head on IMAGINED features: `proxy(feat).mode()` / `.mean` return what the fused
    sd_heads_lambda_fwd call computed for all four heads (run once per distinct feat tensor).  It lives in the agent's
    instance __dict__ and shadows the real frozen module, which stays registered (state_dict keys are unchanged)."""

    def __init__(self, bundle, which, module):
        self.bundle, self.which, self.module = bundle, which, module

    def __call__(self, feat):
        b = self.bundle
        if feat.dim() != 3 or feat.requires_grad or feat.shape[1] != b.agent.imag_horizon + 1 or not feat.is_cuda:
            return self.module(feat)        # replay features (dreamer.py:650-651) keep the reference head
        vals = b.evaluate(feat)[self.which]
        return NS(mode=lambda: vals, mean=vals)

    def __getattr__(self, name):            # parameters(), named_parameters(), ... of the wrapped module
        return getattr(self.__dict__["module"], name)


class _FusedHeads:
    def __init__(self, agent):
        self.agent = agent
        self._key, self._out = None, None

    def evaluate(self, feat):
        key = (feat.data_ptr(), feat._version, tuple(feat.shape))
        if key != self._key:
            a = self.agent
            rew, cont, val, sval, wgt, ret = dreamer_ops.heads_lambda(a._frozen_rssm, feat, a.horizon, a.lamb)
            self._out = {"reward": rew, "cont": cont, "value": val, "slow_value": sval, "weight": wgt, "ret": ret}
            self._key = key
        return self._out

    def lambda_return(self, last, term, reward, value, boot, disc, lamb):
        """Dreamer._lambda_return (dreamer.py:694-707).  The imagined call (reward / value are the fused call's outputs)
        returns the lambda-return the fused kernel already produced; any other call runs sd_lambda_return."""
        a = self.agent
        o = self._out
        if o is not None and value is o["value"] and boot is o["value"] and reward is o["reward"]:
            return o["ret"]
        fr = a._frozen_rssm
        with torch.no_grad():       # the reference's _lambda_return is @torch.no_grad() (dreamer.py:694)
            args = [x.detach().float() for x in (last, term, reward, value, boot)]
            if getattr(fr, "use_custom_ops", False):
                return torch.ops.safedreamer.lambda_return(*args, float(disc), float(lamb), fr._ops_key)
            return dreamer_ops.lambda_return(fr, *args, disc, lamb)


_FROZEN = ("reward", "cont", "value", "slow_value")


def _post_clone(agent, opt):
    """After the reference's clone_and_freeze (dreamer.py:260-322): configure the frozen RSSM mirror, hand it the frozen
    heads (its engine packs their weights) and, in eager mode, shadow the four frozen heads by the fused proxies."""
    fr = agent._frozen_rssm
    fr.precision = opt.imagine_precision
    fr.use_graph = opt.use_graph
    fr.use_custom_ops = opt.custom_ops
    # the frozen copies alias the live parameters' storage through fresh tensors (param_new.data = param_orig.data,
    # dreamer.py:279): their version counters never move, so the packed copies are refreshed on every call
    fr.auto_refresh = True
    from .encoder import ConvEncoder
    for enc in getattr(getattr(agent, "_frozen_encoder", None), "encoders", []):
        if isinstance(enc, ConvEncoder):
            enc.auto_refresh = True      # same aliasing: the frozen encoder behind Dreamer.act (dreamer.py:341)
            if enc.use_custom_ops:
                from . import ops
                enc._ops_key = ops.module_key(enc)
    mods = {nm: agent._modules[f"_frozen_{nm}"] for nm in _FROZEN}
    kw = dict(getattr(getattr(agent._frozen_actor, "_dist", None), "keywords", {}) or {})
    dreamer_ops.attach_heads(fr, actor=agent._frozen_actor, reward=mods["reward"], cont=mods["cont"], value=mods["value"],
                             slow_value=mods["slow_value"], act_kind="onehot" if agent.act_discrete else "cont",
                             min_std=kw.get("min_std", 0.1), max_std=kw.get("max_std", 1.0),
                             act_unimix=kw.get("unimix_ratio", 0.01), bins=mods["reward"].last.out_features)
    if opt.custom_ops:
        from . import ops
        agent.rssm._ops_key = ops.module_key(agent.rssm)
        fr._ops_key = ops.module_key(fr)
    agent.__dict__["_fused"] = _FusedHeads(agent)
    if opt.fuse_heads:
        for nm in _FROZEN:      # instance __dict__ wins over nn.Module's registered children on attribute lookup
            agent.__dict__[f"_frozen_{nm}"] = _HeadProxy(agent._fused, nm, mods[nm])


def install(agent, precision="fp32", imagine_precision="bf16", fuse_heads=None, optimizer=False, return_ema=True,
            custom_ops=None, use_graph=True, encoder=False):
    """Swap the hot path of the reference `agent` (see module docstring).  Returns the agent.
    precision: "fp32" (3xTF32, parity with the fp32 reference) or "bf16" for the posterior path;
    imagine_precision: precision of the (no-grad) imagination rollout and the fused heads."""
    compiled = hasattr(agent._cal_grad, "_torchdynamo_orig_callable")
    if custom_ops is None:
        custom_ops = compiled
    if fuse_heads is None:
        fuse_heads = not custom_ops      # the proxies key on tensor identity: eager only
    opt = NS(imagine_precision=imagine_precision, use_graph=use_graph, custom_ops=bool(custom_ops), fuse_heads=bool(fuse_heads))
    new = mirror_rssm(agent.rssm, agent.embed_size)
    new.precision = precision
    new.use_graph = use_graph
    new.use_custom_ops = bool(custom_ops)
    agent.rssm = new
    if encoder:
        encs = getattr(agent.encoder, "encoders", None)
        shapes = getattr(agent.encoder, "cnn_shapes", None)
        if encs is None or not shapes:
            raise NotImplementedError("install(encoder=True): agent.encoder is not a MultiEncoder with image keys")
        shape = tuple(shapes.values())[0][:2] + (sum(v[-1] for v in shapes.values()),)
        for i, enc in enumerate(encs):
            if type(enc).__name__ == "ConvEncoder":
                encs[i] = mirror_conv_encoder(enc, shape)
                if custom_ops:
                    from . import ops
                    encs[i].use_custom_ops = True
                    encs[i]._ops_key = ops.module_key(encs[i])
    ref_clone = type(agent).clone_and_freeze

    def clone_and_freeze(self):
        ref_clone(self)      # assigning real modules also drops any proxies from the instance __dict__
        _post_clone(self, opt)

    agent.clone_and_freeze = types.MethodType(clone_and_freeze, agent)

    def _imagine(self, start, imag_horizon):
        """Dreamer._imagine (dreamer.py:673-692) on the CUDA rollout; the frozen actor samples inside the scan."""
        stoch, deter = start
        fr = self._frozen_rssm
        N = deter.shape[0]
        hook = getattr(fr, "act_noise_source", None)
        noise = hook((N, imag_horizon, fr._act_dim), deter.device) if hook is not None else None
        if fr.use_custom_ops:
            u = fr._uniform(N, imag_horizon, fr._stoch, fr._discrete)
            if noise is None:
                noise = (torch.rand(N, imag_horizon, fr._act_dim, device=deter.device).clamp_(2.0 ** -24, 1 - 2.0 ** -24)
                         if self.act_discrete else torch.randn(N, imag_horizon, fr._act_dim, device=deter.device))
            with torch.no_grad():
                return torch.ops.safedreamer.imagine(stoch.detach().float(), deter.detach().float(), u, noise, int(imag_horizon),
                                                     fr._ops_key)
        return dreamer_ops.imagine(fr, (stoch, deter), imag_horizon, act_noise=noise)

    agent._imagine = types.MethodType(_imagine, agent)
    agent.clone_and_freeze()
    agent._lambda_return = lambda last, term, reward, value, boot, disc, lamb: agent._fused.lambda_return(
        last, term, reward, value, boot, disc, lamb)
    if return_ema and not custom_ops:
        from .networks import ReturnEMA
        ema = ReturnEMA(device=agent.device, alpha=agent.return_ema.alpha).to(agent.device)
        ema.ema_vals.copy_(agent.return_ema.ema_vals)
        agent.return_ema = ema
    if optimizer:
        from .optim import LaProp, clip_grad_agc_
        old = agent._optimizer
        g = old.param_groups[0]
        new_opt = LaProp(list(agent._named_params.values()), lr=g["lr"], betas=g["betas"], eps=g["eps"],
                         weight_decay=g.get("weight_decay", 0))
        new_opt.load_state_dict(old.state_dict())
        agent._optimizer = new_opt
        agent._scheduler.optimizer = new_opt
        clip = pmin = None
        for cell in (agent._agc.__closure__ or ()):      # the reference closes over its config (dreamer.py:208-211)
            obj = cell.cell_contents
            if hasattr(obj, "agc") and hasattr(obj, "pmin"):
                clip, pmin = float(obj.agc), float(obj.pmin)
        if clip is not None:
            agent._agc = lambda params: clip_grad_agc_(params, clip, pmin, foreach=True)
    return agent
