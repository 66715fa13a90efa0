"""Engine: one libsafedreamer handle + typed wrappers taking torch CUDA tensors.

Thin by design: shape checks, contiguity, the current CUDA stream, and the C call.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (MOD_ACTOR, MOD_CONT, MOD_REWARD, MOD_RSSM, MOD_SLOW_VALUE, MOD_VALUE, SD_FLAG_BF16,
                   SD_FLAG_GRAPH, SD_FLAG_SAVE_TAPE, sd_config)

__all__ = ["Engine", "SD_FLAG_BF16", "SD_FLAG_GRAPH", "SD_FLAG_SAVE_TAPE", "MOD_RSSM", "MOD_ACTOR", "MOD_REWARD",
           "MOD_CONT", "MOD_VALUE", "MOD_SLOW_VALUE"]


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f32c(t, name):
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the RSSM hot path has no CPU implementation)")
    return t.contiguous()


class Engine:
    """Owns one sd_handle (device workspace + packed weights) for a fixed model configuration."""

    def __init__(self, *, D, U, S, K, G, E, A, obs_layers=1, img_layers=2, act_kind=0, units=256,
                 actor_layers=3, value_layers=3, reward_layers=1, cont_layers=1, bins=255, unimix=0.01,
                 act_unimix=0.01, min_std=0.1, max_std=1.0, max_rows=1024, max_steps=64, max_tape_rows=0,
                 device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("safe_dreamer_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cfg = sd_config(D=D, U=U, S=S, K=K, G=G, E=E, A=A, obs_layers=obs_layers, img_layers=img_layers,
                             act_kind=act_kind, units=units, actor_layers=actor_layers, value_layers=value_layers,
                             reward_layers=reward_layers, cont_layers=cont_layers, bins=bins, unimix=unimix,
                             act_unimix=act_unimix, min_std=min_std, max_std=max_std, max_rows=max_rows,
                             max_steps=max_steps, max_tape_rows=max_tape_rows)
        self.SK, self.F = S * K, S * K + D
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sd_create(C.byref(self.cfg), C.byref(h)), "sd_create")
        self.h = h
        self._keep = {}
        # static_outputs=True: outputs live in engine-owned buffers that the next call of the same entry
        # point and shape overwrites.  Pointer-stable outputs keep the CUDA-graph cache (keyed on pointers)
        # hot; with fresh tensors per call the cache only hits when the allocator returns the same blocks.
        self.static_outputs = False
        self.stage_inputs = True    # static mode: stage inputs in engine-owned buffers (False: the caller's are pointer-stable)
        self._outs = {}

    @classmethod
    def from_cfg(cls, c, max_rows, max_steps, max_tape_rows=0, params=None):
        """Engine for a `synth.Cfg`-like size object; `params` = {module: {state_dict name: ndarray}} to load."""
        eng = cls(D=c.D, U=c.U, S=c.S, K=c.K, G=c.G, E=c.E, A=c.A, obs_layers=c.obs_layers, img_layers=c.img_layers,
                  act_kind=0 if c.act_kind == "cont" else 1, units=c.units, actor_layers=c.actor_layers,
                  value_layers=c.value_layers, reward_layers=c.reward_layers, cont_layers=c.cont_layers, bins=c.bins,
                  unimix=c.unimix, act_unimix=c.act_unimix, min_std=c.min_std, max_std=c.max_std, max_rows=max_rows,
                  max_steps=max_steps, max_tape_rows=max_tape_rows)
        if params is not None:
            for mod, key in enumerate(["rssm", "actor", "reward", "cont", "value", "slow_value"]):
                eng.set_weights(mod, {k: torch.from_numpy(v).to(eng.device) for k, v in params[key].items()})
            torch.cuda.synchronize(eng.device)
        return eng

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.sd_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ helpers
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace_bytes(self):
        return int(self.lib.sd_workspace_bytes(C.byref(self.cfg)))

    def weight_names(self, module):
        return list(self._weight_meta(module)[0])

    def _weight_meta(self, module):
        """(names, numels) of a module's tensors; cached (set_weights runs once per training step)."""
        cache = self.__dict__.setdefault("_wmeta", {})
        if module not in cache:
            n = self.lib.sd_weight_count(self.h, module)
            cache[module] = ([self.lib.sd_weight_name(self.h, module, i).decode() for i in range(n)],
                             [int(self.lib.sd_weight_numel(self.h, module, i)) for i in range(n)])
        return cache[module]

    def set_weights(self, module, named):
        """named: mapping state_dict-name -> fp32 CUDA tensor (reference layouts)."""
        names, numels = self._weight_meta(module)
        # fast path (once per training step): the same parameter tensors as last time, updated in place by the optimizer
        fast = self.__dict__.setdefault("_wfast", {}).get(module)
        if fast is not None:
            srcs, ptrs, arr = fast
            if len(srcs) == len(names) and all(named.get(n) is s_ and s_.data_ptr() == p_ for n, s_, p_ in zip(names, srcs, ptrs)):
                _lib.check(self.lib.sd_set_weights(self.h, module, arr, len(srcs), self.stream), "sd_set_weights")
                return
        ts = []
        for n, want in zip(names, numels):
            if n not in named:
                raise KeyError(f"set_weights(module={module}): missing tensor '{n}'")
            t = _f32c(named[n].detach(), n)
            if t.numel() != want:
                raise ValueError(f"set_weights: '{n}' has {t.numel()} elements, expected {want}")
            ts.append(t)
        arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        _lib.check(self.lib.sd_set_weights(self.h, module, arr, len(ts), self.stream), "sd_set_weights")
        self._keep[module] = ts  # keep sources alive until the async repack has been enqueued and run
        srcs = [named[n] for n in names]
        if all(t.data_ptr() == s_.data_ptr() for t, s_ in zip(ts, srcs)):   # no dtype / layout conversion happened
            self._wfast[module] = (srcs, [t.data_ptr() for t in ts], arr)
        else:
            self._wfast.pop(module, None)

    def _stage(self, t, tag):
        """static mode: copy an input into an engine-owned buffer so the pointers seen by the C ABI (and the
        CUDA-graph cache keyed on them) never change between calls."""
        if not self.static_outputs or not self.stage_inputs or t is None:
            return t
        key = ("in_" + tag, tuple(t.shape), t.dtype)
        buf = self._outs.get(key)
        if buf is None:
            buf = self._outs[key] = torch.empty_like(t)
        if buf.data_ptr() != t.data_ptr():
            buf.copy_(t)
        return buf

    def _new(self, *shape, tag=None):
        if self.static_outputs and tag is not None:
            key = (tag, tuple(shape))
            t = self._outs.get(key)
            if t is None:
                t = self._outs[key] = torch.empty(*shape, dtype=torch.float32, device=self.device)
            return t
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ entry points
    # The library keeps ONE observe/imagine tape and ONE prior tape per handle (single slot).  `tape_gen` counts the taped
    # forwards so that a backward can prove the tape it is about to consume is still the one its forward wrote.
    tape_gen = {"scan": 0, "prior": 0}

    def _bump(self, which):
        if self.tape_gen is Engine.tape_gen:
            self.tape_gen = {"scan": 0, "prior": 0}
        self.tape_gen[which] += 1
        return self.tape_gen[which]

    def check_tape(self, which, gen, who):
        if self.tape_gen.get(which, 0) != gen:
            raise RuntimeError(
                f"{who}: the {which} tape of this engine was overwritten by a later taped forward (generation {gen} -> "
                f"{self.tape_gen.get(which, 0)}).  The library keeps one tape per handle: run each backward before the next "
                f"grad-enabled forward of the same kind, or use a second RSSM module / engine for the second graph.")

    def observe(self, embed, action, init_stoch, init_deter, is_first, u, flags=0, out=None):
        B, T = action.shape[:2]
        embed, action, u = _f32c(embed, "embed"), _f32c(action, "action"), _f32c(u, "u")
        init_stoch, init_deter = _f32c(init_stoch, "init_stoch"), _f32c(init_deter, "init_deter")
        first = is_first.reshape(B, T).to(torch.uint8).contiguous()
        embed, action, u, first = (self._stage(embed, "emb"), self._stage(action, "act"), self._stage(u, "u"),
                                   self._stage(first, "first"))
        init_stoch, init_deter = self._stage(init_stoch, "is"), self._stage(init_deter, "id")
        c = self.cfg
        assert embed.shape == (B, T, c.E) and action.shape == (B, T, c.A) and u.numel() == B * T * self.SK
        if out is not None:
            stochs, deters, logits = out
        else:
            stochs, deters, logits = (self._new(B, T, c.S, c.K, tag="obs_s"), self._new(B, T, c.D, tag="obs_d"),
                                      self._new(B, T, c.S, c.K, tag="obs_l"))
        _lib.check(self.lib.sd_observe_fwd(self.h, B, T, _ptr(embed), _ptr(action), _ptr(init_stoch), _ptr(init_deter),
                                           _ptr(first), _ptr(u), _ptr(stochs), _ptr(deters), _ptr(logits), flags,
                                           self.stream), "sd_observe_fwd")
        if flags & SD_FLAG_SAVE_TAPE:
            self._bump("scan")
        return stochs, deters, logits

    def observe_bwd(self, B, T, d_stochs, d_deters, d_logits, want_embed=True, want_init=True, weight_grads=None,
                    flags=0):
        c = self.cfg
        ds = None if d_stochs is None else _f32c(d_stochs, "d_stochs")
        dd = None if d_deters is None else _f32c(d_deters, "d_deters")
        dl = None if d_logits is None else _f32c(d_logits, "d_logits")
        ds, dd, dl = self._stage(ds, "ds"), self._stage(dd, "dd"), self._stage(dl, "dl")
        d_embed = self._new(B, T, c.E, tag="ob_de") if want_embed else None
        d_is = self._new(B, c.S, c.K, tag="ob_dis") if want_init else None
        d_id = self._new(B, c.D, tag="ob_did") if want_init else None
        if weight_grads is not None:
            names = self.weight_names(MOD_RSSM)
            arr = (C.c_void_p * len(names))(*[0 if weight_grads.get(n) is None else weight_grads[n].data_ptr()
                                              for n in names])
        else:
            arr = None
        _lib.check(self.lib.sd_observe_bwd(self.h, B, T, _ptr(ds), _ptr(dd), _ptr(dl), _ptr(d_embed), _ptr(d_is),
                                           _ptr(d_id), arr, flags, self.stream), "sd_observe_bwd")
        return d_embed, d_is, d_id

    def prior(self, deter, u, flags=0):
        c = self.cfg
        lead = deter.shape[:-1]
        deter, u = _f32c(deter, "deter").reshape(-1, c.D), _f32c(u, "u")
        R = deter.shape[0]
        assert u.numel() == R * self.SK
        deter, u = self._stage(deter, "pr_d"), self._stage(u, "pr_u")
        stoch, logit = self._new(R, c.S, c.K, tag="pr_s"), self._new(R, c.S, c.K, tag="pr_l")
        if flags & SD_FLAG_BF16:
            self._imag_feats = None   # sd_prior stages its bf16 operand in the same buffer
        _lib.check(self.lib.sd_prior(self.h, R, _ptr(deter), _ptr(u), _ptr(stoch), _ptr(logit), flags, self.stream),
                   "sd_prior")
        self._bump("prior")   # taped or not: every sd_prior call rewrites the prior tape's buffers
        return stoch.reshape(*lead, c.S, c.K), logit.reshape(*lead, c.S, c.K)

    def prior_bwd(self, R, d_stoch, d_logit, want_deter=True, weight_grads=None, flags=0):
        c = self.cfg
        ds = None if d_stoch is None else self._stage(_f32c(d_stoch, "d_stoch"), "pds")
        dl = None if d_logit is None else self._stage(_f32c(d_logit, "d_logit"), "pdl")
        d_deter = self._new(R, c.D, tag="pr_dd") if want_deter else None
        arr = None
        if weight_grads is not None:
            names = self.weight_names(MOD_RSSM)
            arr = (C.c_void_p * len(names))(*[0 if weight_grads.get(n) is None else weight_grads[n].data_ptr()
                                              for n in names])
        _lib.check(self.lib.sd_prior_bwd(self.h, R, _ptr(ds), _ptr(dl), _ptr(d_deter), arr, flags, self.stream),
                   "sd_prior_bwd")
        return d_deter

    def imagine_with_action(self, stoch, deter, actions, u, flags=0):
        c = self.cfg
        R, T = actions.shape[:2]
        stoch, deter, actions, u = (_f32c(stoch, "stoch"), _f32c(deter, "deter"), _f32c(actions, "actions"),
                                    _f32c(u, "u"))
        assert u.numel() == R * T * self.SK
        stochs, deters = self._new(R, T, c.S, c.K), self._new(R, T, c.D)
        _lib.check(self.lib.sd_imagine_with_action(self.h, R, T, _ptr(stoch), _ptr(deter), _ptr(actions), _ptr(u),
                                                   _ptr(stochs), _ptr(deters), flags, self.stream),
                   "sd_imagine_with_action")
        return stochs, deters

    def imagine(self, stoch0, deter0, u, act_noise, H, flags=0, out=None):
        c = self.cfg
        N = deter0.shape[0]
        stoch0, deter0, u, act_noise = (_f32c(stoch0, "stoch0"), _f32c(deter0, "deter0"), _f32c(u, "u"),
                                        _f32c(act_noise, "act_noise"))
        assert u.numel() == N * H * self.SK and act_noise.numel() == N * H * c.A
        stoch0, deter0, u, act_noise = (self._stage(stoch0, "s0"), self._stage(deter0, "d0"), self._stage(u, "iu"),
                                        self._stage(act_noise, "an"))
        feats, actions = out if out is not None else (self._new(N, H, self.F, tag="im_f"), self._new(N, H, c.A, tag="im_a"))
        _lib.check(self.lib.sd_imagine_fwd(self.h, N, H, _ptr(stoch0), _ptr(deter0), _ptr(u), _ptr(act_noise),
                                           _ptr(feats), _ptr(actions), flags, self.stream), "sd_imagine_fwd")
        if flags & SD_FLAG_SAVE_TAPE:
            self._bump("scan")
        # the library kept a bf16 copy of feats (tcgen05 path, no tape): heads_lambda() may reuse it while feats is unmodified
        self._imag_feats = (feats.data_ptr(), feats._version, N, H) if (flags & SD_FLAG_BF16 and not flags & SD_FLAG_SAVE_TAPE
                                                                       and N >= 128) else None
        return feats, actions

    def imagine_bwd(self, N, H, d_feats, d_actions, flags=0):
        c = self.cfg
        df = None if d_feats is None else _f32c(d_feats, "d_feats")
        da = None if d_actions is None else _f32c(d_actions, "d_actions")
        d_s, d_d = self._new(N, c.S, c.K), self._new(N, c.D)
        _lib.check(self.lib.sd_imagine_bwd(self.h, N, H, _ptr(df), _ptr(da), _ptr(d_s), _ptr(d_d), flags, self.stream),
                   "sd_imagine_bwd")
        return d_s, d_d

    def heads_lambda(self, feats, disc, lamb, flags=0, slow=True, out=None):
        N, H = feats.shape[:2]
        feats = _f32c(feats, "feats")  # (N,H,F) is large: not staged; imagine()'s static output keeps it stable
        if out is None:
            rew, cont, val, wgt = (self._new(N, H, 1, tag=f"hl{i}") for i in range(4))
            sval = self._new(N, H, 1, tag="hl_sv") if slow else None
            ret = self._new(N, H - 1, 1, tag="hl_ret")
        else:
            rew, cont, val, sval, wgt, ret = out
        tag = getattr(self, "_imag_feats", None)
        if (flags & SD_FLAG_BF16) and tag == (feats.data_ptr(), feats._version, N, H):
            flags |= _lib.SD_FLAG_FEATS_FROM_IMAGINE   # same, unmodified tensor: skip the 168 MB re-cast
        else:
            self._imag_feats = None                    # the library's bf16 staging gets overwritten by this call
        _lib.check(self.lib.sd_heads_lambda_fwd(self.h, N, H, _ptr(feats), float(disc), float(lamb), _ptr(rew),
                                                _ptr(cont), _ptr(val), _ptr(sval), _ptr(wgt), _ptr(ret), flags,
                                                self.stream), "sd_heads_lambda_fwd")
        return rew, cont, val, sval, wgt, ret

    def heads_lambda_bwd(self, feats, disc, lamb, d_ret=None, d_reward=None, d_cont=None, d_value=None, flags=0):
        """d(feats) of heads_lambda for cotangents of ret (N,H-1,1) and optionally reward / cont / value (N,H,1)."""
        N, H = feats.shape[:2]
        feats = _f32c(feats, "feats")
        gs = [None if g is None else _f32c(g, "cotangent") for g in (d_ret, d_reward, d_cont, d_value)]
        d_feats = torch.empty_like(feats)
        if flags & SD_FLAG_BF16:
            self._imag_feats = None
        _lib.check(self.lib.sd_heads_lambda_bwd(self.h, N, H, _ptr(feats), float(disc), float(lamb), _ptr(gs[0]), _ptr(gs[1]),
                                                _ptr(gs[2]), _ptr(gs[3]), _ptr(d_feats), flags, self.stream),
                   "sd_heads_lambda_bwd")
        return d_feats

    def lambda_return(self, last, term, reward, value, boot, disc, lamb):
        N, T = reward.shape[:2]
        last, term, reward, value, boot = (_f32c(x, "lambda_return input") for x in (last, term, reward, value, boot))
        out = self._new(N, T - 1, 1)
        _lib.check(self.lib.sd_lambda_return(N, T, _ptr(last), _ptr(term), _ptr(reward), _ptr(value), _ptr(boot),
                                             float(disc), float(lamb), _ptr(out), self.stream), "sd_lambda_return")
        return out

    def kl_loss(self, post_logit, prior_logit, free, entropies=False):
        c = self.cfg
        lead = post_logit.shape[:-2]
        a, b = _f32c(post_logit, "post_logit"), _f32c(prior_logit, "prior_logit")
        R = a.numel() // self.SK
        dyn, rep = self._new(R), self._new(R)
        ep = self._new(R) if entropies else None
        eq = self._new(R) if entropies else None
        _lib.check(self.lib.sd_kl_loss(self.h, R, _ptr(a), _ptr(b), float(free), _ptr(dyn), _ptr(rep), _ptr(ep),
                                       _ptr(eq), self.stream), "sd_kl_loss")
        outs = [dyn.reshape(lead), rep.reshape(lead)]
        if entropies:
            outs += [ep.reshape(lead), eq.reshape(lead)]
        return tuple(outs)

    def kl_loss_bwd(self, post_logit, prior_logit, free, g_dyn, g_rep, want_post=True, want_prior=True):
        """Gradients of (dyn, rep) = kl_loss(...) w.r.t. the raw logits (rssm.py:222-230 detach pattern)."""
        a, b = _f32c(post_logit, "post_logit"), _f32c(prior_logit, "prior_logit")
        R = a.numel() // self.SK
        gd = None if g_dyn is None else _f32c(g_dyn, "g_dyn").reshape(-1)
        gr = None if g_rep is None else _f32c(g_rep, "g_rep").reshape(-1)
        d_post = torch.empty_like(a) if want_post else None
        d_prior = torch.empty_like(b) if want_prior else None
        _lib.check(self.lib.sd_kl_loss_bwd(self.h, R, _ptr(a), _ptr(b), float(free), _ptr(gd), _ptr(gr), _ptr(d_post),
                                           _ptr(d_prior), self.stream), "sd_kl_loss_bwd")
        return d_post, d_prior
