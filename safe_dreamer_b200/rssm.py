"""Drop-in mirror of the reference ``world_model/rssm.py`` whose scans run in libsafedreamer.

Same constructor, method names, argument meaning, parameter names and shapes as the reference
(``RSSM(config.rssm, embed_size, act_dim)``, world_model/rssm.py:78-230), so ``state_dict``s
interchange and ``dreamer.py`` can call it unchanged.  The modules below are parameter
containers: the math runs in CUDA through :class:`safe_dreamer_b200.engine.Engine`.

Noise: the reference draws Gumbel noise inside ``F.gumbel_softmax``; here each call draws
uniforms with ``torch.rand`` on the device (or takes them from ``noise_source`` for parity
tests) and the kernels turn them into Gumbels, g = -log(-log(u)).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import engine as _engine
from ._lib import SD_FLAG_BACKGROUND, SD_FLAG_LAYERWISE, SD_FLAG_PERSISTENT
from .engine import MOD_RSSM, SD_FLAG_BF16, SD_FLAG_GRAPH, SD_FLAG_SAVE_TAPE

_U_LO = 2.0 ** -24


def rpad(x, pad):
    """utils/tools.py:413-416."""
    for _ in range(pad):
        x = x.unsqueeze(-1)
    return x


def weight_init_(m):
    """Same initial distribution as utils/tools.py:76-100: trunc-normal(std=1.1368/sqrt(fan_in)),
    zero bias, unit RMSNorm scale."""
    if isinstance(m, nn.RMSNorm):
        with torch.no_grad():
            m.weight.fill_(1.0)
        return
    w = getattr(m, "weight", None)
    if w is None or w.numel() == 0:
        return
    fan_in = w.shape[1] * (math.prod(w.shape[2:]) if w.dim() > 2 else 1) if w.dim() > 1 else w.shape[0]
    std = 1.1368 * math.sqrt(1.0 / fan_in)
    with torch.no_grad():
        nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std)
        b = getattr(m, "bias", None)
        if b is not None:
            b.fill_(0.0)


class BlockLinear(nn.Module):
    """Parameter container with the reference layout (networks.py:24-41): weight (O/G, I/G, G)."""

    def __init__(self, in_ch, out_ch, blocks, outscale=1.0):
        super().__init__()
        self.in_ch, self.out_ch, self.blocks, self.outscale = int(in_ch), int(out_ch), int(blocks), float(outscale)
        self.weight = nn.Parameter(torch.empty(self.out_ch // self.blocks, self.in_ch // self.blocks, self.blocks))
        self.bias = nn.Parameter(torch.empty(self.out_ch))

    def forward(self, x):  # pragma: no cover - the block GEMMs run inside the fused scan
        raise NotImplementedError("BlockLinear is evaluated inside the CUDA scan; call RSSM.obs_step/img_step")


class Deter(nn.Module):
    """Parameter container mirroring rssm.py:10-34 (same sub-module names => same state_dict keys)."""

    def __init__(self, deter, stoch, act_dim, hidden, blocks, dynlayers, act="SiLU"):
        super().__init__()
        if int(dynlayers) != 1:
            raise NotImplementedError("dyn_layers != 1 is not supported by the CUDA scan (base.yaml:266 uses 1)")
        if act != "SiLU":
            raise NotImplementedError("only act='SiLU' is implemented (base.yaml:129)")
        self.blocks, self.dynlayers = int(blocks), int(dynlayers)
        A = getattr(nn, act)

        def inp(k):
            return nn.Sequential(nn.Linear(k, hidden, bias=True), nn.RMSNorm(hidden, eps=1e-04, dtype=torch.float32), A())

        self._dyn_in0, self._dyn_in1, self._dyn_in2 = inp(deter), inp(stoch), inp(act_dim)
        self._dyn_hid = nn.Sequential()
        in_ch = (3 * hidden + deter // self.blocks) * self.blocks
        self._dyn_hid.add_module("dyn_hid_0", BlockLinear(in_ch, deter, self.blocks))
        self._dyn_hid.add_module("norm_0", nn.RMSNorm(deter, eps=1e-04, dtype=torch.float32))
        self._dyn_hid.add_module("act_0", A())
        self._dyn_gru = BlockLinear(deter, 3 * deter, self.blocks)

    def forward(self, stoch, deter, action):  # pragma: no cover
        raise NotImplementedError("Deter is evaluated inside the CUDA scan; call RSSM.img_step")


class _Runtime:
    """Per-module CUDA state (engine handle, weight signature). Never copied or pickled."""

    def __init__(self):
        self.engine = None
        self.sig = None
        self.limits = (0, 0, 0)
        self.bucket = None
        self.pcache = None     # cached (module id, name -> Parameter) lists (RSSM.cache_params)
        self.ubuf = {}         # pointer-stable noise buffers (RSSM.stage_inputs = False)

    def __deepcopy__(self, memo):
        return _Runtime()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()


class _KLFn(torch.autograd.Function):
    """RSSM.kl_loss (rssm.py:222-230) through the C ABI: rep sends gradient to the posterior logits only, dyn to the prior
    logits only, rows clipped at `free` send none."""

    @staticmethod
    def forward(ctx, rssm, post_logit, prior_logit, free):
        lead = post_logit.shape[:-2]
        rows = 1
        for d in lead:
            rows *= int(d)
        # sd_kl_loss takes up to max_rows * max_steps rows: never grow (= rebuild) the engine here, its tapes may be live
        eng = rssm._get_engine(min(max(rows, 1), rssm.max_rows), 1)
        dyn, rep = eng.kl_loss(post_logit, prior_logit, free)
        ctx.save_for_backward(post_logit, prior_logit)
        ctx.rssm, ctx.free, ctx.rows = rssm, free, rows
        return dyn.reshape(lead).clone(), rep.reshape(lead).clone()

    @staticmethod
    def backward(ctx, g_dyn, g_rep):
        post_logit, prior_logit = ctx.saved_tensors
        eng = ctx.rssm._get_engine(min(max(ctx.rows, 1), ctx.rssm.max_rows), 1)
        d_post, d_prior = eng.kl_loss_bwd(post_logit, prior_logit, ctx.free, g_dyn.contiguous(), g_rep.contiguous(),
                                          ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return None, d_post, d_prior, None


class _ObserveFn(torch.autograd.Function):
    """observe forward/backward through the C ABI (sd_observe_fwd / sd_observe_bwd)."""

    @staticmethod
    def forward(ctx, rssm, embed, action, init_stoch, init_deter, reset, u, *params):
        eng = rssm._get_engine(action.shape[0], action.shape[1], tape=True)
        flags = rssm._flags() | SD_FLAG_SAVE_TAPE
        stochs, deters, logits = eng.observe(embed, action, init_stoch, init_deter, reset, u, flags=flags)
        ctx.rssm, ctx.B, ctx.T = rssm, action.shape[0], action.shape[1]
        ctx.eng, ctx.gen = eng, eng.tape_gen["scan"]
        ctx.need = (ctx.needs_input_grad[1], ctx.needs_input_grad[3] or ctx.needs_input_grad[4],
                    any(ctx.needs_input_grad[7:]))
        ctx.flags = rssm._flags()
        return stochs, deters, logits

    @staticmethod
    def backward(ctx, d_st, d_dt, d_lg):
        rssm = ctx.rssm
        eng = rssm._rt.engine
        if eng is not ctx.eng:
            raise RuntimeError("RSSM.observe backward: the engine that holds this call's tape was rebuilt (a later call needed more "
                               "rows / steps than rssm.max_rows / max_steps); set those limits before the forward")
        eng.check_tape("scan", ctx.gen, "RSSM.observe backward")
        need_embed, need_init, need_w = ctx.need
        names = eng.weight_names(MOD_RSSM)
        wg = None
        if need_w:
            if rssm.static_outputs:   # pointer-stable gradient buffers keep the CUDA-graph cache hot
                from .parallel import GradBucket
                if rssm._rt.bucket is None:
                    rssm._rt.bucket = GradBucket({n: p.shape for n, p in rssm.named_parameters()}, d_dt.device)
                rssm._rt.bucket.zero_()
                wg = rssm._rt.bucket.views
            else:
                wg = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in rssm.named_parameters()}
        d_embed, d_is, d_id = eng.observe_bwd(ctx.B, ctx.T, d_st, d_dt, d_lg, need_embed, need_init, wg, ctx.flags)
        # static mode: gradients are copied out of the pointer-stable bucket -- unless static_grads is set: then p.grad
        # ALIASES the bucket slices (fresh view objects, which autograd adopts without a copy) and is only valid until the
        # next backward of this module (needs zero_grad(set_to_none=True) between backwards; no gradient accumulation)
        pnames = [n for n, _ in rssm._param_dicts()[0][1].items()]
        if wg is None:
            pg = [None] * len(pnames)
        elif rssm.static_outputs:
            pg = [wg[n].view_as(wg[n]) if rssm.static_grads else wg[n].clone() for n in pnames]
        else:
            pg = [wg[n] for n in pnames]
        if rssm.static_outputs:
            d_embed = None if d_embed is None else d_embed.clone()
        assert set(names) == set(pnames)
        return (None, d_embed, None, d_is, d_id, None, None, *pg)


class _PriorFn(torch.autograd.Function):
    """batched prior forward/backward through the C ABI (sd_prior / sd_prior_bwd)."""

    @staticmethod
    def forward(ctx, rssm, deter, u, *params):
        lead = deter.shape[:-1]
        rows = int(math.prod(lead))
        eng = rssm._get_engine(min(rows, rssm.max_rows), 1, tape=True, tape_rows=1)
        stoch, logit = eng.prior(deter, u, flags=rssm._flags() | SD_FLAG_SAVE_TAPE)
        ctx.rssm, ctx.rows, ctx.lead = rssm, rows, lead
        ctx.eng, ctx.gen = eng, eng.tape_gen["prior"]
        ctx.need = (ctx.needs_input_grad[1], any(ctx.needs_input_grad[3:]))
        ctx.flags = rssm._flags()
        return stoch, logit

    @staticmethod
    def backward(ctx, d_stoch, d_logit):
        rssm = ctx.rssm
        eng = rssm._rt.engine
        if eng is not ctx.eng:
            raise RuntimeError("RSSM.prior backward: the engine that holds this call's tape was rebuilt (a later call needed more "
                               "rows / steps than rssm.max_rows / max_steps); set those limits before the forward")
        eng.check_tape("prior", ctx.gen, "RSSM.prior backward")
        need_deter, need_w = ctx.need
        wg = None
        if need_w:
            wg = {n: (torch.zeros_like(p, dtype=torch.float32) if n.startswith("_img_net") else None)
                  for n, p in rssm.named_parameters()}
        d_deter = eng.prior_bwd(ctx.rows, d_stoch, d_logit, need_deter, wg, ctx.flags)
        if d_deter is not None:
            d_deter = d_deter.reshape(*ctx.lead, -1).clone()
        pg = [None if wg is None else wg[n] for n, _ in rssm.named_parameters()]
        return (None, d_deter, None, *pg)


class RSSM(nn.Module):
    """world_model/rssm.py:78-230 with the scans in CUDA."""

    def __init__(self, config, embed_size, act_dim):
        super().__init__()
        self._stoch, self._deter = int(config.stoch), int(config.deter)
        self._hidden, self._discrete = int(config.hidden), int(config.discrete)
        self._unimix_ratio = float(config.unimix_ratio)
        self._initial = str(config.initial)
        self._device = torch.device(config.device)
        self._act_dim, self._embed_size = int(act_dim), int(embed_size)
        self._obs_layers, self._img_layers = int(config.obs_layers), int(config.img_layers)
        self._dyn_layers, self._blocks = int(config.dyn_layers), int(config.blocks)
        self.flat_stoch = self._stoch * self._discrete
        self.feat_size = self.flat_stoch + self._deter
        A = getattr(nn, config.act)
        self._deter_net = Deter(self._deter, self.flat_stoch, act_dim, self._hidden, blocks=self._blocks,
                                dynlayers=self._dyn_layers, act=config.act)
        self._obs_net = nn.Sequential()
        inp = self._deter + embed_size
        for i in range(self._obs_layers):
            self._obs_net.add_module(f"obs_net_{i}", nn.Linear(inp, self._hidden, bias=True))
            self._obs_net.add_module(f"obs_net_n_{i}", nn.RMSNorm(self._hidden, eps=1e-04, dtype=torch.float32))
            self._obs_net.add_module(f"obs_net_a_{i}", A())
            inp = self._hidden
        self._obs_net.add_module("obs_net_logit", nn.Linear(inp, self.flat_stoch, bias=True))
        self._img_net = nn.Sequential()
        inp = self._deter
        for i in range(self._img_layers):
            self._img_net.add_module(f"img_net_{i}", nn.Linear(inp, self._hidden, bias=True))
            self._img_net.add_module(f"img_net_n_{i}", nn.RMSNorm(self._hidden, eps=1e-04, dtype=torch.float32))
            self._img_net.add_module(f"img_net_a_{i}", A())
            inp = self._hidden
        self._img_net.add_module("img_net_logit", nn.Linear(inp, self.flat_stoch))
        self.apply(weight_init_)
        # runtime knobs (not part of the reference signature)
        self.precision = "fp32"        # "bf16" => tcgen05 path when rows >= 128
        self.use_graph = False         # replay scans as cached CUDA graphs
        self.auto_refresh = True       # repack weights on every call (safe with in-place optimizers)
        self.static_outputs = False    # reuse engine-owned output buffers (next same-shape call overwrites them)
        self.stage_inputs = True       # static_outputs: copy inputs into engine-owned buffers (pointer-stable graph keys).
        #                                False = the caller's input tensors are already pointer-stable (persistent device
        #                                buffers filled by copy_); a changed pointer only costs a graph re-capture
        self.static_grads = False      # static_outputs: p.grad aliases engine-owned buffers (valid until the next backward)
        self.cache_params = False      # True = Parameter OBJECTS never change (in-place optimizers): skip the module-tree
        #                                walk (named_parameters) on every call
        self.noise_source = None       # callable(shape, device) -> uniforms, for injected-noise parity
        self.use_custom_ops = False    # True: route through the torch.library operators of safe_dreamer_b200.ops (traceable
        #                                by torch.compile / capturable in CUDA graphs) instead of autograd.Function
        self.max_rows, self.max_steps = 1024, 64
        self.head_modules = {}         # module id -> nn.Module (actor / reward / cont / value / slow value)
        self._rt = _Runtime()

    # ------------------------------------------------------------------ runtime plumbing
    def _key(self):
        """Integer the torch.library operators identify this module by (set once: operator arguments must be plain data)."""
        k = self.__dict__.get("_ops_key")
        if k is None:
            from . import ops
            k = self.__dict__["_ops_key"] = ops.module_key(self)
        return k

    def _flags(self):
        f = SD_FLAG_BF16 if self.precision == "bf16" else 0
        if getattr(self, "background", False):   # calls issued beside latency-critical work on another stream
            f |= SD_FLAG_BACKGROUND
        # imagination rollout: None = library default (persistent kernel where eligible), "persistent" / "layerwise" force one
        path = getattr(self, "imagine_path", None)
        if path == "persistent":
            f |= SD_FLAG_PERSISTENT
        elif path == "layerwise":
            f |= SD_FLAG_LAYERWISE
        return f | (SD_FLAG_GRAPH if self.use_graph else 0)

    def engine_dims(self):
        return dict(D=self._deter, U=self._hidden, S=self._stoch, K=self._discrete, G=self._blocks,
                    E=self._embed_size, A=self._act_dim, obs_layers=self._obs_layers, img_layers=self._img_layers,
                    unimix=self._unimix_ratio)

    def _get_engine(self, rows, steps, tape=False, extra=None, tape_rows=None):
        rt = self._rt
        need = (max(rows, rt.limits[0], self.max_rows), max(steps, rt.limits[1], self.max_steps),
                max((rows if tape_rows is None else tape_rows) if tape else 0, rt.limits[2]))
        if rt.engine is None or need != rt.limits:
            kw = self.engine_dims()
            kw.update(getattr(self, "_head_dims", {}))
            if extra:
                kw.update(extra)
            rt.engine = _engine.Engine(max_rows=max(need[0], 1), max_steps=max(need[1], 1), max_tape_rows=need[2],
                                       device=next(self.parameters()).device, **kw)
            rt.limits, rt.sig = need, None
        rt.engine.static_outputs = self.static_outputs
        rt.engine.stage_inputs = self.stage_inputs
        self.refresh_weights(force=False)
        return rt.engine

    def _param_dicts(self):
        """[(module id, {state_dict name: Parameter})] for the RSSM and the attached heads."""
        rt = self._rt
        pc = rt.pcache
        if self.cache_params and pc is not None and len(pc) == 1 + len(self.head_modules):
            return pc
        pc = [(MOD_RSSM, dict(self.named_parameters()))] + [(mod, dict(m.named_parameters()))
                                                             for mod, m in self.head_modules.items()]
        rt.pcache = pc if self.cache_params else None
        return pc

    def _params(self):
        return list(self._param_dicts()[0][1].values())

    def refresh_weights(self, force=True, heads=True, rssm=True):
        """Repack the (possibly updated in place) parameters into the kernels' layouts.  `rssm` / `heads` select the RSSM's
        own tensors and those of the attached head modules: a training step needs the RSSM's before `observe`, the heads'
        only before the imagination, so a caller may enqueue the two halves at different points (bench.py's end-to-end step)."""
        rt = self._rt
        if rt.engine is None:
            return
        pds = self._param_dicts()
        sig = tuple((p.data_ptr(), p._version) for p in pds[0][1].values())
        if force or self.auto_refresh or sig != rt.sig:
            for mod, named in pds:
                if (mod == MOD_RSSM and rssm) or (mod != MOD_RSSM and heads):
                    rt.engine.set_weights(mod, named)
            if rssm:
                rt.sig = sig

    def _uniform(self, *shape):
        dev = next(self.parameters()).device
        if self.noise_source is not None:
            return self.noise_source(shape, dev)
        if self.static_outputs and not self.stage_inputs:   # pointer-stable noise: regenerate in place
            buf = self._rt.ubuf.get(shape)
            if buf is None or buf.device != dev:
                buf = self._rt.ubuf[shape] = torch.empty(*shape, device=dev, dtype=torch.float32)
            return buf.uniform_().clamp_(_U_LO, 1.0 - _U_LO)
        return torch.rand(*shape, device=dev, dtype=torch.float32).clamp_(_U_LO, 1.0 - _U_LO)

    # ------------------------------------------------------------------ reference API
    def initial(self, batch_size):
        """rssm.py:133-138."""
        dev = next(self.parameters()).device
        deter = torch.zeros(batch_size, self._deter, dtype=torch.float32, device=dev)
        stoch = torch.zeros(batch_size, self._stoch, self._discrete, dtype=torch.float32, device=dev)
        return stoch, deter

    def observe(self, embed, action, initial, reset):
        """rssm.py:140-156: (B,T,E),(B,T,A),((B,S,K),(B,D)),(B,T[,1]) -> stochs, deters, logits."""
        B, T = action.shape[:2]
        stoch, deter = initial
        u = self._uniform(B, T, self._stoch, self._discrete)
        needs_grad = torch.is_grad_enabled() and (
            embed.requires_grad or stoch.requires_grad or deter.requires_grad
            or any(p.requires_grad for p in self._params()))
        if self.use_custom_ops:
            from . import ops
            return torch.ops.safedreamer.observe(embed.float(), action.float(), stoch.float(), deter.float(),
                                                 reset.reshape(B, T).to(torch.uint8), u, self._params(), self._key(),
                                                 bool(needs_grad))
        if needs_grad:
            return _ObserveFn.apply(self, embed.float(), action, stoch.float(), deter.float(), reset, u,
                                    *self._params())
        eng = self._get_engine(B, T)
        return eng.observe(embed, action, stoch, deter, reset, u, flags=self._flags())

    def obs_step(self, stoch, deter, prev_action, embed, reset):
        """rssm.py:158-178 (single posterior step; reset is (B,) or (B,1))."""
        B = deter.shape[0]
        st, dt, lg = self.observe(embed.unsqueeze(1), prev_action.unsqueeze(1), (stoch, deter), reset.reshape(B, 1))
        return st[:, 0], dt[:, 0], lg[:, 0]

    def img_step(self, stoch, deter, prev_action):
        """rssm.py:180-187."""
        st, dt = self.imagine_with_action(stoch, deter, prev_action.unsqueeze(1))
        return st[:, 0], dt[:, 0]

    def prior(self, deter):
        """rssm.py:189-195; also called batched on (B,T,D) (dreamer.py:485)."""
        lead = deter.shape[:-1]
        rows = int(math.prod(lead))
        u = self._uniform(*lead, self._stoch, self._discrete)
        needs_grad = torch.is_grad_enabled() and (deter.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.use_custom_ops:     # (the operator fetches the engine itself: nothing here may touch ctypes under tracing)
            return torch.ops.safedreamer.prior(deter.float(), u, list(self.parameters()), self._key(), bool(needs_grad))
        eng = self._get_engine(min(rows, self.max_rows), 1)
        if rows > eng.cfg.max_rows * eng.cfg.max_steps:
            eng = self._get_engine(self.max_rows, -(-rows // self.max_rows))
        if needs_grad:
            return _PriorFn.apply(self, deter.float(), u, *self.parameters())
        return eng.prior(deter, u, flags=self._flags())

    def imagine_with_action(self, stoch, deter, actions):
        """rssm.py:197-209.  Forward only: the differentiable rollout is `dreamer_ops.imagine_grad` (the attack shape); asking
        this entry for gradients fails loudly instead of silently returning constants."""
        if torch.is_grad_enabled() and (stoch.requires_grad or deter.requires_grad or actions.requires_grad):
            raise RuntimeError("RSSM.img_step / imagine_with_action are forward-only in this build: gradients w.r.t. the state or "
                               "the actions are not propagated.  Use safe_dreamer_b200.dreamer_ops.imagine_grad for the "
                               "differentiable rollout, or call under torch.no_grad() / detach the inputs.")
        R, T = actions.shape[:2]
        eng = self._get_engine(R, T)
        u = self._uniform(R, T, self._stoch, self._discrete)
        return eng.imagine_with_action(stoch, deter, actions, u, flags=self._flags())

    def get_feat(self, stoch, deter):
        """rssm.py:211-217."""
        stoch = stoch.reshape(*stoch.shape[:-2], self._stoch * self._discrete)
        return torch.cat([stoch, deter], -1)

    def get_dist(self, logit):
        """rssm.py:219-220."""
        from .distributions import OneHotDist
        return torch.distributions.independent.Independent(OneHotDist(logit, unimix_ratio=self._unimix_ratio), 1)

    def kl_loss(self, post_logit, prior_logit, free):
        """rssm.py:222-230 -> (dyn_loss, rep_loss).  CUDA tensors go through sd_kl_loss / sd_kl_loss_bwd (values and the
        gradients of the reference's detach pattern); anything else through the torch restatement."""
        if post_logit.is_cuda and prior_logit.is_cuda:
            if self.use_custom_ops:
                from . import ops
                return torch.ops.safedreamer.kl_loss(post_logit.float(), prior_logit.float(), float(free), self._key())
            return _KLFn.apply(self, post_logit.float(), prior_logit.float(), float(free))
        from .distributions import kl
        rep_loss = kl(post_logit, prior_logit.detach()).sum(-1)
        dyn_loss = kl(post_logit.detach(), prior_logit).sum(-1)
        return torch.clip(dyn_loss, min=free), torch.clip(rep_loss, min=free)
