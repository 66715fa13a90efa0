"""Run the UNMODIFIED reference implementation of the hot path (baseline/_ref, see make_ref.py).

Used by `bench.py` only: the `--impl reference` arm (CPU, all host cores) and the `gpu_reference` key (the same
reference modules on the B200, eager and `torch.compile(mode="reduce-overhead")` as `world_model/dreamer.py:231-233`
does).  Nothing of this repository's kernels is on this path; nothing here is imported by the product package.

Import shim: `world_model`, `utils`, `ablations` are pre-registered as bare namespace packages so
their `__init__`s (tensordict / torchrl imports) are skipped; `tensordict.TensorDict` is stubbed.

The hot path is restated from the reference's call sites, calling the reference's own functions:
  observe          RSSM.observe                                   dreamer.py:483
  imagine          Dreamer._imagine (unbound, on a stand-in self) dreamer.py:585, 673-692
  heads + lambda   frozen reward/cont/value/slow-value heads, cumprod weights, Dreamer._lambda_return  dreamer.py:589-602
"""
import os
import sys
import time
import types
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "world_model", "rssm.py"))


_mods = None


def import_reference():
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/make_ref.py` in the build container")
    for pkg in ("world_model", "utils", "ablations"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, pkg)]
        sys.modules[pkg] = m
    if "tensordict" not in sys.modules:
        td = types.ModuleType("tensordict")

        class TensorDict(dict):
            pass

        td.TensorDict = TensorDict
        sys.modules["tensordict"] = td
    sys.path.insert(0, REF)
    import world_model.rssm as rssm
    import world_model.distributions as dists
    import world_model.networks as networks
    import world_model.dreamer as dreamer
    _mods = NS(rssm=rssm, dists=dists, networks=networks, dreamer=dreamer)
    return _mods


def _t(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def build(c, P, dev):
    """Reference RSSM + heads carrying the synthetic weights of safe_dreamer_b200.synth (same as the GPU arm)."""
    M = import_reference()
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device=str(dev), obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G, norm=True)
    R = M.rssm.RSSM(cfg, c.E, c.A).to(dev)
    R.load_state_dict({k: _t(v, dev) for k, v in P["rssm"].items()}, strict=True)
    adist = (NS(name="bounded_normal", min_std=c.min_std, max_std=c.max_std) if c.act_kind == "cont"
             else NS(name="onehot", unimix_ratio=c.act_unimix))
    spec = {
        "actor": ("actor", c.actor_layers, c.A, adist),
        "reward": ("reward", c.reward_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
        "cont": ("cont", c.cont_layers, 1, NS(name="binary")),
        "value": ("value", c.value_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
        "slow_value": ("value", c.value_layers, c.bins, NS(name="symexp_twohot", bin_num=c.bins)),
    }
    heads = {}
    for key, (name, layers, out, dist) in spec.items():
        hc = NS(act="SiLU", symlog_inputs=False, device=str(dev), layers=layers, units=c.units, name=name, dist=dist,
                outscale=1.0, shape=[out], norm=True)
        h = M.networks.MLPHead(hc, c.F).to(dev)
        h.load_state_dict({k: _t(v, dev) for k, v in P[key].items()}, strict=True)
        heads[key] = h
    for k in ("actor", "reward", "cont", "value", "slow_value"):       # the frozen copies of dreamer.py:260-322
        for p in heads[k].parameters():
            p.requires_grad_(False)
    return R, heads


class HotPath:
    """The reference's own modules wired exactly like dreamer.py:483, 580-602."""

    def __init__(self, c, P, dev, B, T, H):
        M = import_reference()
        self.c, self.dev, self.B, self.T, self.H = c, dev, B, T, H
        self.R, self.heads = build(c, P, dev)
        self.fake = NS(_frozen_rssm=self.R, _frozen_actor=self.heads["actor"])
        self.imagine_fn = M.dreamer.Dreamer._imagine.__wrapped__      # the body under @torch.no_grad()
        self.lambda_fn = M.dreamer.Dreamer._lambda_return.__wrapped__
        self.disc = 1 - 1 / c.horizon

    # -- stages ----------------------------------------------------------------------------------------------------
    def observe(self, embed, action, init, is_first):
        return self.R.observe(embed, action, init, is_first)

    def imagine(self, stoch, deter):
        with torch.no_grad():
            return self.imagine_fn(self.fake, (stoch, deter), self.H)

    def heads_lambda(self, feat):
        with torch.no_grad():
            h = self.heads
            rew = h["reward"](feat).mode()
            cont = h["cont"](feat).mean
            val = h["value"](feat).mode()
            slow = h["slow_value"](feat).mode()
            weight = torch.cumprod(cont * self.disc, dim=1)
            ret = self.lambda_fn(self.fake, torch.zeros_like(cont), 1 - cont, rew, val, val, self.disc, self.c.lamb)
            return rew, cont, val, slow, weight, ret

    def observe_fwd_bwd(self, embed, action, init, is_first, g):
        for p in self.R.parameters():
            p.grad = None
        st, dt, lg = self.R.observe(embed, action, init, is_first)
        loss = (st * g[0]).sum() + (dt * g[1]).sum() + (lg.float() * g[2]).sum()
        loss.backward()
        return st, dt, lg

    def step(self, embed, action, init, is_first, g, bwd=True):
        """One pass of the hot path (the unit `bench.py` times): observe fwd(+bwd) -> imagine -> heads + lambda-return."""
        if bwd:
            st, dt, _ = self.observe_fwd_bwd(embed, action, init, is_first, g)
        else:
            with torch.no_grad():
                st, dt, _ = self.R.observe(embed, action, init, is_first)
        n = st.shape[0] * st.shape[1]
        feat, _ = self.imagine(st.detach().reshape(n, self.c.S, self.c.K).float(), dt.detach().reshape(n, self.c.D).float())
        return self.heads_lambda(feat)[-1]


def make_inputs(c, B, T, dev, seed=2):
    from safe_dreamer_b200 import synth as O
    embed, action, reset, _ = O.synth_observe_inputs(c, B, T, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(5)
    gs = [torch.randn(B, T, c.S, c.K, generator=g) * 0.01, torch.randn(B, T, c.D, generator=g) * 0.01,
          torch.randn(B, T, c.S, c.K, generator=g) * 0.01]
    init = (torch.zeros(B, c.S, c.K, device=dev), torch.zeros(B, c.D, device=dev))
    return (_t(embed, dev), _t(action, dev), init, _t(reset, dev)[..., None]), [x.to(dev) for x in gs]


def time_cpu(c, P, B, T, H, steps, warmup, bwd=True, budget_s=150.0):
    """--impl reference: the reference's CPU path, fp32 eager, all host threads.  Each step is a bounded sample of the
    workload: the replay rows are cut to `Bs` of B so that `steps` steps fit the budget."""
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cpu")
    hp = HotPath(c, P, dev, B, T, H)
    # probe with one replay row to size the sample
    (e, a, init, f), g = make_inputs(c, B, T, dev)
    cut = lambda Bs: ((e[:Bs], a[:Bs], (init[0][:Bs], init[1][:Bs]), f[:Bs]), [x[:Bs] for x in g])
    (i1, g1) = cut(1)
    hp.step(*i1, g1, bwd)
    t0 = time.perf_counter()
    hp.step(*i1, g1, bwd)
    t_row = time.perf_counter() - t0
    Bs = int(max(1, min(B, budget_s / ((steps + warmup) * t_row))))
    ins, gs = cut(Bs)
    for _ in range(max(0, warmup - 1)):
        hp.step(*ins, gs, bwd)
    t0 = time.perf_counter()
    for _ in range(steps):
        hp.step(*ins, gs, bwd)
    dt = time.perf_counter() - t0
    return {"units": Bs * T * H * steps, "seconds": dt, "rows": Bs, "cores": os.cpu_count() or 1}


def time_gpu(c, P, B, T, H, dev, iters=5, modes=("eager_fp16", "eager_fp32", "compiled_fp16"), flush=None, log=None):
    """The reference's own CUDA path on this GPU, per stage (ms, CUDA events).  `compiled_*` wraps each stage in
    torch.compile(mode="reduce-overhead") like dreamer.py:231-233 wraps `_cal_grad`; fp16 = the reference's
    `autocast(float16)` (dreamer.py:420); fp32 = TF32 matmuls (`train.py:38` set_float32_matmul_precision("high"))."""
    torch.set_float32_matmul_precision("high")
    hp = HotPath(c, P, dev, B, T, H)
    (e, a, init, f), g = make_inputs(c, B, T, dev)
    N = B * T
    out = {}

    def ev_time(fn, n):
        tot = 0.0
        for _ in range(n):
            if flush is not None:
                flush.fill_(1)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); t.record(); t.synchronize()
            tot += s.elapsed_time(t)
        return tot / n

    for mode in modes:
        kind, prec = mode.split("_")
        ac = lambda: torch.autocast(device_type="cuda", dtype=torch.float16, enabled=(prec == "fp16"))
        obs_fb, imag, heads = hp.observe_fwd_bwd, hp.imagine, hp.heads_lambda
        obs_f = hp.observe
        if kind == "compiled":
            torch._dynamo.reset()
            obs_fb = torch.compile(hp.observe_fwd_bwd, mode="reduce-overhead")
            obs_f = torch.compile(hp.observe, mode="reduce-overhead")
            imag = torch.compile(hp.imagine, mode="reduce-overhead")
            heads = torch.compile(hp.heads_lambda, mode="reduce-overhead")
        res = {}
        try:
            t0 = time.perf_counter()
            with ac():
                with torch.no_grad():
                    st, dt, _ = obs_f(e, a, init, f)
                st0 = st.detach().reshape(N, c.S, c.K).float().clone()
                dt0 = dt.detach().reshape(N, c.D).float().clone()
                feat, _ = imag(st0, dt0)
                feat = feat.clone()
                warm = 3 if kind == "compiled" else 2
                for _ in range(warm):
                    with torch.no_grad():
                        obs_f(e, a, init, f)
                    obs_fb(e, a, init, f, g); imag(st0, dt0); heads(feat)
                torch.cuda.synchronize()
                res["warmup_s"] = time.perf_counter() - t0

                def nograd_obs():
                    with torch.no_grad():
                        obs_f(e, a, init, f)
                res["observe_fwd"] = ev_time(nograd_obs, iters)
                res["observe_fwd_bwd"] = ev_time(lambda: obs_fb(e, a, init, f, g), iters)
                res["imagine_fwd"] = ev_time(lambda: imag(st0, dt0), iters)
                res["heads_lambda"] = ev_time(lambda: heads(feat), iters)
                res["hot_path"] = res["observe_fwd_bwd"] + res["imagine_fwd"] + res["heads_lambda"]
                res["imagined_steps_per_s"] = N * H / (res["hot_path"] * 1e-3)
        except Exception as ex:  # a compile failure must not take the bench down
            res["error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
        out[mode] = res
        if log:
            log(mode, res)
    return out


# ------------------------------------------------------------------------------------------------ whole-agent harness
class AttrDict(dict):
    """Attribute access over nested dicts (what Hydra's DictConfig gives the reference)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _load_base_config():
    """configs/base.yaml of the reference with ${a.b} interpolations resolved by hand and numeric strings ('5e5', '1e-4':
    PyYAML reads those as str) converted."""
    import re

    import yaml
    raw = yaml.safe_load(open(os.path.join(REF, "configs", "base.yaml")))

    def lookup(path):
        node = raw
        for part in path.split("."):
            node = node[part]
        return node

    num = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")

    def resolve(x):
        if isinstance(x, dict):
            return AttrDict({k: resolve(v) for k, v in x.items()})
        if isinstance(x, list):
            return [resolve(v) for v in x]
        if isinstance(x, str):
            m = re.fullmatch(r"\$\{([^}]+)\}", x)
            if m:
                return resolve(lookup(m.group(1)))
            if "${" in x:
                return re.sub(r"\$\{([^}]+)\}", lambda mm: str(resolve(lookup(mm.group(1)))), x)
            if num.match(x):
                return float(x)
        return x

    return resolve(raw)


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class _DictSpace:
    def __init__(self, spaces):
        self.spaces = spaces


class _Discrete:
    def __init__(self, n):
        self.n, self.discrete = n, True


class TensorDictStub(dict):
    """What `_cal_grad` needs of a TensorDict: mapping + `.shape` (B, T) (dreamer.py:467,711,718)."""

    def __init__(self, data, batch_size):
        super().__init__(data)
        self.shape = self.batch_size = tuple(batch_size)


def build_dreamer(dev, kind="proprio", rep_loss="dreamer", act_dim=6, obs_dim=24, discrete_actions=0, compile=False, seed=0):
    """The reference's own `Dreamer` (dreamer.py:22-233) from its own base.yaml.
    kind = "proprio": one vector observation of `obs_dim` through the MLP encoder (config C1: embed size 256);
    kind = "vision": a 64x64x3 image through the ConvEncoder (config C2: embed size 1024)."""
    M = import_reference()
    cfg = _load_base_config()
    model = cfg.model
    model.device = str(dev)
    for sub in ("rssm", "reward", "cont", "actor", "critic"):
        model[sub].device = str(dev)
    model.encoder.mlp.device = str(dev)
    model.decoder.mlp.device = str(dev)
    model.rep_loss = rep_loss
    model.compile = bool(compile)
    if kind == "proprio":
        model.encoder.mlp_keys, model.encoder.cnn_keys = "state", "$^"
        model.decoder.mlp_keys, model.decoder.cnn_keys = "state", "$^"
        spaces = {"state": _Space((obs_dim,))}
    else:
        spaces = {"image": _Space((64, 64, 3))}
    for k in ("is_first", "is_last", "is_terminal"):
        spaces[k] = _Space((1,))
    spaces["reward"] = _Space((1,))
    act_space = _Discrete(discrete_actions) if discrete_actions else _Space((act_dim,))
    torch.manual_seed(seed)
    agent = M.dreamer.Dreamer(model, _DictSpace(spaces), act_space).to(dev)
    return agent, cfg


def perturb_agent(agent, seed=1):
    """the reference initialises reward / critic last layers to zero and RMS scales to one; perturb biases, RMS
    scales and those last layers (peaked two-hot bias) so every number downstream is informative."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in agent.named_parameters():
            if name.startswith("_frozen") or name.startswith("_slow"):
                continue
            if name.endswith("bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g).to(p.device))
            elif p.dim() == 1:
                p.copy_((0.5 + torch.rand(p.shape, generator=g)).to(p.device))
        for head in (agent.reward, agent.value):
            w = head.last.weight
            w.copy_((0.01 * torch.randn(w.shape, generator=g) / (w.shape[1] ** 0.5)).to(w.device))
            n = head.last.bias.shape[0]
            head.last.bias.copy_((-0.5 * (torch.arange(n) - (n - 1) / 2).abs()).to(w.device))
        w = agent.actor.last.weight
        w.copy_((torch.randn(w.shape, generator=g) / (w.shape[1] ** 0.5)).to(w.device))
        for v, s in zip(agent.value.parameters(), agent._slow_value.parameters()):
            s.data.copy_(v.data)
    agent.clone_and_freeze()


def make_batch(agent, B, T, dev, kind="proprio", obs_dim=24, seed=2):
    """Synthetic replay batch in the layout Dreamer.preprocess hands to `_cal_grad` (dreamer.py:709-723)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = agent.act_dim
    data = {
        "action": (torch.rand(B, T, A, generator=g) * 2 - 1),
        "reward": torch.randn(B, T, 1, generator=g),
        "is_first": (torch.rand(B, T, 1, generator=g) < 1 / 64),
        "is_last": torch.zeros(B, T, 1, dtype=torch.bool),
        "is_terminal": (torch.rand(B, T, 1, generator=g) < 1 / 128),
    }
    data["is_first"][:, 0] = True
    if agent.act_discrete:
        idx = torch.randint(0, A, (B, T), generator=g)
        data["action"] = torch.nn.functional.one_hot(idx, A).float()
    if kind == "proprio":
        data["state"] = torch.randn(B, T, obs_dim, generator=g)
    else:
        data["image"] = torch.rand(B, T, 64, 64, 3, generator=g) - 0.5
    data = TensorDictStub({k: v.to(dev) for k, v in data.items()}, (B, T))
    initial = agent.rssm.initial(B)
    return data, initial


class NoiseTape:
    """Injected noise shared by the pure reference and the installed build: uniforms for every categorical draw and
    standard normals for the actor, consumed in the reference's order (per-step draws of `observe`, one batched `prior`
    draw, then per imagination step actor noise followed by the prior uniforms; dreamer.py:483-485,684,688)."""

    def __init__(self, B, T, N, H, S, K, A, discrete_actor, seed=3):
        g = torch.Generator(device="cpu").manual_seed(seed)
        lo = 2.0 ** -24
        self.u_obs = torch.rand(B, T, S, K, generator=g).clamp_(lo, 1 - lo)
        self.u_prior = torch.rand(B, T, S, K, generator=g).clamp_(lo, 1 - lo)
        self.u_img = torch.rand(N, H, S, K, generator=g).clamp_(lo, 1 - lo)
        self.a_noise = (torch.rand(N, H, A, generator=g).clamp_(lo, 1 - lo) if discrete_actor
                        else torch.randn(N, H, A, generator=g))
        self.discrete_actor = discrete_actor
        self.shape = (B, T, N, H)

    def reference_queues(self, dev):
        B, T, N, H = self.shape
        uq = [self.u_obs[:, t].to(dev) for t in range(T)] + [self.u_prior.to(dev)]
        eq = []
        for h in range(H):
            if self.discrete_actor:
                uq.append(self.a_noise[:, h].to(dev))
            else:
                eq.append(self.a_noise[:, h].to(dev))
            uq.append(self.u_img[:, h].to(dev))
        return uq, eq


def patch_reference_noise(uq, eq):
    """Replace the reference's two RNG draws by queue pops (same formulas: F.gumbel_softmax(hard=True) with
    g = -log(-log u); Normal.rsample's standard normal)."""
    M = import_reference()

    def rsample(self, sample_shape=(), temperature=1.0):
        u = uq.pop(0)
        assert u.shape == self.logits.shape, (u.shape, self.logits.shape)
        g = -torch.log(-torch.log(u))
        y = ((self.logits + g) / temperature).softmax(-1)
        index = y.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(self.logits, memory_format=torch.legacy_contiguous_format).scatter_(-1, index, 1.0)
        return y_hard - y.detach() + y

    old = M.dists.OneHotDist.rsample
    M.dists.OneHotDist.rsample = rsample
    import torch.distributions.normal as tdn
    old_n = tdn._standard_normal

    def std_normal(shape, dtype, device):
        e = eq.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype)

    tdn._standard_normal = std_normal

    def undo():
        M.dists.OneHotDist.rsample = old
        tdn._standard_normal = old_n
    return undo
