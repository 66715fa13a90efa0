"""Synthetic, seeded model sizes / weights / inputs for the RSSM hot path (pure numpy, no compute path).

Shared by the benchmark, the tests and the CPU oracle so that all of them see identical weights and inputs
(SURVEY.md section 8d).  Nothing here evaluates the model.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- config
class Cfg:
    """Sizes of the hot path (configs/base.yaml:117-127,252-276,340-420)."""

    def __init__(self, D=2048, U=256, S=32, K=16, G=8, E=1024, A=6, unimix=0.01,
                 img_layers=2, obs_layers=1, act_kind="cont", act_unimix=0.01,
                 min_std=0.1, max_std=1.0, units=256, actor_layers=3,
                 value_layers=3, reward_layers=1, cont_layers=1, bins=255,
                 horizon=333, lamb=0.95):
        self.D, self.U, self.S, self.K, self.G, self.E, self.A = D, U, S, K, G, E, A
        self.unimix = unimix
        self.img_layers, self.obs_layers = img_layers, obs_layers
        self.act_kind, self.act_unimix = act_kind, act_unimix
        self.min_std, self.max_std = min_std, max_std
        self.units = units
        self.actor_layers, self.value_layers = actor_layers, value_layers
        self.reward_layers, self.cont_layers = reward_layers, cont_layers
        self.bins = bins
        self.horizon, self.lamb = horizon, lamb

    @property
    def SK(self):
        return self.S * self.K

    @property
    def F(self):
        return self.S * self.K + self.D

    @property
    def act_out(self):
        return 2 * self.A if self.act_kind == "cont" else self.A

    def as_dict(self):
        return dict(self.__dict__)


# --------------------------------------------------------------------------- params
def rssm_param_shapes(c: Cfg):
    """state_dict names/shapes of reference ``RSSM``."""
    Dg = c.D // c.G
    sh = {
        "_deter_net._dyn_in0.0.weight": (c.U, c.D), "_deter_net._dyn_in0.0.bias": (c.U,),
        "_deter_net._dyn_in0.1.weight": (c.U,),
        "_deter_net._dyn_in1.0.weight": (c.U, c.SK), "_deter_net._dyn_in1.0.bias": (c.U,),
        "_deter_net._dyn_in1.1.weight": (c.U,),
        "_deter_net._dyn_in2.0.weight": (c.U, c.A), "_deter_net._dyn_in2.0.bias": (c.U,),
        "_deter_net._dyn_in2.1.weight": (c.U,),
        "_deter_net._dyn_hid.dyn_hid_0.weight": (Dg, Dg + 3 * c.U, c.G),
        "_deter_net._dyn_hid.dyn_hid_0.bias": (c.D,),
        "_deter_net._dyn_hid.norm_0.weight": (c.D,),
        "_deter_net._dyn_gru.weight": (3 * Dg, Dg, c.G), "_deter_net._dyn_gru.bias": (3 * c.D,),
    }
    inp = c.D + c.E
    for i in range(c.obs_layers):
        sh[f"_obs_net.obs_net_{i}.weight"] = (c.U, inp)
        sh[f"_obs_net.obs_net_{i}.bias"] = (c.U,)
        sh[f"_obs_net.obs_net_n_{i}.weight"] = (c.U,)
        inp = c.U
    sh["_obs_net.obs_net_logit.weight"] = (c.SK, inp)
    sh["_obs_net.obs_net_logit.bias"] = (c.SK,)
    inp = c.D
    for i in range(c.img_layers):
        sh[f"_img_net.img_net_{i}.weight"] = (c.U, inp)
        sh[f"_img_net.img_net_{i}.bias"] = (c.U,)
        sh[f"_img_net.img_net_n_{i}.weight"] = (c.U,)
        inp = c.U
    sh["_img_net.img_net_logit.weight"] = (c.SK, inp)
    sh["_img_net.img_net_logit.bias"] = (c.SK,)
    return sh


def head_param_shapes(name, layers, inp, units, out):
    """state_dict names/shapes of reference ``MLPHead`` (networks.py:313-377)."""
    sh = {}
    for i in range(layers):
        sh[f"mlp.layers.{name}_linear{i}.weight"] = (units, inp)
        sh[f"mlp.layers.{name}_linear{i}.bias"] = (units,)
        sh[f"mlp.layers.{name}_norm{i}.weight"] = (units,)
        inp = units
    sh["last.weight"] = (out, inp)
    sh["last.bias"] = (out,)
    return sh


def all_param_shapes(c: Cfg):
    return {
        "rssm": rssm_param_shapes(c),
        "actor": head_param_shapes("actor", c.actor_layers, c.F, c.units, c.act_out),
        "reward": head_param_shapes("reward", c.reward_layers, c.F, c.units, c.bins),
        "cont": head_param_shapes("cont", c.cont_layers, c.F, c.units, 1),
        "value": head_param_shapes("value", c.value_layers, c.F, c.units, c.bins),
        "slow_value": head_param_shapes("value", c.value_layers, c.F, c.units, c.bins),
    }


def init_params(c: Cfg, seed=0):
    """Deterministic synthetic weights (numpy Philox; reproducible on any box).

    Follows the *spirit* of SURVEY.md section 8(d): fan-in scaled weights,
    non-trivial biases / RMS scales, a peaked bias on the two-hot heads so
    that ``TwoHot.mode`` is O(1) (distributions.py:81-92 cancellation hazard).
    Exact values need not match ``weight_init_``; parity tests load *these*
    arrays into the reference modules.
    """
    rng = np.random.Generator(np.random.Philox(seed))
    out = {}
    for mod, shapes in all_param_shapes(c).items():
        p = {}
        for name, shp in shapes.items():
            if name.endswith("bias"):
                v = (rng.random(shp, dtype=np.float32) * 2 - 1) * 0.1
            elif len(shp) == 1:  # RMSNorm scale
                v = 0.5 + rng.random(shp, dtype=np.float32)
            else:
                fan_in = shp[1]
                std = 1.1368 / np.sqrt(fan_in)
                v = (rng.random(shp, dtype=np.float32) * 2 - 1) * np.float32(std * 1.7)
            p[name] = v.astype(np.float32)
        if mod == "actor":
            p["last.weight"] *= np.float32(0.5)
        if mod in ("reward", "value", "slow_value"):
            n = c.bins
            p["last.weight"] *= np.float32(0.05)
            centre = (n - 1) / 2 + (3.0 if mod == "reward" else -5.0)
            p["last.bias"] = (-0.5 * np.abs(np.arange(n) - centre)).astype(np.float32)
        out[mod] = p
    return out



# --------------------------------------------------------------------------- synthetic inputs
def synth_observe_inputs(c: Cfg, B, T, seed=2, p_reset=1.0 / 64):
    """SURVEY.md section 8(d) inputs: embed~N(0,1), action~U(-1,1), is_first[:,0]=1 + Bernoulli."""
    rng = np.random.Generator(np.random.Philox(seed))
    embed = rng.standard_normal((B, T, c.E), dtype=np.float32)
    if c.act_kind == "cont":
        action = (rng.random((B, T, c.A), dtype=np.float32) * 2 - 1).astype(np.float32)
    else:
        ai = rng.integers(0, c.A, size=(B, T))
        action = np.eye(c.A, dtype=np.float32)[ai]
    reset = rng.random((B, T)) < p_reset
    reset[:, 0] = True
    u = clamp_u(rng.random((B, T, c.S, c.K), dtype=np.float32))
    return embed, action, reset, u


def synth_imagine_inputs(c: Cfg, N, H, seed=3):
    rng = np.random.Generator(np.random.Philox(seed))
    idx = rng.integers(0, c.K, size=(N, c.S))
    stoch = np.eye(c.K, dtype=np.float32)[idx]
    deter = np.tanh(rng.standard_normal((N, c.D), dtype=np.float32)).astype(np.float32)
    u = clamp_u(rng.random((N, H, c.S, c.K), dtype=np.float32))
    if c.act_kind == "cont":
        noise = rng.standard_normal((N, H, c.A), dtype=np.float32)
    else:
        noise = clamp_u(rng.random((N, H, c.A), dtype=np.float32))
    return stoch, deter, u, noise


def clamp_u(u):
    lo = np.float32(2.0 ** -24)
    return np.clip(u, lo, np.float32(1.0) - lo).astype(np.float32)


def cast_params(P, dtype):
    return {m: {k: v.astype(dtype) for k, v in d.items()} for m, d in P.items()}




def encoder_params(depths, cin, k, seed=0, dtype=np.float32):
    """Seeded CNN encoder weights in the reference's state_dict naming (networks.py:192-234: layers.{4i}.weight/bias = conv,
    layers.{4i+2}.weight = RMS scale); shared by the oracle (oracle/cnn_oracle.py re-exports it) and the benchmark."""
    rng = np.random.Generator(np.random.Philox(seed))
    P = {}
    for i, co in enumerate(depths):
        fan = cin * k * k
        P[f"layers.{4 * i}.weight"] = (rng.standard_normal((co, cin, k, k), dtype=np.float32) / np.sqrt(fan)).astype(dtype)
        P[f"layers.{4 * i}.bias"] = (0.1 * rng.standard_normal(co, dtype=np.float32)).astype(dtype)
        P[f"layers.{4 * i + 2}.weight"] = (1.0 + 0.1 * rng.standard_normal(co, dtype=np.float32)).astype(dtype)
        cin = co
    return P
