"""Config C5 (adversarial-patch planning attack, README.md:68-116): the gradient path after the posterior --
grad-enabled imagination -> frozen reward / cont / value heads -> differentiated lambda-return -> objective --
through the Python autograd entries (dreamer_ops.imagine_grad / heads_lambda_grad over sd_imagine_fwd/_bwd and
sd_heads_lambda_fwd/_bwd), against torch.autograd on the unmodified reference (tests/golden/attack_*.npz written by
tests/golden/make_golden.py --attack-only).  fp32 path, tolerance rtol 3e-3 on every gradient (relative to its norm)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import GOLDEN, cu

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _module(c, P):
    from safe_dreamer_b200 import dreamer_ops
    from safe_dreamer_b200.networks import MLPHead
    from safe_dreamer_b200.rssm import RSSM
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned", device="cuda",
             obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    rssm = RSSM(cfg, c.E, c.A).cuda()
    rssm.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
    heads = {}
    for key, (name, layers, out) in {"actor": ("actor", c.actor_layers, c.act_out), "reward": ("reward", c.reward_layers, c.bins),
                                     "cont": ("cont", c.cont_layers, 1), "value": ("value", c.value_layers, c.bins),
                                     "slow_value": ("value", c.value_layers, c.bins)}.items():
        m = MLPHead(name, layers, c.units, c.F, out).cuda()
        m.load_state_dict({k: cu(v) for k, v in P[key].items()})
        heads[key] = m
    for p in list(rssm.parameters()) + [q for m in heads.values() for q in m.parameters()]:
        p.requires_grad_(False)
    dreamer_ops.attach_heads(rssm, **heads, act_kind=c.act_kind, min_std=c.min_std, max_std=c.max_std, act_unimix=c.act_unimix)
    return rssm


def _close(a, b, what, rtol=3e-3):
    num = float(np.linalg.norm((a - b).ravel()))
    den = float(np.linalg.norm(b.ravel()))
    print(f"{what}: ||diff|| / ||ref|| = {num / max(den, 1e-30):.2e}")
    assert num <= rtol * den + 1e-7, (what, num, den)


@pytest.mark.parametrize("tag,kw", [("tiny_cont", dict(D=256, U=64, S=8, K=8, G=4, E=48, units=64, A=3)), ("base_cont", {})])
def test_attack_gradient_chain_matches_reference_autograd(tag, kw):
    from safe_dreamer_b200 import dreamer_ops
    z = np.load(os.path.join(GOLDEN, f"attack_{tag}.npz"))
    c = O.Cfg(**kw)
    P = O.init_params(c, seed=0)
    N, H = int(z["N"]), int(z["H"])
    rssm = _module(c, P)
    rssm.max_rows, rssm.max_steps = max(N, 16), max(H, 16)
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    st = cu(st0).requires_grad_(True)
    dt = cu(dt0).requires_grad_(True)
    feats, acts = dreamer_ops.imagine_grad(rssm, (st, dt), H, act_noise=cu(noise), u=cu(ui))
    feats.retain_grad()
    rew, cont, val, sval, wgt, ret = dreamer_ops.heads_lambda_grad(rssm, feats, c.horizon, c.lamb)
    np.testing.assert_array_equal(_np(feats)[..., :c.SK], z["feats"][..., :c.SK])       # same sampled trajectory
    np.testing.assert_allclose(_np(feats)[..., c.SK:], z["feats"][..., c.SK:], atol=1e-4)
    np.testing.assert_allclose(_np(ret), z["ret"], rtol=3e-4, atol=3e-4)
    loss = (ret * cu(z["c_ret"])).sum() + (rew * cu(z["c_rew"])).sum()
    assert abs(float(loss) - float(z["loss"])) <= 3e-4 * abs(float(z["loss"])) + 3e-4
    loss.backward()
    _close(_np(feats.grad), z["d_feats"], "d(objective)/d(feats)")
    _close(_np(dt.grad), z["d_deter"], "d(objective)/d(deter0)")
    _close(_np(st.grad), z["d_stoch"], "d(objective)/d(stoch0)")


def test_heads_lambda_bwd_linearity_and_bf16_rows():
    """sd_heads_lambda_bwd at a row count that takes the tcgen05 path (N*H >= 128): linear in the cotangent, and the bf16
    gradient agrees with the fp32 one (relative L2 error <= 3 %)."""
    from safe_dreamer_b200 import dreamer_ops
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    rssm = _module(c, P)
    N, H = 64, 16
    rssm.max_rows, rssm.max_steps = N, H
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=5)
    with torch.no_grad():
        feats, _ = dreamer_ops.imagine(rssm, (cu(st0), cu(dt0)), H, act_noise=cu(noise), u=cu(ui))
    feats = feats.clone()
    g = torch.Generator(device="cuda").manual_seed(1)
    c_ret = torch.randn(N, H - 1, 1, device="cuda", generator=g)
    grads = {}
    for prec in ("fp32", "bf16"):
        rssm.precision = prec
        f = feats.clone().requires_grad_(True)
        out = dreamer_ops.heads_lambda_grad(rssm, f, c.horizon, c.lamb)
        (out[-1] * c_ret).sum().backward()
        grads[prec] = f.grad.clone()
        f2 = feats.clone().requires_grad_(True)
        out2 = dreamer_ops.heads_lambda_grad(rssm, f2, c.horizon, c.lamb)
        (out2[-1] * (c_ret * 4.0)).sum().backward()
        torch.testing.assert_close(f2.grad, grads[prec] * 4.0, rtol=1e-5, atol=1e-9)
    rel = float((grads["bf16"] - grads["fp32"]).norm() / grads["fp32"].norm())
    print("heads_lambda_bwd bf16 vs fp32 relative L2 error:", rel)
    assert rel <= 0.03
