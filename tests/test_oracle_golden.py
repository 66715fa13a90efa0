"""CPU: the numpy oracle against golden vectors produced by the real reference.

Tolerances (fp32 numpy/OpenBLAS vs fp32 torch/MKL, different summation order):
  * categorical indices: bit-exact (any mismatch must be a logged near tie < 1e-5; none occur here)
  * deter / logits: |d| <= 2e-5 absolute after up to 6 recurrent steps
  * TwoHot modes / lambda-returns: rtol 2e-4 (255-bin symexp sum, bins reach 4.85e8)
"""
import numpy as np
import pytest

from oracle import rssm_oracle as O
from tests.helpers import golden_initial, golden_params, load_golden

CASES = ["tiny_cont", "tiny_onehot", "base_cont", "base_onehot18", "base_e256", "base_k32"]
ATOL = 2e-5


@pytest.fixture(scope="module", params=CASES)
def case(request):
    c, z = load_golden(request.param)
    return request.param, c, z, golden_params(c, z)


def test_observe_matches_reference(case):
    tag, c, z, P = case
    B, T = int(z["B"]), int(z["T"])
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    st, dt, lg, idx = O.observe(c, P["rssm"], embed, action, golden_initial(c, B), reset, u)
    np.testing.assert_array_equal(idx.astype(np.int8), z["obs_stoch_idx"])
    np.testing.assert_allclose(dt, z["obs_deter"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(lg, z["obs_logit"], atol=ATOL * 5, rtol=0)
    # straight-through forward value is one-hot up to 1 ulp
    assert float(z["obs_stoch_maxdev"]) <= 2.4e-7
    assert np.abs(st - np.eye(c.K, dtype=np.float32)[idx]).max() <= 2.4e-7


def test_prior_kl_entropy(case):
    tag, c, z, P = case
    B, T = int(z["B"]), int(z["T"])
    up = O.clamp_u(np.random.Generator(np.random.Philox(11)).random((B, T, c.S, c.K), dtype=np.float32))
    deters = z["obs_deter"]
    st, plog, pidx = O.prior(c, P["rssm"], deters.reshape(B * T, -1), up.reshape(B * T, c.S, c.K))
    plog = plog.reshape(B, T, c.S, c.K)
    np.testing.assert_allclose(plog, z["prior_logit"], atol=ATOL * 5, rtol=0)
    np.testing.assert_array_equal(pidx.reshape(B, T, c.S).astype(np.int8), z["prior_idx"])
    dyn, rep = O.kl_loss(z["obs_logit"], plog, 1.0)
    np.testing.assert_allclose(dyn, z["kl_dyn"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rep, z["kl_rep"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.onehot_entropy(z["obs_logit"], c.unimix).sum(-1), z["ent_post"], rtol=1e-4)
    np.testing.assert_allclose(O.onehot_entropy(plog, c.unimix).sum(-1), z["ent_prior"], rtol=1e-4)


def test_single_steps(case):
    tag, c, z, P = case
    B, T = int(z["B"]), int(z["T"])
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    s0, d0 = golden_initial(c, B)
    st, dt, lg, idx = O.obs_step(c, P["rssm"], s0, d0, action[:, 0], embed[:, 0], reset[:, :1], u[:, 0])
    np.testing.assert_array_equal(idx.astype(np.int8), z["step_obs_idx"])
    np.testing.assert_allclose(dt, z["step_obs_deter"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(lg, z["step_obs_logit"], atol=ATOL * 5, rtol=0)
    st, dt, _, idx = O.img_step(c, P["rssm"], s0, d0, action[:, 1], u[:, 1])
    np.testing.assert_array_equal(idx.astype(np.int8), z["step_img_idx"])
    np.testing.assert_allclose(dt, z["step_img_deter"], atol=ATOL, rtol=0)
    ws, wd = O.imagine_with_action(c, P["rssm"], s0, d0, action, u)
    np.testing.assert_array_equal(ws.argmax(-1).astype(np.int8), z["iwa_idx"])
    np.testing.assert_allclose(wd, z["iwa_deter"], atol=ATOL, rtol=0)


def test_imagine_heads_lambda(case):
    tag, c, z, P = case
    N, H = int(z["N"]), int(z["H"])
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    feats, acts = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), H, u, noise)
    idx = feats[..., :c.SK].reshape(N, H, c.S, c.K).argmax(-1)
    np.testing.assert_array_equal(idx.astype(np.int8), z["imag_feat_idx"])
    np.testing.assert_allclose(feats[..., c.SK:], z["imag_deter"], atol=ATOL, rtol=0)
    if c.act_kind == "cont":
        np.testing.assert_allclose(acts, z["imag_action"], atol=ATOL, rtol=0)
    else:
        np.testing.assert_array_equal(acts.argmax(-1), z["imag_action"].argmax(-1))
    rew, cont, val, sval, weight, ret = O.heads_lambda(c, P["reward"], P["cont"], P["value"], P["slow_value"], feats)
    for got, key in ((rew, "imag_reward"), (cont, "imag_cont"), (val, "imag_value"), (sval, "imag_slow_value"),
                     (weight, "imag_weight"), (ret, "imag_ret")):
        np.testing.assert_allclose(got, z[key], rtol=2e-4, atol=1e-5, err_msg=key)
    assert bool(z["used_dreamer_py"])  # goldens came through dreamer.py's own _imagine/_lambda_return


def test_twohot_bins_and_cancellation():
    b = O.twohot_bins(255)
    assert b.shape == (255,) and b[127] == 0 and np.all(b[:127] == -b[:127:-1])
    assert abs(b[0] + 4.85165184e8) / 4.85e8 < 1e-5
    # uniform logits: the paired sum gives exactly 0 (SURVEY a11 cancellation hazard)
    assert O.twohot_mode(np.zeros((2, 255), np.float32), b).max() == 0.0


def test_lambda_return_closed_form():
    rng = np.random.default_rng(0)
    N, T = 4, 7
    rew, val = rng.standard_normal((2, N, T, 1)).astype(np.float32)
    cont = rng.random((N, T, 1)).astype(np.float32)
    disc, lamb = 1 - 1 / 333, 0.95
    ret = O.lambda_return(np.zeros_like(cont), 1 - cont, rew, val, val, disc, lamb)
    exp = np.zeros((N, T, 1), np.float64); exp[:, -1] = val[:, -1]
    for i in reversed(range(T - 1)):
        live = cont[:, i + 1].astype(np.float64) * disc
        exp[:, i] = rew[:, i + 1] + live * ((1 - lamb) * val[:, i + 1] + lamb * exp[:, i + 1])
    np.testing.assert_allclose(ret, exp[:, :-1], rtol=1e-5, atol=1e-6)
