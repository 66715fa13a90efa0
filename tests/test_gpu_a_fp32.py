"""GPU parity (fp32 SIMT path) through the C ABI: CUDA vs numpy oracle vs reference goldens.

Tolerances, fp32 CUDA (fmaf, split-K order) vs fp32 numpy/torch:
  indices: bit-exact except logged near ties (top-2 gap < 1e-4); deter/logit |d| <= 5e-5 on
  trajectories without a flipped sample; TwoHot modes / returns rtol 3e-4.
"""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import (assert_indices, cu, golden_initial, golden_params, load_golden, make_engine,
                           perturbed_scores)

pytestmark = pytest.mark.gpu
CASES = ["tiny_cont", "tiny_onehot", "base_cont", "base_onehot18", "base_e256", "base_k32"]
ATOL = 5e-5


@pytest.fixture(scope="module", params=CASES)
def case(request):
    c, z = load_golden(request.param)
    P = golden_params(c, z)
    eng = make_engine(c, P, max_rows=64, max_steps=8)
    return request.param, c, z, P, eng


def _np(t):
    return t.detach().cpu().numpy()


def test_abi_loaded():
    from safe_dreamer_b200 import _lib
    lib = _lib.load()
    assert lib.sd_abi_version() == 1


def test_observe(case):
    tag, c, z, P, eng = case
    B, T = int(z["B"]), int(z["T"])
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    s0, d0 = golden_initial(c, B)
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed, action, (s0, d0), reset, u)
    for flags in (0, 4, 4):  # direct, graph capture, graph replay
        st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=flags)
        torch.cuda.synchronize()
        st, dt, lg = _np(st), _np(dt), _np(lg)
        assert set(np.unique(st)) <= {0.0, 1.0} and np.all(st.sum(-1) == 1.0)
        assert_indices(st.argmax(-1), idx_o, perturbed_scores(lg_o, u, c.unimix), 1e-4, 2e-3, f"{tag} observe")
        # a flipped near-tie sample changes the rest of THAT batch row only: every row without a flip is compared
        # unconditionally, and (the goldens are tie-free at this tolerance) at least half of the rows must be clean
        ok = (st.argmax(-1) == idx_o).reshape(B, -1).all(1)
        assert ok.sum() * 2 >= B, f"{tag}: only {ok.sum()}/{B} trajectories without a flipped sample"
        np.testing.assert_allclose(dt[ok], dt_o[ok], atol=ATOL, rtol=0)
        np.testing.assert_allclose(lg[ok], lg_o[ok], atol=ATOL * 4, rtol=0)
        # and against the reference's own outputs
        np.testing.assert_array_equal(st.argmax(-1).astype(np.int8)[ok], z["obs_stoch_idx"][ok])
        np.testing.assert_allclose(dt[ok], z["obs_deter"][ok], atol=ATOL, rtol=0)
        np.testing.assert_allclose(lg[ok], z["obs_logit"][ok], atol=ATOL * 4, rtol=0)


def test_obs_step_and_img_step(case):
    tag, c, z, P, eng = case
    B, T = int(z["B"]), int(z["T"])
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    s0, d0 = golden_initial(c, B)
    st, dt, lg = eng.observe(cu(embed[:, :1]), cu(action[:, :1]), cu(s0), cu(d0), cu(reset[:, :1]), cu(u[:, :1]))
    np.testing.assert_array_equal(_np(st)[:, 0].argmax(-1).astype(np.int8), z["step_obs_idx"])
    np.testing.assert_allclose(_np(dt)[:, 0], z["step_obs_deter"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(_np(lg)[:, 0], z["step_obs_logit"], atol=ATOL * 4, rtol=0)
    sts, dts = eng.imagine_with_action(cu(s0), cu(d0), cu(action[:, 1:2]), cu(u[:, 1:2]))
    np.testing.assert_array_equal(_np(sts)[:, 0].argmax(-1).astype(np.int8), z["step_img_idx"])
    np.testing.assert_allclose(_np(dts)[:, 0], z["step_img_deter"], atol=ATOL, rtol=0)
    sts, dts = eng.imagine_with_action(cu(s0), cu(d0), cu(action), cu(u))
    np.testing.assert_array_equal(_np(sts).argmax(-1).astype(np.int8), z["iwa_idx"])
    np.testing.assert_allclose(_np(dts), z["iwa_deter"], atol=ATOL, rtol=0)


def test_prior_and_kl(case):
    tag, c, z, P, eng = case
    B, T = int(z["B"]), int(z["T"])
    up = O.clamp_u(np.random.Generator(np.random.Philox(11)).random((B, T, c.S, c.K), dtype=np.float32))
    st, lg = eng.prior(cu(z["obs_deter"]), cu(up))
    np.testing.assert_allclose(_np(lg), z["prior_logit"], atol=ATOL * 4, rtol=0)
    np.testing.assert_array_equal(_np(st).argmax(-1).astype(np.int8), z["prior_idx"])
    dyn, rep, ep, eq = eng.kl_loss(cu(z["obs_logit"]), cu(z["prior_logit"]), 1.0, entropies=True)
    np.testing.assert_allclose(_np(dyn), z["kl_dyn"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(_np(rep), z["kl_rep"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(_np(ep), z["ent_post"], rtol=1e-4)
    np.testing.assert_allclose(_np(eq), z["ent_prior"], rtol=1e-4)


def test_imagine_heads_lambda(case):
    tag, c, z, P, eng = case
    N, H = int(z["N"]), int(z["H"])
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    for flags in (0, 4, 4):
        feats, acts = eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=flags)
        torch.cuda.synchronize()
        feats, acts = _np(feats), _np(acts)
        idx = feats[..., :c.SK].reshape(N, H, c.S, c.K).argmax(-1)
        np.testing.assert_array_equal(idx.astype(np.int8), z["imag_feat_idx"])
        np.testing.assert_allclose(feats[..., c.SK:], z["imag_deter"], atol=ATOL, rtol=0)
        if c.act_kind == "cont":
            np.testing.assert_allclose(acts, z["imag_action"], atol=ATOL, rtol=0)
        else:
            np.testing.assert_array_equal(acts.argmax(-1), z["imag_action"].argmax(-1))
    rew, cont, val, sval, wgt, ret = eng.heads_lambda(cu(feats), 1 - 1 / c.horizon, c.lamb)
    for got, key in ((rew, "imag_reward"), (cont, "imag_cont"), (val, "imag_value"), (sval, "imag_slow_value"),
                     (wgt, "imag_weight"), (ret, "imag_ret")):
        np.testing.assert_allclose(_np(got), z[key], rtol=3e-4, atol=2e-5, err_msg=key)


def test_lambda_return_standalone(case):
    tag, c, z, P, eng = case
    rng = np.random.default_rng(5)
    N, T = 7, 9
    last = (rng.random((N, T, 1)) < 0.2).astype(np.float32)
    term = (rng.random((N, T, 1)) < 0.1).astype(np.float32)
    rew, val, boot = rng.standard_normal((3, N, T, 1)).astype(np.float32)
    exp = O.lambda_return(last, term, rew, val, boot, 1 - 1 / 333, 0.95)
    got = eng.lambda_return(cu(last), cu(term), cu(rew), cu(val), cu(boot), 1 - 1 / 333, 0.95)
    np.testing.assert_allclose(_np(got), exp, rtol=1e-5, atol=1e-6)


def test_errors_are_loud(case):
    tag, c, z, P, eng = case
    with pytest.raises(RuntimeError, match="exceed"):
        eng.imagine_with_action(cu(np.zeros((65, c.S, c.K), np.float32)), cu(np.zeros((65, c.D), np.float32)),
                                cu(np.zeros((65, 1, c.A), np.float32)), cu(np.full((65, 1, c.S, c.K), 0.5, np.float32)))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        eng.prior(torch.zeros(2, c.D), torch.zeros(2, c.S, c.K))
