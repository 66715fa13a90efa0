"""GPU edge cases through the C ABI against the oracle: odd sizes (non-power-of-two widths, 32-class
categories, 2 posterior layers / 3 prior layers), single-row / single-step calls, all-reset and no-reset
sequences, row counts that are not multiples of the 16-row or 128-row tiles (tcgen05 path included)."""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import assert_indices, cu, make_engine, perturbed_scores

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


ODD = dict(D=192, U=48, S=5, K=32, G=3, E=40, A=7, units=80, obs_layers=2, img_layers=3, actor_layers=2,
           value_layers=2, reward_layers=2, cont_layers=1, bins=41)


@pytest.fixture(scope="module")
def odd():
    c = O.Cfg(**ODD)
    P = O.init_params(c, seed=5)
    return c, P, make_engine(c, P, max_rows=40, max_steps=6, max_tape_rows=8)


def test_odd_config_forward_and_backward(odd):
    c, P, eng = odd
    B, T = 7, 5
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=8, p_reset=0.3)
    s0 = np.zeros((B, c.S, c.K), np.float32); d0 = np.zeros((B, c.D), np.float32)
    tapes = []
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed, action, (s0, d0), reset, u, tapes)
    st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=2)
    assert_indices(_np(st).argmax(-1), idx_o, perturbed_scores(lg_o, u, c.unimix), 1e-4, 5e-3, "odd observe")
    np.testing.assert_allclose(_np(dt), dt_o, atol=5e-5, rtol=0)
    np.testing.assert_allclose(_np(lg), lg_o, atol=2e-4, rtol=0)
    g = np.random.Generator(np.random.Philox(9))
    c_st = g.standard_normal(st_o.shape, dtype=np.float32)
    c_dt = g.standard_normal(dt_o.shape, dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal(lg_o.shape, dtype=np.float32) * np.float32(0.1)
    G, d_embed, d_is, d_id = O.observe_bwd(c, P["rssm"], tapes, c_st, c_dt, c_lg)
    wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
    de, dis, did = eng.observe_bwd(B, T, cu(c_st), cu(c_dt), cu(c_lg), True, True, wg)
    np.testing.assert_allclose(_np(de), d_embed, rtol=3e-3, atol=3e-5)
    np.testing.assert_allclose(_np(did), d_id, rtol=3e-3, atol=3e-5)
    for name, ref in G.items():
        gn = float(np.sqrt((ref.astype(np.float64) ** 2).sum()))
        np.testing.assert_allclose(_np(wg[name]), ref, rtol=3e-3, atol=3e-5 * max(1.0, gn), err_msg=name)
    # imagination + heads with the odd head sizes (2-layer actor, 41 bins)
    N, H = 9, 4
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=10)
    feats_o, acts_o = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), H, ui, noise)
    feats, acts = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H)
    np.testing.assert_allclose(_np(feats)[..., c.SK:], feats_o[..., c.SK:], atol=5e-5, rtol=0)
    np.testing.assert_allclose(_np(acts), acts_o, atol=5e-5, rtol=0)
    outs = eng.heads_lambda(feats, 1 - 1 / c.horizon, c.lamb)
    outs_o = O.heads_lambda(c, P["reward"], P["cont"], P["value"], P["slow_value"], feats_o)
    for got, ref in zip(outs, outs_o):
        np.testing.assert_allclose(_np(got), ref, rtol=3e-4, atol=2e-5)


@pytest.mark.parametrize("B,T,mode", [(1, 1, "mixed"), (1, 6, "none"), (5, 1, "all"), (33, 3, "mixed")])
def test_shapes_and_reset_patterns(odd, B, T, mode):
    c, P, eng = odd
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=20 + B + T, p_reset=0.4)
    if mode == "none":
        reset[:] = False
    elif mode == "all":
        reset[:] = True
    rng = np.random.Generator(np.random.Philox(3))
    s0 = np.eye(c.K, dtype=np.float32)[rng.integers(0, c.K, size=(B, c.S))]
    d0 = np.tanh(rng.standard_normal((B, c.D), dtype=np.float32)).astype(np.float32)
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed, action, (s0, d0), reset, u)
    st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u))
    assert_indices(_np(st).argmax(-1), idx_o, perturbed_scores(lg_o, u, c.unimix), 1e-4, 1e-2, f"B={B} T={T} {mode}")
    np.testing.assert_allclose(_np(dt), dt_o, atol=5e-5, rtol=0)


def test_ragged_rows_tcgen05():
    """Row counts that do not fill the last 128-row tile (TMA zero fill + store guards) on the bf16 path."""
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    eng = make_engine(c, P, max_rows=200, max_steps=2)
    R = 130
    rng = np.random.Generator(np.random.Philox(77))
    deter = np.tanh(rng.standard_normal((R, c.D), dtype=np.float32)).astype(np.float32)
    u = O.clamp_u(rng.random((R, c.S, c.K), dtype=np.float32))
    st_o, lg_o, idx_o = O.prior(c, P["rssm"], deter, u)
    guard = torch.full((8, c.S, c.K), 7.0, device="cuda")
    st, lg = eng.prior(cu(deter), cu(u), flags=1)
    torch.cuda.synchronize()
    assert np.abs(_np(lg) - lg_o).max() <= 0.06
    assert_indices(_np(st).argmax(-1), idx_o, perturbed_scores(lg_o, u, c.unimix), 0.25, 0.04, "ragged prior tc")
    assert float(guard.min()) == 7.0
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, R, 2, seed=78)
    feats_o, _ = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), 2, ui, noise)
    feats, _ = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), 2, flags=1)
    assert np.abs(_np(feats)[:, 1, c.SK:] - feats_o[:, 1, c.SK:]).max() <= 0.04
    assert np.isfinite(_np(feats)).all()
