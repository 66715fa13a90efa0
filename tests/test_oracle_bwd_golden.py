"""CPU: the oracle's manual backward against torch.autograd of the real reference (goldens)."""
import numpy as np
import pytest

from oracle import rssm_oracle as O
from tests.helpers import golden_initial, golden_params, load_golden

CASES = ["tiny_cont", "tiny_onehot", "base_cont"]


def _cot(shape_s, shape_d, shape_l):
    g = np.random.Generator(np.random.Philox(13))
    c_st = g.standard_normal(shape_s, dtype=np.float32)
    c_dt = g.standard_normal(shape_d, dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal(shape_l, dtype=np.float32) * np.float32(0.1)
    return c_st, c_dt, c_lg


@pytest.mark.parametrize("tag", CASES)
def test_observe_bwd(tag):
    c, z = load_golden(tag)
    P = golden_params(c, z)
    B, T = int(z["B"]), int(z["T"])
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    tapes = []
    st, dt, lg, idx = O.observe(c, P["rssm"], embed, action, golden_initial(c, B), reset, u, tapes)
    c_st, c_dt, c_lg = _cot(st.shape, dt.shape, lg.shape)
    G, d_embed, d_is, d_id = O.observe_bwd(c, P["rssm"], tapes, c_st, c_dt, c_lg)
    np.testing.assert_allclose(d_embed, z["bwd_d_embed"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(d_is, z["bwd_d_init_stoch"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(d_id, z["bwd_d_init_deter"], rtol=2e-3, atol=2e-5)
    for name in O.rssm_param_shapes(c):
        gn = float(z["bwd_gn/" + name])
        if "bwd_g/" + name in z.files:
            np.testing.assert_allclose(G[name], z["bwd_g/" + name], rtol=2e-3, atol=2e-5 * max(1.0, gn), err_msg=name)
        else:
            sl = G[name].reshape(-1)[:: max(1, G[name].size // 2048)][:2048]
            np.testing.assert_allclose(sl, z["bwd_gs/" + name], rtol=2e-3, atol=2e-5 * max(1.0, gn), err_msg=name)
        assert abs(np.sqrt((G[name].astype(np.float64) ** 2).sum()) - gn) <= 2e-3 * max(gn, 1e-3), name


@pytest.mark.parametrize("tag", CASES)
def test_imagine_bwd(tag):
    c, z = load_golden(tag)
    P = golden_params(c, z)
    N, H = int(z["N"]), int(z["H"])
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    tapes = []
    feats, acts = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), H, u, noise, tapes)
    g2 = np.random.Generator(np.random.Philox(17))
    c_f = g2.standard_normal(feats.shape, dtype=np.float32) * np.float32(0.1)
    c_a = g2.standard_normal(acts.shape, dtype=np.float32)
    ds, dd = O.imagine_bwd(c, P["rssm"], P["actor"], tapes, c_f, c_a)
    np.testing.assert_allclose(ds, z["imag_bwd_d_stoch"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(dd, z["imag_bwd_d_deter"], rtol=2e-3, atol=2e-5)
