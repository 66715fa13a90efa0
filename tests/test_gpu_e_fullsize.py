"""GPU, BASELINE.json's full sizes (B=16, T=64, N=1024, H=16, base.yaml shapes): size-independent properties
(the oracle takes seconds at this size, so only one short oracle comparison is made).
  * determinism: direct launch == CUDA-graph replay == second replay, bit for bit (fixed reduction orders)
  * structure: stoch rows are exact one-hots; feats[:, 0] is the start state; actions finite
  * reset invariance: a row whose is_first is set at step t does not depend on anything before t
  * backward linearity: grads scale exactly with the upstream scale (GradScaler, dreamer.py:227,667) and the
    dgrad-only call returns the same input grads as the full call
  * lambda-return recursion holds on the GPU outputs (dreamer.py:701-706); weight is a running product"""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import cu, make_engine

pytestmark = pytest.mark.gpu
B, T, N, H = 16, 64, 1024, 16


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def full():
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    eng = make_engine(c, P, max_rows=N, max_steps=T, max_tape_rows=B)
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    return c, P, eng, embed, action, reset, u


def test_observe_determinism_structure_and_reset_cut(full):
    c, P, eng, embed, action, reset, u = full
    s0 = np.zeros((B, c.S, c.K), np.float32); d0 = np.zeros((B, c.D), np.float32)
    outs = [tuple(_np(x).copy() for x in eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=f))
            for f in (0, 4, 4)]
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            np.testing.assert_array_equal(a, b)
    st, dt, lg = outs[0]
    assert set(np.unique(st)) == {0.0, 1.0} and np.all(st.sum(-1) == 1.0)
    assert np.isfinite(dt).all() and np.abs(dt).max() <= 1.0 + 1e-6  # convex mix of tanh candidates and the old state
    # first 3 steps against the oracle
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed[:, :3], action[:, :3], (s0, d0), reset[:, :3], u[:, :3])
    np.testing.assert_array_equal(st[:, :3].argmax(-1), idx_o)
    np.testing.assert_allclose(dt[:, :3], dt_o, atol=5e-5, rtol=0)
    # reset invariance: row 3 is reset at step 20 -> from there on it must not depend on the initial state
    r2 = reset.copy(); r2[3, 20] = True
    rng = np.random.Generator(np.random.Philox(1))
    s1 = np.eye(c.K, dtype=np.float32)[rng.integers(0, c.K, size=(B, c.S))]
    d1 = rng.standard_normal((B, c.D), dtype=np.float32)
    r2[:, 0] = False
    a = [_np(x) for x in eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(r2), cu(u))]
    b = [_np(x) for x in eng.observe(cu(embed), cu(action), cu(s1), cu(d1), cu(r2), cu(u))]
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x[3, 20:], y[3, 20:])
    assert np.abs(a[1][3, :20] - b[1][3, :20]).max() > 1e-3


def test_backward_linearity_and_dgrad_only(full):
    c, P, eng, embed, action, reset, u = full
    s0 = np.zeros((B, c.S, c.K), np.float32); d0 = np.zeros((B, c.D), np.float32)
    st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=2)
    g = torch.Generator(device="cuda").manual_seed(0)
    gs = torch.randn(st.shape, device="cuda", generator=g) * 0.01
    gd = torch.randn(dt.shape, device="cuda", generator=g) * 0.01
    gl = torch.randn(lg.shape, device="cuda", generator=g) * 0.01
    names = eng.weight_names(0)

    def run(scale, with_w=True):
        wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in names} if with_w else None
        de, dis, did = eng.observe_bwd(B, T, gs * scale, gd * scale, gl * scale, True, True, wg)
        torch.cuda.synchronize()
        return de.clone(), did.clone(), wg

    de1, did1, w1 = run(1.0)
    de2, did2, w2 = run(65536.0)          # GradScaler's initial scale: a power of two => exact
    torch.testing.assert_close(de2, de1 * 65536.0, rtol=0, atol=0)
    torch.testing.assert_close(did2, did1 * 65536.0, rtol=0, atol=0)
    for n in names:
        torch.testing.assert_close(w2[n], w1[n] * 65536.0, rtol=0, atol=0)
        assert torch.isfinite(w1[n]).all()
    de3, did3, _ = run(1.0, with_w=False)
    torch.testing.assert_close(de3, de1, rtol=0, atol=0)
    assert float(w1["_img_net.img_net_0.weight"].abs().sum()) == 0.0   # prior net takes no part in observe


def test_imagine_heads_properties(full):
    c, P, eng, *_ = full
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    runs = []
    for f in (1, 1 | 4, 1 | 4, 1 | 4 | 16, 1 | 4 | 16):   # 16 = SD_FLAG_BACKGROUND: graph captured without PDL, same results
        feats, acts = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H, flags=f)
        runs.append((_np(feats).copy(), _np(acts).copy()))
    for fa in runs[1:]:
        np.testing.assert_array_equal(runs[0][0], fa[0])
        np.testing.assert_array_equal(runs[0][1], fa[1])
    feats, acts = runs[0]
    np.testing.assert_array_equal(feats[:, 0, :c.SK], st0.reshape(N, -1))
    np.testing.assert_array_equal(feats[:, 0, c.SK:], dt0)
    oh = feats[..., :c.SK].reshape(N, H, c.S, c.K)
    assert np.all(oh.sum(-1) == 1.0) and set(np.unique(oh)) == {0.0, 1.0}
    assert np.isfinite(feats).all() and np.isfinite(acts).all()
    disc, lamb = 1 - 1 / c.horizon, c.lamb
    # (a) heads straight on imagine()'s output tensor: the library reuses the bf16 copy the rollout wrote step by step
    #     (SD_FLAG_FEATS_FROM_IMAGINE); (b) on a copy: re-cast path.  Both must be bit-identical.
    feats_t, _ = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H, flags=1)
    assert eng._imag_feats is not None
    outs_a = [x.clone() for x in eng.heads_lambda(feats_t, disc, lamb, flags=1)]
    outs_b = eng.heads_lambda(cu(feats), disc, lamb, flags=1)
    for xa, xb in zip(outs_a, outs_b):
        assert torch.equal(xa, xb)
    feats_t.add_(0.0)   # an in-place op bumps the tensor version: the copy may be stale, the wrapper must re-cast
    outs_c = eng.heads_lambda(feats_t, disc, lamb, flags=1)
    assert eng._imag_feats is None
    for xa, xc in zip(outs_a, outs_c):
        assert torch.equal(xa, xc)
    rew, cont, val, sval, wgt, ret = [_np(x).astype(np.float64) for x in outs_b]
    # value parity of the 16 384-row head evaluation (wide fused-norm tcgen05 first layers) on the first 256 trajectories
    M = 256
    rew_o, cont_o, val_o, sval_o, wgt_o, ret_o = O.heads_lambda(c, P["reward"], P["cont"], P["value"], P["slow_value"], feats[:M])
    np.testing.assert_allclose(cont[:M], cont_o, atol=0.02)
    np.testing.assert_allclose(rew[:M], rew_o, rtol=0.1, atol=0.05)
    np.testing.assert_allclose(val[:M], val_o, rtol=0.1, atol=0.05)
    np.testing.assert_allclose(sval[:M], sval_o, rtol=0.1, atol=0.05)
    assert ((cont > 0) & (cont < 1)).all()
    np.testing.assert_allclose(wgt, np.cumprod(cont * disc, axis=1), rtol=1e-5)
    # R_t = r_{t+1} + cont_{t+1}*disc*((1-lamb) v_{t+1} + lamb R_{t+1}),  R_{H-1} = v_{H-1}
    nxt = val[:, -1]
    for i in reversed(range(H - 1)):
        live = cont[:, i + 1] * disc
        nxt = rew[:, i + 1] + live * ((1 - lamb) * val[:, i + 1] + lamb * nxt)
        np.testing.assert_allclose(ret[:, i], nxt, rtol=2e-4, atol=1e-4)


def test_heads_chain_ragged_rows_match_fp32_path():
    """Frozen-head trunks as one row-tile resident chain launch per head (csrc/sd_chain.cuh, taken from 4096 rows on the bf16
    path): a row count that is not a multiple of the 128-row tile (300 x 15 = 4500 rows: 35 full tiles + 20 rows) against the
    fp32 path of the same library (which is pinned to the oracle), and the last rows against the oracle directly."""
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    N, H = 300, 15
    eng = make_engine(c, P, max_rows=N, max_steps=H)
    rng = np.random.default_rng(7)
    feats = np.concatenate([np.eye(c.K, dtype=np.float32)[rng.integers(0, c.K, (N, H, c.S))].reshape(N, H, c.SK),
                            np.tanh(rng.standard_normal((N, H, c.D))).astype(np.float32)], axis=-1)
    disc, lamb = 1 - 1 / c.horizon, c.lamb
    out32 = [_np(x).astype(np.float64) for x in eng.heads_lambda(cu(feats), disc, lamb, flags=0)]
    outbf = [_np(x).astype(np.float64) for x in eng.heads_lambda(cu(feats), disc, lamb, flags=1)]
    for name, a, b in zip(("reward", "cont", "value", "slow_value"), outbf, out32):
        assert np.isfinite(a).all(), name
        if name == "cont":
            np.testing.assert_allclose(a, b, atol=0.02, err_msg=name)
        else:
            np.testing.assert_allclose(a, b, rtol=0.1, atol=0.05, err_msg=name)
    M = 8   # the last trajectories live in the ragged tile
    rew_o, cont_o, val_o, sval_o, wgt_o, ret_o = O.heads_lambda(c, P["reward"], P["cont"], P["value"], P["slow_value"], feats[-M:])
    np.testing.assert_allclose(outbf[0][-M:], rew_o, rtol=0.1, atol=0.05)
    np.testing.assert_allclose(outbf[1][-M:], cont_o, atol=0.02)
    np.testing.assert_allclose(outbf[2][-M:], val_o, rtol=0.1, atol=0.05)
