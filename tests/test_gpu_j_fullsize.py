"""GPU parity at BASELINE.json's full sizes, element-wise: the fp32 path of the library against the numpy oracle on the
benchmark's own shapes (posterior scan B=16, T=64; imagination N=1024, H=16; base architecture), plus the transparency of
the reverse scan to non-finite upstream gradients (the reference's autograd lets an inf / NaN reach every weight gradient
it touches, which is what GradScaler's overflow check relies on, dreamer.py:445-447).

Tolerances (fp32 fmaf / split-K order on the GPU vs fp32 numpy): a sampled categorical may flip at a near tie (top-2 gap
of the checker's perturbed logits < 1e-4; counted, printed and bounded); every trajectory WITHOUT a flip is compared
element-wise: deter |d| <= 2e-4 after 64 recurrent steps (5e-5 on the short goldens), logits |d| <= 4e-4, actions 2e-4.
"""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import assert_indices, cu, golden_initial, make_engine, perturbed_scores

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def base():
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    return c, P, make_engine(c, P, max_rows=1024, max_steps=64, max_tape_rows=16)


def test_observe_fp32_full_size(base):
    c, P, eng = base
    B, T = 16, 64
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=41)
    s0, d0 = golden_initial(c, B)
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed, action, (s0, d0), reset, u)
    st, dt, lg = [_np(x) for x in eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=0)]
    assert np.all(st.sum(-1) == 1.0) and set(np.unique(st)) <= {0.0, 1.0}
    idx = st.argmax(-1)
    # a flip changes everything after it in its row, so the near-tie accounting is done on first divergences only
    first = np.full(B, T, np.int64)
    for b in range(B):
        bad = np.nonzero((idx[b] != idx_o[b]).any(-1))[0]
        if bad.size:
            first[b] = bad[0]
    upto = np.arange(T)[None, :] <= first[:, None]              # steps up to and including the first flip
    sc = perturbed_scores(lg_o, u, c.unimix)
    assert_indices(idx[upto], idx_o[upto], sc[upto], 1e-4, 2e-3, "full-size observe (up to the first flip of each row)")
    ok = first == T
    print(f"full-size observe: {int(ok.sum())}/{B} trajectories without a flipped sample over {T} steps")
    assert ok.sum() * 2 >= B
    print("  |d deter| =", np.abs(dt[ok] - dt_o[ok]).max(), " |d logit| =", np.abs(lg[ok] - lg_o[ok]).max())
    np.testing.assert_allclose(dt[ok], dt_o[ok], atol=2e-4, rtol=0)
    np.testing.assert_allclose(lg[ok], lg_o[ok], atol=4e-4, rtol=0)
    # the steps BEFORE a row's first flip are still exact trajectories
    pre = np.arange(T)[None, :] < first[:, None]
    np.testing.assert_allclose(dt[pre], dt_o[pre], atol=2e-4, rtol=0)


def test_imagine_fp32_full_size(base):
    c, P, eng = base
    N, H = 1024, 16
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=43)
    feats_o, acts_o = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), H, u, noise)
    feats, acts = [_np(x) for x in eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=0)]
    idx = feats[..., :c.SK].reshape(N, H, c.S, c.K).argmax(-1)
    idx_o = feats_o[..., :c.SK].reshape(N, H, c.S, c.K).argmax(-1)
    np.testing.assert_array_equal(feats[:, 0], feats_o[:, 0])
    ok = (idx == idx_o).reshape(N, -1).all(1)
    n_flip = int((~ok).sum())
    print(f"full-size imagine: {N - n_flip}/{N} trajectories without a flipped sample over {H} steps")
    assert n_flip <= 0.02 * N, f"{n_flip} of {N} trajectories diverged"
    print("  |d deter| =", np.abs(feats[ok][..., c.SK:] - feats_o[ok][..., c.SK:]).max(),
          " |d action| =", np.abs(acts[ok] - acts_o[ok]).max())
    np.testing.assert_allclose(feats[ok][..., c.SK:], feats_o[ok][..., c.SK:], atol=2e-4, rtol=0)
    np.testing.assert_allclose(acts[ok], acts_o[ok], atol=2e-4, rtol=0)
    # diverged rows: exact up to the step of their first flip (the flip itself must be a near tie of the oracle)
    for n in np.nonzero(~ok)[0]:
        t = int(np.nonzero((idx[n] != idx_o[n]).any(-1))[0][0])
        np.testing.assert_allclose(feats[n, :t, c.SK:], feats_o[n, :t, c.SK:], atol=2e-4, rtol=0)
    # heads + lambda-return on the full tensor
    rew, cont, val, sval, wgt, ret = [_np(x) for x in eng.heads_lambda(cu(feats_o), 1 - 1 / c.horizon, c.lamb)]
    exp = O.heads_lambda(c, P["reward"], P["cont"], P["value"], P["slow_value"], feats_o)
    for got, want, name in zip((rew, cont, val, sval, wgt, ret), exp, ("reward", "cont", "value", "slow_value", "weight", "ret")):
        np.testing.assert_allclose(got, want, rtol=3e-4, atol=2e-5, err_msg=name)


@pytest.mark.parametrize("bad", [np.inf, np.nan])
@pytest.mark.parametrize("flags", [0, 1])
def test_observe_bwd_is_transparent_to_nonfinite_grads(base, bad, flags):
    """One non-finite upstream element at (b=3, t=5): every weight gradient on its path and the input gradients of row 3 at
    steps <= 5 must come out non-finite (nothing clamps, masks or zeroes it), while the other rows' input gradients, which
    do not depend on it, stay finite and equal to a run without the bad element."""
    c, P, eng = base
    B, T, b0, t0 = 8, 8, 3, 5
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=47)
    reset = np.zeros_like(reset)
    s0, d0 = golden_initial(c, B)
    g = np.random.Generator(np.random.Philox(49))
    c_dt = g.standard_normal((B, T, c.D), dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal((B, T, c.S, c.K), dtype=np.float32) * np.float32(0.1)
    names = eng.weight_names(0)

    def run(dd):
        eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=2 | flags)
        wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in names}
        d_embed, d_is, d_id = eng.observe_bwd(B, T, None, cu(dd), cu(c_lg), True, True, wg, flags=flags)
        torch.cuda.synchronize()
        return _np(d_embed).copy(), _np(d_id).copy(), {n: _np(v).copy() for n, v in wg.items()}

    de0, did0, wg0 = run(c_dt)
    assert np.isfinite(de0).all() and all(np.isfinite(v).all() for v in wg0.values())
    poisoned = c_dt.copy()
    poisoned[b0, t0, 17] = bad
    de1, did1, wg1 = run(poisoned)
    others = np.arange(B) != b0
    assert np.isfinite(de1[others]).all() and np.isfinite(did1[others]).all()
    tol = dict(rtol=0, atol=0) if flags == 0 else dict(rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(de1[others], de0[others], **tol)
    np.testing.assert_allclose(de1[b0, t0 + 1:], de0[b0, t0 + 1:], **tol)     # later steps do not see it either
    assert not np.isfinite(de1[b0, :t0 + 1]).all(), "d_embed of the poisoned row stayed finite"
    assert not np.isfinite(did1[b0]).all(), "d_initial_deter of the poisoned row stayed finite"
    hit = [n for n in names if not np.isfinite(wg1[n]).all()]
    print(f"non-finite weight grads ({bad}, flags={flags}): {len(hit)}/{len(names)}")
    # deter at (b0, t0) is produced by the deter net and consumed by the posterior head: all of both are on the path
    for n in names:
        if n.startswith(("_deter_net.", "_obs_net.")):
            assert n in hit, f"{n}: gradient stayed finite although a non-finite upstream gradient reaches it"
