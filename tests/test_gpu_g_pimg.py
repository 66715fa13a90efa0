"""GPU tests of the bf16 imagination rollout (Dreamer._imagine, dreamer.py:673-692) over the FULL horizon: the
persistent team-resident kernel (csrc/sd_pimg.cuh; the default for the base architecture and the path the headline
benchmark runs) and the layer-by-layer tcgen05 launch sequence (SD_FLAG_LAYERWISE).

Both are bf16-operand / fp32-accumulate, so element-wise parity is checked teacher-forced (test_gpu_b_tc.py and the
first test below) and over the full horizon statistically:
  * the two bf16 paths against each other and against the fp32 path of the same library:
    step-0 actions, step-1 deter, per-step index agreement, per-step action / deter statistics, and the
    mean / 5 % / 95 % quantiles of the lambda-return computed from the rolled-out features
  * structure: stoch rows exact one-hots, feats[:, 0] the start state, deter a convex mix (|d| <= 1)
  * determinism: direct launch == CUDA-graph replay == second replay, bit for bit; a row's trajectory does not
    depend on how many other rows are in the call (ragged row counts, several 128-row groups per team)
"""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import cu, make_engine

pytestmark = pytest.mark.gpu
BF16, GRAPH, PERSIST, LAYERWISE = 1, 4, 32, 64


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def full():
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    eng = make_engine(c, P, max_rows=1280, max_steps=16)
    return c, P, eng


def _idx(feats, c):
    return feats[..., :c.SK].reshape(*feats.shape[:-1], c.S, c.K).argmax(-1)


def test_pimg_matches_layerwise_and_oracle_first_steps(full):
    c, P, eng = full
    N, H = 384, 3
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=31)
    fp, ap = [_np(x).copy() for x in eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=BF16 | PERSIST)]
    fl, al = [_np(x).copy() for x in eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=BF16 | LAYERWISE)]
    torch.cuda.synchronize()
    np.testing.assert_array_equal(fp[:, 0], fl[:, 0])
    print("pimg vs layerwise: |dact0| =", np.abs(ap[:, 0] - al[:, 0]).max(), " |ddeter1| =", np.abs(fp[:, 1, c.SK:] - fl[:, 1, c.SK:]).max())
    assert np.abs(ap[:, 0] - al[:, 0]).max() <= 0.03
    assert np.abs(fp[:, 1, c.SK:] - fl[:, 1, c.SK:]).max() <= 0.04
    mis = (_idx(fp[:, 1], c) != _idx(fl[:, 1], c)).mean()
    print("pimg vs layerwise: step-1 index mismatch rate", mis)
    assert mis <= 0.02
    # teacher-forced against the fp32 oracle
    feats_o, acts_o = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), 2, u, noise)
    assert np.abs(ap[:, 0] - acts_o[:, 0]).max() <= 0.03
    assert np.abs(fp[:, 1, c.SK:] - feats_o[:, 1, c.SK:]).max() <= 0.04
    mis_o = (_idx(fp[:, 1], c) != _idx(feats_o[:, 1], c)).mean()
    print("pimg vs oracle: step-1 index mismatch rate", mis_o)
    assert mis_o <= 0.02


def test_pimg_fullsize_structure_and_determinism(full):
    c, P, eng = full
    N, H = 1024, 16
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=33)
    ins = [cu(x) for x in (st0, dt0, u, noise)]
    outs = []
    for flags in (BF16 | PERSIST, BF16 | PERSIST | GRAPH, BF16 | PERSIST | GRAPH):
        f, a = eng.imagine(*ins, H, flags=flags)
        torch.cuda.synchronize()
        outs.append((_np(f).copy(), _np(a).copy()))
    for f, a in outs[1:]:
        np.testing.assert_array_equal(f, outs[0][0])
        np.testing.assert_array_equal(a, outs[0][1])
    f, a = outs[0]
    np.testing.assert_array_equal(f[:, 0, :c.SK], st0.reshape(N, -1))
    np.testing.assert_array_equal(f[:, 0, c.SK:], dt0)
    oh = f[..., :c.SK].reshape(N, H, c.S, c.K)
    assert set(np.unique(oh)) == {0.0, 1.0} and np.all(oh.sum(-1) == 1.0)
    assert np.isfinite(f).all() and np.isfinite(a).all()
    assert np.abs(f[..., c.SK:]).max() <= 1.0 + 1e-5
    # ragged row count spanning several groups per team (1100 rows = 9 groups on 8-9 teams): rows are independent
    M = 1100
    st1, dt1, u1, n1 = O.synth_imagine_inputs(c, M, H, seed=33)
    np.testing.assert_array_equal(st1[:N], st0)        # Philox stream: the first N rows are the same inputs
    if np.array_equal(u1[:N], u) and np.array_equal(n1[:N], noise) and np.array_equal(dt1[:N], dt0):
        f2, a2 = eng.imagine(cu(st1), cu(dt1), cu(u1), cu(n1), H, flags=BF16 | PERSIST)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(_np(f2)[:N], f)
        np.testing.assert_array_equal(_np(a2)[:N], a)
    else:   # the generator interleaves rows: cut the big batch instead
        f2, a2 = eng.imagine(cu(st1), cu(dt1), cu(u1), cu(n1), H, flags=BF16 | PERSIST)
        f3, a3 = eng.imagine(cu(st1[:777]), cu(dt1[:777]), cu(u1[:777]), cu(n1[:777]), H, flags=BF16 | PERSIST)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(_np(f3), _np(f2)[:777])
        np.testing.assert_array_equal(_np(a3), _np(a2)[:777])


def test_pimg_h16_statistical_parity(full):
    """Full-horizon parity of the bf16 persistent rollout against the fp32 path (which is pinned to the oracle / the
    reference goldens element-wise): trajectories diverge after the first flipped sample, so the comparison is
    distributional, with these stated bounds over 1024 rows x 16 steps."""
    c, P, eng = full
    N, H = 1024, 16
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=35)
    ins = [cu(x) for x in (st0, dt0, u, noise)]
    f32, a32 = [_np(x).copy() for x in eng.imagine(*ins, H, flags=0)]
    fb, ab = [_np(x).copy() for x in eng.imagine(*ins, H, flags=BF16 | PERSIST)]
    fw, aw = [_np(x).copy() for x in eng.imagine(*ins, H, flags=BF16 | LAYERWISE)]
    torch.cuda.synchronize()
    flip = [(_idx(fb[:, t], c) != _idx(f32[:, t], c)).mean() for t in range(H)]
    flip_w = [(_idx(fw[:, t], c) != _idx(f32[:, t], c)).mean() for t in range(H)]
    print("index mismatch rate per step, persistent vs fp32:", np.round(flip, 4).tolist())
    print("index mismatch rate per step, layerwise  vs fp32:", np.round(flip_w, 4).tolist())
    assert flip[0] == 0.0 and flip[1] <= 0.02           # one step of bf16 rounding: near ties only
    assert flip[-1] <= max(0.35, 1.5 * flip_w[-1])      # chaotic growth, no worse than the layer-by-layer bf16 path
    assert flip_w[0] == 0.0 and flip_w[1] <= 0.02 and flip_w[-1] <= 0.35
    disc = 1 - 1 / c.horizon
    r32 = _np(eng.heads_lambda(cu(f32), disc, c.lamb, flags=0)[-1])
    q32 = np.quantile(r32, [0.05, 0.5, 0.95])
    scale = max(1.0, float(q32[2] - q32[0]))
    for name, fx, ax in (("persistent", fb, ab), ("layer-by-layer", fw, aw)):
        for t in range(H):
            da = np.abs(ax[:, t].mean(0) - a32[:, t].mean(0)).max()
            ds = np.abs(ax[:, t].std(0) - a32[:, t].std(0)).max()
            dd = abs(fx[:, t, c.SK:].mean() - f32[:, t, c.SK:].mean())
            dr = abs(np.sqrt((fx[:, t, c.SK:] ** 2).mean()) - np.sqrt((f32[:, t, c.SK:] ** 2).mean()))
            assert da <= 0.05 and ds <= 0.05, (name, t, da, ds)
            assert dd <= 0.01 and dr <= 0.01, (name, t, dd, dr)
        rb = _np(eng.heads_lambda(cu(fx), disc, c.lamb, flags=0)[-1])
        qb = np.quantile(rb, [0.05, 0.5, 0.95])
        print(f"lambda-return mean / q05 / q50 / q95  fp32: {r32.mean():.5f} {q32}   bf16 {name}: {rb.mean():.5f} {qb}")
        assert abs(rb.mean() - r32.mean()) <= 0.03 * scale, name
        assert np.abs(qb - q32).max() <= 0.06 * scale, name


def test_layerwise_chain_rollout_large_n_matches_small_n_and_oracle():
    """From 4096 rows the launch-sequence rollout runs the actor and img_net trunks as row-tile resident chain kernels
    (csrc/sd_chain.cuh; SD_CHAIN unset).  Rows are independent, so the first rows of a 4096-row call must agree with the same
    rows run as a 384-row call (no chain kernels) within the teacher-forced bf16 bounds, and with the fp32 oracle."""
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    N, M, H = 4096, 384, 3
    eng = make_engine(c, P, max_rows=N, max_steps=H)
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=41)
    fc, ac = [_np(x).copy() for x in eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=BF16 | LAYERWISE)]
    fs, as_ = [_np(x).copy() for x in eng.imagine(cu(st0[:M]), cu(dt0[:M]), cu(u[:M]), cu(noise[:M]), H, flags=BF16 | LAYERWISE)]
    torch.cuda.synchronize()
    oh = fc[..., :c.SK].reshape(N, H, c.S, c.K)
    assert set(np.unique(oh)) == {0.0, 1.0} and np.all(oh.sum(-1) == 1.0)
    assert np.isfinite(fc).all() and np.isfinite(ac).all()
    np.testing.assert_array_equal(fc[:, 0, :c.SK], st0.reshape(N, -1))
    print("chain vs launch sequence: |dact0| =", np.abs(ac[:M, 0] - as_[:, 0]).max(), " |ddeter1| =", np.abs(fc[:M, 1, c.SK:] - fs[:, 1, c.SK:]).max())
    assert np.abs(ac[:M, 0] - as_[:, 0]).max() <= 0.03
    assert np.abs(fc[:M, 1, c.SK:] - fs[:, 1, c.SK:]).max() <= 0.04
    mis = (_idx(fc[:M, 1], c) != _idx(fs[:, 1], c)).mean()
    print("chain vs launch sequence: step-1 index mismatch rate", mis)
    assert mis <= 0.02
    feats_o, acts_o = O.imagine(c, P["rssm"], P["actor"], (st0[:M], dt0[:M]), 2, u[:M], noise[:M])
    assert np.abs(ac[:M, 0] - acts_o[:, 0]).max() <= 0.03
    assert np.abs(fc[:M, 1, c.SK:] - feats_o[:, 1, c.SK:]).max() <= 0.04
    mis_o = (_idx(fc[:M, 1], c) != _idx(feats_o[:, 1], c)).mean()
    print("chain vs oracle: step-1 index mismatch rate", mis_o)
    assert mis_o <= 0.02
    # the last row tile (rows 3968 .. 4095) went through the same kernels: same statistics as the first one
    assert abs(np.abs(ac[-128:]).mean() - np.abs(ac[:128]).mean()) <= 0.1
