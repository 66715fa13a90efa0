"""CNN encoder (networks.py:59-98,192-234): the numpy oracle against the golden written by the reference's own
ConvEncoder under torch autograd (tests/golden/make_golden.py:run_cnn_encoder).  There is no CUDA implementation of this row
yet; this pins the oracle it will be built against (forward embedding, d(obs), all weight / bias / RMS-scale gradients)."""
import os

import numpy as np
import pytest

from oracle import cnn_oracle as CO

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "cnn_encoder.npz"))


def _inputs(hw, n):
    # restated from tests/golden/make_golden.py:cnn_encoder_inputs
    rng = np.random.Generator(np.random.Philox(5150 + hw))
    return rng.random((n, 2, hw, hw, 3), dtype=np.float32)


@pytest.mark.parametrize("tag,hw,depth,n", [("tiny", 32, 4, 2), ("base", 64, 16, 1)])
def test_encoder_matches_reference(golden, tag, hw, depth, n):
    P = CO.encoder_params([depth * m for m in (2, 3, 4, 4)], 3, 5, seed=77 + hw)
    obs = _inputs(hw, n)
    tape = []
    emb = CO.encoder_fwd(P, obs, tape=tape)
    ref = golden[f"{tag}/emb"]
    assert emb.shape == ref.shape == (n, 2, depth * 4 * (hw // 16) ** 2)
    np.testing.assert_allclose(emb, ref, rtol=2e-4, atol=2e-5)
    g = golden[f"{tag}/g"]
    d_obs, G = CO.encoder_bwd(P, tape, g.reshape(-1, g.shape[-1]))
    ref_d = golden[f"{tag}/d_obs"].reshape(d_obs.shape)
    np.testing.assert_allclose(d_obs, ref_d, rtol=2e-3, atol=2e-5 * float(np.abs(ref_d).max()))
    for name in P:
        r = golden[f"{tag}/grad/{name}"]
        np.testing.assert_allclose(G[name], r, rtol=2e-3, atol=2e-5 * float(np.abs(r).max()), err_msg=name)


def test_same_padding_and_pool_edges():
    """Conv2dSamePad pads k-1 split floor/ceil (networks.py:62-75); MaxPool2d(2,2) drops an odd last row/column."""
    assert CO.same_pad(64, 5) == 4 and CO.same_pad(7, 5) == 4 and CO.same_pad(7, 4) == 3
    x = np.arange(2 * 5 * 5 * 1, dtype=np.float32).reshape(2, 5, 5, 1)
    p, arg = CO.maxpool2(x)
    assert p.shape == (2, 2, 2, 1) and (arg == 3).all()
    np.testing.assert_array_equal(p[0, :, :, 0], [[6, 8], [16, 18]])
    dx = CO.maxpool2_bwd(np.ones_like(p), arg, x.shape)
    assert dx.sum() == p.size and dx[0, 4].sum() == 0 and dx[0, :, 4].sum() == 0
    # even kernel: one more pad cell after than before
    xp, (p0, p1) = CO._pad_same(np.zeros((1, 7, 7, 1), np.float32), 4)
    assert xp.shape == (1, 10, 10, 1) and (p0, p1) == (1, 1)


def test_bf16_operand_model_stays_close_to_fp32(golden):
    """oracle.round_bf16 models the CUDA path's bf16 convolution operands: nearest-even, idempotent, and the rounded encoder
    stays within bf16 distance of the reference embedding (the GPU parity tests compare against this rounded oracle)."""
    x = np.array([1.0, 1.00390625, 1.001953125, -3.1415927, 1e-30, 65504.0], np.float32)
    r = CO.round_bf16(x)
    np.testing.assert_array_equal(CO.round_bf16(r), r)
    assert r[0] == 1.0 and r[1] == 1.0 and r[2] == 1.0          # ties / below half an ulp (2^-8) round to even
    assert abs(r[3] + 3.140625) < 1e-6
    P = CO.encoder_params([8, 12, 16, 16], 3, 5, seed=77 + 32)
    emb = CO.encoder_fwd(P, _inputs(32, 2), rnd=CO.round_bf16)
    d = np.abs(emb - golden["tiny/emb"])
    assert d.max() <= 0.03 and d.mean() <= 3e-3
