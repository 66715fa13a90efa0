"""TwoHot.log_prob (distributions.py:100-129, bins from symexp_twohot :242-251): numpy oracle (forward + backward) vs the real
reference with autograd (golden), and the CUDA kernels behind sd_twohot_logprob / sd_twohot_logprob_bwd vs both."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs(bins):
    # restated from tests/golden/make_golden.py:twohot_inputs
    rng = np.random.Generator(np.random.Philox(777))
    R, n = 96, len(bins)
    logits = (rng.standard_normal((R, n), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    target = (rng.standard_normal(R, dtype=np.float32) * np.float32(30.0)).astype(np.float32)
    target[:8] = bins[[0, 1, n // 2, n // 2 + 1, n - 2, n - 1, 17, 200]]
    target[8:12] = np.array([-1e9, 1e9, bins[0] * 2, bins[-1] * 2], np.float32)
    target[12:16] = np.array([0.0, 1e-6, -1e-6, 0.5], np.float32)
    g = rng.standard_normal(R, dtype=np.float32)
    return logits, target.astype(np.float32), g


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "twohot_logprob.npz"))


def test_oracle_twohot_logprob_matches_reference(golden):
    bins = golden["bins"]                                   # the reference's own bin positions (torch expm1)
    np.testing.assert_allclose(O.twohot_bins(255), bins, rtol=4e-6)
    logits, target, g = _inputs(bins)
    lp, mixed = O.twohot_logprob(logits, bins, target)
    np.testing.assert_allclose(mixed.sum(-1), 1.0, atol=1e-6)
    np.testing.assert_allclose(lp, golden["log_prob"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(O.twohot_logprob_bwd(logits, mixed, g), golden["d_logits"], rtol=2e-5, atol=2e-7)


@pytest.mark.gpu
def test_cuda_twohot_logprob(golden):
    import torch
    from safe_dreamer_b200.distributions import TwoHot, symexp_twohot
    bins = golden["bins"]
    logits, target, g = _inputs(bins)
    lg = torch.from_numpy(logits).cuda().requires_grad_(True)
    np.testing.assert_allclose(symexp_twohot(lg, 255).bins.cpu().numpy(), bins, rtol=4e-6)
    dist = TwoHot(lg, torch.from_numpy(bins).cuda())          # exact-hit targets need the very same bin values
    lp = dist.log_prob(torch.from_numpy(target).cuda()[..., None])
    np.testing.assert_allclose(lp.detach().cpu().numpy(), golden["log_prob"], rtol=2e-6, atol=2e-6)
    (lp * torch.from_numpy(g).cuda()).sum().backward()
    np.testing.assert_allclose(lg.grad.cpu().numpy(), golden["d_logits"], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(dist.mode().detach().cpu().numpy(), golden["mode"], rtol=3e-4, atol=1e-3)
    # (B, T, bins) shaped logits / (B, T, 1) targets as used at dreamer.py:571,654-660
    lg3 = torch.from_numpy(logits.reshape(8, 12, -1)).cuda()
    lp3 = TwoHot(lg3, torch.from_numpy(bins).cuda()).log_prob(torch.from_numpy(target.reshape(8, 12, 1)).cuda())
    assert lp3.shape == (8, 12)
    np.testing.assert_allclose(lp3.reshape(-1).cpu().numpy(), golden["log_prob"], rtol=2e-6, atol=2e-6)
