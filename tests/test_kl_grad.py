"""Backward of RSSM.kl_loss (rssm.py:222-230): numpy oracle vs autograd of the real reference (golden), and the CUDA kernel
behind sd_kl_loss_bwd (+ the module-level autograd wiring) vs both."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    # restated from tests/golden/make_golden.py:kl_grad_inputs (that script imports torch + the reference)
    rng = np.random.Generator(np.random.Philox(4242))
    R, S, K = 12, 32, 16
    post = rng.standard_normal((R, S, K), dtype=np.float32) * np.linspace(0.05, 1.5, R, dtype=np.float32)[:, None, None]
    prior = rng.standard_normal((R, S, K), dtype=np.float32) * np.linspace(0.05, 1.5, R, dtype=np.float32)[:, None, None]
    g_dyn = rng.standard_normal(R, dtype=np.float32)
    g_rep = rng.standard_normal(R, dtype=np.float32)
    return post.astype(np.float32), prior.astype(np.float32), g_dyn, g_rep


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "kl_grad.npz"))


def test_oracle_kl_backward_matches_reference_autograd(golden):
    post, prior, g_dyn, g_rep = _inputs()
    dyn, rep = O.kl_loss(post, prior, 1.0)
    np.testing.assert_allclose(dyn, golden["dyn"], rtol=2e-5, atol=1e-6)
    assert 0 < int((golden["dyn"] <= 1.0).sum()) < post.shape[0]      # the fixture exercises both sides of the clip
    d_post, d_prior = O.kl_loss_bwd(post, prior, 1.0, g_dyn, g_rep)
    np.testing.assert_allclose(d_post, golden["d_post"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(d_prior, golden["d_prior"], rtol=2e-4, atol=2e-6)


@pytest.mark.gpu
def test_cuda_kl_backward(golden):
    import torch
    from tests.helpers import cu, make_engine
    c = O.Cfg()
    P = O.init_params(c, seed=0)
    eng = make_engine(c, P, max_rows=16, max_steps=2)
    post, prior, g_dyn, g_rep = _inputs()
    dyn, rep = eng.kl_loss(cu(post), cu(prior), 1.0)
    np.testing.assert_allclose(dyn.cpu().numpy(), golden["dyn"], rtol=2e-5, atol=1e-6)
    d_post, d_prior = eng.kl_loss_bwd(cu(post), cu(prior), 1.0, cu(g_dyn), cu(g_rep))
    np.testing.assert_allclose(d_post.cpu().numpy(), golden["d_post"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(d_prior.cpu().numpy(), golden["d_prior"], rtol=2e-4, atol=2e-6)
    # module surface: RSSM.kl_loss is differentiable through the CUDA backward
    from types import SimpleNamespace as NS
    from safe_dreamer_b200.rssm import RSSM
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cuda", obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    rssm = RSSM(cfg, c.E, c.A).cuda()
    a, b = cu(post).requires_grad_(True), cu(prior).requires_grad_(True)
    dyn_t, rep_t = rssm.kl_loss(a, b, 1.0)
    ((dyn_t * cu(g_dyn)).sum() + (rep_t * cu(g_rep)).sum()).backward()
    np.testing.assert_allclose(a.grad.cpu().numpy(), golden["d_post"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(b.grad.cpu().numpy(), golden["d_prior"], rtol=2e-4, atol=2e-6)
    # batch-shaped (B, T, S, K) inputs and a no-grad call keep working
    with torch.no_grad():
        d2, r2 = rssm.kl_loss(cu(post).reshape(3, 4, c.S, c.K), cu(prior).reshape(3, 4, c.S, c.K), 1.0)
    assert d2.shape == (3, 4)
    np.testing.assert_allclose(d2.reshape(-1).cpu().numpy(), golden["dyn"], rtol=2e-5, atol=1e-6)
