"""Barlow-twins loss of dreamer.py:525-532: numpy oracle (value + gradient) vs the reference statements under torch
autograd (golden), and the CUDA path behind sd_barlow_loss (dreamer_ops.barlow_loss) vs both, incl. the full (1024, 1024) size."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    # restated from tests/golden/make_golden.py:barlow_inputs
    rng = np.random.Generator(np.random.Philox(2718))
    N, E = 64, 48
    z = rng.standard_normal((N, E), dtype=np.float32)
    x1 = (z * np.float32(1.5) + np.float32(0.3) * rng.standard_normal((N, E), dtype=np.float32) + np.float32(0.7)).astype(np.float32)
    x2 = (z @ (np.eye(E, dtype=np.float32) + np.float32(0.1) * rng.standard_normal((E, E), dtype=np.float32))
          + np.float32(0.5) * rng.standard_normal((N, E), dtype=np.float32)).astype(np.float32)
    return x1, x2


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "barlow.npz"))


def test_oracle_barlow_matches_reference_autograd(golden):
    x1, x2 = _inputs()
    loss, dx = O.barlow_loss(x1, x2, float(golden["lambd"]))
    np.testing.assert_allclose(loss, golden["loss"], rtol=2e-5)
    np.testing.assert_allclose(dx, golden["d_x1"], rtol=2e-3, atol=2e-6 * float(np.abs(golden["d_x1"]).max()))


@pytest.mark.gpu
def test_cuda_barlow(golden):
    import torch
    from safe_dreamer_b200.dreamer_ops import barlow_loss
    x1, x2 = _inputs()
    a = torch.from_numpy(x1).cuda().requires_grad_(True)
    loss = barlow_loss(a, torch.from_numpy(x2).cuda(), float(golden["lambd"]))
    (3.0 * loss).backward()
    np.testing.assert_allclose(float(loss.detach()), golden["loss"], rtol=2e-5)
    np.testing.assert_allclose(a.grad.cpu().numpy() / 3.0, golden["d_x1"], rtol=2e-3, atol=2e-6 * float(np.abs(golden["d_x1"]).max()))
    # the real size: (B*T, E) = (1024, 1024), against the fp64 oracle
    rng = np.random.Generator(np.random.Philox(99))
    z = rng.standard_normal((1024, 1024), dtype=np.float32)
    y1 = (z + np.float32(0.5) * rng.standard_normal((1024, 1024), dtype=np.float32)).astype(np.float32)
    y2 = (z * np.float32(2.0) + np.float32(1.0) + rng.standard_normal((1024, 1024), dtype=np.float32)).astype(np.float32)
    loss_o, dx_o = O.barlow_loss(y1.astype(np.float64), y2.astype(np.float64), 5e-4)
    b = torch.from_numpy(y1).cuda().requires_grad_(True)
    l2 = barlow_loss(b, torch.from_numpy(y2).cuda(), 5e-4)
    l2.backward()
    np.testing.assert_allclose(float(l2.detach()), loss_o, rtol=1e-4)
    np.testing.assert_allclose(b.grad.cpu().numpy(), dx_o, rtol=5e-3, atol=5e-5 * float(np.abs(dx_o).max()))
    # no-grad call: value only
    with torch.no_grad():
        l3 = barlow_loss(torch.from_numpy(y1).cuda(), torch.from_numpy(y2).cuda(), 5e-4)
    assert float(l3) == float(l2.detach())
