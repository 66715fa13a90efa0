"""ReturnEMA (networks.py:405-422): oracle vs goldens produced by the real reference (CPU), and the CUDA kernel
behind the C ABI (sd_return_ema) vs both, bit for bit (exact order statistics + torch's fp32 rank / lerp arithmetic)."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs(case, call):
    # same seeded generator as tests/golden/make_golden.py:return_ema_inputs (restated: that script imports torch + the reference)
    n_rows, scale, shift, ties = [(1024, 3.0, 1.0, False), (7, 1.0, 0.0, False), (1, 1.0, 2.0, False), (333, 50.0, -20.0, True),
                                  (9001, 0.01, 0.0, False)][case]
    rng = np.random.Generator(np.random.Philox(900 + 17 * case + call))
    x = (rng.standard_normal((n_rows, 15, 1), dtype=np.float32) * np.float32(scale) + np.float32(shift + 0.5 * call)).astype(np.float32)
    return np.round(x) if ties else x


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "return_ema.npz"))


def test_oracle_matches_reference_goldens(golden):
    for case in range(5):
        ema = np.zeros(2, np.float32)
        for call in range(4):
            ema, off, scl = O.return_ema(_inputs(case, call), ema, 1e-2)
            np.testing.assert_array_equal(ema, golden[f"c{case}_{call}_ema"], err_msg=f"case {case} call {call}")
            assert off == golden[f"c{case}_{call}_offset"] and scl == golden[f"c{case}_{call}_scale"]


@pytest.mark.gpu
def test_cuda_return_ema_bit_exact(golden):
    import torch
    from safe_dreamer_b200.networks import ReturnEMA
    for case in range(5):
        m = ReturnEMA(device="cuda")
        ema_o = np.zeros(2, np.float32)
        for call in range(4):
            x = _inputs(case, call)
            off, scl = m(torch.from_numpy(x).cuda())
            ema_o, off_o, scl_o = O.return_ema(x, ema_o, 1e-2)
            got = m.ema_vals.cpu().numpy()
            np.testing.assert_array_equal(got, golden[f"c{case}_{call}_ema"], err_msg=f"case {case} call {call}")
            np.testing.assert_array_equal(got, ema_o)
            assert float(off) == float(golden[f"c{case}_{call}_offset"]) and float(scl) == float(golden[f"c{case}_{call}_scale"])
    # a large tensor with negative / repeated / extreme values against torch.quantile itself
    g = torch.Generator(device="cuda").manual_seed(5)
    big = torch.randn(8192 * 15, device="cuda", generator=g) * 100.0
    big[::7] = -3.5
    big[1::1001] = 1e30
    m = ReturnEMA(device="cuda", alpha=0.25)
    off, scl = m(big.reshape(8192, 15, 1))
    q = torch.quantile(big, torch.tensor([0.05, 0.95], device="cuda"))
    want = 0.25 * q + 0.75 * torch.zeros(2, device="cuda")
    assert torch.equal(m.ema_vals, want)
    assert float(scl) == float(torch.clip(want[1] - want[0], min=1.0)) and float(off) == float(want[0])
    assert set(m.state_dict().keys()) == {"ema_vals"}
