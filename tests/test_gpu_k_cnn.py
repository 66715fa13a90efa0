"""GPU parity of the CNN encoder (networks.py:192-234) through the C ABI: implicit-GEMM convolutions on tcgen05 (bf16
operands, fp32 accumulate) with the MaxPool / RMSNorm / SiLU epilogue, against the numpy oracle and the golden written by
the reference's own ConvEncoder (tests/golden/cnn_encoder.npz).

Tolerance: activations cross the four stages as bf16 and every contraction has bf16 operands (the reference's own GPU path
is fp16 autocast), so the embedding is compared at |d| <= 0.03 + 0.03 |ref| with a mean |d| <= 4e-3; a pooling arg-max may
differ from the fp32 oracle only where the two best window members are closer than the bf16 rounding of the conv output."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn_oracle as CO

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("tiny", 32, 4, 2), ("base", 64, 16, 1)]


def _inputs(hw, n):
    rng = np.random.Generator(np.random.Philox(5150 + hw))
    return rng.random((n, 2, hw, hw, 3), dtype=np.float32)


def _engine(hw, depth, frames, tape=0):
    from safe_dreamer_b200.encoder import CnnEngine
    depths = [depth * m for m in (2, 3, 4, 4)]
    P = CO.encoder_params(depths, 3, 5, seed=77 + hw)
    eng = CnnEngine(hw, hw, 3, depths, 5, max_frames=frames, max_tape_frames=tape)
    ts = []
    for i in range(4):
        ts += [P[f"layers.{4 * i}.weight"], P[f"layers.{4 * i}.bias"], P[f"layers.{4 * i + 2}.weight"]]
    eng.set_weights([torch.from_numpy(t).cuda() for t in ts])
    return eng, P


@pytest.mark.parametrize("tag,hw,depth,n", CASES)
def test_forward_matches_oracle_and_reference(tag, hw, depth, n):
    golden = np.load(os.path.join(HERE, "golden", "cnn_encoder.npz"))
    eng, P = _engine(hw, depth, 2 * n)
    obs = _inputs(hw, n)
    emb = eng.forward(torch.from_numpy(obs).cuda()).cpu().numpy()
    ref = golden[f"{tag}/emb"]
    assert emb.shape == ref.shape
    d = np.abs(emb - ref)
    print(f"{tag}: max|d| {d.max():.4f} mean|d| {d.mean():.5f} (|ref| max {np.abs(ref).max():.3f})")
    assert np.all(d <= 0.03 + 0.03 * np.abs(ref)) and d.mean() <= 4e-3
    emb_o = CO.encoder_fwd(P, obs)
    assert np.abs(emb - emb_o).mean() <= 4e-3


def test_forward_full_batch_properties():
    """1024 frames (B=16, T=64) at base sizes: finite, frame-order equivariant, identical frames give identical rows."""
    eng, P = _engine(64, 16, 1024)
    g = torch.Generator(device="cuda").manual_seed(5)
    obs = torch.rand(1024, 64, 64, 3, device="cuda", generator=g)
    obs[7] = obs[900]
    emb = eng.forward(obs)
    assert emb.shape == (1024, 1024) and bool(torch.isfinite(emb).all())
    assert torch.equal(emb[7], emb[900])
    perm = torch.randperm(1024, device="cuda", generator=g)
    emb_p = eng.forward(obs[perm])
    assert torch.equal(emb_p, emb[perm])
    sub = CO.encoder_fwd(P, obs[:3].cpu().numpy())
    assert np.abs(emb[:3].cpu().numpy() - sub).mean() <= 4e-3
