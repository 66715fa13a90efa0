"""Debug driver of the CNN encoder kernels: per-case difference against the numpy oracle (test infrastructure; python
profiles/cnn_debug.py [time])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cnn_oracle as CO
from safe_dreamer_b200.encoder import CnnEngine

for tag, hw, depth, n in (("tiny", 32, 4, 2), ("base", 64, 16, 1)):
    depths = [depth * m for m in (2, 3, 4, 4)]
    P = CO.encoder_params(depths, 3, 5, seed=77 + hw)
    for L in (1, 2, 3, 4):
        eng = CnnEngine(hw, hw, 3, depths[:L], 5, max_frames=2 * n)
        ts = []
        for i in range(L):
            ts += [P[f"layers.{4 * i}.weight"], P[f"layers.{4 * i}.bias"], P[f"layers.{4 * i + 2}.weight"]]
        eng.set_weights([torch.from_numpy(t).cuda() for t in ts])
        rng = np.random.Generator(np.random.Philox(5150 + hw))
        obs = rng.random((n, 2, hw, hw, 3), dtype=np.float32)
        emb = eng.forward(torch.from_numpy(obs).cuda()).cpu().numpy()
        ref = CO.encoder_fwd(P, obs, n_layers=L)
        d = np.abs(emb - ref)
        print(f"{tag} layers={L}: shape {emb.shape} max|d| {d.max():.4f} mean|d| {d.mean():.5f} |ref| mean {np.abs(ref).mean():.3f} max {np.abs(ref).max():.3f}", flush=True)
if len(sys.argv) > 1:
    depths = [32, 48, 64, 64]
    P = CO.encoder_params(depths, 3, 5, seed=1)
    eng = CnnEngine(64, 64, 3, depths, 5, max_frames=1024)
    ts = []
    for i in range(4):
        ts += [P[f"layers.{4 * i}.weight"], P[f"layers.{4 * i}.bias"], P[f"layers.{4 * i + 2}.weight"]]
    eng.set_weights([torch.from_numpy(t).cuda() for t in ts])
    obs = torch.rand(1024, 64, 64, 3, device="cuda")
    for _ in range(3):
        eng.forward(obs)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        eng.forward(obs)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"encoder forward, 1024 frames 64x64x3: {ms:.3f} ms  ({0.1507 * 1024 / ms:.1f} TFLOP/s)")
    for L in (1, 2, 3):
        e2 = CnnEngine(64, 64, 3, depths[:L], 5, max_frames=1024)
        e2.set_weights([torch.from_numpy(t).cuda() for t in ts[:3 * L]])
        for _ in range(3):
            e2.forward(obs)
        torch.cuda.synchronize(); a.record()
        for _ in range(20):
            e2.forward(obs)
        b.record(); b.synchronize()
        print(f"  first {L} stage(s): {a.elapsed_time(b) / 20:.3f} ms")
if "bwd" in sys.argv:
    golden = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cnn_encoder.npz"))
    for tag, hw, depth, n in (("tiny", 32, 4, 2), ("base", 64, 16, 1)):
        depths = [depth * m for m in (2, 3, 4, 4)]
        P = CO.encoder_params(depths, 3, 5, seed=77 + hw)
        names = []
        for i in range(4):
            names += [f"layers.{4 * i}.weight", f"layers.{4 * i}.bias", f"layers.{4 * i + 2}.weight"]
        eng = CnnEngine(hw, hw, 3, depths, 5, max_frames=2 * n, max_tape_frames=2 * n)
        eng.set_weights([torch.from_numpy(P[k]).cuda() for k in names])
        rng = np.random.Generator(np.random.Philox(5150 + hw))
        obs = rng.random((n, 2, hw, hw, 3), dtype=np.float32)
        eng.forward(torch.from_numpy(obs).cuda(), tape=True)
        wg = [torch.zeros(P[k].shape, device="cuda") for k in names]
        d_obs = eng.backward(torch.from_numpy(golden[f"{tag}/g"]).cuda(), want_obs_grad=True, weight_grads=wg)
        torch.cuda.synchronize()
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        print(f"{tag}: d_obs rel err {rel(d_obs.cpu().numpy().reshape(golden[f'{tag}/d_obs'].shape), golden[f'{tag}/d_obs']):.4f}", flush=True)
        for k, t in zip(names, wg):
            print(f"{tag}: {k:24s} rel err {rel(t.cpu().numpy(), golden[f'{tag}/grad/{k}']):.4f}  |ref| {np.linalg.norm(golden[f'{tag}/grad/{k}']):.3e}", flush=True)
if "bwdtime" in sys.argv:
    depths = [32, 48, 64, 64]
    P = CO.encoder_params(depths, 3, 5, seed=1)
    names = []
    for i in range(4):
        names += [f"layers.{4 * i}.weight", f"layers.{4 * i}.bias", f"layers.{4 * i + 2}.weight"]
    eng = CnnEngine(64, 64, 3, depths, 5, max_frames=1024, max_tape_frames=1024)
    eng.set_weights([torch.from_numpy(P[k]).cuda() for k in names])
    obs = torch.rand(1024, 64, 64, 3, device="cuda")
    g = torch.randn(1024, 1024, device="cuda")
    wg = [torch.zeros(P[k].shape, device="cuda") for k in names]
    def step(dobs):
        eng.forward(obs, tape=True)
        eng.backward(g, want_obs_grad=dobs, weight_grads=wg)
    for dobs in (False, True):
        for _ in range(3):
            step(dobs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(10):
            step(dobs)
        b.record(); b.synchronize()
        print(f"encoder forward(tape) + backward (d_obs={dobs}), 1024 frames: {a.elapsed_time(b) / 10:.3f} ms", flush=True)
