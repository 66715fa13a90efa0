"""Drop-in proof: the REFERENCE's own `Dreamer` (baseline/_ref, unmodified dreamer.py) runs on top of this
library after `safe_dreamer_b200.install(agent)`.

One `_cal_grad` call (dreamer.py:452-670: encoder -> observe -> prior -> KL -> decoder / heads losses -> imagination ->
frozen heads -> lambda-return -> actor / critic / replay-value losses -> backward) is run twice from identical weights,
batch and injected noise: on the pure reference (fp32 cuBLAS, eager) and on the installed build (fp32 paths of the
library).  Every entry of the returned metrics dict and every RSSM parameter gradient must agree:
    metrics: |a - b| <= 2e-3 * |b| + 2e-4     gradients: ||a - b|| <= 5e-3 * ||b|| per tensor (loss-scaled by GradScaler's 65536)
Configs: C1 (proprio observation through the MLP encoder, rep_loss=dreamer) and the r2dreamer loss on the same encoder
(config C2's loss; its CNN encoder is outside the swapped path).  A third test runs the installed agent under
torch.compile(mode="reduce-overhead") as base.yaml:172 does, through the torch.library operators.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _harness():
    from baseline import ref_harness as RH
    if not RH.available():
        pytest.skip("baseline/_ref not present (run python baseline/make_ref.py in the build container)")
    return RH


def _run_cal_grad(agent, data, initial):
    for p in agent.parameters():
        p.grad = None
    (st, dt), mets = agent._cal_grad(data, initial)
    torch.cuda.synchronize()
    mets = {k: float(v) for k, v in mets.items()}
    grads = {n: p.grad.detach().clone() for n, p in agent.rssm.named_parameters() if p.grad is not None}
    return st.detach().clone(), dt.detach().clone(), mets, grads


@pytest.mark.parametrize("rep_loss", ["dreamer", "r2dreamer"])
def test_reference_dreamer_runs_on_the_installed_build(rep_loss):
    RH = _harness()
    import safe_dreamer_b200
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    dev = torch.device("cuda")
    B, T = 4, 16
    agent, cfg = RH.build_dreamer(dev, kind="proprio", rep_loss=rep_loss, compile=False)
    RH.perturb_agent(agent)
    data, initial = RH.make_batch(agent, B, T, dev)
    H = agent.imag_horizon + 1
    S, K, A = agent.rssm._stoch, agent.rssm._discrete, agent.act_dim
    nt = RH.NoiseTape(B, T, B * T, H, S, K, A, False)
    # ---- pure reference
    uq, eq = nt.reference_queues(dev)
    undo = RH.patch_reference_noise(uq, eq)
    try:
        st_r, dt_r, m_r, g_r = _run_cal_grad(agent, data, initial)
    finally:
        undo()
    assert not uq and not eq
    agent.return_ema.ema_vals.zero_()
    # ---- installed build: same agent object, same parameters
    safe_dreamer_b200.install(agent, precision="fp32", imagine_precision="fp32")
    q_live = [nt.u_obs.to(dev), nt.u_prior.to(dev)]
    agent.rssm.noise_source = lambda shape, d: q_live.pop(0).reshape(shape)
    q_img = [nt.u_img.to(dev)]
    agent._frozen_rssm.noise_source = lambda shape, d: q_img.pop(0).reshape(shape)
    agent._frozen_rssm.act_noise_source = lambda shape, d: nt.a_noise.to(d).reshape(shape)
    st_i, dt_i, m_i, g_i = _run_cal_grad(agent, data, initial)
    assert not q_live and not q_img
    # posterior: identical samples, deter to fp32 accuracy
    mism = (st_i.argmax(-1) != st_r.argmax(-1)).float().mean().item()
    print(f"[{rep_loss}] posterior index mismatch rate {mism:.2e}; max |d deter| {float((dt_i - dt_r).abs().max()):.2e}")
    assert mism == 0.0
    assert float((dt_i - dt_r).abs().max()) <= 1e-4
    assert set(m_i) == set(m_r)
    bad = []
    for k in sorted(m_r):
        a, b = m_i[k], m_r[k]
        ok = abs(a - b) <= 2e-3 * abs(b) + 2e-4
        print(f"  {k:28s} reference {b: .6e}  installed {a: .6e}{'' if ok else '   <-- MISMATCH'}")
        if not ok:
            bad.append(k)
    assert not bad, bad
    assert set(g_i) == set(g_r)
    for n in sorted(g_r):
        num = float((g_i[n] - g_r[n]).norm())
        den = float(g_r[n].norm())
        assert num <= 5e-3 * den + 1e-6 * 65536, (n, num, den)


def test_installed_agent_under_torch_compile():
    """configs/base.yaml:172 `compile: True`: dreamer.py:231-233 wraps `_cal_grad` in torch.compile(mode="reduce-overhead").
    The installed RSSM then goes through the torch.library operators (fake implementations for tracing, the C ABI on
    the stream torch captures its CUDA graph on).  Checked: three compiled calls run, every metric is finite, the
    losses agree with an eager call on the same batch within sampling noise, and the RSSM receives gradients."""
    RH = _harness()
    import safe_dreamer_b200
    dev = torch.device("cuda")
    B, T = 4, 16
    agent, cfg = RH.build_dreamer(dev, kind="proprio", rep_loss="dreamer", compile=True)
    RH.perturb_agent(agent)
    data, initial = RH.make_batch(agent, B, T, dev)
    safe_dreamer_b200.install(agent, precision="fp32", imagine_precision="bf16")
    assert agent.rssm.use_custom_ops and agent._frozen_rssm.use_custom_ops
    outs = []
    for it in range(3):
        torch.compiler.cudagraph_mark_step_begin()
        for p in agent.parameters():
            p.grad = None
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            (st, dt), mets = agent._cal_grad(data, initial)
        torch.cuda.synchronize()
        outs.append({k: float(v) for k, v in mets.items()})
        gn = sum(float(p.grad.float().norm()) for p in agent.rssm.parameters() if p.grad is not None)
        assert np.isfinite(gn) and gn > 0
    for m in outs:
        assert all(np.isfinite(v) for v in m.values()), m
    # same batch, fresh noise each call: the KL / reconstruction losses of the three calls agree to sampling noise
    for k in ("loss/dyn", "loss/rep", "loss/rew", "loss/con"):
        vals = [m[k] for m in outs]
        assert max(vals) - min(vals) <= 0.1 * abs(np.mean(vals)) + 0.05, (k, vals)


def test_installed_cnn_encoder_in_the_reference_dreamer():
    """Config C2 (vision, r2dreamer): the reference Dreamer with `install(agent, encoder=True)` -- its ConvEncoder replaced by
    the tcgen05 encoder (same Parameter objects), forward AND backward inside the reference's own `_cal_grad`.
    The encoder computes with bf16 operands, so sampled posterior indices may differ from the fp32 reference at near ties
    and every downstream number moves a little: losses are compared at 10 %, and the encoder's gradients (which now come
    from sd_cnn_backward) against the reference's autograd at 25 % rel. L2 per tensor (tests/test_gpu_k_cnn.py is the tight
    parity gate of the kernels; this test proves the wiring)."""
    RH = _harness()
    import safe_dreamer_b200
    from safe_dreamer_b200.encoder import ConvEncoder
    dev = torch.device("cuda")
    B, T = 4, 16
    agent, cfg = RH.build_dreamer(dev, kind="vision", rep_loss="r2dreamer", compile=False)
    RH.perturb_agent(agent)
    data, initial = RH.make_batch(agent, B, T, dev, kind="vision")
    H = agent.imag_horizon + 1
    S, K, A = agent.rssm._stoch, agent.rssm._discrete, agent.act_dim
    nt = RH.NoiseTape(B, T, B * T, H, S, K, A, False)

    def enc_grads():
        return {n: p.grad.detach().clone() for n, p in agent.encoder.named_parameters() if p.grad is not None}

    uq, eq = nt.reference_queues(dev)
    undo = RH.patch_reference_noise(uq, eq)
    try:
        st_r, dt_r, m_r, _ = _run_cal_grad(agent, data, initial)
        ge_r = enc_grads()
    finally:
        undo()
    agent.return_ema.ema_vals.zero_()
    safe_dreamer_b200.install(agent, precision="fp32", imagine_precision="fp32", encoder=True)
    assert any(isinstance(e, ConvEncoder) for e in agent.encoder.encoders)
    assert any(isinstance(e, ConvEncoder) and e.auto_refresh for e in agent._frozen_encoder.encoders)
    q_live = [nt.u_obs.to(dev), nt.u_prior.to(dev)]
    agent.rssm.noise_source = lambda shape, d: q_live.pop(0).reshape(shape)
    q_img = [nt.u_img.to(dev)]
    agent._frozen_rssm.noise_source = lambda shape, d: q_img.pop(0).reshape(shape)
    agent._frozen_rssm.act_noise_source = lambda shape, d: nt.a_noise.to(d).reshape(shape)
    st_i, dt_i, m_i, _ = _run_cal_grad(agent, data, initial)
    ge_i = enc_grads()
    mism = (st_i.argmax(-1) != st_r.argmax(-1)).float().mean().item()
    print(f"posterior index mismatch rate with the bf16 encoder: {mism:.3e}")
    assert mism <= 0.05
    assert set(m_i) == set(m_r) and all(np.isfinite(v) for v in m_i.values())
    for k in sorted(m_r):
        if k.startswith("loss/") or k.endswith("_loss"):
            print(f"  {k:28s} reference {m_r[k]: .5e}  installed {m_i[k]: .5e}")
            assert abs(m_i[k] - m_r[k]) <= 0.1 * abs(m_r[k]) + 1e-2, k
    assert set(ge_i) == set(ge_r) and len(ge_i) == 12
    for n in sorted(ge_r):
        rel = float((ge_i[n] - ge_r[n]).norm() / ge_r[n].norm().clamp_min(1e-30))
        print(f"  encoder grad {n:22s} rel L2 diff {rel:.3f}")
        assert rel <= 0.25, n


def test_installed_cnn_encoder_under_torch_compile():
    """The vision agent built with config.compile: `install(agent, encoder=True)` routes the encoder through
    torch.ops.safedreamer.cnn_encoder / cnn_encoder_bwd inside the compiled `_cal_grad` (CUDA-graph captured)."""
    RH = _harness()
    import safe_dreamer_b200
    from safe_dreamer_b200.encoder import ConvEncoder
    dev = torch.device("cuda")
    B, T = 4, 16
    agent, cfg = RH.build_dreamer(dev, kind="vision", rep_loss="r2dreamer", compile=True)
    RH.perturb_agent(agent)
    data, initial = RH.make_batch(agent, B, T, dev, kind="vision")
    safe_dreamer_b200.install(agent, precision="fp32", imagine_precision="bf16", encoder=True)
    enc = [e for e in agent.encoder.encoders if isinstance(e, ConvEncoder)][0]
    assert enc.use_custom_ops
    outs = []
    for it in range(3):
        torch.compiler.cudagraph_mark_step_begin()
        for p in agent.parameters():
            p.grad = None
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            (st, dt), mets = agent._cal_grad(data, initial)
        torch.cuda.synchronize()
        outs.append({k: float(v) for k, v in mets.items()})
        gn = sum(float(p.grad.float().norm()) for p in agent.encoder.parameters() if p.grad is not None)
        assert np.isfinite(gn) and gn > 0
        assert all(np.isfinite(v) for v in outs[-1].values())
    for k in ("loss/dyn", "loss/rep", "loss/barlow"):
        assert abs(outs[2][k] - outs[0][k]) <= 0.05 * abs(outs[0][k]) + 1e-3, (k, outs[0][k], outs[2][k])
