"""Fused AGC + LaProp (utils/optim/agc.py, utils/optim/laprop.py): numpy oracle vs three steps of the real reference
(golden), and the CUDA path behind sd_agc_laprop_step (LaProp / clip_grad_agc_ mirrors) vs both."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(48, 200), (256,), (8, 12, 8), (12, 64), (1,), (255, 16), (3, 5, 7)]
LR, B1, B2, EPS, CLIP, PMIN = 4e-5, 0.9, 0.999, 1e-20, 0.3, 1e-3


def _inputs(step):
    # restated from tests/golden/make_golden.py:optim_inputs
    rng = np.random.Generator(np.random.Philox(31337 + (step if step >= 0 else 1000)))
    out = []
    for i, shp in enumerate(SHAPES):
        x = rng.standard_normal(shp, dtype=np.float32)
        if step < 0:
            x = x * np.float32(1e-5 if i == 6 else 0.05)
        else:
            x = x * np.float32([1e-3, 5.0, 1e-2, 1e-4, 0.3, 2e-2, 1e-3][i])
        out.append(x.astype(np.float32))
    return out


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "optim.npz"))


def _close(a, b, what):
    # fp32 tolerance: 2e-5 relative, plus 2e-6 of the tensor's scale for elements that cancel towards zero
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-6 * float(np.abs(b).max()) + 1e-30, err_msg=what)


def test_oracle_optimiser_matches_reference(golden):
    ps = _inputs(-1)
    ms = [np.zeros_like(p) for p in ps]
    vs = [np.zeros_like(p) for p in ps]
    sts = [dict(step=0, lr1=0.0, lr2=0.0) for _ in ps]
    clipped = 0
    for step in range(3):
        gs = _inputs(step)
        for i in range(len(ps)):
            sc = O.agc_scale(ps[i], gs[i], CLIP, PMIN)
            clipped += int(sc < 1.0)
            g = (gs[i] * sc).astype(np.float32)
            _close(g, golden[f"s{step}_g{i}"], f"agc step {step} tensor {i}")
            ps[i], ms[i], vs[i] = O.laprop_step(ps[i], g, ms[i], vs[i], sts[i], LR, B1, B2, EPS)
            _close(ps[i], golden[f"s{step}_p{i}"], f"p step {step} tensor {i}")
            _close(ms[i], golden[f"s{step}_m{i}"], f"m step {step} tensor {i}")
            _close(vs[i], golden[f"s{step}_v{i}"], f"v step {step} tensor {i}")
    assert 0 < clipped < 3 * len(ps)      # the fixture exercises both sides of the clip


@pytest.mark.gpu
def test_cuda_fused_agc_laprop(golden):
    import torch
    from safe_dreamer_b200.optim import LaProp, clip_grad_agc_
    params = [torch.nn.Parameter(torch.from_numpy(x.copy()).cuda()) for x in _inputs(-1)]
    opt = LaProp(params, lr=LR, betas=(B1, B2), eps=EPS, agc=CLIP, pmin=PMIN)        # AGC folded into the step
    params2 = [torch.nn.Parameter(torch.from_numpy(x.copy()).cuda()) for x in _inputs(-1)]
    opt2 = LaProp(params2, lr=LR, betas=(B1, B2), eps=EPS)                           # reference call pattern: agc, then step
    for step in range(3):
        for p_, q_, g_ in zip(params, params2, _inputs(step)):
            p_.grad = torch.from_numpy(g_.copy()).cuda()
            q_.grad = torch.from_numpy(g_.copy()).cuda()
        clip_grad_agc_(params2, CLIP, PMIN, foreach=True)
        opt.step()
        opt2.step()
        for i, (p_, q_) in enumerate(zip(params, params2)):
            _close(q_.grad.cpu().numpy(), golden[f"s{step}_g{i}"], f"agc step {step} tensor {i}")
            _close(p_.grad.cpu().numpy(), golden[f"s{step}_g{i}"], f"fused agc step {step} tensor {i}")
            for o_, x_ in ((opt, p_), (opt2, q_)):
                _close(x_.detach().cpu().numpy(), golden[f"s{step}_p{i}"], f"p step {step} tensor {i}")
                _close(o_.state[x_]["exp_avg"].cpu().numpy(), golden[f"s{step}_m{i}"], f"m step {step} tensor {i}")
                _close(o_.state[x_]["exp_avg_sq"].cpu().numpy(), golden[f"s{step}_v{i}"], f"v step {step} tensor {i}")
            assert torch.equal(p_.detach(), q_.detach())      # folding AGC into the step changes nothing, bit for bit
    assert set(opt.state[params[0]].keys()) == {"step", "exp_avg", "exp_avg_lr_1", "exp_avg_lr_2", "exp_avg_sq"}
    # GradScaler semantics: a non-finite gradient raises found_inf and leaves parameters and moments untouched
    before = [p_.detach().clone() for p_ in params]
    for p_, g_ in zip(params, _inputs(0)):
        p_.grad = torch.from_numpy(g_.copy()).cuda()
    params[2].grad[0, 0, 0] = float("inf")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    opt.step(inv_scale=1.0 / 65536.0, found_inf=flag)
    assert int(flag.item()) == 1
    for p_, b_ in zip(params, before):
        assert torch.equal(p_.detach(), b_)


@pytest.mark.gpu
def test_cuda_fused_step_unscales_before_the_clip(golden):
    """GradScaler folded into the step (dreamer.py:422,432-433): gradients arrive multiplied by the loss scale; they must be
    unscaled BEFORE the AGC clip.  A power-of-two scale is exact, so the goldens of the unscaled run apply unchanged."""
    import torch
    from safe_dreamer_b200.optim import LaProp
    params = [torch.nn.Parameter(torch.from_numpy(x.copy()).cuda()) for x in _inputs(-1)]
    opt = LaProp(params, lr=LR, betas=(B1, B2), eps=EPS, agc=CLIP, pmin=PMIN)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(3):
        for p_, g_ in zip(params, _inputs(step)):
            p_.grad = torch.from_numpy(g_.copy()).cuda() * 65536.0
        opt.step(inv_scale=1.0 / 65536.0, found_inf=flag, sync=True)
        assert int(flag.item()) == 0
        for i, p_ in enumerate(params):
            _close(p_.grad.cpu().numpy(), golden[f"s{step}_g{i}"], f"unscaled + clipped grad, step {step} tensor {i}")
            _close(p_.detach().cpu().numpy(), golden[f"s{step}_p{i}"], f"p step {step} tensor {i}")
            _close(opt.state[p_]["exp_avg"].cpu().numpy(), golden[f"s{step}_m{i}"], f"m step {step} tensor {i}")
            _close(opt.state[p_]["exp_avg_sq"].cpu().numpy(), golden[f"s{step}_v{i}"], f"v step {step} tensor {i}")


@pytest.mark.gpu
def test_cuda_overflow_skips_every_bucket_and_rolls_back():
    """Two parameter groups with different learning rates are two launches: an overflow in the SECOND one must leave the
    first untouched as well, and the host-side step counters must not advance (GradScaler.step skips optimizer.step())."""
    import torch
    from safe_dreamer_b200.optim import LaProp
    a = [torch.nn.Parameter(torch.from_numpy(x.copy()).cuda()) for x in _inputs(-1)[:3]]
    b = [torch.nn.Parameter(torch.from_numpy(x.copy()).cuda()) for x in _inputs(-1)[3:]]
    opt = LaProp([{"params": a, "lr": LR}, {"params": b, "lr": 3 * LR}], betas=(B1, B2), eps=EPS, agc=CLIP, pmin=PMIN)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for p_, g_ in zip(a + b, _inputs(0)):
        p_.grad = torch.from_numpy(g_.copy()).cuda()
    opt.step(found_inf=flag, sync=True)                   # a clean step first (different bias corrections per group from now on)
    assert int(flag.item()) == 0 and opt.state[a[0]]["step"] == 1
    before = [p_.detach().clone() for p_ in a + b]
    moments = [opt.state[p_]["exp_avg"].clone() for p_ in a + b]
    for p_, g_ in zip(a + b, _inputs(1)):
        p_.grad = torch.from_numpy(g_.copy()).cuda()
    b[-1].grad.view(-1)[0] = float("nan")
    opt.step(found_inf=flag, sync=True)
    assert int(flag.item()) == 1
    for p_, b_, m_ in zip(a + b, before, moments):
        assert torch.equal(p_.detach(), b_) and torch.equal(opt.state[p_]["exp_avg"], m_)
        assert opt.state[p_]["step"] == 1                 # rolled back
