"""Times one optimiser step over all RSSM + head parameters (11.8 M elements): the fused AGC + LaProp path
(sd_agc_laprop_step, 3 launches) against a torch restatement of the reference's call pattern (foreach AGC, then a Python
loop of per-tensor ops as in utils/optim/laprop.py:46-118).  python profiles/optim_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S
from safe_dreamer_b200.optim import LaProp

c = S.Cfg()
shapes = [shp for mod in S.all_param_shapes(c).values() for shp in mod.values()]
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda: [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g) * 0.05) for s in shapes]
grads = [torch.randn(s, device="cuda", generator=g) * 1e-3 for s in shapes]
LR, B1, B2, EPS, CLIP, PMIN = 4e-5, 0.9, 0.999, 1e-20, 0.3, 1e-3


def reference_style(params, state):
    ps = [p for p in params]
    gs = [p.grad for p in params]
    pnorm = torch._foreach_norm(ps, 2); gnorm = torch._foreach_norm(gs, 2)
    upper = torch._foreach_mul(torch._foreach_maximum(pnorm, PMIN), CLIP)
    scale = torch._foreach_reciprocal(torch._foreach_maximum(torch._foreach_div(gnorm, upper), 1.0))
    torch._foreach_mul_(gs, scale)
    for p in params:
        st = state.setdefault(p, dict(lr1=0.0, lr2=0.0, m=torch.zeros_like(p.data), v=torch.zeros_like(p.data)))
        grad = p.grad
        st["v"].mul_(B2).addcmul_(grad, grad, value=1 - B2)
        st["lr1"] = st["lr1"] * B1 + (1 - B1) * LR
        st["lr2"] = st["lr2"] * B2 + (1 - B2)
        step_size = 1 / (st["lr1"] / LR)
        denom = st["v"].div(st["lr2"]).sqrt_().add_(EPS)
        st["m"].mul_(B1).add_(grad / denom, alpha=(1 - B1) * LR)
        p.data.add_(st["m"], alpha=-step_size)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, 1e3 * (time.perf_counter() - t0) / iters


pa, pb = mk(), mk()
for p, q, gr in zip(pa, pb, grads):
    p.grad = gr.clone(); q.grad = gr.clone()
opt = LaProp(pa, lr=LR, betas=(B1, B2), eps=EPS, agc=CLIP, pmin=PMIN)
state = {}
with torch.no_grad():
    dev_f, wall_f = timed(lambda: opt.step())
    dev_r, wall_r = timed(lambda: reference_style(pb, state))
n = sum(p.numel() for p in pa)
print(f"optimiser step over {len(pa)} tensors / {n/1e6:.2f} M params: fused {dev_f:.3f} ms device ({wall_f:.3f} ms wall), "
      f"reference-style torch loop {dev_r:.3f} ms device ({wall_r:.3f} ms wall)")
