"""Time the grad-enabled imagination (attack shape: frozen weights, dgrad-only) fwd+bwd at N rows, H=16."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200.engine import Engine
H = 16
c = O.Cfg(); P = O.init_params(c, seed=0)
for N in (64, 1024):
    eng = Engine.from_cfg(c, N, H, N, P)
    eng.static_outputs = True
    st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    cu = lambda x: torch.from_numpy(x).cuda()
    args = [cu(x) for x in (st0, dt0, ui, noise)]
    df = torch.randn(N, H, c.F, device="cuda") * 0.01
    da = torch.randn(N, H, c.A, device="cuda") * 0.01
    for bf in (0, 1):
        def step():
            eng.imagine(*args, H, flags=2 | 4 | bf)
            eng.imagine_bwd(N, H, df, da, flags=4 | bf)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            step()
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"N={N} H={H} fwd+bwd ({"bf16 tcgen05" if bf else "fp32"}): {ms:.2f} ms -> {N * H / ms * 1e3:.0f} imagined steps/s fwd+bwd")
    del eng
