"""How much of the posterior backward is the weight-gradient tail?  Times sd_observe_bwd (B=16, T=64, CUDA graph) with and
without weight gradients.  python profiles/bwd_tail_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from safe_dreamer_b200 import synth as O
from safe_dreamer_b200.engine import Engine

c = O.Cfg()
P = O.init_params(c, seed=0)
B, T = 16, 64
eng = Engine.from_cfg(c, 1024, 64, 16, P)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
s0 = torch.zeros(B, c.S, c.K, device="cuda"); d0 = torch.zeros(B, c.D, device="cuda")
ins = [cu(embed), cu(action), s0, d0, cu(reset), cu(u)]
g_dt = torch.randn(B, T, c.D, device="cuda") * 0.1
g_lg = torch.randn(B, T, c.S, c.K, device="cuda") * 0.1
names = eng.weight_names(0)
wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in names}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


eng.observe(*ins, flags=4 | 2)
t_full = timed(lambda: eng.observe_bwd(B, T, None, g_dt, g_lg, True, True, wg, flags=4))
t_scan = timed(lambda: eng.observe_bwd(B, T, None, g_dt, g_lg, True, True, None, flags=4))
print(f"observe_bwd with weight gradients {t_full:.3f} ms, scan + dgrad only {t_scan:.3f} ms -> weight-gradient tail {t_full - t_scan:.3f} ms "
      f"(SD_WGRAD_TC={os.environ.get('SD_WGRAD_TC', '1')})")
