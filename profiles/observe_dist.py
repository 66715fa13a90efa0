"""Distribution of sd_observe_fwd times (fp32 persistent scan, CUDA graph, L2 flushed): python profiles/observe_dist.py [iters]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
B, T = 16, 64
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=16, max_steps=T, max_tape_rows=B)
embed, action, reset, u = S.synth_observe_inputs(c, B, T, seed=2)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for fl, name in ((4, "fwd"), (4 | 2, "fwd+tape")):
    for _ in range(3): eng.observe(*args, flags=fl)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.observe(*args, flags=fl); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts = np.array(ts); s = np.sort(ts)
    print(f"{name}: min {s[0]:.3f} p25 {s[len(s)//4]:.3f} median {s[len(s)//2]:.3f} mean {ts.mean():.3f} p90 {s[int(len(s)*0.9)]:.3f} max {s[-1]:.3f} ms; first 8: {np.round(ts[:8],3).tolist()}")
