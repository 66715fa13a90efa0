"""Debug / timing driver of the persistent imagination kernel: compares it with the layer-by-layer bf16 path of the same
library (product path only) and prints per-quantity differences and timings.  python profiles/pimg_debug.py [N] [H]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from safe_dreamer_b200 import synth as O
from safe_dreamer_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = O.Cfg()
P = O.init_params(c, seed=0)
eng = Engine.from_cfg(c, max(N, 128), max(H, 2), 0, P)
st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=31)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
ins = [cu(x) for x in (st0, dt0, u, noise)]
fl, al = [x.clone() for x in eng.imagine(*ins, H, flags=1 | 64)]
torch.cuda.synchronize()
t0 = time.time()
fp, ap = [x.clone() for x in eng.imagine(*ins, H, flags=1 | 32)]
torch.cuda.synchronize()
print(f"persistent call returned after {time.time() - t0:.3f} s")
SK = c.SK
idx = lambda f: f[..., :SK].reshape(*f.shape[:-1], c.S, c.K).argmax(-1)
for t in range(H):
    da = (ap[:, t] - al[:, t]).abs().max().item()
    dd = (fp[:, t, SK:] - fl[:, t, SK:]).abs().max().item()
    mis = (idx(fp[:, t]) != idx(fl[:, t])).float().mean().item()
    onehot = bool((fp[:, t, :SK].reshape(N, c.S, c.K).sum(-1) == 1).all())
    print(f"step {t}: |dact|={da:.4f} |ddeter|={dd:.4f} idx mismatch={mis:.4f} onehot={onehot} finite={bool(torch.isfinite(fp[:, t]).all())}")
if len(sys.argv) > 3:
    def timed(flags, iters=20):
        for _ in range(3):
            eng.imagine(*ins, H, flags=flags)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(iters):
            eng.imagine(*ins, H, flags=flags)
        b.record(); b.synchronize()
        return a.elapsed_time(b) / iters
    print(f"N={N} H={H}: persistent {timed(1 | 4 | 32):.3f} ms   layer-by-layer {timed(1 | 4 | 64):.3f} ms")
