"""Times sd_observe_fwd (fp32 path, CUDA graph) and fwd+bwd: python profiles/observe_time.py [B] [T]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=max(B, 16), max_steps=T, max_tape_rows=B)
embed, action, reset, u = S.synth_observe_inputs(c, B, T, seed=2)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
g = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
f = timed(lambda: eng.observe(*args, flags=4))
def fb():
    st, dt, lg = eng.observe(*args, flags=4 | 2)
    eng.observe_bwd(B, T, torch.ones_like(st), torch.ones_like(dt), torch.ones_like(lg), True, True, g, flags=4)
ft = timed(lambda: eng.observe(*args, flags=4 | 2))
fbt = timed(fb)
print(f"observe B={B} T={T} SD_PSCAN={os.environ.get('SD_PSCAN','1')}: fwd {f:.3f} ms ({1e3*f/T:.1f} us/step), fwd+tape {ft:.3f} ms, fwd+bwd {fbt:.3f} ms")
