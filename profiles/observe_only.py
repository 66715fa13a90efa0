"""Driver: posterior scan only (fp32 path), T steps, direct launches (for SD_TRACE_G / ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from profiles._common import O, cu, make_engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 16
c = O.Cfg(); P = O.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=B, max_steps=max(T, 2), max_tape_rows=B)
embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
for it in range(2):
    st, dt, lg = eng.observe(*args, flags=2)
    torch.cuda.synchronize()
print("ok", float(dt.mean()))
