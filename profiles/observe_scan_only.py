"""ncu driver: posterior scan forward only (fp32, persistent kernel), direct launches.  python profiles/observe_scan_only.py [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = 16
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=16, max_steps=T, max_tape_rows=0)
embed, action, reset, u = S.synth_observe_inputs(c, B, T, seed=2)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
for it in range(2):
    st, dt, lg = eng.observe(*args, flags=0)
    torch.cuda.synchronize()
print("ok", float(dt.mean()))
