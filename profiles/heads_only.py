"""ncu driver: heads + lambda-return on 16 384 imagined rows (tcgen05 path), direct launches.  python profiles/heads_only.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
N, H = 1024, 16
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=H)
st0, dt0, ui, noise = S.synth_imagine_inputs(c, N, H, seed=3)
feats, acts = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H, flags=1)
for it in range(2):
    outs = eng.heads_lambda(feats, 1 - 1 / c.horizon, c.lamb, flags=1)
    torch.cuda.synchronize()
print("ok", float(outs[-1].mean()))
