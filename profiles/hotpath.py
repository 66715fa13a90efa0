"""Small driver for ncu: one direct (non-graph) pass of the hot path at the bench shapes with a
shortened scan (T_OBS posterior steps, H_IMAG imagination steps) so the launch list stays short.
  python profiles/hotpath.py [T_OBS] [H_IMAG] [bwd]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from profiles._common import O, cu, make_engine

T_OBS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H_IMAG = int(sys.argv[2]) if len(sys.argv) > 2 else 3
BWD = len(sys.argv) > 3 and sys.argv[3] == "bwd"
B, N = 16, 1024
c = O.Cfg()
P = O.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=max(T_OBS, H_IMAG, 2), max_tape_rows=B if BWD else 0)
embed, action, reset, u = O.synth_observe_inputs(c, B, T_OBS, seed=2)
st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H_IMAG, seed=3)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
iargs = [cu(x) for x in (st0, dt0, ui, noise)]
for it in range(2):
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.nvtx.range_push("hotpath")
    st, dt, lg = eng.observe(*args, flags=2 if BWD else 0)
    if BWD:
        g = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
        eng.observe_bwd(B, T_OBS, torch.ones_like(st), torch.ones_like(dt), torch.ones_like(lg), True, True, g)
    feats, acts = eng.imagine(*iargs, H_IMAG, flags=1)
    outs = eng.heads_lambda(feats, 1 - 1 / 333, 0.95, flags=1)
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.nvtx.range_pop()
print("ok", float(outs[-1].mean()))
