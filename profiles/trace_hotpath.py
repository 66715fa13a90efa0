"""In-stream per-kernel timing with SD_TRACE=1 (CUDA events between launches, direct non-graph run):
  SD_TRACE=1 python profiles/trace_hotpath.py
Prints, for each entry point, the time per kernel type including launch gaps."""
import os
import sys

os.environ["SD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from profiles._common import O, cu, make_engine

B, T, N, H = 16, 64, 1024, 16
c = O.Cfg()
P = O.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=T, max_tape_rows=B)
embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
iargs = [cu(x) for x in (st0, dt0, ui, noise)]
g = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
for it in range(2):
    print(f"==== pass {it}", file=sys.stderr)
    st, dt, lg = eng.observe(*args, flags=2)
    eng.observe_bwd(B, T, torch.ones_like(st), torch.ones_like(dt), torch.ones_like(lg), True, True, g)
    feats, acts = eng.imagine(*iargs, H, flags=1)
    outs = eng.heads_lambda(feats, 1 - 1 / 333, 0.95, flags=1)
    torch.cuda.synchronize()
print("ok")
