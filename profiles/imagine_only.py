"""ncu driver: imagination rollout only (tcgen05 path), H steps, direct launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from profiles._common import O, cu, make_engine
H = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 1024
c = O.Cfg(); P = O.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=max(H, 2))
st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
iargs = [cu(x) for x in (st0, dt0, ui, noise)]
for it in range(2):
    feats, acts = eng.imagine(*iargs, H, flags=1)
    torch.cuda.synchronize()
print("ok", float(feats.mean()))
