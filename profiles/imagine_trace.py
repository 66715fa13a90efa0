"""SD_TRACE=1 python profiles/imagine_trace.py : per-kernel in-stream times of the imagination scan (direct launches)."""
import os, sys
os.environ["SD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
N, H = 1024, 16
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=H)
st0, dt0, ui, noise = S.synth_imagine_inputs(c, N, H, seed=3)
iargs = [cu(x) for x in (st0, dt0, ui, noise)]
for it in range(2):
    print(f"==== pass {it}", file=sys.stderr)
    feats, acts = eng.imagine(*iargs, H, flags=1)
    torch.cuda.synchronize()
