"""Times RSSM.refresh_weights (sd_set_weights for the RSSM + 5 heads) on the device: python profiles/refresh_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import _lib
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=1024, max_steps=64, max_tape_rows=16)
Pd = {m: {k: cu(v) for k, v in P[m].items()} for m in P}
mods = {"rssm": 0, "actor": 1, "reward": 2, "cont": 3, "value": 4, "slow_value": 5}
def refresh():
    for name, mod in mods.items():
        eng.set_weights(mod, Pd[name])
for _ in range(3): refresh()
torch.cuda.synchronize()
l0 = _lib.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record(); refresh(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"refresh_weights: device {a.elapsed_time(b):.3f} ms, host enqueue {1e3*(t1-t0):.3f} ms, {_lib.launch_count()-l0} kernel launches")
