"""Helper of tests/test_gpu_l_wgrad_tc.py: one taped posterior scan + backward at B=16, T=64 (1024 taped rows) and one batched
prior + backward, weight gradients written to an .npz.  SD_WGRAD_TC in the environment selects the weight-gradient kernel."""
import sys

import numpy as np
import torch

from oracle import rssm_oracle as O
from tests.helpers import cu, golden_initial, make_engine

out = sys.argv[1]
c = O.Cfg()
P = O.init_params(c, seed=0)
B, T = 16, 64
eng = make_engine(c, P, max_rows=1024, max_steps=64, max_tape_rows=1024)
embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=51)
s0, d0 = golden_initial(c, B)
g = np.random.Generator(np.random.Philox(53))
c_dt = g.standard_normal((B, T, c.D), dtype=np.float32) * np.float32(0.1)
c_lg = g.standard_normal((B, T, c.S, c.K), dtype=np.float32) * np.float32(0.1)
st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=2)
names = eng.weight_names(0)
wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in names}
eng.observe_bwd(B, T, None, cu(c_dt), cu(c_lg), True, True, wg)
up = O.clamp_u(np.random.Generator(np.random.Philox(55)).random((B * T, c.S, c.K), dtype=np.float32))
pst, plg = eng.prior(dt.reshape(B * T, c.D), cu(up), flags=2)
d_pl = g.standard_normal((B * T, c.S, c.K), dtype=np.float32) * np.float32(0.1)
eng.prior_bwd(B * T, None, cu(d_pl), True, wg)
torch.cuda.synchronize()
np.savez(out, **{n: v.cpu().numpy() for n, v in wg.items()})
