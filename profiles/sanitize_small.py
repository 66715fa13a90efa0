"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from profiles._common import O, cu, make_engine
from safe_dreamer_b200.dreamer_ops import barlow_loss
from safe_dreamer_b200.distributions import symexp_twohot
from safe_dreamer_b200.networks import ReturnEMA
from safe_dreamer_b200.optim import LaProp

c = O.Cfg(); P = O.init_params(c, seed=0)
B, T, N, H = 5, 3, 130, 2
eng = make_engine(c, P, max_rows=N, max_steps=T, max_tape_rows=B)
embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
args = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
st, dt, lg = eng.observe(*args, flags=0)                       # persistent scan, no tape
st, dt, lg = eng.observe(*args, flags=2)                       # persistent scan + tape
g = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
eng.observe_bwd(B, T, torch.ones_like(st), torch.ones_like(dt), torch.ones_like(lg), True, True, g)
st0, dt0, ui, noise = O.synth_imagine_inputs(c, N, H, seed=3)
feats, acts = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H, flags=1)      # tcgen05 path, ragged last tile
outs = eng.heads_lambda(feats, 1 - 1 / c.horizon, c.lamb, flags=1)
feats32, _ = eng.imagine(cu(st0[:7]), cu(dt0[:7]), cu(ui[:7]), cu(noise[:7]), H, flags=0)   # fp32 path
pst, plog = eng.prior(dt.reshape(B * T, c.D), cu(O.clamp_u(np.random.default_rng(0).random((B * T, c.S, c.K), dtype=np.float32))))
dyn, rep = eng.kl_loss(lg, plog.reshape(B, T, c.S, c.K), 1.0)
eng.kl_loss_bwd(lg, plog.reshape(B, T, c.S, c.K), 1.0, torch.ones_like(dyn), torch.ones_like(rep))
ema = ReturnEMA(device="cuda"); ema(outs[-1])
lgt = torch.randn(33, 255, device="cuda", requires_grad=True)
symexp_twohot(lgt, 255).log_prob(torch.randn(33, 1, device="cuda") * 10).sum().backward()
ps = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.05) for s in [(37, 19), (9000,), (1,)]]
for p_ in ps:
    p_.grad = torch.randn_like(p_) * 1e-3
LaProp(ps, lr=4e-5, eps=1e-20, agc=0.3).step()
x1 = torch.randn(36, 32, device="cuda", requires_grad=True)
barlow_loss(x1, torch.randn(36, 32, device="cuda"), 5e-4).backward()
torch.cuda.synchronize()
print("sanitize_small ok", float(feats.mean()), float(outs[-1].mean()))
