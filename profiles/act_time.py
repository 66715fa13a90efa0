"""Latency of one policy-inference step after the encoder (dreamer_ops.act == dreamer.py:345-357) at B = env_num."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from types import SimpleNamespace as NS
import numpy as np
import torch
from safe_dreamer_b200 import dreamer_ops, synth as S
from safe_dreamer_b200.networks import MLPHead
from safe_dreamer_b200.rssm import RSSM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c = S.Cfg(); P = S.init_params(c, seed=0)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned", device="cuda",
         obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
rssm = RSSM(cfg, c.E, c.A).cuda()
rssm.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
actor = MLPHead("actor", c.actor_layers, c.units, c.F, 2 * c.A).cuda()
actor.load_state_dict({k: cu(v) for k, v in P["actor"].items()})
dreamer_ops.attach_heads(rssm, actor=actor)
rssm.use_graph, rssm.auto_refresh, rssm.static_outputs = True, False, True
state = (torch.zeros(B, c.S, c.K, device="cuda"), torch.zeros(B, c.D, device="cuda"), torch.zeros(B, c.A, device="cuda"))
emb = torch.randn(B, c.E, device="cuda"); first = torch.zeros(B, dtype=torch.bool, device="cuda")
for _ in range(5):
    action, state = dreamer_ops.act(rssm, emb, state, first)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(100):
    action, state = dreamer_ops.act(rssm, emb, state, first)
b.record(); torch.cuda.synchronize()
print(f"act step B={B}: {a.elapsed_time(b) * 10:.1f} us device, {1e4 * (time.perf_counter() - t0):.1f} us wall per step")
