"""Heads + lambda-return on the 16 384 imagined rows of the headline config (bf16 tcgen05 path, CUDA graph, L2 flushed): time per
call and the outputs' checksum (SD_HEADS_CHAIN=0/1 selects one launch per layer / the row-tile resident trunk chain).
python profiles/heads_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import cu, make_engine
N, H = 1024, 16
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=H)
st0, dt0, ui, noise = S.synth_imagine_inputs(c, N, H, seed=3)
feats, acts = eng.imagine(cu(st0), cu(dt0), cu(ui), cu(noise), H, flags=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
disc = 1 - 1 / c.horizon
for _ in range(3):
    outs = eng.heads_lambda(feats, disc, c.lamb, flags=1 | 4)
tot = 0.0
for _ in range(20):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); outs = eng.heads_lambda(feats, disc, c.lamb, flags=1 | 4); b.record(); b.synchronize()
    tot += a.elapsed_time(b)
print(f"heads_lambda N={N} H={H} SD_HEADS_CHAIN={os.environ.get('SD_HEADS_CHAIN', '1')}: {tot / 20:.4f} ms; "
      f"mean ret {float(outs[-1].double().mean()):.6f} rew {float(outs[0].double().mean()):.6f} val {float(outs[2].double().mean()):.6f} "
      f"cont {float(outs[1].double().mean()):.6f}")
