"""Replay latent write-back at BASELINE sizes (B*T = 1024 rows of 32x16 + 2048): sd_latent_writeback / sd_latent_gather
vs the reference's formulation (two index assignments into one-hot fp32 storage) in torch on the same GPU.
python profiles/latent_store_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200.replay import LatentStore
B, T, S, K, D, n_len, n_env = 16, 64, 32, 16, 2048, 4096, 16
ls = LatentStore(n_len, n_env, S, K, D)
g = torch.Generator().manual_seed(1)
t0 = torch.randint(0, n_len - T, (B,), generator=g)
time = (t0[:, None] + torch.arange(T)[None]).cuda(); env = torch.arange(B)[:, None].expand(B, T).contiguous().cuda()
stoch = torch.nn.functional.one_hot(torch.randint(0, K, (B, T, S), generator=g), K).float().cuda()
deter = torch.randn(B, T, D, generator=g).cuda()
ref_st = torch.zeros(n_len, n_env, S, K, device="cuda"); ref_dt = torch.zeros(n_len, n_env, D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
def ref_update():
    ref_st[time.reshape(-1), env.reshape(-1)] = stoch.reshape(-1, S, K)
    ref_dt[time.reshape(-1), env.reshape(-1)] = deter.reshape(-1, D)
def ref_initial():
    return ref_st[time[:, 0], env[:, 0]], ref_dt[time[:, 0], env[:, 0]]
w = timed(lambda: ls.update([env, time], stoch, deter, validate=False))
r = timed(lambda: ls.initial([env[:, :1], time[:, :1]], validate=False))
wr = timed(ref_update); rr = timed(ref_initial)
bytes_w = B * T * (S * K * 4 + D * 4 + 16 + S + D * 4)
print(f"write-back 1024 rows: sd_latent_writeback {w:.1f} us ({bytes_w / w / 1e3:.0f} GB/s algorithmic) vs torch index_put x2 {wr:.1f} us")
print(f"initial gather 16 rows: sd_latent_gather {r:.1f} us vs torch indexing x2 {rr:.1f} us")
print(f"storage per slot: {S + D * 4} B (uint8 classes + fp32 deter) vs reference {S * K * 4 + D * 4} B")
