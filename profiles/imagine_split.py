"""Experiment: imagination of N rows as S independent row groups on S streams (one engine handle each).
python profiles/imagine_split.py [N] [H] [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S_
from profiles._common import O, cu, make_engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = S_.Cfg(); P = S_.init_params(c, seed=0)
n = N // S
engs = [make_engine(c, P, max_rows=n, max_steps=H) for _ in range(S)]
st0, dt0, ui, noise = S_.synth_imagine_inputs(c, N, H, seed=3)
args = [[cu(x[i * n:(i + 1) * n]) for x in (st0, dt0, ui, noise)] for i in range(S)]
streams = [torch.cuda.Stream() for _ in range(S)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    outs = []
    for i in range(S):
        streams[i].wait_event(ev)
        with torch.cuda.stream(streams[i]):
            outs.append(engs[i].imagine(*args[i], H, flags=5))
        e2 = torch.cuda.Event(); e2.record(streams[i]); cur.wait_event(e2)
    return outs
for it in range(3): run()
torch.cuda.synchronize()
ts = []
for it in range(20):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); outs = run(); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); ms = ts[len(ts) // 2]
print(f"imagine N={N} H={H} groups={S} SD_CHAIN={os.environ.get('SD_CHAIN','0')}: median {ms:.3f} ms -> {N*H*11674624/ms/1e9:.1f} TFLOP/s")
