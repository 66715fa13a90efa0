"""Times sd_imagine_fwd (tcgen05 path, CUDA graph) with CUDA events: python profiles/imagine_time.py [N] [H] [iters]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import O, cu, make_engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N, max_steps=H)
st0, dt0, ui, noise = S.synth_imagine_inputs(c, N, H, seed=3)
iargs = [cu(x) for x in (st0, dt0, ui, noise)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    feats, acts = eng.imagine(*iargs, H, flags=5)
torch.cuda.synchronize()
ts = []
for it in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); feats, acts = eng.imagine(*iargs, H, flags=5); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
print(f"imagine N={N} H={H} SD_CHAIN={os.environ.get('SD_CHAIN','0')}: median {ms:.3f} ms  min {ts[0]:.3f}  "
      f"-> {N*H*11674624/ms/1e9:.1f} TFLOP/s; feats mean {float(feats.mean()):.5f}")
