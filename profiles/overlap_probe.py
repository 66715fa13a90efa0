"""Where does the two-stream schedule lose time?  python profiles/overlap_probe.py
Stream A: posterior fwd (tape) + reverse-time backward (fp32, R = 16 rows).  Stream B (starts after the posterior fwd):
imagination N = 1024, H = 16 + heads.  Prints each stream alone and, overlapped, when each stream finishes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from safe_dreamer_b200 import synth as S
from profiles._common import cu, make_engine
B, T, H = 16, 64, 16
N = B * T
c = S.Cfg(); P = S.init_params(c, seed=0)
eng = make_engine(c, P, max_rows=N * H, max_steps=max(T, H), max_tape_rows=B)
embed, action, reset, u = S.synth_observe_inputs(c, B, T, seed=2)
oargs = [cu(x) for x in (embed, action, np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32), reset, u)]
_, _, ui, noise = S.synth_imagine_inputs(c, N, H, seed=3)
ui, noise = cu(ui), cu(noise)
g = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
disc = 1.0 - 1.0 / c.horizon if hasattr(c, "horizon") else 0.997
HI = 0
BG = int(os.environ.get("PROBE_BG", "1")) * 16   # SD_FLAG_BACKGROUND on the side stream's calls
LW = int(os.environ.get("PROBE_LW", "1")) * 64   # SD_FLAG_LAYERWISE: the launch-sequence rollout (bench.py's default schedule)
side = torch.cuda.Stream()
st, dt, lg = eng.observe(*oargs, flags=4 | 2)
gs, gd, gl = torch.ones_like(st), torch.ones_like(dt), torch.ones_like(lg)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def A_fwd():
    return eng.observe(*oargs, flags=4 | 2 | HI)
def A_bwd():
    eng.observe_bwd(B, T, gs, gd, gl, True, True, g, flags=4 | HI)
def B_all(st, dt):
    f, a = eng.imagine(st.reshape(N, c.S, c.K), dt.reshape(N, c.D), ui, noise, H, flags=1 | 4 | BG | LW)
    eng.heads_lambda(f, disc, 0.95, flags=1 | 4 | BG)

src_buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
dst_buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
small = torch.empty(8 << 20, dtype=torch.float32, device="cuda")
def B_memcpy():      # copy-engine traffic only: 24 x 64 MB device-to-device copies, no SM work
    for _ in range(24):
        dst_buf.copy_(src_buf, non_blocking=True)
def B_l2():          # L2-resident traffic from SM kernels: an 32 MB buffer scaled in place 40 times
    for _ in range(40):
        small.mul_(1.0001)
def B_imag(st, dt):
    eng.imagine(st.reshape(N, c.S, c.K), dt.reshape(N, c.D), ui, noise, H, flags=1 | 4 | BG | LW)

def run(mode):
    main = torch.cuda.current_stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    flush.zero_()
    ev[0].record(main)
    st, dt, lg = A_fwd()
    ev[1].record(main)
    if mode == "A":
        A_bwd(); ev[2].record(main); ev[3].record(main)
    elif mode == "B":
        B_all(st, dt); ev[2].record(main); ev[3].record(main)
    elif mode == "seq":
        A_bwd(); ev[2].record(main); B_all(st, dt); ev[3].record(main)
    else:
        side.wait_event(ev[1])
        with torch.cuda.stream(side):
            if mode == "overlap": B_all(st, dt)
            elif mode == "ov_memcpy": B_memcpy()
            elif mode == "ov_l2": B_l2()
            elif mode == "ov_imag": B_imag(st, dt)
            ev[3].record(side)
        A_bwd(); ev[2].record(main)
        main.wait_event(ev[3])
    ev[4].record(main)
    torch.cuda.synchronize()
    return [ev[0].elapsed_time(e) for e in ev[1:]]

for mode in ("A", "B", "seq", "overlap", "ov_imag", "ov_memcpy", "ov_l2"):
    for _ in range(3): run(mode)
    rs = np.median(np.array([run(mode) for _ in range(15)]), axis=0)
    print(f"{mode:8s} fwd done {rs[0]:.3f}  main-stream work done {rs[1]:.3f}  side/B done {rs[2]:.3f}  all {rs[3]:.3f} ms")
