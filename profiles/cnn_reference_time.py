"""Times the UNMODIFIED reference ConvEncoder (baseline/_ref, networks.py:192-234) on this GPU: forward and
forward + backward on 1024 frames of 64x64x3 (B=16, T=64), eager fp32 (TF32), eager fp16 autocast (dreamer.py:420), and
torch.compile(mode="reduce-overhead") fp16 (configs/base.yaml:172), next to this library's encoder.  Prints one JSON line.
python profiles/cnn_reference_time.py [--no-compile]"""
import json
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from baseline import ref_harness as RH

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.set_float32_matmul_precision("high")
mods = RH.import_reference()
cfg = NS(act="SiLU", norm=True, kernel_size=5, minres=4, depth=16, mults=[2, 3, 4, 4])
dev = torch.device("cuda")
torch.manual_seed(0)
enc = mods.networks.ConvEncoder(cfg, (64, 64, 3)).to(dev)
obs = torch.rand(16, 64, 64, 64, 3, device=dev)
g = torch.randn(16, 64, 1024, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def fwd(m, amp):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        return m(obs)


def fwd_bwd(m, amp):
    for p in m.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        e = m(obs)
    e.backward(g.to(e.dtype))


out = {}
out["ref_eager_fp32"] = {"fwd": timed(lambda: fwd(enc, False)), "fwd_bwd": timed(lambda: fwd_bwd(enc, False))}
out["ref_eager_fp16"] = {"fwd": timed(lambda: fwd(enc, True)), "fwd_bwd": timed(lambda: fwd_bwd(enc, True))}
if "--no-compile" not in sys.argv:
    try:
        cenc = torch.compile(enc, mode="reduce-overhead")
        out["ref_compiled_fp16"] = {"fwd": timed(lambda: fwd(cenc, True), warm=5), "fwd_bwd": timed(lambda: fwd_bwd(cenc, True), warm=5)}
    except Exception as e:  # noqa: BLE001
        out["ref_compiled_fp16"] = {"error": repr(e)[:200]}
try:
    from safe_dreamer_b200.encoder import ConvEncoder
    mine = ConvEncoder(cfg, (64, 64, 3)).to(dev)
    mine.load_state_dict(enc.state_dict())
    out["ours_bf16"] = {"fwd": timed(lambda: fwd(mine, False))}
    try:
        out["ours_bf16"]["fwd_bwd"] = timed(lambda: fwd_bwd(mine, False))
    except NotImplementedError:
        pass
    with torch.no_grad():
        d = (mine(obs) - enc(obs)).abs()
    out["ours_bf16"]["max_abs_diff_vs_ref_fp32"] = float(d.max())
    out["ours_bf16"]["mean_abs_diff_vs_ref_fp32"] = float(d.mean())
except Exception as e:  # noqa: BLE001
    out["ours_bf16"] = {"error": repr(e)[:300]}
print(json.dumps({"cnn_encoder_ms": out, "frames": 1024, "frame": [64, 64, 3], "gflop_fwd": 154.3}))
