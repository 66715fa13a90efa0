"""Drop-in mirrors of utils/optim/laprop.py (LaProp) and utils/optim/agc.py (clip_grad_agc_): same constructor /
call signatures and state_dict layout (`step`, `exp_avg`, `exp_avg_lr_1`, `exp_avg_lr_2`, `exp_avg_sq` per parameter),
but the whole update of every tensor is three CUDA launches (sd_agc_laprop_step) instead of a Python loop of ~8 kernels
per tensor.  `LaProp(..., agc=clip, pmin=pmin)` additionally folds the adaptive gradient clipping that dreamer.py:432 runs
right before the step into the same launches."""
import ctypes as C

import torch
from torch.optim import Optimizer

from . import _lib


class _Fused:
    """Device scratch + launch helper shared by LaProp and clip_grad_agc_."""

    def __init__(self):
        self.table = None
        self.scratch = None
        self.key = None

    def run(self, entries, mode, clip, pmin, inv_scale, beta1, beta2, lr_term, step_size, bc2, eps, wd, found_inf):
        lib = _lib.load()
        n = len(entries)
        dev = entries[0][0].device
        # the same tensors every step: the ctypes table and the device scratch are rebuilt only when a pointer changes
        key = (dev,) + tuple(x.data_ptr() if x is not None else 0 for e in entries for x in e)
        if key != self.key:
            arr = (_lib.sd_opt_tensor * n)()
            for i, (p, g, m, v) in enumerate(entries):
                arr[i].param, arr[i].grad = p.data_ptr(), g.data_ptr()
                arr[i].exp_avg = m.data_ptr() if m is not None else None
                arr[i].exp_avg_sq = v.data_ptr() if v is not None else None
                arr[i].numel = p.numel()
            self.arr = arr
            self.table = torch.empty(int(lib.sd_opt_table_bytes(n)), dtype=torch.uint8, device=dev)
            self.scratch = torch.empty(int(lib.sd_opt_scratch_bytes(arr, n)), dtype=torch.uint8, device=dev)
            self.key = key
        arr = self.arr
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.sd_agc_laprop_step(arr, n, mode, float(clip), float(pmin), float(inv_scale), float(beta1), float(beta2),
                                          float(1 - beta2), float(lr_term), float(step_size), float(bc2), float(eps), float(wd),
                                          self.table.data_ptr(), self.scratch.data_ptr(),
                                          found_inf.data_ptr() if found_inf is not None else None, stream),
                   "sd_agc_laprop_step")


def _check(t, what):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"{what}: the fused optimiser needs contiguous fp32 CUDA tensors (no CPU implementation)")


_AGC = _Fused()


@torch.no_grad()
def clip_grad_agc_(parameters, clip, pmin, foreach=None):
    """utils/optim/agc.py:15-60: per tensor g *= 1 / max(||g|| / (clip * max(||p||, pmin)), 1), in place."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    entries = []
    for p in parameters:
        if p.grad is not None:
            _check(p.data, "clip_grad_agc_ parameter")
            _check(p.grad, "clip_grad_agc_ gradient")
            entries.append((p.data, p.grad, None, None))
    if entries:
        _AGC.run(entries, 1, clip, pmin, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, None)


class LaProp(Optimizer):
    def __init__(self, params, lr=4e-4, betas=(0.9, 0.999), eps=1e-15, weight_decay=0, amsgrad=False, centered=False,
                 agc=None, pmin=1e-3):
        if amsgrad or centered:
            raise NotImplementedError("LaProp(amsgrad/centered): not used by the reference configs, not implemented in CUDA")
        # same argument checks and messages as the reference constructor (laprop.py:36-43)
        for ok, msg in ((lr >= 0.0, f"Invalid learning rate: {lr}"), (eps >= 0.0, f"Invalid epsilon value: {eps}"),
                        (0.0 <= betas[0] < 1.0, f"Invalid beta parameter at index 0: {betas[0]}"),
                        (0.0 <= betas[1] < 1.0, f"Invalid beta parameter at index 1: {betas[1]}")):
            if not ok:
                raise ValueError(msg)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                                      centered=centered))
        self.agc, self.pmin = agc, pmin
        self._fused, self._check, self._more = _Fused(), _Fused(), {}

    @torch.no_grad()
    def step(self, inv_scale=1.0, found_inf=None, sync=False):
        """One optimisation step (laprop.py:46-118).  inv_scale / found_inf: GradScaler's unscale factor and (device int32)
        overflow flag, for callers that fold `scaler.unscale_` + `scaler.step` in (dreamer.py:422,433): gradients are unscaled
        BEFORE the AGC clip, and a raised flag skips every tensor update of the step (the finite check runs over all
        tensors before the first update launch).  The scalar state (`step`, `exp_avg_lr_1/2`) lives on the host as in the
        reference: with sync=True the flag is read back after the launches and that state is rolled back when the step
        was skipped (what GradScaler.step does by not calling optimizer.step()); with sync=False (no host sync) call
        `rollback()` yourself once you have seen the flag."""
        self._undo = []
        todo = []
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p.data, "LaProp parameter")
                _check(p.grad, "LaProp gradient")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_lr_1"] = 0.0
                    state["exp_avg_lr_2"] = 0.0
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                self._undo.append((state, state["step"], state["exp_avg_lr_1"], state["exp_avg_lr_2"]))
                state["step"] += 1
                state["exp_avg_lr_1"] = state["exp_avg_lr_1"] * beta1 + (1 - beta1) * group["lr"]
                state["exp_avg_lr_2"] = state["exp_avg_lr_2"] * beta2 + (1 - beta2)
                bc1 = state["exp_avg_lr_1"] / group["lr"] if group["lr"] != 0.0 else 1.0
                key = (1 / bc1, state["exp_avg_lr_2"])
                buckets.setdefault(key, []).append((p.data, p.grad, state["exp_avg"], state["exp_avg_sq"]))
            for (step_size, bc2), entries in buckets.items():
                todo.append((entries, beta1, beta2, (1 - beta1) * group["lr"], step_size, bc2, group["eps"], group["weight_decay"]))
        if found_inf is not None and len(todo) > 1:   # several launches: no tensor may be updated if ANY gradient overflowed
            every = [e for t in todo for e in t[0]]
            self._check.run(every, 2, 0.0, self.pmin, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, found_inf)
        for i, (entries, beta1, beta2, lr_term, step_size, bc2, eps, wd) in enumerate(todo):
            fused = self._fused if i == 0 else self._more.setdefault(i, _Fused())
            fused.run(entries, 0, self.agc if self.agc else 0.0, self.pmin, inv_scale, beta1, beta2, lr_term, step_size, bc2,
                      eps, wd, found_inf)
        if sync and found_inf is not None and int(found_inf.item()) != 0:
            self.rollback()

    def rollback(self):
        """Undo the host-side scalar state of the last step() (call when its found_inf flag was raised)."""
        for state, st, l1, l2 in getattr(self, "_undo", []):
            state["step"], state["exp_avg_lr_1"], state["exp_avg_lr_2"] = st, l1, l2
        self._undo = []
