"""Device-side store of the replay buffer's latents: the B200 counterpart of `Buffer.update` and of the `initial` read in
`Buffer.sample` (reference utils/buffer.py:40,44-53; write-back call site dreamer.py:450).

The reference keeps stoch as a (length, envs, S, K) one-hot and deter as (length, envs, D) inside torchrl's
LazyTensorStorage and scatters the B*T freshly inferred rows back after every update.  Here the scatter is one CUDA launch
(`sd_latent_writeback`) into plain device tensors, with stoch held as uint8 class indices (S bytes per row instead of
4*S*K), and the gather (`sd_latent_gather`) decodes them back to exact one-hots.  No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class LatentStore:
    """`index` follows the reference: a pair [index0, index1] of integer tensors of shape (B, T) (or flat), where the slot
    of a row is storage[index[1], index[0]] -- index[1] addresses the length dimension, index[0] the environment."""

    def __init__(self, length, envs, stoch, discrete, deter, device="cuda", keep_onehot=False):
        self.length, self.envs, self.S, self.K, self.D = int(length), int(envs), int(stoch), int(discrete), int(deter)
        if self.K > 256:
            raise ValueError("class indices are stored as uint8: discrete must be <= 256")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("LatentStore lives on the GPU (no CPU fallback)")
        self.idx = torch.zeros(self.length, self.envs, self.S, dtype=torch.uint8, device=dev)
        self.deter = torch.zeros(self.length, self.envs, self.D, dtype=torch.float32, device=dev)
        self.onehot = (torch.zeros(self.length, self.envs, self.S, self.K, dtype=torch.float32, device=dev)
                       if keep_onehot else None)
        self._bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lib = _lib.load()

    def _index(self, index):
        i0 = index[1].reshape(-1).to(device=self.idx.device, dtype=torch.int64).contiguous()
        i1 = index[0].reshape(-1).to(device=self.idx.device, dtype=torch.int64).contiguous()
        if i0.numel() != i1.numel():
            raise ValueError("index[0] and index[1] must have the same number of elements")
        return i0, i1

    def _check(self, validate, what):
        if validate and int(self._bad.item()) != 0:     # one device->host read; the reference raises IndexError here
            n = int(self._bad.item())
            self._bad.zero_()
            raise IndexError(f"{what}: {n} row(s) address a slot outside the ({self.length}, {self.envs}) storage")

    def update(self, index, stoch, deter, validate=True):
        """Buffer.update: stoch (B, T, S, K), deter (B, T, D); the row with the largest flat position wins a repeated slot."""
        i0, i1 = self._index(index)
        R = i0.numel()
        st = stoch.detach().reshape(-1, self.S, self.K).to(torch.float32).contiguous()
        dt = deter.detach().reshape(-1, self.D).to(torch.float32).contiguous()
        if st.shape[0] != R or dt.shape[0] != R:
            raise ValueError("index, stoch and deter disagree on the number of rows")
        if not (st.is_cuda and dt.is_cuda):
            raise RuntimeError("LatentStore.update needs CUDA tensors (no CPU fallback)")
        stream = torch.cuda.current_stream(self.idx.device).cuda_stream
        _lib.check(self.lib.sd_latent_writeback(_ptr(i1), _ptr(i0), R, _ptr(st), _ptr(dt), self.S, self.K, self.D,
                                                self.length, self.envs, _ptr(self.idx), _ptr(self.onehot),
                                                _ptr(self.deter), _ptr(self._bad), stream), "sd_latent_writeback")
        self._check(validate, "LatentStore.update")

    def initial(self, index, validate=True):
        """(stoch (R, S, K) exact one-hot, deter (R, D)) of the rows at storage[index[1], index[0]]."""
        i0, i1 = self._index(index)
        R = i0.numel()
        st = torch.empty(R, self.S, self.K, dtype=torch.float32, device=self.idx.device)
        dt = torch.empty(R, self.D, dtype=torch.float32, device=self.idx.device)
        stream = torch.cuda.current_stream(self.idx.device).cuda_stream
        _lib.check(self.lib.sd_latent_gather(_ptr(i1), _ptr(i0), R, self.S, self.K, self.D, self.length, self.envs,
                                             _ptr(self.idx), _ptr(self.deter), _ptr(st), _ptr(dt), _ptr(self._bad), stream),
                   "sd_latent_gather")
        self._check(validate, "LatentStore.initial")
        return st, dt
