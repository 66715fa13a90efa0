"""CUDA replacement of the reference's CNN encoder (world_model/networks.py:192-234, ConvEncoder).

`ConvEncoder(config, input_shape)` has the reference's constructor, attribute names (`depths`, `kernel_size`, `out_dim`,
`layers`) and state_dict (`layers.{4i}.weight/bias` = conv, `layers.{4i+2}.weight` = RMS scale), so checkpoints load
unchanged; the modules inside `layers` are parameter containers only -- `forward` goes through the C ABI
(sd_cnn_forward / sd_cnn_backward: implicit-GEMM convolutions on tcgen05).  CUDA only, no CPU path."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib

SD_FLAG_SAVE_TAPE = _lib.SD_FLAG_SAVE_TAPE


class CnnEngine:
    """One sd_cnn handle: frame size, depths and capacity are fixed at creation."""

    def __init__(self, height, width, channels, depths, kernel=5, max_frames=1024, max_tape_frames=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("safe_dreamer_b200: the CNN encoder needs a CUDA device (there is no CPU path)")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cfg = _lib.sd_cnn_config()
        cfg.height, cfg.width, cfg.channels, cfg.layers, cfg.kernel = int(height), int(width), int(channels), len(depths), int(kernel)
        for i, d in enumerate(depths):
            cfg.depths[i] = int(d)
        cfg.max_frames, cfg.max_tape_frames = int(max_frames), int(max_tape_frames)
        self.cfg = cfg
        self.depths = tuple(int(d) for d in depths)
        self.frame = (int(height), int(width), int(channels))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sd_cnn_create(C.byref(cfg), C.byref(h)), "sd_cnn_create")
        self.h = h
        self.embed_size = int(self.lib.sd_cnn_embed_size(self.h))
        self.max_frames, self.max_tape_frames = int(max_frames), int(max_tape_frames)
        self._wkey = None
        self.tape_gen = 0
        self._tape_obs, self._tape_frames = None, 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.sd_cnn_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_weights(self, tensors):
        """tensors: [conv weight, conv bias, RMS scale] per stage (fp32 CUDA, reference layouts)."""
        keep = [t.detach().to(self.device, torch.float32).contiguous() for t in tensors]
        arr = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
        _lib.check(self.lib.sd_cnn_set_weights(self.h, arr, len(keep), self.stream), "sd_cnn_set_weights")

    def forward(self, obs, tape=False, keep_obs=True):
        """obs (..., H, W, C) fp32 CUDA in [0, 1] -> embed (..., embed_size) fp32.
        keep_obs=False: do not hold on to the frames (the caller hands them to `backward` itself; needed under CUDA-graph
        capture, where nothing may keep pool tensors alive behind torch's back)."""
        if not obs.is_cuda:
            raise RuntimeError("safe_dreamer_b200: expected a CUDA tensor (there is no CPU path)")
        lead = obs.shape[:-3]
        if tuple(obs.shape[-3:]) != self.frame:
            raise ValueError(f"frame shape {tuple(obs.shape[-3:])} != {self.frame}")
        x = obs.to(torch.float32).contiguous()
        frames = int(x.numel() // (self.frame[0] * self.frame[1] * self.frame[2]))
        out = torch.empty(frames, self.embed_size, device=x.device, dtype=torch.float32)
        _lib.check(self.lib.sd_cnn_forward(self.h, frames, x.data_ptr(), out.data_ptr(), SD_FLAG_SAVE_TAPE if tape else 0,
                                           self.stream), "sd_cnn_forward")
        if tape:
            self._tape_obs, self._tape_frames = (x if keep_obs else None), frames   # sd_cnn_backward re-reads the frames (stage-1 weight gradient)
            self.tape_gen += 1
        return out.reshape(*lead, self.embed_size)

    def backward(self, d_embed, want_obs_grad=False, weight_grads=None, obs=None):
        """Backward of the last tape=True forward.  d_embed (..., embed_size) -> d_obs (frames, H, W, C) or None;
        weight_grads: list of 3 * layers fp32 CUDA tensors (or None entries) the gradients are ACCUMULATED into."""
        frames = self._tape_frames
        g = d_embed.to(torch.float32).contiguous()
        if g.numel() != frames * self.embed_size:
            raise ValueError(f"d_embed has {g.numel()} elements, the taped forward produced {frames * self.embed_size}")
        d_obs = torch.empty(frames, *self.frame, device=g.device, dtype=torch.float32) if want_obs_grad else None
        arr = None
        if weight_grads is not None:
            arr = (C.c_void_p * len(weight_grads))(*[0 if t is None else t.data_ptr() for t in weight_grads])
        if obs is not None:
            obs = obs.to(torch.float32).contiguous()
        elif self._tape_obs is None and weight_grads is not None:
            raise RuntimeError("CnnEngine.backward: the forward ran with keep_obs=False, pass obs=")
        _lib.check(self.lib.sd_cnn_backward(self.h, frames, g.data_ptr(), None if obs is None else obs.data_ptr(),
                                            None if d_obs is None else d_obs.data_ptr(), arr, self.stream), "sd_cnn_backward")
        return d_obs


class ConvEncoder(nn.Module):
    """Drop-in for networks.ConvEncoder (same constructor and state_dict)."""

    def __init__(self, config, input_shape):
        super().__init__()
        if str(config.act) != "SiLU" or not bool(config.norm):
            raise NotImplementedError("ConvEncoder: only act=SiLU with norm=True (configs/base.yaml) has kernels")
        h, w, input_ch = input_shape
        self.depths = tuple(int(config.depth) * int(m) for m in list(config.mults))
        self.kernel_size = int(config.kernel_size)
        self._input_shape = (int(h), int(w), int(input_ch))
        layers, in_dim = [], input_ch
        for depth in self.depths:
            layers += [nn.Conv2d(in_dim, depth, self.kernel_size, stride=1, bias=True), nn.MaxPool2d(2, 2),
                       nn.RMSNorm(depth, eps=1e-4, dtype=torch.float32), nn.SiLU()]
            in_dim = depth
            h, w = h // 2, w // 2
        self.out_dim = self.depths[-1] * h * w
        self.layers = nn.Sequential(*layers)          # parameter containers: names / shapes of the reference
        self.max_frames = 1024
        self.auto_refresh = False      # True: repack the weights on every call (frozen copies alias live storage, dreamer.py:279)
        self.use_custom_ops = False    # True: go through torch.ops.safedreamer.cnn_encoder (traceable by torch.compile)
        self._ops_key = None
        self._eng = None
        self._wkey = None

    def _tensors(self):
        out = []
        for i in range(len(self.depths)):
            out += [self.layers[4 * i].weight, self.layers[4 * i].bias, self.layers[4 * i + 2].weight]
        return out

    def __deepcopy__(self, memo):
        new = ConvEncoder.__new__(ConvEncoder)
        nn.Module.__init__(new)
        import copy
        for k, v in self.__dict__.items():
            if k in ("_eng", "_wkey", "_ops_key"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _engine(self, frames, tape):
        ts = self._tensors()
        dev = ts[0].device
        eng = self._eng
        need_tape = frames if tape else 0
        if eng is None or eng.device != dev or frames > eng.max_frames or need_tape > eng.max_tape_frames:
            mf = max(frames, self.max_frames, eng.max_frames if eng else 0)
            mt = max(need_tape, eng.max_tape_frames if eng else 0)
            eng = self._eng = CnnEngine(*self._input_shape, self.depths, self.kernel_size, mf, mt, device=dev)
            self._wkey = None
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if key != self._wkey or self.auto_refresh:
            eng.set_weights(ts)
            self._wkey = key
        return eng

    def forward(self, obs):
        """(B, T, H, W, C) in [0, 1] -> (B, T, out_dim) (networks.py:218-234)."""
        if self.use_custom_ops:
            ts = self._tensors()
            taped = torch.is_grad_enabled() and (obs.requires_grad or any(t.requires_grad for t in ts))
            return torch.ops.safedreamer.cnn_encoder(obs.float(), ts, self._ops_key, taped)
        frames = int(obs.numel() // (self._input_shape[0] * self._input_shape[1] * self._input_shape[2]))
        need_grad = torch.is_grad_enabled() and (obs.requires_grad or any(t.requires_grad for t in self._tensors()))
        if need_grad:
            return _EncoderFn.apply(self, obs, *self._tensors())
        return self._engine(frames, False).forward(obs)


class _EncoderFn(torch.autograd.Function):
    """sd_cnn_forward with a tape / sd_cnn_backward.  One tape per engine: a second taped forward before this one's backward
    is detected (generation counter) instead of silently differentiating the wrong frames."""

    @staticmethod
    def forward(ctx, module, obs, *weights):
        frames = int(obs.numel() // (module._input_shape[0] * module._input_shape[1] * module._input_shape[2]))
        eng = module._engine(frames, True)
        out = eng.forward(obs, tape=True)
        ctx.eng, ctx.gen, ctx.obs_shape = eng, eng.tape_gen, tuple(obs.shape)
        ctx.wshapes = [tuple(w.shape) for w in weights]
        return out

    @staticmethod
    def backward(ctx, g):
        eng = ctx.eng
        if eng.tape_gen != ctx.gen:
            raise RuntimeError("ConvEncoder backward: another grad-enabled forward ran on this encoder since this one; "
                               "call backward before the next forward (one activation tape per encoder)")
        need_w = ctx.needs_input_grad[2:]
        wg = [torch.zeros(s, device=g.device, dtype=torch.float32) if n else None for s, n in zip(ctx.wshapes, need_w)]
        d_obs = eng.backward(g, want_obs_grad=ctx.needs_input_grad[1], weight_grads=wg if any(need_w) else None)
        return (None, None if d_obs is None else d_obs.reshape(ctx.obs_shape), *wg)
