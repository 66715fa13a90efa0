"""CPU restatement (numpy) of the reference CNN encoder -- TEST INFRASTRUCTURE ONLY.

Follows world_model/networks.py:
  * Conv2dSamePad            (:59-86)   TensorFlow 'SAME' padding, stride 1: pad = k - 1, split floor / ceil
  * RMSNorm2D                (:89-98)   RMSNorm over the channel dimension, eps = 1e-4 as constructed at :212
  * ConvEncoder              (:192-234) obs - 0.5 -> 4 x [conv k x k SAME -> MaxPool2d(2,2) -> RMSNorm2D -> SiLU] -> flatten (C,H,W)
Pinned by tests/golden/cnn_encoder.npz, written by tests/golden/make_golden.py:run_cnn_encoder from the reference's own
ConvEncoder (forward values and autograd gradients).  No CUDA implementation exists yet: this file and its goldens are
the parity gate the B200 kernels of that row will be built against.  Nothing in the product path imports it.

Layout: activations are NHWC here (the layout the B200 kernels will use); weights keep the reference layout
(Cout, Cin, kh, kw); the returned embedding is flattened in the reference's (C, H, W) order."""
import numpy as np

RMS_EPS = 1e-4


def same_pad(i, k, s=1, d=1):
    """networks.py:62-64."""
    return max(((i + s - 1) // s - 1) * s + (k - 1) * d + 1 - i, 0)


def _pad_same(x, k):
    ph, pw = same_pad(x.shape[1], k), same_pad(x.shape[2], k)
    return np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0))), (ph // 2, pw // 2)


def conv_same(x, w, b):
    """x (N,H,W,Cin), w (Cout,Cin,k,k), b (Cout) -> (N,H,W,Cout); networks.py:66-86 with stride 1."""
    k = w.shape[-1]
    xp, _ = _pad_same(x, k)
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(1, 2))      # (N,H,W,Cin,k,k)
    return np.einsum("nhwcij,ocij->nhwo", win, w, optimize=True).astype(x.dtype) + b


def conv_same_bwd(x, w, dy):
    """-> (dx, dw, db) of conv_same."""
    k = w.shape[-1]
    xp, (p0, p1) = _pad_same(x, k)
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(1, 2))
    dw = np.einsum("nhwcij,nhwo->ocij", win, dy, optimize=True).astype(x.dtype)
    db = dy.sum((0, 1, 2)).astype(x.dtype)
    dxp = np.zeros_like(xp)
    H, W = x.shape[1], x.shape[2]
    for i in range(k):
        for j in range(k):
            dxp[:, i:i + H, j:j + W, :] += np.einsum("nhwo,oc->nhwc", dy, w[:, :, i, j], optimize=True)
    return dxp[:, p0:p0 + H, p1:p1 + W, :].astype(x.dtype), dw, db


def maxpool2(x):
    """MaxPool2d(2, 2) (networks.py:209): returns the pooled map and the flat arg-max (first maximum, as torch) per window."""
    N, H, W, C = x.shape
    xr = x[:, :H // 2 * 2, :W // 2 * 2].reshape(N, H // 2, 2, W // 2, 2, C).transpose(0, 1, 3, 5, 2, 4).reshape(N, H // 2, W // 2, C, 4)
    arg = xr.argmax(-1)
    return np.take_along_axis(xr, arg[..., None], -1)[..., 0], arg


def maxpool2_bwd(dy, arg, shape):
    N, H, W, C = shape
    d = np.zeros(dy.shape + (4,), dy.dtype)
    np.put_along_axis(d, arg[..., None], dy[..., None], -1)
    dx = np.zeros(shape, dy.dtype)
    dx[:, :H // 2 * 2, :W // 2 * 2] = d.reshape(N, H // 2, W // 2, C, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(N, H // 2 * 2, W // 2 * 2, C)
    return dx


def silu(x):
    return x / (1.0 + np.exp(-x))


def norm_act(x, g):
    """RMSNorm over channels (eps 1e-4) then SiLU (networks.py:89-98,210-213)."""
    rho = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + x.dtype.type(RMS_EPS))
    return silu(x * rho * g).astype(x.dtype)


def norm_act_bwd(x, g, dy):
    """-> (dx, dg); same algebra as the MLP layers."""
    f = x.dtype.type
    rho = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + f(RMS_EPS))
    n = x * rho
    m = n * g
    sg = 1.0 / (1.0 + np.exp(-m))
    dm = dy * (sg * (1.0 + m * (1.0 - sg)))
    dg = (dm * n).sum((0, 1, 2)).astype(x.dtype)
    dn = dm * g
    dx = rho * (dn - n * (dn * n).mean(-1, keepdims=True))
    return dx.astype(x.dtype), dg


from safe_dreamer_b200.synth import encoder_params  # noqa: E402,F401  (seeded weights: shared with the benchmark)


def round_bf16(x):
    """fp32 -> nearest-even bfloat16 -> fp32: the rounding the tcgen05 path applies to every conv operand."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def encoder_fwd(P, obs, n_layers=4, tape=None, rnd=None):
    """ConvEncoder.forward (networks.py:218-234): obs (..., H, W, C) in [0, 1] -> (..., Cf*Hf*Wf).
    rnd (e.g. round_bf16): applied to every convolution operand (stage inputs and conv weights), which models the CUDA
    path's bf16 operands / fp32 accumulation; None = the reference's fp32 arithmetic."""
    rnd = rnd or (lambda a: a)
    lead = obs.shape[:-3]
    x = rnd((obs - obs.dtype.type(0.5)).reshape((-1,) + obs.shape[-3:]))
    for i in range(n_layers):
        w, b, g = rnd(P[f"layers.{4 * i}.weight"]), P[f"layers.{4 * i}.bias"], P[f"layers.{4 * i + 2}.weight"]
        y = conv_same(x, w, b)
        p, arg = maxpool2(y)
        if tape is not None:
            tape.append((x, y.shape, arg, p))
        x = norm_act(p, g)
        if i + 1 < n_layers:
            x = rnd(x)
    out = x.transpose(0, 3, 1, 2).reshape(x.shape[0], -1)        # flatten in (C, H, W) order
    if tape is not None:
        tape.append(x.shape)
    return out.reshape(lead + (out.shape[-1],))


def encoder_bwd(P, tape, d_out, n_layers=4, rnd=None):
    """-> (d_obs, {name: grad}) given d(loss)/d(embedding).  rnd: as in encoder_fwd (conv weights and the conv-output
    gradient dy are the rounded operands of the backward convolutions; the tape already holds the rounded stage inputs)."""
    rnd = rnd or (lambda a: a)
    N, Hf, Wf, Cf = tape[-1]
    dx = d_out.reshape(N, Cf, Hf, Wf).transpose(0, 2, 3, 1)
    G = {}
    for i in reversed(range(n_layers)):
        x, yshape, arg, p = tape[i]
        w, g = rnd(P[f"layers.{4 * i}.weight"]), P[f"layers.{4 * i + 2}.weight"]
        dp, G[f"layers.{4 * i + 2}.weight"] = norm_act_bwd(p, g, dx)
        dy = rnd(maxpool2_bwd(dp, arg, yshape))
        dx, G[f"layers.{4 * i}.weight"], G[f"layers.{4 * i}.bias"] = conv_same_bwd(x, w, dy)
    return dx, G
