"""Parameter containers mirroring world_model/networks.py:313-377 (MLP / MLPHead) so head
state_dicts interchange with the reference.  Their *frozen* evaluation on the hot path (in-loop
actor, reward/cont/value on imagined feats) runs in CUDA via dreamer_ops."""
import torch
from torch import nn

from .rssm import weight_init_


class MLP(nn.Module):
    def __init__(self, name, layers, units, inp_dim):
        super().__init__()
        self.layers = nn.Sequential()
        for i in range(layers):
            self.layers.add_module(f"{name}_linear{i}", nn.Linear(inp_dim, units, bias=True))
            self.layers.add_module(f"{name}_norm{i}", nn.RMSNorm(units, eps=1e-04, dtype=torch.float32))
            self.layers.add_module(f"{name}_act{i}", nn.SiLU())
            inp_dim = units
        self.out_dim = units


class MLPHead(nn.Module):
    """networks.py:339-377: `mlp.layers.{name}_linear{i}` / `_norm{i}` + `last`."""

    def __init__(self, name, layers, units, inp_dim, out_dim, outscale=1.0):
        super().__init__()
        self.mlp = MLP(name, layers, units, inp_dim)
        self.last = nn.Linear(units, out_dim, bias=True)
        self.mlp.apply(weight_init_)
        self.last.apply(weight_init_)
        if outscale != 1.0:
            with torch.no_grad():
                self.last.weight.mul_(outscale)
        self.n_layers, self.units, self.out_dim = layers, units, out_dim


class ReturnEMA(nn.Module):
    """networks.py:405-422: running 5 % / 95 % return quantiles.  Same constructor, buffer name (`ema_vals`) and call
    signature as the reference; the quantile + EMA update runs in one CUDA kernel (sd_return_ema: exact radix select,
    torch.quantile's fp32 rank / lerp arithmetic), with no host synchronisation."""

    def __init__(self, device, alpha=1e-2):
        super().__init__()
        self.device = device
        self.alpha = alpha
        self.register_buffer("ema_vals", torch.zeros(2, dtype=torch.float32, device=self.device))

    def __call__(self, x):
        from . import _lib
        x = x.detach()
        if not x.is_cuda:
            raise RuntimeError("ReturnEMA: expected a CUDA tensor (the hot path has no CPU implementation)")
        flat = x.reshape(-1).float().contiguous()
        out = torch.empty(2, dtype=torch.float32, device=x.device)   # [offset, scale]
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().sd_return_ema(flat.data_ptr(), flat.numel(), float(self.alpha), self.ema_vals.data_ptr(),
                                             out.data_ptr(), out.data_ptr() + 4, stream), "sd_return_ema")
        return out[0].detach(), out[1].detach()
