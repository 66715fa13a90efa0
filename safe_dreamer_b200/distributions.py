"""Host-side mirror of the distribution helpers the RSSM surface returns
(world_model/distributions.py:16-36, 266-271).  Sampling on the hot path happens inside the
CUDA kernels; these classes exist so `rssm.get_dist(logit).entropy()` / `kl_loss` keep working
for callers (metrics and the differentiable KL terms, dreamer.py:486,575-576)."""
import torch
from torch import distributions as torchd
from torch.nn import functional as F


class OneHotDist(torchd.one_hot_categorical.OneHotCategorical):
    """Unimix categorical (distributions.py:16-36)."""

    def __init__(self, logits, unimix_ratio=0.0):
        probs = F.softmax(logits.float(), dim=-1)
        probs = probs * (1.0 - unimix_ratio) + unimix_ratio / probs.shape[-1]
        super().__init__(logits=torch.log(probs))

    @property
    def mode(self):
        m = F.one_hot(torch.argmax(self.logits, dim=-1), self.logits.shape[-1])
        return m.detach() + self.logits - self.logits.detach()

    def sample(self, **kwargs):
        raise NotImplementedError


def kl(logits_left, logits_right):
    """distributions.py:266-271."""
    lp, rp = torch.log_softmax(logits_left, -1), torch.log_softmax(logits_right, -1)
    return (torch.softmax(logits_left, -1) * (lp - rp)).sum(-1)


def symexp(x):
    return torch.sign(x) * torch.expm1(torch.abs(x))


class _TwoHotLogProb(torch.autograd.Function):
    """TwoHot.log_prob through the C ABI (sd_twohot_logprob / sd_twohot_logprob_bwd)."""

    @staticmethod
    def forward(ctx, logits, bins, target):
        from . import _lib
        lib = _lib.load()
        n = logits.shape[-1]
        lg = logits.reshape(-1, n).float().contiguous()
        tg = target.reshape(-1).float().contiguous()
        out = torch.empty(lg.shape[0], dtype=torch.float32, device=lg.device)
        stream = torch.cuda.current_stream(lg.device).cuda_stream
        _lib.check(lib.sd_twohot_logprob(lg.data_ptr(), n, bins.data_ptr(), n, tg.data_ptr(), lg.shape[0], out.data_ptr(), stream),
                   "sd_twohot_logprob")
        ctx.save_for_backward(lg, bins, tg)
        ctx.shape = logits.shape
        return out.reshape(logits.shape[:-1])

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        lib = _lib.load()
        lg, bins, tg = ctx.saved_tensors
        n = lg.shape[-1]
        gg = g.reshape(-1).float().contiguous()
        d = torch.empty_like(lg)
        stream = torch.cuda.current_stream(lg.device).cuda_stream
        _lib.check(lib.sd_twohot_logprob_bwd(lg.data_ptr(), n, bins.data_ptr(), n, tg.data_ptr(), gg.data_ptr(), lg.shape[0],
                                             d.data_ptr(), n, stream), "sd_twohot_logprob_bwd")
        return d.reshape(ctx.shape), None, None


class TwoHot:
    """distributions.py:67-129 (squash = identity, as symexp_twohot builds it): same constructor and `log_prob(target)` /
    `mode()` surface; log_prob and its gradient run in CUDA."""

    def __init__(self, logits, bins, squash=None, unsquash=None):
        if squash is not None or unsquash is not None:
            raise NotImplementedError("TwoHot with a squash function (the reference never builds one)")
        self.logits = logits.float()
        assert self.logits.shape[-1] == len(bins), (self.logits.shape, len(bins))
        self.bins = bins.float().contiguous()
        self.probs = F.softmax(self.logits, dim=-1)

    def log_prob(self, target):
        assert target.dtype == self.probs.dtype
        return _TwoHotLogProb.apply(self.logits, self.bins, target.squeeze(-1).detach())

    def mode(self):
        n = self.logits.shape[-1]
        p, b = self.probs, self.bins
        if n % 2 == 1:
            m = (n - 1) // 2
            return (p[..., m:m + 1] * b[m:m + 1]).sum(-1, keepdim=True) + (
                (p[..., :m] * b[:m]).flip(dims=(-1,)) + p[..., m + 1:] * b[m + 1:]).sum(-1, keepdim=True)
        h = n // 2
        return ((p[..., :h] * b[:h]).flip(dims=(-1,)) + p[..., h:] * b[h:]).sum(-1, keepdim=True)


def symexp_twohot(logits, bin_num, **kwargs):
    """distributions.py:242-251."""
    if bin_num % 2 == 1:
        half = symexp(torch.linspace(-20, 0, (bin_num - 1) // 2 + 1, dtype=torch.float32, device=logits.device))
        bins = torch.concatenate([half, -half[:-1].flip(dims=(0,))], 0)
    else:
        half = symexp(torch.linspace(-20, 0, bin_num // 2, dtype=torch.float32, device=logits.device))
        bins = torch.concatenate([half, -half.flip(dims=(0,))], 0)
    return TwoHot(logits.float(), bins)
