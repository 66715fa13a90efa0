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
