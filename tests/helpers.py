"""Shared helpers for the tests: golden loading and near-tie accounting."""
import os

import numpy as np

from oracle import rssm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)
    kv = dict(zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist()))
    kw = {}
    for k, v in kv.items():
        if k in ("act_kind",):
            kw[k] = v
        elif k in ("unimix", "act_unimix", "min_std", "max_std", "lamb"):
            kw[k] = float(v)
        else:
            kw[k] = int(v)
    return O.Cfg(**kw), z


def golden_params(c, z):
    """Weights for a golden case: stored ones when present, else regenerated from the seed."""
    P = O.init_params(c, seed=0)
    stored = [k for k in z.files if k.startswith("P/")]
    for k in stored:
        _, mod, name = k.split("/", 2)
        np.testing.assert_array_equal(P[mod][name], z[k], err_msg=f"numpy RNG drift in {k}")
    return P


def golden_initial(c, B):
    rng = np.random.Generator(np.random.Philox(7))
    idx = rng.integers(0, c.K, size=(B, c.S))
    stoch = np.eye(c.K, dtype=np.float32)[idx]
    deter = np.tanh(rng.standard_normal((B, c.D), dtype=np.float32)).astype(np.float32)
    return stoch, deter


def top2_gap(score):
    """Gap between best and second-best of the last axis (near-tie detector)."""
    s = np.sort(score, axis=-1)
    return s[..., -1] - s[..., -2]


def index_mismatch_report(idx_a, idx_b, score, tol):
    """Count mismatching categorical indices and how many are NOT explained by a near tie.

    score: perturbed logits (l + g) of the checker, same leading shape as idx.
    Returns (n_mismatch, n_unexplained, n_total).
    """
    mism = idx_a != idx_b
    gap = top2_gap(score)
    unexplained = mism & (gap > tol)
    return int(mism.sum()), int(unexplained.sum()), int(mism.size)


# ----------------------------------------------------------------------------- GPU-side helpers
def make_engine(c, P, max_rows, max_steps, max_tape_rows=0):
    """Engine for config `c` with weights `P` loaded (CUDA only)."""
    from safe_dreamer_b200.engine import Engine
    return Engine.from_cfg(c, max_rows, max_steps, max_tape_rows, P)


def cu(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def assert_indices(idx_gpu, idx_ref, score_ref, tol, max_rate, what):
    """Categorical indices must match except at near ties (top-2 gap of the checker's perturbed
    logits below `tol`); the near-tie rate is printed (logged) and bounded by `max_rate`."""
    n_mis, n_bad, n = index_mismatch_report(np.asarray(idx_gpu), np.asarray(idx_ref), score_ref, tol)
    print(f"[near-tie log] {what}: {n_mis}/{n} indices differ ({n_mis / n:.2e}), {n_bad} not explained by a gap < {tol}")
    assert n_bad == 0, f"{what}: {n_bad} index mismatches are not near ties"
    assert n_mis / n <= max_rate, f"{what}: mismatch rate {n_mis / n:.3e} > {max_rate}"
    return n_mis


def perturbed_scores(logit, u, unimix):
    """l + g of the oracle (what argmax is taken over)."""
    l = O.unimix_logits(logit, unimix)
    return l + (-np.log(-np.log(u.astype(logit.dtype))))
