"""Helpers shared by the profiling drivers (no oracle / tests imports: these scripts exercise the product path only)."""
import numpy as np
import torch

from safe_dreamer_b200 import synth as O   # seeded synthetic sizes / weights / inputs (no compute)
from safe_dreamer_b200.engine import Engine


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def make_engine(c, P, max_rows, max_steps, max_tape_rows=0):
    return Engine.from_cfg(c, max_rows, max_steps, max_tape_rows, P)
