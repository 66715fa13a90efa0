"""Weight gradients of the posterior scan / batched prior on tcgen05 (two-term bf16 split, MN-major operands; csrc/sd_wgrad_tc.cuh) against
the 3xTF32 mma.sync kernel (fp32-class accuracy, itself pinned to the reference's autograd by tests/test_gpu_c_bwd.py) at the
benchmark's size: B=16, T=64 = 1024 taped rows.  The kernel is chosen by SD_WGRAD_TC, read once per process, so each variant
runs in its own interpreter (tests/_wgrad_dump.py).
Tolerance: the two-term bf16 split keeps 16 mantissa bits per operand and drops only the lo*lo products (2^-16 relative;
the reference.s own fp32 mode is single-pass TF32, train.py:38): rel. L2 difference <= 1e-4 per tensor, |d| <= 2e-4 ||g||_inf."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dump(tmp_path, tc):
    out = str(tmp_path / f"wg_{tc}.npz")
    env = dict(os.environ, SD_WGRAD_TC=str(tc), PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_wgrad_dump.py"), out], env=env, cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


def test_split_bf16_tensor_core_wgrad_matches_3xtf32(tmp_path):
    ref, tc = _dump(tmp_path, 0), _dump(tmp_path, 1)
    assert sorted(ref.files) == sorted(tc.files)
    checked = 0
    for n in ref.files:
        a, b = tc[n].astype(np.float64), ref[n].astype(np.float64)
        if not np.any(b):
            assert not np.any(a), n
            continue
        rel = np.linalg.norm(a - b) / np.linalg.norm(b)
        worst = np.abs(a - b).max() / np.abs(b).max()
        print(f"{n:40s} rel L2 {rel:.2e}  max |d| / max |g| {worst:.2e}")
        assert rel <= 1e-4 and worst <= 2e-4, n
        checked += 1
    assert checked >= 20
