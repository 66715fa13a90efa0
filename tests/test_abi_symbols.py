"""CPU: the C-ABI library builds, loads, and exports every symbol include/safedreamer.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "safedreamer.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from safe_dreamer_b200 import _lib
    so = _lib.build()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in safedreamer.h but not exported"
    assert set(names) == set(_lib.exported_symbols()), "ctypes signature table out of sync with the header"
    assert lib.sd_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from safe_dreamer_b200 import _lib
    from safe_dreamer_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(D=64, U=16, S=4, K=4, G=2, E=8, A=2)
    lib = _lib.load()
    cfg = _lib.sd_config(D=64, U=16, S=4, K=4, G=2, E=8, A=2, obs_layers=1, img_layers=2, act_kind=0, units=16,
                         actor_layers=3, value_layers=3, reward_layers=1, cont_layers=1, bins=255, unimix=0.01,
                         act_unimix=0.01, min_std=0.1, max_std=1.0, max_rows=4, max_steps=4, max_tape_rows=0)
    assert lib.sd_workspace_bytes(ctypes.byref(cfg)) > 0
    h = ctypes.c_void_p()
    rc = lib.sd_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.sd_last_error_string()
    bad = _lib.sd_config(D=64, U=16, S=4, K=64, G=2, E=8, A=2, obs_layers=1, img_layers=2, act_kind=0, units=16,
                         actor_layers=3, value_layers=3, reward_layers=1, cont_layers=1, bins=255, max_rows=4, max_steps=4)
    assert lib.sd_create(ctypes.byref(bad), ctypes.byref(h)) == -1


def test_module_mirror_state_dict_matches_reference_names():
    """The drop-in RSSM exposes exactly the reference's parameter names/shapes."""
    from types import SimpleNamespace as NS
    from oracle import rssm_oracle as O
    from safe_dreamer_b200.rssm import RSSM
    c = O.Cfg(D=64, U=16, S=4, K=4, G=2, E=8, A=3)
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=0.01, initial="learned",
             device="cpu", obs_layers=1, img_layers=2, dyn_layers=1, blocks=c.G)
    m = RSSM(cfg, c.E, c.A)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in O.rssm_param_shapes(c).items()}
    assert m.feat_size == c.F and m.flat_stoch == c.SK
    import copy
    assert copy.deepcopy(m)._rt.engine is None
    st, dt = m.initial(5)
    assert st.shape == (5, c.S, c.K) and dt.shape == (5, c.D) and float(st.abs().sum() + dt.abs().sum()) == 0
    assert m.get_feat(st, dt).shape == (5, c.F)
