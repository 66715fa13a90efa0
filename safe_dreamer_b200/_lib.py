"""ctypes binding of libsafedreamer.so (include/safedreamer.h) + the in-tree nvcc build.

PyTorch is plumbing here: it owns device memory and streams; every compute call
goes through the C ABI with raw device pointers.  There is no CPU fallback: if the
shared library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libsafedreamer.so")
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "safedreamer.h")
# translation units of the library and the headers each one includes (a unit is recompiled when any of them is newer
# than its object file under csrc/_obj/)
_UNITS = {
    "sd_api.cu": ["sd_kernels.cuh", "sd_tc.cuh", "sd_bwd.cuh", "sd_chain.cuh", "sd_tc2.cuh", "sd_scan.cuh", "sd_pimg.cuh", "sd_wgrad_tc.cuh",
                  "sd_internal.h"],
    "sd_cnn.cu": ["sd_cnn.cuh", "sd_cnn_bwd.cuh", "sd_tc.cuh", "sd_internal.h"],
}
_SOURCES = sorted(set(_UNITS) | {h for hs in _UNITS.values() for h in hs})
_OBJ = os.path.join(_CSRC, "_obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

SD_FLAG_BF16, SD_FLAG_SAVE_TAPE, SD_FLAG_GRAPH, SD_FLAG_FEATS_FROM_IMAGINE, SD_FLAG_BACKGROUND = 1, 2, 4, 8, 16
SD_FLAG_PERSISTENT, SD_FLAG_LAYERWISE = 32, 64
MOD_RSSM, MOD_ACTOR, MOD_REWARD, MOD_CONT, MOD_VALUE, MOD_SLOW_VALUE = range(6)


class sd_opt_tensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]


class sd_cnn_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("height", "width", "channels", "layers", "kernel")] + \
               [("depths", C.c_int32 * 8), ("max_frames", C.c_int32), ("max_tape_frames", C.c_int32)]


class sd_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("D", "U", "S", "K", "G", "E", "A", "obs_layers", "img_layers", "act_kind", "units",
                 "actor_layers", "value_layers", "reward_layers", "cont_layers", "bins")] + \
               [(n, C.c_float) for n in ("unimix", "act_unimix", "min_std", "max_std")] + \
               [(n, C.c_int32) for n in ("max_rows", "max_steps", "max_tape_rows")]


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    srcs = [os.path.join(_CSRC, s) for s in _SOURCES] + [_INCLUDE]
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU). Returns the .so path.

    One process per GPU means several ranks may get here at once: the build runs under an exclusive file lock, writes to
    a temporary file and renames it into place, so nobody ever loads a half-written library."""
    if not force and not _stale():
        return _SO
    import fcntl
    with open(_SO + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another rank built it while we waited
                return _SO
            nvcc = os.environ.get("NVCC", "nvcc")
            os.makedirs(_OBJ, exist_ok=True)
            jobs = []
            for unit, headers in _UNITS.items():
                obj = os.path.join(_OBJ, unit + ".o")
                deps = [os.path.join(_CSRC, unit), _INCLUDE] + [os.path.join(_CSRC, h) for h in headers]
                if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
                    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("SD_NVCC_EXTRA", "").split() + ["-c", os.path.join(_CSRC, unit), "-o", obj]
                    if verbose:
                        print(" ".join(cmd))
                    jobs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            for unit, proc in jobs:      # the units compile side by side
                out, _ = proc.communicate()
                if proc.returncode != 0:
                    obj = os.path.join(_OBJ, unit + ".o")
                    if os.path.exists(obj):
                        os.remove(obj)
                    raise RuntimeError(f"nvcc failed on {unit}:\n" + out)
                if verbose and out.strip():
                    print(out)
            tmp = f"{_SO}.tmp.{os.getpid()}"
            cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + \
                  [os.path.join(_OBJ, u + ".o") for u in _UNITS] + ["-o", tmp]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, _SO)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return _SO


_lib = None
_lock = threading.Lock()

_P = C.c_void_p
_SIGS = {
    "sd_abi_version": (C.c_int, []),
    "sd_last_error_string": (C.c_char_p, []),
    "sd_workspace_bytes": (C.c_size_t, [C.POINTER(sd_config)]),
    "sd_create": (C.c_int, [C.POINTER(sd_config), C.POINTER(_P)]),
    "sd_destroy": (C.c_int, [_P]),
    "sd_weight_count": (C.c_int, [_P, C.c_int]),
    "sd_weight_name": (C.c_char_p, [_P, C.c_int, C.c_int]),
    "sd_weight_numel": (C.c_int64, [_P, C.c_int, C.c_int]),
    "sd_set_weights": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.c_int, _P]),
    "sd_observe_fwd": (C.c_int, [_P, C.c_int, C.c_int] + [_P] * 9 + [C.c_uint32, _P]),
    "sd_observe_bwd": (C.c_int, [_P, C.c_int, C.c_int] + [_P] * 6 + [C.POINTER(_P), C.c_uint32, _P]),
    "sd_prior": (C.c_int, [_P, C.c_int] + [_P] * 4 + [C.c_uint32, _P]),
    "sd_prior_bwd": (C.c_int, [_P, C.c_int] + [_P] * 3 + [C.POINTER(_P), C.c_uint32, _P]),
    "sd_imagine_with_action": (C.c_int, [_P, C.c_int, C.c_int] + [_P] * 6 + [C.c_uint32, _P]),
    "sd_imagine_fwd": (C.c_int, [_P, C.c_int, C.c_int] + [_P] * 6 + [C.c_uint32, _P]),
    "sd_imagine_bwd": (C.c_int, [_P, C.c_int, C.c_int] + [_P] * 4 + [C.c_uint32, _P]),
    "sd_heads_lambda_fwd": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_float, C.c_float] + [_P] * 6 + [C.c_uint32, _P]),
    "sd_heads_lambda_bwd": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_float, C.c_float] + [_P] * 5 + [C.c_uint32, _P]),
    "sd_lambda_return": (C.c_int, [C.c_int, C.c_int] + [_P] * 5 + [C.c_float, C.c_float, _P, _P]),
    "sd_kl_loss": (C.c_int, [_P, C.c_int, _P, _P, C.c_float] + [_P] * 4 + [_P]),
    "sd_kl_loss_bwd": (C.c_int, [_P, C.c_int, _P, _P, C.c_float] + [_P] * 4 + [_P]),
    "sd_twohot_logprob": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P]),
    "sd_twohot_logprob_bwd": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P]),
    "sd_barlow_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sd_barlow_loss": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "sd_opt_table_bytes": (C.c_size_t, [C.c_int]),
    "sd_opt_scratch_bytes": (C.c_size_t, [C.POINTER(sd_opt_tensor), C.c_int]),
    "sd_agc_laprop_step": (C.c_int, [C.POINTER(sd_opt_tensor), C.c_int, C.c_int] + [C.c_float] * 11 + [_P, _P, _P, _P]),
    "sd_return_ema": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P, _P, _P]),
    "sd_latent_writeback": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [_P] * 5),
    "sd_latent_gather": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [_P] * 6),
    "sd_launch_count": (C.c_uint64, []),
    "sd_scan_mode": (C.c_int, [C.c_int]),
    "sd_cnn_create": (C.c_int, [C.POINTER(sd_cnn_config), C.POINTER(_P)]),
    "sd_cnn_destroy": (C.c_int, [_P]),
    "sd_cnn_embed_size": (C.c_int64, [_P]),
    "sd_cnn_set_weights": (C.c_int, [_P, C.POINTER(_P), C.c_int, _P]),
    "sd_cnn_forward": (C.c_int, [_P, C.c_int, _P, _P, C.c_uint32, _P]),
    "sd_cnn_backward": (C.c_int, [_P, C.c_int, _P, _P, _P, C.POINTER(_P), _P]),
}


def exported_symbols():
    return sorted(_SIGS)


def load():
    """Load the shared library (building it if the sources are newer). Raises if unavailable."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            build()
        lib = C.CDLL(_SO)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.sd_abi_version() != 1:
            raise RuntimeError("libsafedreamer ABI mismatch")
        _lib = lib
        return lib


def last_error():
    return load().sd_last_error_string().decode()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def launch_count():
    return int(load().sd_launch_count())


def scan_mode(set=-1):
    """include/safedreamer.h: sd_scan_mode (hand-off mode of the persistent posterior scan on the current device)."""
    return int(load().sd_scan_mode(int(set)))
