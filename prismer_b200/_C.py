"""ctypes binding of ``libprismer_sm100.so`` (the C-ABI declared in ``include/prismer_sm100.h``).

There is NO fallback: if the library is missing, or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_longlong, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRISMER_LIB selects another in-tree BUILD of the same library (e.g. the -DPRISMER_PDL variant); there is still no fallback
LIB_PATH = os.environ.get("PRISMER_LIB") or os.path.join(_HERE, "libprismer_sm100.so")

ERRORS = {-1: "bad shape / argument", -2: "misaligned pointer or leading dimension", -3: "unsupported architecture (needs sm_100)",
          -4: "CUDA runtime error", -5: "CUDA driver entry point (cuTensorMapEncodeTiled) unavailable"}


class PrismerError(RuntimeError):
    pass


class GemmArgs(Structure):
    _fields_ = [
        ("A", c_void_p), ("B", c_void_p), ("C", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("lda", c_longlong), ("ldb", c_longlong), ("ldc", c_longlong),
        ("transA", c_int), ("transB", c_int),
        ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_longlong),
        ("aux_out", c_void_p), ("aux_in", c_void_p), ("ldaux", c_longlong),
        ("act", c_int), ("act_grad", c_int), ("out_fp32", c_int), ("accumulate", c_int),
        ("alpha", c_float), ("drop_p", c_float), ("seed", c_void_p), ("rng_stream", c_uint32),
        ("force_bn", c_int), ("max_ctas", c_int), ("force_splits", c_int),
    ]


_lib = None
PROFILE = None   # set to a list to record (entry point, start_event, end_event) around every C-ABI call (bench.py / tools)


class _ProfilingLib:
    """Proxy used while PROFILE is a list: brackets every C-ABI call with CUDA events on the current stream."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith("prismer_") or name in ("prismer_abi_version", "prismer_check_device"):
            return fn
        import torch

        def wrapped(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            PROFILE.append((name, e0, e1))
            return rc
        return wrapped


def lib():
    real = _real_lib()
    return _ProfilingLib(real) if PROFILE is not None else real


def _real_lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:     # in-tree build (needs nvcc); this is NOT a fallback path -- without the CUDA library nothing runs
                from . import build as _build
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise PrismerError(f"{LIB_PATH} is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise PrismerError(
                f"{LIB_PATH} not found: build it with `python -m prismer_b200.build` "
                "(or `__graft_entry__.build()`); prismer_b200 has no CPU / PyTorch fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.prismer_abi_version.restype = c_int
    L.prismer_check_device.restype = c_int
    L.prismer_gemm_bf16.argtypes = [POINTER(GemmArgs), c_void_p]
    L.prismer_gemm_bf16.restype = c_int
    for fn in ("prismer_attention_fwd", "prismer_attention_bwd"):
        getattr(L, fn).argtypes = [POINTER(AttnArgs), c_void_p]
        getattr(L, fn).restype = c_int
    from . import _C_decl
    _C_decl.declare(L)


CALLS = 0   # number of C-ABI calls issued (each enqueues >= 1 kernel); bench.py reports it as gpu_launches


def check(rc: int, what: str = ""):
    global CALLS
    CALLS += 1
    if rc != 0:
        raise PrismerError(f"libprismer_sm100 {what} failed: {ERRORS.get(rc, rc)} (code {rc})")


class AttnArgs(Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p),
        ("q_bs", c_longlong), ("q_rs", c_longlong), ("k_bs", c_longlong), ("k_rs", c_longlong),
        ("v_bs", c_longlong), ("v_rs", c_longlong), ("o_bs", c_longlong), ("o_rs", c_longlong),
        ("lse", c_void_p), ("key_mask", c_void_p),
        ("B", c_int), ("H", c_int), ("Lq", c_int), ("Lk", c_int), ("d", c_int), ("causal", c_int),
        ("scale", c_float), ("drop_p", c_float), ("seed", c_void_p), ("rng_stream", c_uint32),
        ("dout", c_void_p), ("do_bs", c_longlong), ("do_rs", c_longlong),
        ("dq", c_void_p), ("dq_bs", c_longlong), ("dq_rs", c_longlong),
        ("dk", c_void_p), ("dk_bs", c_longlong), ("dk_rs", c_longlong),
        ("dv", c_void_p), ("dv_bs", c_longlong), ("dv_rs", c_longlong),
        ("delta", c_void_p),
        ("kv_div", c_int),
    ]
