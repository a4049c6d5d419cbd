"""argtypes / restype declarations for the remaining C-ABI entry points (kept next to the header)."""
from ctypes import c_float, c_int, c_longlong, c_uint32, c_ulonglong, c_void_p

P, I, L, F, U32, U64 = c_void_p, c_int, c_longlong, c_float, c_uint32, c_ulonglong

SIGNATURES = {
    "prismer_layernorm_fwd": [P, L, P, P, P, L, P, P, I, I, F, P],
    "prismer_layernorm_bwd": [P, L, P, L, P, P, P, P, L, P, L, P, L, P, P, I, I, F, P, U32, P],
}


def declare(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
