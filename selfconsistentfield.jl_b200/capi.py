"""ctypes binding of libscf_b200.so — the same exported symbols the Julia glue
(`julia/JKBuilderCUDA.jl`) reaches with `ccall`, with the same memory layout
(column-major fp64).  There is NO fallback: if the shared library is missing or a call
fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCFB_LIB overrides the library path (kernel experiments only)
LIB_PATH = os.environ.get("SCFB_LIB") or os.path.join(_HERE, "libscf_b200.so")

OK = 0
ERR_ARG, ERR_CUDA, ERR_CUSOLVER, ERR_NCCL, ERR_OOM, ERR_STATE = -1, -2, -3, -4, -5, -6
LAYOUT_FULL, LAYOUT_PACKED8 = 0, 1
ERI_HASH, ERI_CHAIN = 0, 1

_ERR_NAMES = {ERR_ARG: "bad argument", ERR_CUDA: "CUDA error", ERR_CUSOLVER: "cuSOLVER/cuBLAS error",
              ERR_NCCL: "NCCL error", ERR_OOM: "out of device memory", ERR_STATE: "call-sequence error"}


class ScfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libscf_b200: {_ERR_NAMES.get(code, 'error')} ({code}): {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)
_h = C.c_void_p

# name -> (argtypes); every function returns int unless listed in _RESTYPE
_SIGNATURES = {
    "scfb_version": [],
    "scfb_create": [C.POINTER(_h), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    "scfb_create_multi": [C.POINTER(_h), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)],
    "scfb_destroy": [_h],
    "scfb_last_error": [_h],
    "scfb_sync": [_h],
    "scfb_set_stream": [_h, C.c_void_p],
    "scfb_shard_range": [_h, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "scfb_eri_bytes": [_h, C.POINTER(C.c_uint64)],
    "scfb_eri_upload": [_h, _dp],
    "scfb_eri_upload_slabs": [_h, _dp, C.c_int, C.c_int],
    "scfb_eri_mark_ready": [_h],
    "scfb_eri_generate": [_h, C.c_int, C.c_uint64, _dp],
    "scfb_eri_download": [_h, _dp],
    "scfb_allow_partial": [_h, C.c_int],
    "scfb_fock_jk": [_h, C.c_int, _dp, _dp, _dp],
    "scfb_fock_jk_dev": [_h, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    "scfb_jk_split": [_h, C.c_int, _dp, _dp, _dp, _dp],
    "scfb_last_build_ms": [_h, _dp, _dp],
    "scfb_build_ms_history": [_h, C.c_int, _dp, _dp, C.POINTER(C.c_int)],
    "scfb_launch_count": [_h, C.POINTER(C.c_uint64)],
    "scfb_full_plan_query": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)],
    "scfb_scf_setup": [_h, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    "scfb_scf_n_spin": [_h, C.POINTER(C.c_int)],
    "scfb_scf_n_densities": [_h, C.POINTER(C.c_int)],
    "scfb_scf_set_density": [_h, C.c_int, _dp],
    "scfb_scf_set_fock": [_h, C.c_int, _dp],
    "scfb_scf_fock": [_h, _dp, _dp],
    "scfb_scf_roothaan": [_h],
    "scfb_scf_diis_push": [_h, C.c_int, C.POINTER(C.c_int), _dp, _dp],
    "scfb_scf_diis_extrapolate": [_h, C.c_int, C.c_int, _dp],
    "scfb_scf_diis_purge": [_h, C.c_int, C.c_int],
    "scfb_scf_diis_reset": [_h],
    "scfb_scf_accept_fock": [_h],
    "scfb_scf_get": [_h, C.c_int, _dp],
    "scfb_pulay_error": [_h, C.c_int, _dp, _dp, _dp, _dp],
    "scfb_sygvd": [_h, C.c_int, C.c_int, _dp, _dp, _dp, _dp],
    "scfb_syev_jacobi": [_h, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)],
    "scfb_density": [_h, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp],
    "scfb_fock_assemble": [_h, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp],
    "scfb_comm_unique_id": [C.c_char_p],
    "scfb_comm_init": [_h, C.c_char_p],
    "scfb_peer_export": [_h, C.c_char_p],
    "scfb_peer_import": [_h, C.c_char_p],
}
_RESTYPE = {"scfb_last_error": C.c_char_p}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
    return _lib


def check(rc, handle=None):
    if rc != OK:
        msg = lib().scfb_last_error(handle)
        raise ScfbError(rc, msg.decode() if msg else "?")


def as_f64_colmajor(a, shape=None):
    """Dense column-major fp64 view/copy of `a` (what Julia arrays are in memory)."""
    a = np.asarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return np.asfortranarray(a)


def ptr(a):
    return a.ctypes.data_as(_dp)
