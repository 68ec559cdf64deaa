"""ctypes wrapper of oracle/jk_oracle.c (CPU ORACLE — test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libjk_oracle.so")
_dp = C.POINTER(C.c_double)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_PATH)
        _lib.jk_literal_ld.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        _lib.jk_literal_ld_slabs.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        _lib.jk_literal_ld_slabs.restype = None
        _lib.permute_mban_slabs.argtypes = [_dp, C.c_int, C.c_int, _dp]
        _lib.permute_mban_slabs.restype = None
        _lib.jk_onepass_omp.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
        _lib.hash_eri_slabs.argtypes = [_dp, C.c_int, C.c_uint64, C.c_int, C.c_int]
        _lib.chain_eri_slabs.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int]
        for f in (_lib.jk_literal_ld, _lib.jk_onepass_omp, _lib.hash_eri_slabs, _lib.chain_eri_slabs):
            f.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def jk_literal_ld(eri, density, total_density):
    """(J, K) of the literal contraction, long-double accumulation (arbiter at larger n)."""
    n, n_spin = density.shape[0], density.shape[2]
    e = np.asfortranarray(eri, dtype=np.float64)
    d = np.asfortranarray(density, dtype=np.float64)
    t = np.asfortranarray(total_density, dtype=np.float64)
    j = np.zeros((n, n), order="F")
    k = np.zeros((n, n, n_spin), order="F")
    lib().jk_literal_ld(_p(e), n, n_spin, _p(d), _p(t), _p(j), _p(k))
    return j, k


def jk_literal_ld_slabs(eri_slabs, density, total_density):
    """Columns J[:, nu_s] (n, S) and K[:, nu_s, spin] (n, S, n_spin) of the literal contraction for the
    S nu-slabs in `eri_slabs` ((n,n,n,S), F-ordered), long-double accumulation."""
    n, n_spin = density.shape[0], density.shape[2]
    e = np.asfortranarray(eri_slabs, dtype=np.float64)
    n_slabs = e.shape[3]
    d = np.asfortranarray(density, dtype=np.float64)
    t = np.asfortranarray(total_density, dtype=np.float64)
    j = np.zeros((n, n_slabs), order="F")
    k = np.zeros((n, n_slabs, n_spin), order="F")
    lib().jk_literal_ld_slabs(_p(e), n, n_spin, n_slabs, _p(d), _p(t), _p(j), _p(k))
    return j, k


def permute_mban(eri_slabs, out=None):
    """P[a,b,m,v] = eri_slabs[m,b,a,v] (blocked OpenMP transpose; the K temporary of the reference)."""
    e = np.asfortranarray(eri_slabs, dtype=np.float64)
    n, n_slabs = e.shape[0], e.shape[3]
    if out is None:
        out = np.empty_like(e, order="F")
    lib().permute_mban_slabs(_p(e), n, n_slabs, _p(out))
    return out


def jk_onepass(eri_slabs, density, total_density, nu_lo, nu_hi):
    n, n_spin = density.shape[0], density.shape[2]
    d = np.asfortranarray(density, dtype=np.float64)
    t = np.asfortranarray(total_density, dtype=np.float64)
    j = np.zeros((n, n), order="F")
    k = np.zeros((n, n, n_spin), order="F")
    lib().jk_onepass_omp(_p(eri_slabs), n, n_spin, _p(d), _p(t), nu_lo, nu_hi, _p(j), _p(k))
    return j, k


def hash_eri(n, seed, nu_lo=0, nu_hi=None):
    nu_hi = n if nu_hi is None else nu_hi
    out = np.empty((n, n, n, nu_hi - nu_lo), order="F")
    lib().hash_eri_slabs(_p(out), n, seed, nu_lo, nu_hi)
    return out


def chain_eri(n, aux, nu_lo=0, nu_hi=None):
    nu_hi = n if nu_hi is None else nu_hi
    a = np.ascontiguousarray(aux, dtype=np.float64)
    out = np.empty((n, n, n, nu_hi - nu_lo), order="F")
    lib().chain_eri_slabs(_p(out), n, _p(a), nu_lo, nu_hi)
    return out
