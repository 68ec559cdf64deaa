"""CPU ORACLE (test infrastructure only) — host twins of the synthetic input families.

The product generates these tensors directly in HBM
(`selfconsistentfield.jl_b200/csrc/eri_gen.cu`); this module is the independent NumPy
statement of the same frozen formulas (DESIGN.md "synthetic inputs") used
by the tests to check the device generators bit for bit at small n, and to feed the
oracle the same tensor the device holds.  Index order is the reference's:
eri[a,b,c,d] = (ab|cd), the pairs that `JKBuilderFromTensor.jl:37` contracts.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SEED_ERI, SEED_DA, SEED_DB = 20260101, 20260102, 20260103


def _splitmix(z):
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _uniform_pm1(idx, seed):
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed + 1) * GOLDEN
    z = _splitmix(z)
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0


def _pair(i, j):
    hi = np.maximum(i, j).astype(np.uint64)
    lo = np.minimum(i, j).astype(np.uint64)
    return hi * (hi + np.uint64(1)) // np.uint64(2) + lo


def hash_eri(n, seed=SEED_ERI, nu_lo=0, nu_hi=None):
    """Dense 8-fold-symmetric tensor, every element O(1) uniform(-1,1); returns the
    nu-slabs [nu_lo, nu_hi) as an F-ordered (n,n,n,hi-lo) array."""
    nu_hi = n if nu_hi is None else nu_hi
    a = np.arange(n)
    p1 = _pair(a[:, None], a[None, :])                      # (a,b)
    p2 = _pair(a[:, None], np.arange(nu_lo, nu_hi)[None, :])  # (c,d)
    p = np.maximum(p1[:, :, None, None], p2[None, None, :, :])
    q = np.minimum(p1[:, :, None, None], p2[None, None, :, :])
    idx = p * (p + np.uint64(1)) // np.uint64(2) + q
    return np.asfortranarray(_uniform_pm1(idx, seed))


def hash_symmetric_matrix(n, seed):
    """Symmetric n x n matrix, elements uniform(-1,1) from the same hash."""
    a = np.arange(n)
    return np.asfortranarray(_uniform_pm1(_pair(a[:, None], a[None, :]), seed))


# ---- CHAIN family: soft-Coulomb dimerised 1-D chain -------------------
CHAIN = dict(d_in=1.4, d_out=3.0, a=1.0, soft=1.0, Z=1.0)


def chain_geometry(n, d_in=1.4, d_out=3.0, **_):
    mu = np.arange(n)
    return (mu // 2) * (d_in + d_out) + (mu % 2) * d_in


def chain_matrices(n, **params):
    """(x, S, T, V, E_nn) of the chain model."""
    p = dict(CHAIN, **params)
    x = chain_geometry(n, **p)
    dx = x[:, None] - x[None, :]
    a, soft, z = p["a"], p["soft"], p["Z"]
    s = np.exp(-a * dx ** 2 / 2)
    t = s * (a / 2) * (1 - a * dx ** 2)
    r = (x[:, None] + x[None, :]) / 2
    v = np.zeros((n, n))
    for xa in x:
        v += -z * s / np.sqrt(soft + (r - xa) ** 2)
    enn = 0.0
    for i in range(n):
        for j in range(i):
            enn += z * z / np.sqrt(soft + (x[i] - x[j]) ** 2)
    return x, s, t, v, enn


def chain_eri(n, nu_lo=0, nu_hi=None, **params):
    """(ab|cd) = S_ab S_cd / sqrt(soft + (R_ab - R_cd)^2); |v| < 1e-200 -> 0."""
    p = dict(CHAIN, **params)
    nu_hi = n if nu_hi is None else nu_hi
    x, s, _, _, _ = chain_matrices(n, **params)
    r = 0.5 * (x[:, None] + x[None, :])
    d = r[:, :, None, None] - r[None, None, :, nu_lo:nu_hi]
    v = (s[:, :, None, None] * s[None, None, :, nu_lo:nu_hi]) / np.sqrt(p["soft"] + d * d)
    v[np.abs(v) < 1e-200] = 0.0
    return np.asfortranarray(v)


def chain_aux(n, **params):
    """The `aux` array scfb_eri_generate(CHAIN) expects: [soft, x[n], S[n*n] col-major]."""
    p = dict(CHAIN, **params)
    x, s, _, _, _ = chain_matrices(n, **params)
    return np.concatenate([[p["soft"]], x, np.asfortranarray(s).reshape(-1, order="F")])


def hash_eri_elements(a, b, c, d, seed=SEED_ERI):
    """Selected elements eri[a,b,c,d] of the HASH family (broadcasting index arrays) —
    O(1) each, so known answers exist at any n (used by the full-size tests)."""
    a, b, c, d = (np.asarray(v, dtype=np.uint64) for v in (a, b, c, d))
    p1, p2 = _pair(a, b), _pair(c, d)
    p, q = np.maximum(p1, p2), np.minimum(p1, p2)
    return _uniform_pm1(p * (p + np.uint64(1)) // np.uint64(2) + q, seed)


def random_symmetric_eri(n, rng):
    """Random tensor with the full 8-fold permutational symmetry (ij|kl) = (ji|kl) = (ij|lk)
    = (kl|ij) (test input for the packed layout)."""
    t = rng.standard_normal((n, n, n, n))
    t = t + t.transpose(1, 0, 2, 3)
    t = t + t.transpose(0, 1, 3, 2)
    t = t + t.transpose(2, 3, 0, 1)
    return np.asfortranarray(t / 8)


PACKED8_SEG_PAD = 8          # every k-segment is zero-padded to a multiple of 8 elements (64 bytes)


def packed8_seg_len(l_count):
    return -(-l_count // PACKED8_SEG_PAD) * PACKED8_SEG_PAD


def packed8_seg_off(k):
    """Padded elements of the k-segments 0..k-1 of a packed row."""
    return sum(packed8_seg_len(m + 1) for m in range(k))


def packed8_total(n):
    """Stored (padded) elements of the packed8 layout = B[n] of DESIGN.md."""
    return sum((i + 1) * packed8_seg_off(i) + packed8_seg_off(i + 1) for i in range(n))


def pack8(eri, i_lo=0, i_hi=None):
    """Oracle statement of the packed8 storage: for i >= j the row (i,j) holds the
    segments k = 0..i with l = 0..k (k < i) or l = 0..j (k == i), each zero-padded to a multiple
    of 8 elements; value (ij|kl) * 0.5^{[i=j]+[k=l]+[pr(i,j)=pr(k,l)]}."""
    n = eri.shape[0]
    i_hi = n if i_hi is None else i_hi
    out = []
    for i in range(i_lo, i_hi):
        for j in range(i + 1):
            for k in range(i + 1):
                lmax = k if k < i else j
                for l in range(lmax + 1):
                    v = eri[i, j, k, l] * 0.5 ** ((i == j) + (k == l) + (i == k and j == l))
                    out.append(v)
                out.extend([0.0] * (packed8_seg_len(lmax + 1) - (lmax + 1)))
    return np.array(out)
