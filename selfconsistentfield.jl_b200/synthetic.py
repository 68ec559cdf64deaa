"""Synthetic problems for configs 2-5 of BASELINE.json (host-side O(n^2) setup only; the
O(n^4) tensor is generated on the device by `scfb_eri_generate`).

Two closed-form families, frozen in DESIGN.md ("synthetic inputs"):
  hash   dense, 8-fold symmetric, every element uniform(-1,1) from splitmix64 of the
         canonical orbit index — for build parity and throughput;
  chain  soft-Coulomb dimerised 1-D chain — physical, the SCF converges (RHF closed shell).
"""
from __future__ import annotations

import numpy as np

SEED_ERI, SEED_DA, SEED_DB = 20260101, 20260102, 20260103
CHAIN_PARAMS = dict(d_in=1.4, d_out=3.0, a=1.0, soft=1.0, Z=1.0)

_M64 = (1 << 64) - 1


def _splitmix_uniform(idx: np.ndarray, seed: int) -> np.ndarray:
    z = idx.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(((seed + 1) * 0x9E3779B97F4A7C15) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0


def hash_density(n: int, seed: int) -> np.ndarray:
    """Symmetric (n,n) test density, uniform(-1,1), column-major."""
    i = np.arange(n, dtype=np.uint64)
    hi = np.maximum(i[:, None], i[None, :])
    lo = np.minimum(i[:, None], i[None, :])
    return np.asfortranarray(_splitmix_uniform(hi * (hi + np.uint64(1)) // np.uint64(2) + lo, seed))


def packed8_total(n: int) -> int:
    """Stored (padded) elements of the packed8 layout (csrc/jk_packed.cu: B[n])."""
    seg = lambda k: 8 * (k // 8 + 1) * (4 * (k // 8) + k % 8)       # segments padded to 8 elements
    return sum((i + 1) * seg(i) + seg(i + 1) for i in range(n))


def unit_symmetric(n: int, a: int, b: int) -> np.ndarray:
    """E_ab + E_ba (or E_aa): the symmetric density that picks single tensor elements."""
    m = np.zeros((n, n), order="F")
    m[a, b] = 1.0
    m[b, a] = 1.0
    return m


def hash_eri_elements(a, b, c, d, seed: int = SEED_ERI) -> np.ndarray:
    """eri[a,b,c,d] of the HASH family for broadcastable index arrays (closed form, O(1))."""
    a, b, c, d = (np.asarray(v, dtype=np.uint64) for v in (a, b, c, d))
    one, two = np.uint64(1), np.uint64(2)

    def pair(i, j):
        hi, lo = np.maximum(i, j), np.minimum(i, j)
        return hi * (hi + one) // two + lo

    p1, p2 = pair(a, b), pair(c, d)
    p, q = np.maximum(p1, p2), np.minimum(p1, p2)
    return _splitmix_uniform(p * (p + one) // two + q, seed)


class ChainModel:
    """x, S, T, V, E_nn of the chain family plus the aux array for the device generator."""

    def __init__(self, n, **params):
        p = dict(CHAIN_PARAMS, **params)
        self.n, self.params = n, p
        mu = np.arange(n)
        self.x = (mu // 2) * (p["d_in"] + p["d_out"]) + (mu % 2) * p["d_in"]
        dx = self.x[:, None] - self.x[None, :]
        a, soft, z = p["a"], p["soft"], p["Z"]
        self.overlap = np.exp(-a * dx ** 2 / 2)
        self.kinetic = self.overlap * (a / 2) * (1 - a * dx ** 2)
        r = (self.x[:, None] + self.x[None, :]) / 2
        v = np.zeros((n, n))
        for xa in self.x:
            v += -z * self.overlap / np.sqrt(soft + (r - xa) ** 2)
        self.nuclear_attraction = v
        d = np.abs(self.x[:, None] - self.x[None, :])
        iu = np.triu_indices(n, 1)
        self.nuclear_repulsion = float(np.sum(z * z / np.sqrt(soft + d[iu] ** 2)))

    @property
    def aux(self):
        return np.concatenate([[self.params["soft"]], self.x,
                               np.asfortranarray(self.overlap).reshape(-1, order="F")])

    def n_elec(self, unrestricted=False):
        n = self.n
        return (n // 2 + 1, n // 2 - 1) if unrestricted else (n // 2, n // 2)
