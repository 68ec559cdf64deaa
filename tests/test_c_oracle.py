"""The C restatement of the oracle agrees with the NumPy one (both are test infrastructure)."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import scf_oracle as o
from oracle import synth


@pytest.mark.parametrize("n,n_spin", [(1, 1), (5, 2), (12, 1), (17, 2)])
def test_literal_long_double_matches_numpy_exact(n, n_spin):
    rng = np.random.default_rng(n)
    eri = rng.standard_normal((n, n, n, n))          # no symmetry
    dens = rng.standard_normal((n, n, n_spin))
    total = rng.standard_normal((n, n))
    j, k = c_oracle.jk_literal_ld(eri, dens, total)
    jr, kr = o.jk_exact(eri, dens, total)
    assert np.abs(j - jr).max() <= 2e-16 * np.abs(jr).max()
    assert np.abs(k - kr).max() <= 2e-16 * np.abs(kr).max()


def test_generators_are_bitwise_twins_of_numpy():
    n = 11
    assert np.array_equal(c_oracle.hash_eri(n, synth.SEED_ERI), synth.hash_eri(n))
    assert np.array_equal(c_oracle.hash_eri(n, 7, 3, 8), synth.hash_eri(n, 7, 3, 8))
    assert np.array_equal(c_oracle.chain_eri(n, synth.chain_aux(n)), synth.chain_eri(n))


def test_onepass_kernel_matches_literal():
    n = 24
    eri = synth.hash_eri(n)
    da, db = synth.hash_symmetric_matrix(n, 2), synth.hash_symmetric_matrix(n, 3)
    dens = np.stack([da, db], axis=2)
    j, k = c_oracle.jk_onepass(eri, dens, da + db, 0, n)
    jr, kr = c_oracle.jk_literal_ld(eri, dens, da + db)
    assert np.abs(j - jr).max() < 1e-13 * np.abs(jr).max()
    assert np.abs(k - kr).max() < 1e-13 * np.abs(kr).max()
