"""CPU tests of the product's host logic (no GPU, no library calls): the DIIS bookkeeping of
scf.py against the oracle's literal restatement of cDIIS.jl / EDIIS.jl / parts/diis.jl, the
convergence rule, the HDF5 reader against the committed fixtures, and the synthetic-problem
setup against the oracle twins."""
import os

import numpy as np
import pytest

import scf_b200
from oracle import scf_oracle as o
from oracle import synth
from scf_b200 import scf as S

REF_DATA = "/root/reference/test/data"


def _history(rng, n, n_spin, steps):
    out = []
    for _ in range(steps):
        f = rng.standard_normal((n, n, n_spin)); f = f + f.transpose(1, 0, 2)
        d = rng.standard_normal((n, n, n_spin)); d = d + d.transpose(1, 0, 2)
        e = rng.standard_normal((n, n, n_spin)) * 1e-2
        out.append((f, e, d, {"total": float(rng.standard_normal()), "1e": 0.0, "2e": 0.0, "0e": 0.0}))
    return out


@pytest.mark.parametrize("n_spin,sync", [(1, True), (2, True), (2, False)])
def test_cdiis_extrapolation_matches_oracle(n_spin, sync):
    rng = np.random.default_rng(10 + n_spin)
    n = 6
    acc_p = S.cDIIS(n_diis_size=4, sync_spins=sync)
    acc_o = o.cDIIS(n_diis_size=4, sync_spins=sync)
    for f, e, d, en in _history(rng, n, n_spin, 9):       # longer than the history: exercises the ring
        ip = S.compute_next_iterate(acc_p, S.FockIterState(f, e, en, None, None, d))
        io = o.compute_next_iterate(acc_o, o.FockIterState(f, e, en, None, None, d))
        assert np.abs(ip.fock - io.fock).max() < 1e-10 * np.abs(io.fock).max()
        assert ip.energies is en and ip.error_pulay is e    # update_iterate_matrix keeps the rest


def test_cdiis_purges_like_the_oracle_on_linear_dependence():
    """Identical error vectors make the bordered matrix singular: eigenvalues are masked and
    count+1 oldest entries purged (cDIIS.jl:38-62, DiisState.jl:40-50)."""
    rng = np.random.default_rng(3)
    n = 5
    f0 = rng.standard_normal((n, n, 1)); e0 = rng.standard_normal((n, n, 1))
    acc_p, acc_o = S.cDIIS(), o.cDIIS()
    for step in range(6):
        f = f0 + 1e-3 * step
        en = {"total": -1.0, "1e": 0.0, "2e": 0.0, "0e": 0.0}
        ip = S.compute_next_iterate(acc_p, S.FockIterState(f, e0, en, None, None, f))
        io = o.compute_next_iterate(acc_o, o.FockIterState(f, e0, en, None, None, f))
        assert len(acc_p.host[0]) == len(acc_o.state[0].iterate)
        assert acc_p.b[0].shape[0] == len(acc_o.state[0].iterate)
        assert np.abs(ip.fock - io.fock).max() < 1e-8 * np.abs(io.fock).max()


def test_ediis_matrix_and_fixed_point_match_oracle():
    rng = np.random.default_rng(5)
    n = 5
    acc_p, acc_o = S.EDIIS(n_diis_size=3), o.EDIIS(n_diis_size=3)
    for f, e, d, en in _history(rng, n, 2, 5):
        S.compute_next_iterate(acc_p, S.FockIterState(f, e, en, None, None, d))
        o.compute_next_iterate(acc_o, o.FockIterState(f, e, en, None, None, d))
    # same B matrix (tr((Fi-Fj)(Di-Dj)), EDIIS.jl:87, spin blocks added, :126-128) and energies
    m = acc_p.b[0].shape[0]
    assert m == 3
    bo = np.zeros((m, m))
    for sp in range(2):
        st = acc_o.state[sp]
        for i in range(m):
            for j in range(m):
                bo[i, j] += np.trace((st.iterate[i] - st.iterate[j]) @ (st.density[i] - st.density[j]))
    assert np.abs(acc_p.merge(acc_p.system(0), acc_p.system(1)) - bo).max() < 1e-10
    assert acc_p.energies == [en["total"] for en in acc_o.state[0].energies.items]
    # first EDIIS step is trivial: c = [1] (EDIIS.jl:37-41)
    first = S.EDIIS()
    first.push_energy({"total": -3.0})
    c, purge = first.solve(np.zeros((1, 1)))
    assert c[0] == 1.0 and purge == 0


def test_convergence_rule_is_signed():
    """check_convergence.jl:27-35 compares signed changes, not magnitudes."""
    old = {"total": -1.0, "1e": -2.0, "2e": 1.0, "0e": 0.0}
    better = {"total": -1.5, "1e": -2.5, "2e": 1.0, "0e": 0.0}     # large DEcrease passes
    worse = {"total": -1.0 + 2e-7, "1e": -2.0, "2e": 1.0, "0e": 0.0}
    assert S._convergence(1e-8, old, better).is_converged
    assert not S._convergence(1e-8, old, worse).is_converged
    assert not S._convergence(1e-6, old, better).is_converged        # error norm threshold 5e-7
    ref = o.check_convergence(o.FockIterState(None, None, old, None, None, None),
                              o.FockIterState(None, np.full((2, 2, 1), 5e-9), better, None, None, None))
    assert ref.is_converged


def test_fold_small_coefficients_mutates_like_the_reference():
    c = np.array([0.6, 0.4 - 3e-7, 1e-7, 2e-7])
    used = S._fold_small_coefficients(c, 1e-6)
    assert used[2] == 0 and used[3] == 0 and abs(used.sum() - 1.0) < 1e-15
    assert abs(c[0] - (0.6 + 3e-7)) < 1e-15                           # argmax entry mutated in place
    c2 = np.array([1e-9, 1.0])                                        # index 1 (newest) always kept
    used2 = S._fold_small_coefficients(c2, 1e-6)
    assert used2[0] == 1e-9


def test_chain_model_matches_oracle_twin():
    for n in (8, 19):
        m = scf_b200.synthetic.ChainModel(n)
        x, s, t, v, enn = synth.chain_matrices(n)
        assert np.array_equal(m.x, x) and np.array_equal(m.overlap, s) and np.array_equal(m.kinetic, t)
        assert np.abs(m.nuclear_attraction - v).max() < 1e-13 and abs(m.nuclear_repulsion - enn) < 1e-12
        assert np.array_equal(m.aux, synth.chain_aux(n))
    assert scf_b200.synthetic.packed8_total(23) == synth.packed8_total(23)
    i = np.arange(9)
    assert np.array_equal(scf_b200.synthetic.hash_eri_elements(i[:, None], 2, 7, i[None, :]),
                          synth.hash_eri(9)[:, 2, 7, :])


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not present")
@pytest.mark.parametrize("stem", ["integrals_be_3-21g", "integrals_c_3-21g"])
def test_minih5_reads_the_reference_dumps_like_the_committed_fixtures(stem, golden_dir):
    from scf_b200.minih5 import MiniH5
    h = MiniH5(os.path.join(REF_DATA, stem + ".hdf5"))
    z = np.load(os.path.join(golden_dir, stem + ".npz"))
    for key in ("electron_repulsion_bbbb", "kinetic_bb", "nuclear_attraction_bb", "overlap_bb"):
        assert np.array_equal(h.read(key), z[key])
    assert list(h.read("system/nelec")) == list(z["nelec"])
    assert h.read("discretisation/basis_type") == "gaussian"
    assert sorted(h.keys("/")) == ["discretisation", "electron_repulsion_bbbb", "kinetic_bb",
                                   "nuclear_attraction_bb", "overlap_bb", "system"]


def test_minih5_rejects_garbage(tmp_path):
    from scf_b200.minih5 import H5FormatError, MiniH5
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(H5FormatError):
        MiniH5(str(p))


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not present")
def test_minih5_row_ranges_equal_the_full_read():
    """Slab streaming: rows [lo, hi) of the chunked ERI dataset (chunks of 5 along that axis)."""
    from scf_b200.minih5 import MiniH5
    h = MiniH5(os.path.join(REF_DATA, "integrals_c_3-21g.hdf5"))
    assert h.shape("electron_repulsion_bbbb") == (9, 9, 9, 9) and h.shape("system/nelec") == (2,)
    full = h.read("electron_repulsion_bbbb")
    for lo, hi in ((0, 9), (0, 1), (3, 7), (4, 6), (8, 9), (5, 5)):
        assert np.array_equal(h.read_rows("electron_repulsion_bbbb", lo, hi), full[lo:hi])
    assert np.array_equal(h.read_rows("system/coords", 0, 1), h.read("system/coords"))
    with pytest.raises(ValueError):
        h.read_rows("electron_repulsion_bbbb", 3, 12)
