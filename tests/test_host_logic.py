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


# ---- result dump (reference src/util/dump_molsturm_hdf5.jl) ----------------------------------
def test_h5_writer_round_trips_through_the_reader(tmp_path):
    from scf_b200.minih5 import MiniH5, write_h5, _dtype_message
    rng = np.random.default_rng(3)
    data = {"n_alpha": np.int64(2), "restricted": np.uint8(1), "energy": -14.4868202421763,
            "overlap_bb": rng.standard_normal((6, 6)), "orben_f": rng.standard_normal(12),
            "ints": np.arange(5), "f32": np.float32(2.5), "fortran": np.asfortranarray(rng.standard_normal((3, 4)))}
    path = str(tmp_path / "t.h5")
    write_h5(path, data, {"format_version": "0.1.0", "format_type": "hdf5", "n": np.int64(7)})
    h = MiniH5(path)
    assert h.keys() == sorted(data)
    assert h.attrs() == {"format_version": "0.1.0", "format_type": "hdf5", "n": 7}
    for k, v in data.items():
        r = h.read(k)
        assert r.shape == np.shape(v) and r.dtype == np.asarray(v).dtype and np.array_equal(r, v)
    assert np.array_equal(h.read_rows("overlap_bb", 2, 5), data["overlap_bb"][2:5])
    many = {f"k{i:03d}": np.full(2, i) for i in range(70)}     # more names than a default symbol node holds
    write_h5(path, many)
    h = MiniH5(path)
    assert h.keys() == sorted(many) and all(h.read(k)[0] == int(k[1:]) for k in many)
    write_h5(path, {})
    assert MiniH5(path).keys() == []
    with pytest.raises(Exception):
        write_h5(path, {"a/b": np.zeros(2)})


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not present")
def test_h5_writer_encodes_datatypes_like_libhdf5():
    """The datatype messages of the writer are byte-identical to the ones libhdf5 wrote into the
    reference's integral dumps (f8, u8, i8) — the one cross-check against a real HDF5 file that is
    possible without an HDF5 library."""
    from scf_b200.minih5 import MiniH5, _dtype_message
    h = MiniH5(os.path.join(REF_DATA, "integrals_be_3-21g.hdf5"))
    for path, dt in (("kinetic_bb", "<f8"), ("system/nelec", "<u8"), ("system/atom_numbers", "<i8")):
        ref = [bytes(p) for t, _, p in h._messages(h._lookup(path)["ohdr"] + h.base) if t == 0x0003][0]
        mine = _dtype_message(np.dtype(dt))
        assert ref[:len(mine)] == mine and not any(ref[len(mine):])


@pytest.mark.parametrize("restricted", [True, False])
def test_dump_molsturm_hdf5_contents(golden_dir, tmp_path, restricted):
    """dump_molsturm_hdf5.jl:83-143 on a converged Be/3-21G result (the oracle's SCF stands in for
    run_scf here; the GPU path produces the same dict)."""
    from scf_b200.dump import dump_molsturm_hdf5, build_ab_matrix, concat_spin
    from scf_b200.minih5 import MiniH5
    system, integrals = o.load_integral_npz(os.path.join(golden_dir, "integrals_be_3-21g.npz"))
    problem = o.assemble_hf_problem(system, integrals, restricted=restricted)
    guess = o.compute_guess_hcore(problem)
    res = o.run_scf(problem, guess)
    assert res["converged"]
    da, db = res["density"][:, :, 0], res["density"][:, :, -1]
    ej, ek = o.termwise_jk_energies(integrals["electron_repulsion_bbbb"], np.stack([da, db], axis=2))
    res["energies"].update(coulomb=ej, exchange=ek, nuclear_repulsion=0.0,
                           kinetic=float(np.sum((da + db) * integrals["kinetic_bb"])),
                           nuclear_attraction=float(np.sum((da + db) * integrals["nuclear_attraction_bb"])))
    path = str(tmp_path / "state.hdf5")
    dump_molsturm_hdf5(integrals, res, path)
    h = MiniH5(path)
    assert h.attrs() == {"format_version": "0.1.0", "format_type": "hdf5"}
    expected = {"n_alpha", "n_beta", "n_bas", "restricted", "overlap_bb", "n_orbs_alpha", "n_orbs_beta",
                "orben_f", "fock_bb", "fock_ff", "hcore_ff", "orbcoeff_bf", "n_mtx_applies", "n_iter",
                "energy_coulomb", "energy_exchange", "energy_kinetic", "energy_nuclear_attraction",
                "energy_nuclear_repulsion", "energy_ground_state", "final_1e_energy_change",
                "final_error_norm", "final_tot_energy_change"}
    assert set(h.keys()) == expected
    n, n_orb = 9, problem.n_orb
    assert h.read("n_alpha") == 2 and h.read("n_beta") == 2 and h.read("n_bas") == n
    assert h.read("restricted") == int(restricted) and h.read("restricted").dtype == np.uint8
    assert h.read("n_iter") == res["n_iter"] and np.isnan(h.read("n_mtx_applies"))
    assert abs(h.read("energy_ground_state") - (-14.4868202421763)) < 1e-8
    # the file holds Julia's column-major bytes under reversed dims: a C-order reader sees the transpose
    s2 = h.read("overlap_bb")
    assert s2.shape == (2 * n, 2 * n) and np.array_equal(s2[:n, :n], problem.overlap.T) and not s2[:n, n:].any()
    assert np.array_equal(h.read("orben_f"), concat_spin(res["orben"]))
    cbf = h.read("orbcoeff_bf")                       # Julia (2 n_orb, n_bas) -> file (n_bas, 2 n_orb)
    assert cbf.shape == (n, 2 * n_orb) and np.array_equal(cbf[:, :n_orb], res["orbcoeff"][:, :, 0])
    assert np.array_equal(cbf[:, n_orb:], res["orbcoeff"][:, :, -1])
    # in the orbital basis a converged Fock matrix is diagonal with the orbital energies on it — to
    # the SCF threshold only: the returned orbitals diagonalise the previous (extrapolated) Fock
    # matrix, not the returned one (run_scf.jl:40 vs :48)
    fff = h.read("fock_ff")
    assert fff.shape == (2 * n_orb, 2 * n_orb)
    assert np.abs(np.diag(fff) - concat_spin(res["orben"])).max() < 1e-4
    assert np.abs(fff - np.diag(np.diag(fff))).max() < 1e-4
    assert np.array_equal(h.read("fock_bb"), build_ab_matrix(res["fock"]).T)
    e = res["energies"]
    assert abs(e["kinetic"] + e["nuclear_attraction"] + e["coulomb"] + e["exchange"] - e["total"]) < 1e-7
    with pytest.raises(NotImplementedError):
        dump_molsturm_hdf5(integrals, res, path, export_eri_bbbb=True)
