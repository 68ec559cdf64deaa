"""Pins the CPU oracle against the reference's own golden vectors
(/root/reference/test/functionality_hartee_fock.jl:16-38: four converged total energies,
rtol 1e-9, plus `converged`) on the reference's bundled integral dumps
(tests/golden/*.npz, converted by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import scf_oracle as o
from oracle import synth


def _run(golden_dir, case):
    system, integrals = o.load_integral_npz(os.path.join(golden_dir, case["file"] + ".npz"))
    problem = o.assemble_hf_problem(system, integrals, restricted=case["restricted"])
    guess = o.compute_guess_hcore(problem)
    if (not problem.restricted) and problem.n_elec[0] == problem.n_elec[1]:
        o.break_spin_symmetry(guess)
    return o.run_scf(problem, guess)


@pytest.mark.parametrize("name", ["rhf_be_321g", "uhf_be_321g", "rohf_c_321g", "uhf_c_321g"])
def test_reference_golden_energies(golden_dir, name):
    case = json.load(open(os.path.join(golden_dir, "golden_energies.json")))[name]
    res = _run(golden_dir, case)
    assert res["converged"]
    # the reference's own assertion: `≈` with rtol=1e-9 on the (stale-iterate) energy
    assert res["energies"]["total"] == pytest.approx(case["e_total"], rel=1e-9, abs=0)
    # and the newest Fock build agrees much tighter
    assert abs(res["energies_newest"]["total"] - case["e_total"]) < 5e-11


def test_fixture_symmetry_and_shapes(golden_dir):
    for stem in ("integrals_be_3-21g", "integrals_c_3-21g"):
        _, integrals = o.load_integral_npz(os.path.join(golden_dir, stem + ".npz"))
        eri = integrals["electron_repulsion_bbbb"]
        assert eri.shape == (9, 9, 9, 9)
        assert np.abs(eri - eri.transpose(1, 0, 2, 3)).max() < 1e-15
        assert np.abs(eri - eri.transpose(2, 3, 0, 1)).max() < 1e-15
        assert np.allclose(np.diag(integrals["overlap_bb"]), 1.0)


def test_literal_vs_exact_vs_blas_forms():
    rng = np.random.default_rng(1)
    n = 7
    eri = rng.standard_normal((n, n, n, n))        # no symmetry at all
    d = rng.standard_normal((n, n, 2))             # non-symmetric densities
    dtot = rng.standard_normal((n, n))             # caller-supplied, unrelated to d
    a = o.add_2e_term(np.zeros((n, n, 2)), eri, d, dtot)
    b = o.add_2e_term_exact(np.zeros((n, n, 2)), eri, d, dtot)
    c = o.add_2e_term_blas(np.zeros((n, n, 2)), np.asfortranarray(eri), d, dtot)
    assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()
    assert np.abs(c - b).max() < 1e-12 * np.abs(b).max()
    # literal loops (JKBuilderFromTensor.jl:37): J[m,v]=eri[a,b,m,v]Dtot[a,b]; K[m,v]=eri[m,b,a,v]D[a,b]
    j = np.zeros((n, n))
    k = np.zeros((n, n))
    for m in range(n):
        for v in range(n):
            j[m, v] = sum(eri[a_, b_, m, v] * dtot[a_, b_] for a_ in range(n) for b_ in range(n))
            k[m, v] = sum(eri[m, b_, a_, v] * d[a_, b_, 0] for a_ in range(n) for b_ in range(n))
    assert np.abs((j - k) - a[:, :, 0]).max() < 1e-12 * np.abs(a).max()


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
def test_shard_sum_equals_unsharded(n_ranks):
    rng = np.random.default_rng(2)
    n = 10  # ragged for 3, 4, 8
    eri = rng.standard_normal((n, n, n, n))
    d = rng.standard_normal((n, n, 1))
    dtot = 2 * d[:, :, 0]
    full = o.add_2e_term(np.zeros((n, n, 1)), eri, d, dtot)
    shard = o.add_2e_term_sharded(np.zeros((n, n, 1)), eri, d, dtot, n_ranks)
    assert np.abs(full - shard).max() < 1e-13 * np.abs(full).max()
    ranges = [o.shard_range(n, g, n_ranks) for g in range(n_ranks)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(n_ranks - 1))


def test_chain_family_regression_energies():
    """SURVEY.md Appendix F anchors (derived, not from the reference)."""
    for n, e_ref in ((8, -5.3238691536), (16, -10.6493090426)):
        x, s, t, v, enn = synth.chain_matrices(n)
        problem = o.ScfProblem(s, (n // 2, n // 2), n, True, {"nuclear_repulsion": enn},
                               {"kinetic": t, "nuclear_attraction": v},
                               {"coulomb+exchange": o.JKBuilderFromTensor(synth.chain_eri(n))})
        res = o.run_scf(problem, o.compute_guess_hcore(problem), max_error_norm=1e-10, max_iter=200)
        assert res["converged"]
        assert abs(res["energies_newest"]["total"] - e_ref) < 5e-10


def test_hash_family_is_8fold_symmetric_and_sliceable():
    n = 6
    h = synth.hash_eri(n)
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)):
        assert np.array_equal(h, h.transpose(perm))
    assert np.array_equal(h[:, :, :, 2:5], synth.hash_eri(n, nu_lo=2, nu_hi=5))
    assert -1 <= h.min() < -0.9 and 0.9 < h.max() < 1
