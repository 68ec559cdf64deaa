"""GPU parity of the 8-fold-packed J/K path (K2) through the C ABI.  Precondition of the
packed layout: 8-fold symmetric tensor, symmetric densities (include/scf_b200.h)."""
import numpy as np
import pytest

import scf_b200
from oracle import scf_oracle as o
from oracle import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_err(x, ref):
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


def sym(rng, n):
    a = rng.standard_normal((n, n))
    return a + a.T


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 16, 17, 24, 31, 33, 40])
@pytest.mark.parametrize("n_spin", [1, 2])
def test_packed_matches_oracle_random_symmetric(n, n_spin):
    rng = np.random.default_rng(77 * n + n_spin)
    eri = synth.random_symmetric_eri(n, rng)
    density = np.stack([sym(rng, n) for _ in range(n_spin)], axis=2)
    total = sym(rng, n)                       # caller-supplied, symmetric
    b = scf_b200.JKBuilderCUDA.from_tensor(eri, layout="packed8")
    assert b.eri_bytes == 8 * synth.packed8_total(n)
    j, k = b.jk(density, total)
    jr, kr = o.jk_exact(eri, density, total)
    assert rel_err(j, jr) < RTOL
    assert rel_err(k, kr) < RTOL
    out = np.asfortranarray(rng.standard_normal((n, n, n_spin)))
    ref = o.add_2e_term_exact(out.copy(), eri, density, total)
    b.add_2e_term(out, density, total)
    assert rel_err(out, ref) < RTOL
    b.close()


def test_packed_storage_matches_oracle_pack8():
    rng = np.random.default_rng(3)
    n = 7
    eri = synth.random_symmetric_eri(n, rng)
    b = scf_b200.JKBuilderCUDA.from_tensor(eri, layout="packed8")
    assert np.array_equal(b.eri_shard_raw(), synth.pack8(eri))
    b.close()
    g = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout="packed8")
    assert np.array_equal(g.eri_shard_raw(), synth.pack8(synth.hash_eri(n)))
    g.close()
    model = scf_b200.synthetic.ChainModel(8)
    c = scf_b200.JKBuilderCUDA.synthetic(8, "chain", 0, aux=model.aux, layout="packed8")
    assert np.array_equal(c.eri_shard_raw(), synth.pack8(synth.chain_eri(8)))
    c.close()


@pytest.mark.parametrize("n,n_spin", [(64, 1), (96, 2), (130, 1)])
def test_packed_equals_full_layout_on_hash_family(n, n_spin):
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    density = np.stack([da, db][:n_spin], axis=2)
    total = 2 * da if n_spin == 1 else da + db
    full = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI)
    jf, kf = full.jk(density, total)
    full.close()
    packed = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout="packed8")
    jp, kp = packed.jk(density, total)
    packed.close()
    assert rel_err(jp, jf) < RTOL and rel_err(kp, kf) < RTOL


@pytest.mark.parametrize("n,n_ranks", [(12, 2), (20, 4), (9, 8)])
def test_packed_shard_partials_sum_to_full(n, n_ranks):
    rng = np.random.default_rng(n + n_ranks)
    eri = synth.random_symmetric_eri(n, rng)
    density = np.stack([sym(rng, n), sym(rng, n)], axis=2)
    total = density[:, :, 0] + density[:, :, 1]
    jr, kr = o.jk_exact(eri, density, total)
    js, ks = np.zeros_like(jr), np.zeros_like(kr)
    bounds = []
    nbytes = 0
    for g in range(n_ranks):
        part = scf_b200.JKBuilderCUDA.from_tensor(eri, layout="packed8", rank=g, n_ranks=n_ranks, allow_partial=True)
        bounds.append(part.shard_range)
        nbytes += part.eri_bytes
        j, k = part.jk(density, total)
        js += j
        ks += k
        part.close()
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    assert all(bounds[g][1] == bounds[g + 1][0] for g in range(n_ranks - 1))
    assert nbytes == 8 * synth.packed8_total(n)
    assert rel_err(js, jr) < RTOL and rel_err(ks, kr) < RTOL


def test_packed_full_scf_chain_matches_full_layout():
    n = 32
    model = scf_b200.synthetic.ChainModel(n)
    energies = []
    for layout in ("full", "packed8"):
        builder = scf_b200.JKBuilderCUDA.synthetic(n, "chain", 0, aux=model.aux, layout=layout)
        problem = scf_b200.ScfProblem(model.overlap, model.n_elec(), n, True,
                                      {"nuclear_repulsion": model.nuclear_repulsion},
                                      {"kinetic": model.kinetic,
                                       "nuclear_attraction": model.nuclear_attraction},
                                      {"coulomb+exchange": builder})
        res = scf_b200.run_scf(problem, scf_b200.compute_guess(problem), verbose=False,
                               max_error_norm=1e-9, max_energy_total_change=1e-11, max_iter=200)
        assert res["converged"]
        energies.append(res["energies_newest"]["total"])
        builder.close()
    assert abs(energies[0] - energies[1]) < 1e-10
    assert abs(energies[0] - (-21.3001879351)) < 1e-9       # SURVEY.md Appendix F anchor
