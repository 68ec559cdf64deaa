"""ONE handle over several GPUs of one process (scfb_create_multi; SURVEY.md 8b/8e: the caller of
/root/reference/src/run_scf.jl:4-65 stays single-threaded).  Skipped below 2 visible GPUs; the
world-8 run of tests/multi_handle_check.py covers the BASELINE sizes."""
import numpy as np
import pytest

import scf_b200
from oracle import scf_oracle as o
from oracle import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def n_gpus():
    import torch
    return torch.cuda.device_count()


def rel_err(x, ref):
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.fixture(scope="module")
def devices():
    g = n_gpus()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    return list(range(min(g, 8)))


@pytest.mark.parametrize("n,n_spin,layout", [(10, 1, "full"), (33, 2, "full"), (64, 2, "full"), (7, 2, "full"),
                                             (40, 2, "packed8"), (9, 1, "packed8")])
def test_multi_handle_matches_oracle(devices, n, n_spin, layout):
    rng = np.random.default_rng(n)
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout=layout, devices=devices)
    assert b.shard_range == (0, n)
    eri = synth.hash_eri(n)
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    dens = np.stack([da, db][:n_spin], axis=2)
    total = 2 * da if n_spin == 1 else da + db
    for _ in range(4):                      # repeated builds: epochs / parity double buffering
        j, k = b.jk(dens, total)
        jr, kr = o.jk_exact(eri, dens, total)
        assert rel_err(j, jr) < RTOL and rel_err(k, kr) < RTOL
        out = np.asfortranarray(rng.standard_normal((n, n, n_spin)))
        ref = o.add_2e_term_exact(out.copy(), eri, dens, total)
        b.add_2e_term(out, dens, total)
        assert rel_err(out, ref) < RTOL
    b.close()


def test_multi_handle_from_tensor_and_download(devices):
    n = 12
    rng = np.random.default_rng(3)
    eri = np.asfortranarray(rng.standard_normal((n, n, n, n)))
    dens = rng.standard_normal((n, n, 2))
    total = rng.standard_normal((n, n))
    b = scf_b200.JKBuilderCUDA.from_tensor(eri, devices=devices)
    assert np.array_equal(b.eri_shard(), eri)                 # members' shards back to back
    j, k = b.jk(dens, total)
    jr, kr = o.jk_exact(eri, dens, total)
    assert rel_err(j, jr) < RTOL and rel_err(k, kr) < RTOL
    b.close()


def test_multi_handle_device_resident_scf(devices):
    """The device-resident SCF step on member 0 drives every GPU's shard: chain RHF energy equals the
    single-GPU run (SURVEY.md Appendix F anchor)."""
    n = 32
    model = scf_b200.synthetic.ChainModel(n)
    builder = scf_b200.JKBuilderCUDA.synthetic(n, "chain", 0, aux=model.aux, devices=devices)
    problem = scf_b200.ScfProblem(model.overlap, model.n_elec(), n, True,
                                  {"nuclear_repulsion": model.nuclear_repulsion},
                                  {"kinetic": model.kinetic, "nuclear_attraction": model.nuclear_attraction},
                                  {"coulomb+exchange": builder})
    res = scf_b200.run_scf(problem, scf_b200.compute_guess(problem), verbose=False,
                           max_error_norm=1e-9, max_energy_total_change=1e-11, max_iter=200)
    assert res["converged"] and res["mode"] == "device"
    assert abs(res["energies_newest"]["total"] - (-21.3001879351)) < 1e-9
    builder.close()
