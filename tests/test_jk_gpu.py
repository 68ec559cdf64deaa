"""GPU parity tests proper: the CUDA J/K path, called through the C ABI
(`JKBuilderCUDA` = ctypes over libscf_b200.so), against the CPU oracle on the same inputs.

Tolerance (north star): J and K within 1e-12 relative of the reference contraction,
metric max|Δ| / max|J or K| against the extended-precision oracle."""
import json
import os

import numpy as np
import pytest

import scf_b200
from oracle import scf_oracle as o
from oracle import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_err(x, ref):
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


def check_jk(builder, eri, density, total, exact=True):
    j, k = builder.jk(density, total)
    if exact:
        jr, kr = o.jk_exact(eri, density, total)
    else:
        g = o.add_2e_term_blas  # noqa: F841  (large n: fp64 BLAS evaluation)
        jr = o.coulomb(eri, total)
        kr = np.stack([o.exchange(eri, density[:, :, s]) for s in range(density.shape[2])], axis=2)
    assert rel_err(j, jr) < RTOL
    for s in range(density.shape[2]):
        assert rel_err(k[:, :, s], kr[:, :, s]) < RTOL
    return j, k


# ---- the reference's own fixtures ------------------------------------------------------
@pytest.mark.parametrize("name", ["rhf_be_321g", "uhf_be_321g", "rohf_c_321g", "uhf_c_321g"])
def test_golden_fixture_first_fock_build(golden_dir, name):
    case = json.load(open(os.path.join(golden_dir, "golden_energies.json")))[name]
    system, integrals = o.load_integral_npz(os.path.join(golden_dir, case["file"] + ".npz"))
    problem = o.assemble_hf_problem(system, integrals, restricted=case["restricted"])
    density = o.compute_guess_hcore(problem)
    if (not problem.restricted) and problem.n_elec[0] == problem.n_elec[1]:
        o.break_spin_symmetry(density)     # makes the beta density non-trivially different
    n_spin = density.shape[2]
    total = 2 * density[:, :, 0] if n_spin == 1 else density[:, :, 0] + density[:, :, 1]
    eri = integrals["electron_repulsion_bbbb"]
    b = scf_b200.JKBuilderCUDA.from_tensor(eri)
    check_jk(b, eri, density, total)
    # add_2e_term! accumulates into a pre-filled `out` (JKBuilderFromTensor.jl:39,48-49)
    rng = np.random.default_rng(0)
    base = rng.standard_normal((9, 9, n_spin))
    out = np.asfortranarray(base.copy())
    ret = b.add_2e_term(out, density, total)
    assert ret is out
    ref = o.add_2e_term_exact(base.copy(), eri, density, total)
    assert rel_err(out, ref) < RTOL
    # C-ordered `out` takes the copy-in/copy-out path and must give the same answer
    out_c = np.ascontiguousarray(base.copy())
    b.add_2e_term(out_c, density, total)
    assert np.array_equal(out_c, out)
    b.close()


def test_golden_energy_through_cuda_builder(golden_dir):
    """Full SCF with the oracle's loop but the CUDA builder plugged in at the
    add_2e_term! boundary -> the reference's golden RHF/UHF energies."""
    cases = json.load(open(os.path.join(golden_dir, "golden_energies.json")))
    for name in ("rhf_be_321g", "uhf_c_321g"):
        case = cases[name]
        system, integrals = o.load_integral_npz(os.path.join(golden_dir, case["file"] + ".npz"))
        integrals["jk_builder_bb"] = scf_b200.JKBuilderCUDA.from_tensor(
            integrals["electron_repulsion_bbbb"])
        problem = o.assemble_hf_problem(system, integrals, restricted=case["restricted"])
        guess = o.compute_guess_hcore(problem)
        res = o.run_scf(problem, guess)
        assert res["converged"]
        assert res["energies"]["total"] == pytest.approx(case["e_total"], rel=1e-9, abs=0)
        assert abs(res["energies_newest"]["total"] - case["e_total"]) < 1e-10


# ---- arbitrary tensors / densities: the literal index patterns ---------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 15, 16, 17, 24, 33, 40])
@pytest.mark.parametrize("n_spin", [1, 2])
def test_random_unsymmetric_inputs(n, n_spin):
    rng = np.random.default_rng(1000 * n + n_spin)
    eri = rng.standard_normal((n, n, n, n))          # no permutational symmetry
    density = rng.standard_normal((n, n, n_spin))    # not symmetric
    total = rng.standard_normal((n, n))              # caller-supplied, unrelated to density
    b = scf_b200.JKBuilderCUDA.from_tensor(eri)
    check_jk(b, eri, density, total)
    b.close()


def test_hypothesis_small_cases():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(n=st.integers(2, 24), n_spin=st.sampled_from([1, 2]), seed=st.integers(0, 2 ** 31))
    def run(n, n_spin, seed):
        rng = np.random.default_rng(seed)
        eri = rng.uniform(-1, 1, (n, n, n, n))
        density = rng.uniform(-1, 1, (n, n, n_spin))
        total = rng.uniform(-1, 1, (n, n))
        b = scf_b200.JKBuilderCUDA.from_tensor(eri)
        try:
            check_jk(b, eri, density, total)
        finally:
            b.close()

    run()


# ---- device generators == host twins, bit for bit ----------------------------------------
@pytest.mark.parametrize("n", [6, 17, 32])
def test_hash_generator_matches_host_twin(n):
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI)
    assert np.array_equal(b.eri_shard(), synth.hash_eri(n))
    b.close()


@pytest.mark.parametrize("n", [8, 18])
def test_chain_generator_matches_host_twin(n):
    model = scf_b200.synthetic.ChainModel(n)
    x, s, t, v, enn = synth.chain_matrices(n)
    assert np.array_equal(model.overlap, s) and np.array_equal(model.kinetic, t)
    assert np.allclose(model.nuclear_attraction, v, rtol=0, atol=1e-13)
    assert abs(model.nuclear_repulsion - enn) < 1e-12
    b = scf_b200.JKBuilderCUDA.synthetic(n, "chain", 0, aux=model.aux)
    dev, host = b.eri_shard(), synth.chain_eri(n)
    assert np.array_equal(dev, host)
    b.close()


# ---- medium sizes against the fp64 oracle ------------------------------------------------
@pytest.mark.parametrize("n,n_spin", [(64, 1), (64, 2), (96, 2), (100, 1), (127, 2)])
def test_hash_family_medium(n, n_spin):
    """Arbiter at these sizes: the C restatement with long-double accumulation."""
    from oracle import c_oracle
    eri = c_oracle.hash_eri(n, synth.SEED_ERI)
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    density = np.stack([da, db][:n_spin], axis=2)
    total = 2 * da if n_spin == 1 else da + db
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI)
    j, k = b.jk(density, total)
    jr, kr = c_oracle.jk_literal_ld(eri, density, total)
    assert rel_err(j, jr) < RTOL
    assert rel_err(k, kr) < RTOL
    # 8-fold symmetric tensor + symmetric D  =>  J and K symmetric
    assert rel_err(j, j.T) < RTOL and rel_err(k[:, :, 0], k[:, :, 0].T) < RTOL
    # and an unsymmetric density against the literal index patterns
    rng = np.random.default_rng(n)
    du = rng.standard_normal((n, n, n_spin))
    tu = rng.standard_normal((n, n))
    j, k = b.jk(du, tu)
    jr, kr = c_oracle.jk_literal_ld(eri, du, tu)
    assert rel_err(j, jr) < RTOL and rel_err(k, kr) < RTOL
    b.close()


@pytest.mark.parametrize("split", ["2", "4"])
def test_split_items_fold_to_the_same_sums(split, monkeypatch):
    """Every work item cut along b (the persistent kernel folds a group's pieces before the slab finisher):
    dense densities against the long-double oracle, and bit-equal to nothing but itself — the order of
    the partial sums differs from the unsplit plan, the 1e-12 gate does not."""
    from oracle import c_oracle
    n = 96
    monkeypatch.setenv("SCFB_FULL_BSPLIT", split)
    eri = c_oracle.hash_eri(n, synth.SEED_ERI)
    rng = np.random.default_rng(int(split))
    du, tu = rng.standard_normal((n, n, 2)), rng.standard_normal((n, n))
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI)
    j, k = b.jk(du, tu)
    jr, kr = c_oracle.jk_literal_ld(eri, du, tu)
    assert rel_err(j, jr) < RTOL and rel_err(k, kr) < RTOL
    out = np.asfortranarray(rng.standard_normal((n, n, 2)))
    ref = out + jr[:, :, None] - kr
    b.add_2e_term(out, du, tu)           # out += J - K through the host-pointer boundary
    assert rel_err(out, ref) < RTOL
    b.close()


# ---- sharding: sum of per-rank partials == unsharded (no NCCL needed) --------------------
@pytest.mark.parametrize("n,n_ranks", [(12, 2), (10, 4), (9, 8), (5, 8)])
def test_shard_partials_sum_to_full(n, n_ranks):
    rng = np.random.default_rng(n * 10 + n_ranks)
    eri = rng.standard_normal((n, n, n, n))
    density = rng.standard_normal((n, n, 2))
    total = density[:, :, 0] + density[:, :, 1]
    full = scf_b200.JKBuilderCUDA.from_tensor(eri)
    jf, kf = full.jk(density, total)
    js, ks = np.zeros_like(jf), np.zeros_like(kf)
    covered = []
    for g in range(n_ranks):
        part = scf_b200.JKBuilderCUDA.from_tensor(eri, rank=g, n_ranks=n_ranks, allow_partial=True)
        lo, hi = part.shard_range
        assert (lo, hi) == o.shard_range(n, g, n_ranks)
        covered.append((lo, hi))
        j, k = part.jk(density, total)
        mask = np.ones(n, bool)
        mask[lo:hi] = False
        assert not j[:, mask].any() and not k[:, mask, :].any()   # zero outside owned columns
        js += j
        ks += k
        part.close()
    assert np.array_equal(js, jf) and np.array_equal(ks, kf)     # columns are disjoint: exact
    full.close()


# ---- determinism, error behaviour -------------------------------------------------------
def test_bitwise_reproducible():
    n = 48
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", 7)
    d = synth.hash_symmetric_matrix(n, 11)[:, :, None]
    j1, k1 = b.jk(d, 2 * d[:, :, 0])
    for _ in range(3):
        j2, k2 = b.jk(d, 2 * d[:, :, 0])
        assert np.array_equal(j1, j2) and np.array_equal(k1, k2)
    b.close()


def test_error_behaviour():
    b = scf_b200.JKBuilderCUDA(6)
    d = np.zeros((6, 6, 1))
    with pytest.raises(scf_b200.ScfbError) as ei:      # ERI not uploaded yet
        b.jk(d, d[:, :, 0])
    assert ei.value.code == scf_b200.capi.ERR_STATE
    b.close()
    b = scf_b200.JKBuilderCUDA.synthetic(6, "hash", 1)
    with pytest.raises(AssertionError):                 # JKBuilderFromTensor.jl:30-32
        b.add_2e_term(np.zeros((6, 6, 3)), np.zeros((6, 6, 3)), np.zeros((6, 6)))
    with pytest.raises(AssertionError):
        b.add_2e_term(np.zeros((6, 6, 1)), np.zeros((6, 6, 1)), np.zeros((5, 5)))
    with pytest.raises(AssertionError):
        b.add_2e_term(np.zeros((6, 6, 2)), np.zeros((6, 6, 1)), np.zeros((6, 6)))
    b.close()
    with pytest.raises(scf_b200.ScfbError):
        scf_b200.JKBuilderCUDA(6, device=99)


# ---- rows wider than one 256-thread pass (even n > 512, odd n > 255): 512-thread variant -----
@pytest.mark.parametrize("n,n_ranks", [(520, 130), (301, 43)])
def test_wide_rows_known_answers_on_a_shard(n, n_ranks):
    """The full tensor would be 0.6 TB / 66 GB; rank 0 of `n_ranks` owns only a few nu-slabs,
    whose J/K columns have exact closed-form answers for unit densities (hash family)."""
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, rank=0, n_ranks=n_ranks, allow_partial=True)
    lo, hi = b.shard_range
    assert lo == 0 and 0 < hi <= 8
    idx = np.arange(n)
    unit = lambda a, c: (lambda m: (m.__setitem__((a, c), 1.0), m)[1])(np.zeros((n, n)))
    a0, b0, a1, b1 = 5, n - 3, n - 1, 7
    j, k = b.jk(unit(a1, b1)[:, :, None], unit(a0, b0))
    own = np.arange(lo, hi)
    assert np.array_equal(j[:, lo:hi], synth.hash_eri_elements(a0, b0, idx[:, None], own[None, :]))
    assert np.array_equal(k[:, lo:hi, 0], synth.hash_eri_elements(idx[:, None], b1, a1, own[None, :]))
    assert not j[:, hi:].any() and not k[:, hi:, :].any()
    # a dense density: linearity against the two halves
    d = synth.hash_symmetric_matrix(n, 5)
    d1 = np.triu(d)
    j_full, k_full = b.jk(d[:, :, None], d)
    j_a, k_a = b.jk(d1[:, :, None], d1)
    j_b, k_b = b.jk((d - d1)[:, :, None], d - d1)
    assert rel_err(j_a + j_b, j_full) < RTOL and rel_err(k_a + k_b, k_full) < RTOL
    b.close()


# ---- 1/8 shards of the bench sizes: the b-split grid (2-4 CTAs per work item) --------------------
@pytest.mark.parametrize("n,rank,split,tma", [(256, 5, None, "1"), (256, 5, "2", "1"), (128, 7, "4", "1"),
                                              (64, 0, "2", "1"), (128, 3, None, "0"), (64, 1, "2", "0")])
def test_eighth_shard_with_b_splits_known_answers(n, rank, split, tma, monkeypatch):
    """What one GPU of eight runs at n_bf=256 / 128, with the work items split along b
    (csrc/jk_full.cu scfb_full_plan: the LDG kernel's rule, forced for the TMA kernel whose default splits only the tail): J
    becomes a per-CTA partial too and the persistent kernel's queue hands out chunk x slab x split
    items.  Unit densities pin the owned columns bit for bit (UHF: both K accumulators)."""
    monkeypatch.setenv("SCFB_FULL_TMA", tma)           # both read when the handle is created / launched
    if split:
        monkeypatch.setenv("SCFB_FULL_BSPLIT", split)
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, rank=rank, n_ranks=8, allow_partial=True)
    lo, hi = b.shard_range
    assert hi - lo == n // 8
    idx, own = np.arange(n), np.arange(lo, hi)
    unit = lambda a, c: (lambda m: (m.__setitem__((a, c), 1.0), m)[1])(np.zeros((n, n)))
    for (a0, b0), (a1, b1), (a2, b2) in [((1, n - 2), (n - 1, 0), (n // 2, n // 2 + 1)),
                                         ((n - 1, n - 1), (0, n - 1), (7, 8))]:
        j, k = b.jk(np.stack([unit(a1, b1), unit(a2, b2)], axis=2), unit(a0, b0))
        assert np.array_equal(j[:, lo:hi], synth.hash_eri_elements(a0, b0, idx[:, None], own[None, :]))
        assert np.array_equal(k[:, lo:hi, 0], synth.hash_eri_elements(idx[:, None], b1, a1, own[None, :]))
        assert np.array_equal(k[:, lo:hi, 1], synth.hash_eri_elements(idx[:, None], b2, a2, own[None, :]))
        assert not j[:, :lo].any() and not j[:, hi:].any()
    # dense densities: run-to-run bit-reproducibility of the fixed-order partial sums
    da, db = synth.hash_symmetric_matrix(n, synth.SEED_DA), synth.hash_symmetric_matrix(n, synth.SEED_DB)
    first = b.jk(np.stack([da, db], axis=2), da + db)
    for _ in range(3):
        again = b.jk(np.stack([da, db], axis=2), da + db)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])
    b.close()


# ---- device-pointer twin on the caller's stream (what bench.py times) ------------------------
@pytest.mark.parametrize("layout", ["full", "packed8"])
def test_device_pointer_api_on_callers_stream(layout):
    import torch
    n, n_spin = 48, 2
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    dens = np.asfortranarray(np.stack([da, db], axis=2))
    total = np.asfortranarray(da + db)
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout=layout)
    ref = np.zeros((n, n, n_spin), order="F")
    b.add_2e_term(ref, dens, total)                       # host-pointer path
    stream = torch.cuda.Stream()
    b.set_stream(stream.cuda_stream)
    # torch tensors viewing column-major arrays: tensor[s, v, m] == array[m, v, s]
    d_dens = torch.from_numpy(dens.T.copy()).cuda()
    d_total = torch.from_numpy(total.T.copy()).cuda()
    d_out = torch.zeros_like(d_dens)
    before = b.launch_count
    with torch.cuda.stream(stream):
        for _ in range(3):                                # accumulates: 3 x (J - K)
            b.add_2e_term_device(d_out.data_ptr(), d_dens.data_ptr(), d_total.data_ptr(), n_spin)
    stream.synchronize()
    assert b.launch_count - before >= 3                   # >= 1 kernel per build (full layout: the fused K1 alone)
    out = d_out.cpu().numpy().T
    assert rel_err(out, 3 * ref) < RTOL
    stream_ms, total_ms = b.last_build_ms()
    assert 0 < stream_ms <= total_ms
    b.close()


def test_chunked_slab_upload_equals_single_upload():
    rng = np.random.default_rng(8)
    n = 14
    eri = np.asfortranarray(rng.standard_normal((n, n, n, n)))
    dens = rng.standard_normal((n, n, 2))
    total = rng.standard_normal((n, n))
    whole = scf_b200.JKBuilderCUDA.from_tensor(eri)
    jw, kw = whole.jk(dens, total)
    whole.close()
    for rank, n_ranks in ((0, 1), (1, 3)):
        b = scf_b200.JKBuilderCUDA(n, rank=rank, n_ranks=n_ranks, allow_partial=True)
        with pytest.raises(scf_b200.ScfbError):               # not ready yet
            b.jk(dens, total)
        for first in range(4, n, 4):                          # 4 slabs at a time, ragged tail
            b.upload_slabs(eri[:, :, :, first:first + 4], first)
        lo, hi = b.shard_range
        if lo < 4:                                            # slabs 0..3 are still missing: refused
            with pytest.raises(scf_b200.ScfbError) as exc:
                b.mark_ready()
            assert exc.value.code == scf_b200.capi.ERR_STATE
        b.upload_slabs(eri[:, :, :, 0:4], 0)
        b.mark_ready()
        j, k = b.jk(dens, total)
        lo, hi = b.shard_range
        assert np.array_equal(j[:, lo:hi], jw[:, lo:hi]) and np.array_equal(k[:, lo:hi], kw[:, lo:hi])
        assert np.array_equal(b.eri_shard(), eri[:, :, :, lo:hi])
        b.close()


def test_multi_rank_handle_without_exchange_refuses_to_build():
    """a rank of a multi-rank builder with neither communicator nor peer exchange must not
    silently return its partial (include/scf_b200.h: SCFB_ERR_STATE unless scfb_allow_partial)."""
    n = 8
    rng = np.random.default_rng(5)
    eri = np.asfortranarray(rng.standard_normal((n, n, n, n)))
    dens = rng.standard_normal((n, n, 1))
    b = scf_b200.JKBuilderCUDA.from_tensor(eri, rank=1, n_ranks=2)
    with pytest.raises(scf_b200.ScfbError) as exc:
        b.jk(dens, 2 * dens[:, :, 0])
    assert exc.value.code == scf_b200.capi.ERR_STATE
    out = np.zeros((n, n, 1), order="F")
    with pytest.raises(scf_b200.ScfbError):
        b.add_2e_term(out, dens, 2 * dens[:, :, 0])
    assert not out.any()
    b.allow_partial()
    j, k = b.jk(dens, 2 * dens[:, :, 0])
    lo, hi = b.shard_range
    assert j[:, lo:hi].any() and not j[:, :lo].any()
    b.close()
