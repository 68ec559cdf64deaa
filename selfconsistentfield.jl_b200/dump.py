"""Result dump in the molsturm HDF5 layout — host-side mirror of
/root/reference/src/util/dump_molsturm_hdf5.jl:4-143.

The reference writes through HDF5.jl, which hands libhdf5 the column-major bytes of a Julia array
together with REVERSED dimensions; every array here is therefore written as `A.T` in C order, so
that a file produced from the same SCF result has the same datasets, shapes and bytes a Julia run
would produce (a Python reader sees `orbcoeff_bf` as (2 n_orb, n_bas)ᵀ exactly like the comment at
dump_molsturm_hdf5.jl:107-111 describes).  Written with this package's own minimal HDF5 writer
(`minih5.write_h5`; no libhdf5 / h5py exists in this environment), which also makes the
reference's `sanitise_hdf5_hack` (a Python/h5py post-processing call, :150-175) unnecessary: the
two string attributes are written as plain fixed-length strings and `restricted` as a one-byte
integer from the start.
"""
from __future__ import annotations

import numpy as np

from .minih5 import write_h5


def build_ab_matrix(mat):
    """dump_molsturm_hdf5.jl:8-27: alpha-beta block-diagonal matrix from 1 or 2 spin blocks."""
    mat = np.asarray(mat)
    if mat.ndim == 2:
        ma = mb = mat
    else:
        assert mat.ndim == 3 and mat.shape[2] in (1, 2)
        ma, mb = mat[:, :, 0], mat[:, :, -1]
    n = ma.shape[0]
    out = np.zeros((2 * n, 2 * ma.shape[1]))
    out[:n, :ma.shape[1]] = ma
    out[n:, ma.shape[1]:] = mb
    return out


def concat_spin(mat):
    """dump_molsturm_hdf5.jl:33-51: alpha block followed by beta block (rows for vectors,
    columns for matrices)."""
    mat = np.asarray(mat)
    assert mat.ndim in (2, 3)
    if mat.ndim == 2:
        assert mat.shape[1] in (1, 2)
        return np.concatenate([mat[:, 0], mat[:, -1]])
    assert mat.shape[2] in (1, 2)
    return np.concatenate([mat[:, :, 0], mat[:, :, -1]], axis=1)


def matrix_b2f(mat, orbcoeff):
    """dump_molsturm_hdf5.jl:57-75: C' M C per spin block, as an alpha-beta block matrix."""
    mat, orbcoeff = np.asarray(mat), np.asarray(orbcoeff)
    n_bas, n_orb, n_spin = orbcoeff.shape
    if mat.ndim == 2:
        mat = np.stack([mat] * n_spin, axis=2)
    assert mat.shape == (n_bas, n_bas, n_spin)
    ff = np.stack([orbcoeff[:, :, s].T @ mat[:, :, s] @ orbcoeff[:, :, s] for s in range(n_spin)], axis=2)
    return build_ab_matrix(ff)


def dump_molsturm_hdf5(integrals, scfres, file, export_eri_bbbb=False, export_eri_ffff=False):
    """dump_molsturm_hdf5.jl:83-143.  `scfres` is what `run_scf` returns after
    `compute_termwise_hf_energies` has added the per-term energies (examples/HartreeFock.jl:128-132)."""
    if export_eri_bbbb:
        raise NotImplementedError("Exporting eri in basis functions not yet implemented.")
    if export_eri_ffff:
        raise NotImplementedError("Exporting eri in orbitals not yet implemented.")
    problem, energies, change = scfres["problem"], scfres["energies"], scfres["energy_change"]
    if scfres.get("orbcoeff") is None or scfres.get("orben") is None:
        # run_scf.jl:40 breaks before :48: a run that converges in its very first iteration returns
        # the guess iterate, which has no orbitals yet (the reference would fail inside concat_spin)
        raise ValueError("scfres holds no orbitals (SCF converged before the first Roothaan step was "
                         "accepted): nothing to dump; rerun with tighter thresholds or another guess")
    fock, orbcoeff = np.asarray(scfres["fock"]), np.asarray(scfres["orbcoeff"])
    n_bas = fock.shape[0]
    julia = {                                   # name -> array with Julia index semantics
        "n_alpha": np.int64(problem.n_elec[0]), "n_beta": np.int64(problem.n_elec[1]),
        "n_bas": np.int64(n_bas), "restricted": np.uint8(bool(problem.restricted)),   # plain 0/1 byte; the reference's
        # sanitise step re-creates it as h5py's bool enum over int8 (same storage size and values, a
        # different HDF5 datatype class) — readers that test the value are unaffected
        "overlap_bb": build_ab_matrix(problem.overlap),
        "n_orbs_alpha": np.int64(problem.n_orb), "n_orbs_beta": np.int64(problem.n_orb),
        "orben_f": concat_spin(scfres["orben"]),
        "fock_bb": build_ab_matrix(fock),
        "fock_ff": matrix_b2f(fock, orbcoeff),
        "hcore_ff": matrix_b2f(problem.h_core, orbcoeff),
        "orbcoeff_bf": concat_spin(orbcoeff).T,            # the explicit adjoint at :111
        "n_mtx_applies": np.float64(scfres["n_applies"]),  # NaN in the reference (run_scf.jl:13)
        "n_iter": np.int64(scfres["n_iter"]),
        "energy_coulomb": np.float64(energies["coulomb"]),
        "energy_exchange": np.float64(energies["exchange"]),
        "energy_kinetic": np.float64(energies["kinetic"]),
        "energy_nuclear_attraction": np.float64(energies["nuclear_attraction"]),
        "energy_nuclear_repulsion": np.float64(energies["nuclear_repulsion"]),
        "energy_ground_state": np.float64(energies["total"]),
        "final_1e_energy_change": np.float64(change["1e"]),
        "final_error_norm": np.float64(scfres["error_norm"]),
        "final_tot_energy_change": np.float64(change["total"]),
    }
    # HDF5.jl: column-major bytes + reversed dims == the transposed array in C order
    write_h5(file, {k: np.asarray(v).T for k, v in julia.items()},
             {"format_version": "0.1.0", "format_type": "hdf5"})
