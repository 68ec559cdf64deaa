"""Host-side mirror of the reference's problem / Fock-term / algorithm API — the objects
that `run_scf` and `examples/HartreeFock.jl` drive — with every O(n^4) and O(n^3) step
executed by libscf_b200.so on the GPU.  (The reference is Julia; Julia is not available in
this environment, so the host side is Python over the same C ABI the Julia glue binds.)

Same names, argument meaning and defaults as the reference (paths relative to
/root/reference):
  ScfProblem, assemble_hf_problem      src/types/ScfProblem.jl:14-85, src/parts/assemble_hf_problem.jl
  compute_fock_matrix                  src/parts/compute_fock_matrix.jl:8-71
  compute_pulay_error                  src/parts/compute_pulay_error.jl
  compute_orbitals / compute_density   src/parts/compute_orbitals.jl, compute_density.jl
  check_convergence                    src/parts/check_convergence.jl:13-37
  cDIIS / EDIIS / compute_next_iterate src/algorithms/cDIIS.jl, EDIIS.jl, src/parts/diis.jl
  roothan_step, run_scf                src/algorithms/Roothaan.jl, src/run_scf.jl
  compute_guess, break_spin_symmetry   src/util/guess.jl, break_spin_symmetry.jl
  load_integral_hdf5                   src/util/load_integral_hdf5.jl

Two execution modes behind the same `run_scf`:
  * device-resident (default when the problem's only two-electron term is a JKBuilderCUDA
    and the SCF is RHF closed-shell or UHF): S, h_core, D, F, the Pulay error, the orbitals
    and the DIIS history live in HBM (`scfb_scf_*`); per iteration only scalars, DIIS rows
    and <= 6 coefficients cross PCIe.
  * generic (any TwoElectronBuilder mix, ROHF): arrays live on the host, the builders are
    called through the `add_2e_term` boundary (CUDA for JKBuilderCUDA), the commutator and
    eigensolve go through the stateless `scfb_pulay_error` / `scfb_sygvd` calls.
There is no CPU implementation of the two-electron contraction anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .builder import JKBuilderCUDA, TwoElectronBuilder

# --------------------------------------------------------------------------------------
# Problem
# --------------------------------------------------------------------------------------


@dataclass
class ScfProblem:
    """src/types/ScfProblem.jl:14-85."""
    overlap: np.ndarray
    n_elec: tuple
    n_orb: int
    restricted: bool
    terms_0e: dict
    terms_1e: dict
    terms_2e: dict
    h_core: np.ndarray = field(init=False)

    def __post_init__(self):
        assert len(self.n_elec) == 2, "n_elec needs to have exactly two elements"
        n_bas = self.overlap.shape[0]
        assert self.overlap.shape == (n_bas, n_bas), "Overlap must be a quadratic matrix"
        h_core = np.zeros((n_bas, n_bas))
        for label, term in self.terms_1e.items():
            assert term.shape == (n_bas, n_bas), f"Term '{label}' must have a shape of {n_bas} x {n_bas}"
            h_core = h_core + term
        self.h_core = h_core


def is_closed_shell(obj):
    n_elec = obj["n_elec"] if isinstance(obj, dict) else obj.n_elec
    return n_elec[0] == n_elec[1]


def compute_nuclear_repulsion(system):
    """assemble_hf_problem.jl:8-18."""
    total = 0.0
    coords, z = np.asarray(system["coords"], dtype=float), system["atom_numbers"]
    for a in range(len(z)):
        for b in range(a):
            total += z[a] * z[b] / np.linalg.norm(coords[a] - coords[b])
    return total


def assemble_hf_problem(system, integrals, n_orb=None, restricted=None, **kwargs):
    """assemble_hf_problem.jl:37-78.  Extra keyword terms are sorted into the 2e / 1e / 0e
    dictionaries by type, exactly like the reference."""
    overlap = integrals["overlap_bb"]
    if restricted is None:
        restricted = is_closed_shell(system)
    if n_orb is None:
        n_orb = overlap.shape[0]
    terms_0e = {"nuclear_repulsion": compute_nuclear_repulsion(system)}
    terms_1e = {"kinetic": integrals["kinetic_bb"],
                "nuclear_attraction": integrals["nuclear_attraction_bb"]}
    terms_2e = {"coulomb+exchange": integrals["jk_builder_bb"]}
    for label, term in kwargs.items():
        if isinstance(term, TwoElectronBuilder):
            terms_2e[label] = term
        elif isinstance(term, np.ndarray) and term.ndim == 2:
            terms_1e[label] = term
        elif isinstance(term, (int, float)):
            terms_0e[label] = term
        else:
            raise AssertionError(f"the term labelled {label} does not have a recognised type.")
    return ScfProblem(overlap, (int(system["n_elec"][0]), int(system["n_elec"][1])), int(n_orb),
                      bool(restricted), terms_0e, terms_1e, terms_2e)


def load_integral_hdf5(path, slab_batch=None, **builder_kw):
    """load_integral_hdf5.jl:6-55 with the ERI going straight into a JKBuilderCUDA.  The file
    is parsed by the package's own minimal HDF5 reader (no libhdf5 here); dims are reversed
    like HDF5.jl does (Julia eri[i,j,k,l] == file[l,k,j,i]).

    slab_batch=N streams the tensor to the device N nu-slabs at a time (file row v IS Julia's
    column-major slab eri[:,:,:,v]), reading only the rows this builder's shard owns — the
    whole tensor never exists on the host (the reference's TODO at load_integral_hdf5.jl:9-10);
    `integrals["electron_repulsion_bbbb"]` is then None.  With layout="packed8" the slabs are
    packed on the device as they arrive (csrc/jk_packed.cu: pack_slabs_kernel), so the n^4 tensor
    never exists on the device either."""
    from .minih5 import MiniH5
    f = MiniH5(path)
    if slab_batch:
        n = f.shape("electron_repulsion_bbbb")[0]
        builder = JKBuilderCUDA(n, **builder_kw)
        lo, hi = builder.shard_range
        for first in range(lo, hi, slab_batch):
            rows = f.read_rows("electron_repulsion_bbbb", first, min(first + slab_batch, hi))
            builder.upload_slabs(rows.T, first)           # (S,n,n,n) C-order == (n,n,n,S) F-order
        builder.mark_ready()
        small = {k: f.read(k) for k in ("kinetic_bb", "nuclear_attraction_bb", "overlap_bb")}
        system, integrals = _integrals_from_arrays(
            dict(small, electron_repulsion_bbbb=None), f.read("system/nelec"),
            f.read("system/atom_numbers"), f.read("system/coords"), f.read("discretisation/basis_type"),
            f.read("discretisation/basis_set_name") if "basis_set_name" in f.keys("discretisation") else None,
            builder_kw, builder=builder)
        return system, integrals
    return _integrals_from_arrays(
        {k: f.read(k) for k in ("electron_repulsion_bbbb", "kinetic_bb", "nuclear_attraction_bb",
                                "overlap_bb")},
        f.read("system/nelec"), f.read("system/atom_numbers"), f.read("system/coords"),
        f.read("discretisation/basis_type"),
        f.read("discretisation/basis_set_name") if "basis_set_name" in f.keys("discretisation") else None,
        builder_kw)


def load_integral_npz(path, **builder_kw):
    """Same as load_integral_hdf5 for the converted fixtures under tests/golden/."""
    z = np.load(path, allow_pickle=False)
    return _integrals_from_arrays(z, z["nelec"], z["atom_numbers"], z["coords"], str(z["basis_type"]),
                                  str(z["basis_set_name"]), builder_kw)


def _integrals_from_arrays(a, nelec, atnums, coords, basis_type, basis_set_name, builder_kw,
                           builder=None):
    eri = None if builder is not None else np.asarray(a["electron_repulsion_bbbb"]).T  # dims reversed
    system = {"coords": np.asarray(coords, dtype=float), "atom_numbers": np.asarray(atnums),
              "n_elec": (int(nelec[0]), int(nelec[1]))}
    integrals = {
        "kinetic_bb": np.asarray(a["kinetic_bb"]).T.copy(),
        "nuclear_attraction_bb": np.asarray(a["nuclear_attraction_bb"]).T.copy(),
        "overlap_bb": np.asarray(a["overlap_bb"]).T.copy(),
        "electron_repulsion_bbbb": eri,
        "jk_builder_bb": builder if builder is not None else JKBuilderCUDA.from_tensor(eri, **builder_kw),
        "discretisation": {"basis_type": basis_type, "basis_set_name": basis_set_name},
    }
    return system, integrals


# --------------------------------------------------------------------------------------
# Generic (host-array) parts; heavy steps go through the C ABI
# --------------------------------------------------------------------------------------
def _any_cuda_builder(problem) -> JKBuilderCUDA:
    for b in problem.terms_2e.values():
        if isinstance(b, JKBuilderCUDA):
            return b
    raise RuntimeError("this package needs a JKBuilderCUDA among terms_2e: the dense steps run "
                       "on its GPU (there is no CPU fallback)")


def compute_pulay_error(fock, density, overlap, *, builder):
    """S D F - F D S per spin block on the GPU (compute_pulay_error.jl:1-25); 2-D in, 2-D out."""
    two_d = fock.ndim == 2
    f3 = fock[:, :, None] if two_d else fock
    d3 = density[:, :, None] if two_d else density
    n, _, n_spin = d3.shape
    assert n_spin == 1 or n_spin == 2
    assert f3.shape == (n, n, n_spin) and overlap.shape == (n, n)
    f, d, s = capi.as_f64_colmajor(f3), capi.as_f64_colmajor(d3), capi.as_f64_colmajor(overlap)
    e = np.empty((n, n, n_spin), order="F")
    capi.check(capi.lib().scfb_pulay_error(builder._h, n_spin, capi.ptr(f), capi.ptr(d), capi.ptr(s),
                                           capi.ptr(e)), builder._h)
    return e[:, :, 0] if two_d else e


def compute_roothaan_effective_fock(fock, density, overlap):
    """compute_fock_matrix.jl:82-123 for HOST arrays (the generic mode with several two-electron
    builders); the device-resident loop builds the same matrix on the GPU (csrc/dense.cu:
    rohf_effective_fock)."""
    da, db = density[:, :, 0], density[:, :, 1]
    fa, fb = fock[:, :, 0], fock[:, :, 1]
    s = overlap
    fc = (fa + fb) / 2
    pc, po, pv = db @ s, (da - db) @ s, np.eye(s.shape[0]) - da @ s
    feff = (pc.T @ fc @ pc / 2 + po.T @ fb @ pc + po.T @ fc @ po / 2
            + pv.T @ fc @ pc + pv.T @ fa @ po + pv.T @ fc @ pv / 2)
    return feff + feff.T


def compute_fock_matrix(problem, density, **kwargs):
    """compute_fock_matrix.jl:8-71 -> (fock, error_pulay, energies)."""
    overlap, h_core = problem.overlap, problem.h_core
    n_bas = h_core.shape[0]
    n_spin = density.shape[2]
    assert density.shape == (n_bas, n_bas, n_spin)
    assert n_spin == 1 or n_spin == 2
    total_density = 2 * density[:, :, 0] if n_spin == 1 else density[:, :, 0] + density[:, :, 1]
    g = np.zeros((n_bas, n_bas, n_spin), order="F")
    for _label, builder in problem.terms_2e.items():
        builder.add_2e_term(g, density, total_density)        # the plugin boundary
    dev = _any_cuda_builder(problem)
    fock = np.empty((n_bas, n_bas, n_spin), order="F")
    e1e, e2e = C.c_double(), C.c_double()
    gd, hd = capi.as_f64_colmajor(g), capi.as_f64_colmajor(h_core)
    dd, td = capi.as_f64_colmajor(density), capi.as_f64_colmajor(total_density)
    capi.check(capi.lib().scfb_fock_assemble(dev._h, n_spin, capi.ptr(gd), capi.ptr(hd), capi.ptr(dd),
                                             capi.ptr(td), capi.ptr(fock), C.byref(e1e), C.byref(e2e)),
               dev._h)
    energies = {"0e": sum(problem.terms_0e.values()), "1e": e1e.value, "2e": e2e.value}
    energies["total"] = energies["0e"] + energies["1e"] + energies["2e"]
    if n_spin == 2 and problem.restricted:
        fock2 = compute_roothaan_effective_fock(fock, density, overlap)
        error_pulay = compute_pulay_error(fock2, total_density, overlap, builder=dev)
        fock = fock2.reshape(n_bas, n_bas, 1)
    else:
        error_pulay = compute_pulay_error(fock, density, overlap, builder=dev)
    return fock, error_pulay, energies


def compute_orbitals(problem, fock, **kwargs):
    """compute_orbitals.jl:4-19 via cuSOLVER Dsygvd -> (orben (n_orb,n_spin), orbcoeff (n,n_orb,n_spin))."""
    n_bas, _, n_spin = fock.shape
    assert fock.shape == (n_bas, n_bas, n_spin), "fock should be a square matrix"
    n_orb = problem.n_orb
    dev = _any_cuda_builder(problem)
    f, s = capi.as_f64_colmajor(fock), capi.as_f64_colmajor(problem.overlap)
    orben = np.empty((n_orb, n_spin), order="F")
    orbcoeff = np.empty((n_bas, n_orb, n_spin), order="F")
    capi.check(capi.lib().scfb_sygvd(dev._h, n_spin, n_orb, capi.ptr(f), capi.ptr(s), capi.ptr(orben),
                                     capi.ptr(orbcoeff)), dev._h)
    return orben, orbcoeff


def compute_density(problem, orbcoeff):
    """compute_density.jl:4-35."""
    n_elec = problem.n_elec
    n_bas, n_orb, n_spin = orbcoeff.shape
    assert n_spin == 1 or n_spin == 2, "Only one or two spin components allowed"
    assert n_elec[0] < n_orb and n_elec[1] < n_orb, "n_elec of alpha and beta should be less than n_orb"
    n_densities = n_spin
    if (not problem.restricted) or n_elec[0] != n_elec[1]:
        n_densities = 2
    density = np.empty((n_bas, n_bas, n_densities), order="F")
    dev = _any_cuda_builder(problem)
    c = capi.as_f64_colmajor(orbcoeff)
    capi.check(capi.lib().scfb_density(dev._h, n_spin, n_densities, n_elec[0], n_elec[1], n_orb,
                                       capi.ptr(c), capi.ptr(density)), dev._h)
    return density


@dataclass
class ScfConvergence:
    """check_convergence.jl:4-8."""
    error_norm: float
    energy_change: dict
    is_converged: bool


def _convergence(error_norm, old_energies, new_energies, max_error_norm=5e-7,
                 max_energy_total_change=1.25e-7, max_energy_1e_change=5e-5, **kwargs):
    """check_convergence.jl:13-37 — signed (not absolute) energy-change tests."""
    change = {k: new_energies[k] - v for k, v in old_energies.items()}
    converged = True
    if error_norm >= max_error_norm:
        converged = False
    if change["total"] >= max_energy_total_change:
        converged = False
    if change["1e"] >= max_energy_1e_change:
        converged = False
    return ScfConvergence(float(error_norm), change, converged)


def check_convergence(olditerate, newiterate, **kwargs):
    return _convergence(np.linalg.norm(newiterate.error_pulay.reshape(-1)), olditerate.energies,
                        newiterate.energies, **kwargs)


@dataclass
class FockIterState:
    """types/ScfIterState.jl:10-27."""
    fock: np.ndarray
    error_pulay: np.ndarray
    energies: dict
    orbcoeff: object
    orben: object
    density: object


# --------------------------------------------------------------------------------------
# Accelerators.  The coefficient logic (bordered system, eigenvalue mask, folding, purge
# counts, EDIIS minimisation) is host code on <= 6x6 data; the matrices it needs come from a
# history backend: host arrays (generic mode) or the HBM-resident history (device mode).
# --------------------------------------------------------------------------------------
class _HostHistory:
    """Per-spin newest-first history of (F, e, D) host arrays."""

    def __init__(self, cap):
        self.cap = cap
        self.f, self.e, self.d = [], [], []

    def push(self, f, e, d):
        for lst, x in ((self.f, f), (self.e, e), (self.d, d)):
            lst.insert(0, None if x is None else np.array(x, copy=True))
            del lst[self.cap:]
        rc = [float(np.sum(self.e[0] * ej)) for ej in self.e]
        re = [float(np.sum((self.f[0] - fj) * (self.d[0] - dj).T)) if self.d[0] is not None else 0.0
              for fj, dj in zip(self.f, self.d)]
        return rc, re

    def combine(self, c, out):
        out[...] = 0.0
        for ci, fi in zip(c, self.f):
            if ci != 0.0:
                out += ci * fi

    def purge(self, count):
        if count:
            for lst in (self.f, self.e, self.d):
                del lst[max(0, len(lst) - (count + 1)):]

    def __len__(self):
        return len(self.f)


class Accelerator:
    """types/Accelerator.jl:2-18."""


class _Diis(Accelerator):
    uses = "cdiis"

    def __init__(self, problem=None, n_diis_size=5, sync_spins=True, conditioning_threshold=1e-14,
                 coefficient_threshold=1e-6, **kwargs):
        self.n_diis_size = n_diis_size
        self.sync_spins = sync_spins
        self.conditioning_threshold = conditioning_threshold
        self.coefficient_threshold = coefficient_threshold
        self.b = [np.zeros((0, 0)), np.zeros((0, 0))]     # per-spin B matrices, newest first
        self.energies = []                                 # totals, newest first (state[1].energies)
        self.host = (_HostHistory(n_diis_size), _HostHistory(n_diis_size))

    # the row caching of cDIIS.jl:123-151 / EDIIS.jl:89-109 amounts to: new row/column in
    # front, everything else shifts by one, the oldest drops out at capacity
    def insert_row(self, spin, row):
        m = len(row)
        old = self.b[spin]
        new = np.zeros((m, m))
        new[0, :] = row
        new[:, 0] = row
        k = m - 1
        new[1:, 1:] = old[:k, :k]
        self.b[spin] = new

    def purge_rows(self, spin, count):
        if count:
            m = max(0, self.b[spin].shape[0] - (count + 1))
            self.b[spin] = self.b[spin][:m, :m]

    def push_energy(self, energies):
        self.energies.insert(0, energies["total"])
        del self.energies[self.n_diis_size:]


class cDIIS(_Diis):
    """algorithms/cDIIS.jl."""
    uses = "cdiis"

    def system(self, spin):
        b = self.b[spin]
        m = b.shape[0]
        a = np.zeros((m + 1, m + 1))
        a[:m, :m] = b
        a[:m, m] = 1
        a[m, :m] = 1
        return a

    def merge(self, a1, a2):
        return a1          # cDIIS.jl:69-72: the sum is computed and discarded

    def solve(self, a):
        """cDIIS.jl:33-63 -> (c, n_purge)."""
        rhs = np.zeros(a.shape[0])
        rhs[-1] = 1
        lam, u = np.linalg.eigh(a)
        mask = np.abs(lam) > self.conditioning_threshold
        if not mask.all():
            print(f"   Removing {int((~mask).sum())} of {len(mask)} eigenvalues from DIIS linear system.")
        if not mask.any():
            print("All eigenvalues are under the threshold! Skipping iterate modification…")
            c = np.zeros(a.shape[0] - 1)
            c[0] = 1
            return c, 0
        c = (u[:, mask] * (1.0 / lam[mask])) @ (u[:, mask].T @ rhs)
        return c[:-1].copy(), int((~mask).sum())


class EDIIS(_Diis):
    """algorithms/EDIIS.jl."""
    uses = "ediis"

    def system(self, spin):
        return self.b[spin].copy()

    def merge(self, b1, b2):
        return b1 + b2     # EDIIS.jl:126-128

    def solve(self, b):
        """EDIIS.jl:20-44: minimise E'c - c'Bc/2 over c = x^2/sum(x^2) (BFGS) -> (c, 0)."""
        import scipy.optimize
        m = b.shape[0]
        e = np.asarray(self.energies[:m])

        def f(x):
            c = x ** 2 / np.sum(x ** 2)
            return e @ c - 0.5 * c @ b @ c

        def grad(x):   # exact derivative (the reference uses forward-mode AD, EDIIS.jl:33)
            s2 = np.sum(x ** 2)
            c = x ** 2 / s2
            g_c = e - b @ c                                   # d f / d c   (b symmetric)
            return 2 * x / s2 * (g_c - c @ g_c)               # chain rule through c = x^2 / sum x^2

        guess = np.ones(m) / m
        guess[0] = 1
        res = scipy.optimize.minimize(f, guess, jac=grad, method="BFGS", options={"xrtol": 1e-4})
        c = res.x ** 2 / np.sum(res.x ** 2)
        if res.nit == 0:
            c = np.zeros(m)
            c[0] = 1
        return c, 0


def _fold_small_coefficients(c, threshold):
    """parts/diis.jl:45-47: mutates `c` (the argmax entry) exactly like the reference; returns
    the vector actually used in the linear combination (masked-out entries zeroed)."""
    mask = np.abs(c) >= threshold
    mask[0] = True
    c[int(np.argmax(c))] += c[~mask].sum()
    return np.where(mask, c, 0.0)


def _next_iterate(acc, n_spin, push, combine, purge):
    """parts/diis.jl:4-36 over an abstract history: push(s)->(row_c,row_e), combine(s,c),
    purge(s,count)."""
    for s in range(n_spin):
        rc, re = push(s)
        acc.insert_row(s, rc if acc.uses == "cdiis" else re)
    if acc.sync_spins and n_spin == 2:
        a = acc.merge(acc.system(0), acc.system(1))
        c, n_purge = acc.solve(a)
        for s in range(n_spin):
            combine(s, _fold_small_coefficients(c, acc.coefficient_threshold))
        for s in range(n_spin):
            purge(s, n_purge)
            acc.purge_rows(s, n_purge)
    else:
        for s in range(n_spin):
            c, n_purge = acc.solve(acc.system(s))
            combine(s, _fold_small_coefficients(c, acc.coefficient_threshold))
            purge(s, n_purge)
            acc.purge_rows(s, n_purge)


def compute_next_iterate(acc, iterstate):
    """parts/diis.jl:4-36 on host arrays (generic mode)."""
    if acc is None:
        return iterstate
    fock = iterstate.fock
    n_spin = fock.shape[-1]
    new_iterate = np.zeros(fock.shape, order="F")
    acc.push_energy(iterstate.energies)

    def push(s):   # misc.jl:8-10 `spin`: last-dimension view (a column for the 2-D ROHF error)
        return acc.host[s].push(fock[..., s], iterstate.error_pulay[..., s],
                                iterstate.density[..., s] if iterstate.density is not None else None)

    _next_iterate(acc, n_spin, push,
                  lambda s, c: acc.host[s].combine(c, new_iterate[..., s]),
                  lambda s, k: acc.host[s].purge(k))
    return FockIterState(new_iterate, iterstate.error_pulay, iterstate.energies, iterstate.orbcoeff,
                         iterstate.orben, iterstate.density)


def roothan_step(problem, iterate, **kwargs):
    """algorithms/Roothaan.jl:4-9."""
    orben, orbcoeff = compute_orbitals(problem, iterate.fock, **kwargs)
    density = compute_density(problem, orbcoeff)
    fock, error_pulay, energies = compute_fock_matrix(problem, density, **kwargs)
    return FockIterState(fock, error_pulay, energies, orbcoeff, orben, density)


# --------------------------------------------------------------------------------------
# Guess
# --------------------------------------------------------------------------------------
def compute_guess_hcore(problem, coords=None, atom_numbers=None, **kwargs):
    """util/guess.jl:11-22."""
    n_bas = problem.h_core.shape[0]
    _, orbcoeff = compute_orbitals(problem, problem.h_core.reshape(n_bas, n_bas, 1), **kwargs)
    return compute_density(problem, orbcoeff)


def compute_guess(problem, coords=None, atom_numbers=None, method="hcore", **kwargs):
    """util/guess.jl:91-105.  Only `hcore` is built (the only method the reference's tests
    and example use for Gaussian bases); `loewdin`/`random` are one-off O(n^3) host steps
    outside the hot-path scope."""
    avail = {"hcore": compute_guess_hcore}
    if method not in avail:
        raise ValueError(f"Method {method} is not an available guess method, try one of {' '.join(avail)}")
    return avail[method](problem, coords, atom_numbers, **kwargs)


def break_spin_symmetry(density):
    """util/break_spin_symmetry.jl:5-18 (in place)."""
    if density.shape[2] == 1:
        return density
    density[:, :, 1] = np.triu(np.tril(density[:, :, 1], 1), -1)
    return density


# --------------------------------------------------------------------------------------
# Device-resident SCF step
# --------------------------------------------------------------------------------------
class DeviceScf:
    """Thin wrapper of the `scfb_scf_*` calls on one builder's handle."""
    FOCK, FOCK_NEW, DENSITY, ERROR, ORBCOEFF, ORBEN, ORBCOEFF_PREV, ORBEN_PREV, DENSITY_PREV = range(9)

    def __init__(self, problem, builder, n_diis_size=5):
        self.problem, self.builder = problem, builder
        self._h = builder._h
        self.n = builder.n_bas
        s = capi.as_f64_colmajor(problem.overlap, (self.n, self.n))
        hc = capi.as_f64_colmajor(problem.h_core, (self.n, self.n))
        e0 = float(sum(problem.terms_0e.values()))
        self._call("scfb_scf_setup", capi.ptr(s), capi.ptr(hc), C.c_double(e0), problem.n_elec[0],
                   problem.n_elec[1], problem.n_orb, int(problem.restricted), n_diis_size)
        ns, nd = C.c_int(), C.c_int()
        self._call("scfb_scf_n_spin", C.byref(ns))
        self._call("scfb_scf_n_densities", C.byref(nd))
        self.n_spin, self.n_dens = ns.value, nd.value     # ROHF: one Fock block, two densities

    def _call(self, name, *args):
        capi.check(getattr(capi.lib(), name)(self._h, *args), self._h)

    def set_density(self, density):
        d = capi.as_f64_colmajor(density, (self.n, self.n, self.n_dens))
        self._call("scfb_scf_set_density", self.n_dens, capi.ptr(d))

    def set_fock(self, fock):
        f = capi.as_f64_colmajor(fock, (self.n, self.n, self.n_spin))
        self._call("scfb_scf_set_fock", self.n_spin, capi.ptr(f))

    def fock(self):
        e = (C.c_double * 4)()
        norm = C.c_double()
        self._call("scfb_scf_fock", e, C.byref(norm))
        return {"0e": e[0], "1e": e[1], "2e": e[2], "total": e[3]}, norm.value

    def roothaan(self):
        self._call("scfb_scf_roothaan")

    def diis_push(self, spin):
        m = C.c_int()
        rc, re = (C.c_double * 16)(), (C.c_double * 16)()
        self._call("scfb_scf_diis_push", spin, C.byref(m), rc, re)
        return list(rc[:m.value]), list(re[:m.value])

    def diis_extrapolate(self, spin, c):
        arr = (C.c_double * len(c))(*[float(x) for x in c])
        self._call("scfb_scf_diis_extrapolate", spin, len(c), arr)

    def diis_purge(self, spin, count):
        self._call("scfb_scf_diis_purge", spin, int(count))

    def diis_reset(self):
        self._call("scfb_scf_diis_reset")

    def accept_fock(self):
        self._call("scfb_scf_accept_fock")

    def get(self, what):
        n, ns = self.n, self.n_spin
        if what in (self.DENSITY, self.DENSITY_PREV):
            ns = self.n_dens
        shape = (n, ns) if what in (self.ORBEN, self.ORBEN_PREV) else (n, n, ns)
        out = np.empty(shape, order="F")
        self._call("scfb_scf_get", what, capi.ptr(out))
        return out


def _device_eligible(problem):
    if len(problem.terms_2e) != 1:
        return None
    b = next(iter(problem.terms_2e.values()))
    if not isinstance(b, JKBuilderCUDA):
        return None
    return b                 # RHF, UHF and ROHF (Roothaan effective Fock on the device) all qualify


# --------------------------------------------------------------------------------------
# run_scf
# --------------------------------------------------------------------------------------
def run_scf(problem, guess_density, max_iter=100, ediis_max_error_norm=0.1, mode="auto",
            verbose=True, **kwargs):
    """run_scf.jl:4-65.  Same control flow, kwargs splatting and result dictionary (including
    the stale-result-on-convergence behaviour of :40 vs :48); extra key "energies_newest"
    holds the energies of the newest Fock build.  mode: "auto" | "device" | "generic"."""
    builder = _device_eligible(problem) if mode in ("auto", "device") else None
    if mode == "device" and builder is None:
        raise ValueError("device mode needs exactly one JKBuilderCUDA term")
    if builder is not None:
        return _run_scf_device(problem, builder, guess_density, max_iter, ediis_max_error_norm,
                               verbose, **kwargs)
    return _run_scf_generic(problem, guess_density, max_iter, ediis_max_error_norm, verbose, **kwargs)


def _print_header(verbose):
    if verbose:
        print("%5s %14s %14s %14s %15s %12s" % ("iter", "e1e", "e2e", "etot", "scf_error", "n_applies"))


def _print_row(verbose, i, e, err):
    if verbose:   # run_scf.jl:35-37 (n_applies is NaN there)
        print(" %4d %14.8f %14.8f %14.8f %16.9g %12s" % (i, e["1e"], e["2e"], e["total"], err, "NaN"))


def _run_scf_generic(problem, guess_density, max_iter, ediis_max_error_norm, verbose, **kwargs):
    damping = EDIIS(problem, **kwargs)
    scfconv, switched, n_iter = None, False, 0
    fock, error_pulay, energies = compute_fock_matrix(problem, guess_density, **kwargs)
    iterate = FockIterState(fock, error_pulay, energies, None, None, guess_density)
    newiterate = iterate
    _print_header(verbose)
    for i in range(1, max_iter + 1):
        n_iter = i
        iterate = compute_next_iterate(damping, iterate)
        newiterate = roothan_step(problem, iterate, **kwargs)
        scfconv = check_convergence(iterate, newiterate, **kwargs)
        _print_row(verbose, i, newiterate.energies, scfconv.error_norm)
        if scfconv.is_converged:
            break
        if scfconv.error_norm < ediis_max_error_norm and not switched:
            if verbose:
                print("**** Switching algorithms EDIIS -> cDIIS ****")
            damping = cDIIS(problem, **kwargs)
            switched = True
        iterate = newiterate
    return {
        "n_iter": n_iter, "n_applies": math.nan, "problem": problem, "orben": iterate.orben,
        "orbcoeff": iterate.orbcoeff,
        "density": compute_density(problem, iterate.orbcoeff) if iterate.orbcoeff is not None else None,
        "converged": scfconv.is_converged, "fock": iterate.fock, "energies": iterate.energies,
        "error_norm": scfconv.error_norm, "energy_change": scfconv.energy_change,
        "energies_newest": newiterate.energies, "mode": "generic",
    }


def _run_scf_device(problem, builder, guess_density, max_iter, ediis_max_error_norm, verbose,
                    **kwargs):
    n_diis_size = kwargs.get("n_diis_size", 5)
    dev = DeviceScf(problem, builder, n_diis_size)
    n_spin = dev.n_spin
    assert guess_density.shape == (dev.n, dev.n, dev.n_dens)
    damping = EDIIS(problem, **kwargs)
    scfconv, switched, n_iter = None, False, 0
    dev.set_density(guess_density)
    energies, _ = dev.fock()                              # Fock build #0 (run_scf.jl:16)
    new_energies = energies
    first = True      # orbitals of the *iterate* do not exist before the first Roothaan step
    broke = False
    _print_header(verbose)
    for i in range(1, max_iter + 1):
        n_iter = i
        damping.push_energy(energies)                     # compute_next_iterate (parts/diis.jl)
        _next_iterate(damping, n_spin, dev.diis_push, lambda s, c: dev.diis_extrapolate(s, c),
                      dev.diis_purge)
        dev.roothaan()                                    # roothan_step (Roothaan.jl:5-7)
        new_energies, error_norm = dev.fock()
        scfconv = _convergence(error_norm, energies, new_energies, **kwargs)
        _print_row(verbose, i, new_energies, scfconv.error_norm)
        if scfconv.is_converged:
            broke = True
            break
        if scfconv.error_norm < ediis_max_error_norm and not switched:
            if verbose:
                print("**** Switching algorithms EDIIS -> cDIIS ****")
            damping = cDIIS(problem, **kwargs)
            dev.diis_reset()
            switched = True
        energies = new_energies
        first = False
    # run_scf.jl:40 breaks before :48 `iterate = newiterate`: on convergence the result holds
    # the extrapolated iterate with the previous step's orbitals/energies
    if broke:
        orbcoeff = None if first else dev.get(dev.ORBCOEFF_PREV)[:, :problem.n_orb, :]
        orben = None if first else dev.get(dev.ORBEN_PREV)[:problem.n_orb, :]
        density = None if first else dev.get(dev.DENSITY_PREV)
        fock, res_energies = dev.get(dev.FOCK), energies
    else:
        orbcoeff = dev.get(dev.ORBCOEFF)[:, :problem.n_orb, :]
        orben = dev.get(dev.ORBEN)[:problem.n_orb, :]
        density = dev.get(dev.DENSITY)
        fock, res_energies = dev.get(dev.FOCK_NEW), new_energies
    return {
        "n_iter": n_iter, "n_applies": math.nan, "problem": problem, "orben": orben,
        "orbcoeff": orbcoeff, "density": density, "converged": scfconv.is_converged, "fock": fock,
        "energies": res_energies, "error_norm": scfconv.error_norm,
        "energy_change": scfconv.energy_change, "energies_newest": new_energies, "mode": "device",
        "density_newest": dev.get(dev.DENSITY),
    }


# --------------------------------------------------------------------------------------
# Post-SCF energy split (examples/HartreeFock.jl:67-98) — SURVEY.md §8f rank 1
# --------------------------------------------------------------------------------------
def compute_termwise_hf_energies(res, problem):
    """Adds per-term energies incl. "coulomb" and "exchange" to res["energies"].  The
    reference asserts `isa(JKBuilder, JKBuilderFromTensor)` and contracts `.eri` six more
    times; here the same scalars come from one more J/K pass of the CUDA builder."""
    density = res["density"]
    n_spin = density.shape[2]
    energies = res["energies"]
    da, db = density[:, :, 0], density[:, :, n_spin - 1]
    for label, term in problem.terms_0e.items():
        energies[label] = term
    for label, term in problem.terms_1e.items():
        energies[label] = float(np.sum(da * term.T) + np.sum(db * term.T))
    assert len(problem.terms_2e) == 1
    builder = problem.terms_2e["coulomb+exchange"]
    assert isinstance(builder, JKBuilderCUDA)
    dens2 = np.stack([da, db], axis=2)
    j, k = builder.jk(dens2, da + db)
    # E_J = 1/2 sum_{s,t} D_s[mu,nu] J[D_t][mu,nu];  E_K = -1/2 sum_s D_s[mu,nu] K[D_s][mu,nu]
    energies["coulomb"] = 0.5 * float(np.sum((da + db) * j))
    energies["exchange"] = -0.5 * float(np.sum(da * k[:, :, 0]) + np.sum(db * k[:, :, 1]))
    return energies
