"""CPU ORACLE — test infrastructure only, never the product path.

NumPy/SciPy restatement of the Fock-build hot path of SelfConsistentField.jl and
of the SCF loop that drives it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module, and
only as the checker (or the CPU baseline), never as what is shipped or measured
as "ours".  The product (`selfconsistentfield.jl_b200`) never imports it and
fails loudly when its CUDA library is missing.

Parity status: PINNED through the reference's own four golden total energies
(`/root/reference/test/functionality_hartee_fock.jl:18,24,30,36`, rtol 1e-9),
reproduced by `tests/test_oracle_golden.py` from the reference's bundled integral
dumps (converted by `tests/golden/make_golden.py`).  The reference cannot run
here (no Julia), and its tests pin nothing finer than those energies; element
level J/K values are therefore arbitrated by the extended-precision evaluation
of the reference's literal index patterns (`add_2e_term_exact`).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Arrays use Julia's index order: `density[mu, nu, spin]`,
`eri[a, b, c, d]`; memory order is irrelevant here (einsum), the C-ABI callers
convert to column-major explicitly.

Third-party pieces the reference delegates to packages that are not vendored
(and not version-pinned: Project.toml has no [compat]):
  * `@tensor` (TensorOperations.jl)   -> numpy einsum on the same index strings
  * `eigen(Hermitian(F), Hermitian(S))` (LAPACK dsygvd) -> scipy.linalg.eigh(F, S)
  * `DataStructures.CircularBuffer`   -> `CircularBuffer` below (same semantics)
  * `Optim.optimize(..., BFGS())`     -> scipy.optimize.minimize(method="BFGS");
    the EDIIS *trajectory* is therefore not bit-comparable, only fixed points.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.optimize


# --------------------------------------------------------------------------------------
# DataStructures.CircularBuffer semantics used by the reference
# --------------------------------------------------------------------------------------
class CircularBuffer:
    """pushfirst! on a full buffer drops the last element, push! drops the first,
    fill! appends up to capacity without overwriting, pop! removes the last."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.items = []

    def pushfirst(self, x):
        self.items.insert(0, x)
        if len(self.items) > self.capacity:
            self.items.pop()

    def push(self, x):
        self.items.append(x)
        if len(self.items) > self.capacity:
            self.items.pop(0)

    def fill(self, x):
        while len(self.items) < self.capacity:
            self.items.append(x)

    def pop(self):
        return self.items.pop()

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):  # 0-based here
        return self.items[i]


# --------------------------------------------------------------------------------------
# Problem definition  (src/types/ScfProblem.jl:14-85, src/parts/assemble_hf_problem.jl)
# --------------------------------------------------------------------------------------
class JKBuilderFromTensor:
    """src/algorithms/JKBuilderFromTensor.jl:7-9 — dense in-memory ERI."""

    def __init__(self, eri):
        self.eri = np.asarray(eri, dtype=np.float64)

    def add_2e_term(self, out, density, total_density):
        return add_2e_term(out, self.eri, density, total_density)


@dataclass
class ScfProblem:
    overlap: np.ndarray
    n_elec: tuple
    n_orb: int
    restricted: bool
    terms_0e: dict
    terms_1e: dict
    terms_2e: dict
    h_core: np.ndarray = field(init=False)

    def __post_init__(self):  # ScfProblem.jl:63-84
        n = self.overlap.shape[0]
        assert self.overlap.shape == (n, n)
        assert len(self.n_elec) == 2
        self.h_core = np.zeros((n, n))
        for label, term in self.terms_1e.items():
            assert term.shape == (n, n), label
            self.h_core = self.h_core + term


def compute_nuclear_repulsion(coords, atom_numbers):
    """assemble_hf_problem.jl:8-18."""
    total = 0.0
    for a, za in enumerate(atom_numbers):
        for b in range(a):
            dist = np.linalg.norm(np.asarray(coords[a]) - np.asarray(coords[b]))
            total += za * atom_numbers[b] / dist
    return total


def assemble_hf_problem(system, integrals, n_orb=None, restricted=None, **extra):
    """assemble_hf_problem.jl:37-78.  `system`/`integrals` are dicts with the
    reference's field names; integrals["jk_builder_bb"] is any object with
    `add_2e_term(out, density, total_density)`."""
    overlap = integrals["overlap_bb"]
    if restricted is None:
        restricted = system["n_elec"][0] == system["n_elec"][1]
    if n_orb is None:
        n_orb = overlap.shape[0]
    terms_0e = {"nuclear_repulsion": compute_nuclear_repulsion(system["coords"],
                                                               system["atom_numbers"])}
    terms_1e = {"kinetic": integrals["kinetic_bb"],
                "nuclear_attraction": integrals["nuclear_attraction_bb"]}
    terms_2e = {"coulomb+exchange": integrals["jk_builder_bb"]}
    for label, term in extra.items():
        if hasattr(term, "add_2e_term"):
            terms_2e[label] = term
        elif isinstance(term, np.ndarray):
            terms_1e[label] = term
        elif isinstance(term, (int, float)):
            terms_0e[label] = term
        else:
            raise AssertionError(f"the term labelled {label} does not have a recognised type.")
    return ScfProblem(overlap, tuple(int(x) for x in system["n_elec"]), n_orb, restricted,
                      terms_0e, terms_1e, terms_2e)


def load_integral_npz(path):
    """Mirror of load_integral_hdf5.jl:6-55 on the converted fixtures.  Applies the
    dims reversal HDF5.jl applies (Julia eri[i,j,k,l] == file[l,k,j,i]); coords are
    re-transposed back to (n_atoms, 3) as the reference does at :47."""
    z = np.load(path, allow_pickle=False)
    eri = np.ascontiguousarray(z["electron_repulsion_bbbb"].T)
    system = {"coords": z["coords"], "atom_numbers": z["atom_numbers"],
              "n_elec": tuple(int(x) for x in z["nelec"])}
    integrals = {
        "kinetic_bb": z["kinetic_bb"].T.copy(),
        "nuclear_attraction_bb": z["nuclear_attraction_bb"].T.copy(),
        "overlap_bb": z["overlap_bb"].T.copy(),
        "electron_repulsion_bbbb": eri,
        "jk_builder_bb": JKBuilderFromTensor(eri),
        "discretisation": {"basis_type": str(z["basis_type"]),
                           "basis_set_name": str(z["basis_set_name"])},
    }
    return system, integrals


# --------------------------------------------------------------------------------------
# THE HOT PATH: J/K contraction  (src/algorithms/JKBuilderFromTensor.jl:22-51)
# --------------------------------------------------------------------------------------
def coulomb(eri, dtot):
    """J[mu,nu] = eri[a,b,mu,nu] * Dtot[a,b]   (JKBuilderFromTensor.jl:37,44)."""
    return np.einsum("abmn,ab->mn", eri, dtot)


def exchange(eri, d):
    """K[mu,nu] = eri[mu,b,a,nu] * D[a,b]      (JKBuilderFromTensor.jl:37,45-46)."""
    return np.einsum("mban,ab->mn", eri, d)


def add_2e_term(out, eri, density, total_density):
    """out[:,:,s] += J[Dtot] - K[D_s]   (JKBuilderFromTensor.jl:22-51), literal index
    patterns, valid for non-symmetric D and tensors lacking symmetry."""
    n, _, n_spin = density.shape
    assert n_spin in (1, 2)
    assert total_density.shape == (n, n)
    j = coulomb(eri, total_density)
    for s in range(n_spin):
        out[:, :, s] += j - exchange(eri, density[:, :, s])
    return out


def jk_exact(eri, density, total_density):
    """(J, K[:,:,s]) of the same literal contraction evaluated in extended precision
    (np.longdouble, 64-bit mantissa on x86) and rounded once to fp64 — the arbiter
    for the 1e-12 relative J/K parity gate."""
    e = np.asarray(eri, dtype=np.longdouble)
    j = np.einsum("abmn,ab->mn", e, np.asarray(total_density, dtype=np.longdouble))
    k = np.stack([np.einsum("mban,ab->mn", e, np.asarray(density[:, :, s], dtype=np.longdouble))
                  for s in range(density.shape[2])], axis=2)
    return j.astype(np.float64), k.astype(np.float64)


def add_2e_term_exact(out, eri, density, total_density):
    j, k = jk_exact(eri, density, total_density)
    for s in range(density.shape[2]):
        out[:, :, s] += j - k[:, :, s]
    return out


def add_2e_term_sharded(out, eri, density, total_density, n_ranks):
    """Emulates the (mu nu)-row-block sharding of SURVEY.md §8e on the CPU: rank g
    owns the nu-slabs [lo, hi) of `shard_range`, computes J/K columns for them, the
    partials are summed (the allreduce).  Must equal add_2e_term exactly up to
    summation order (here: exactly, columns are disjoint)."""
    n = density.shape[0]
    total = np.zeros_like(out)
    for g in range(n_ranks):
        lo, hi = shard_range(n, g, n_ranks)
        part = np.zeros_like(out)
        if hi > lo:
            sub = eri[:, :, :, lo:hi]
            j = np.einsum("abmn,ab->mn", sub, total_density)
            for s in range(density.shape[2]):
                part[:, lo:hi, s] = j - np.einsum("mban,ab->mn", sub, density[:, :, s])
        total += part
    out += total
    return out


def shard_range(n, rank, n_ranks):
    """Contiguous nu-slab ownership; ragged remainders go to the first ranks."""
    base, rem = divmod(n, n_ranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# --------------------------------------------------------------------------------------
# Fock assembly, Pulay error  (src/parts/compute_fock_matrix.jl, compute_pulay_error.jl)
# --------------------------------------------------------------------------------------
def compute_pulay_error(fock, density, overlap):
    """S D F - F D S per spin (compute_pulay_error.jl:1-25); 2-D in -> 2-D out."""
    if fock.ndim == 2:
        return overlap @ density @ fock - fock @ density @ overlap
    return np.stack([compute_pulay_error(fock[:, :, s], density[:, :, s], overlap)
                     for s in range(density.shape[2])], axis=2)


def compute_roothaan_effective_fock(fock, density, overlap):
    """compute_fock_matrix.jl:82-123 (ROHF; only needed for golden #3)."""
    da, db = density[:, :, 0], density[:, :, 1]
    fa, fb = fock[:, :, 0], fock[:, :, 1]
    s = overlap
    fc = (fa + fb) / 2
    pc = db @ s
    po = (da - db) @ s
    pv = np.eye(s.shape[0]) - da @ s
    feff = (pc.T @ fc @ pc / 2
            + po.T @ fb @ pc + po.T @ fc @ po / 2
            + pv.T @ fc @ pc + pv.T @ fa @ po + pv.T @ fc @ pv / 2)
    return feff + feff.T


def compute_fock_matrix(problem, density, **kwargs):
    """compute_fock_matrix.jl:8-71.  Returns (fock, error_pulay, energies)."""
    overlap = problem.overlap
    h_core = problem.h_core
    n = h_core.shape[0]
    n_spin = density.shape[2]
    assert density.shape == (n, n, n_spin) and n_spin in (1, 2)
    if n_spin == 1:
        total_density = 2 * density[:, :, 0]
    else:
        total_density = density[:, :, 0] + density[:, :, 1]
    g = np.zeros((n, n, n_spin))
    for _label, builder in problem.terms_2e.items():
        builder.add_2e_term(g, density, total_density)
    energies = {"0e": sum(problem.terms_0e.values()),
                "1e": np.trace(h_core @ total_density)}
    if n_spin == 1:
        energies["2e"] = np.trace(g[:, :, 0] @ density[:, :, 0])
    else:
        energies["2e"] = 0.5 * (np.trace(g[:, :, 0] @ density[:, :, 0])
                                + np.trace(g[:, :, 1] @ density[:, :, 1]))
    energies["total"] = energies["0e"] + energies["1e"] + energies["2e"]
    fock = g + h_core[:, :, None]
    if n_spin == 2 and problem.restricted:
        fock2 = compute_roothaan_effective_fock(fock, density, overlap)
        error_pulay = compute_pulay_error(fock2, total_density, overlap)  # stays 2-D
        fock = fock2.reshape(n, n, 1)
    else:
        error_pulay = compute_pulay_error(fock, density, overlap)
    return fock, error_pulay, energies


# --------------------------------------------------------------------------------------
# Orbitals, density, convergence
# --------------------------------------------------------------------------------------
def compute_orbitals(problem, fock, **kwargs):
    """compute_orbitals.jl:4-19 — generalized symmetric eigenproblem per spin."""
    n, _, n_spin = fock.shape
    n_orb = problem.n_orb
    orbcoeff = np.empty((n, n_orb, n_spin))
    orben = np.empty((n_orb, n_spin))
    for s in range(n_spin):
        w, v = scipy.linalg.eigh(fock[:, :, s], problem.overlap)
        orben[:, s] = w[:n_orb]
        orbcoeff[:, :, s] = v[:, :n_orb]
    return orben, orbcoeff


def compute_density(problem, orbcoeff):
    """compute_density.jl:4-35."""
    n_elec = problem.n_elec
    n, n_orb, n_spin = orbcoeff.shape
    assert n_spin in (1, 2)
    assert n_elec[0] < n_orb and n_elec[1] < n_orb
    cocc = (orbcoeff[:, :n_elec[0], 0], orbcoeff[:, :n_elec[1], n_spin - 1])
    n_densities = n_spin
    if (not problem.restricted) or n_elec[0] != n_elec[1]:
        n_densities = 2
    density = np.empty((n, n, n_densities))
    for s in range(n_densities):
        density[:, :, s] = cocc[s] @ cocc[s].T
    return density


@dataclass
class ScfConvergence:
    error_norm: float
    energy_change: dict
    is_converged: bool


def check_convergence(old, new, max_error_norm=5e-7, max_energy_total_change=1.25e-7,
                      max_energy_1e_change=5e-5, **kwargs):
    """check_convergence.jl:13-37 — note the *signed* energy tests."""
    error_norm = float(np.linalg.norm(new.error_pulay.reshape(-1)))
    change = {k: new.energies[k] - v for k, v in old.energies.items()}
    converged = True
    if error_norm >= max_error_norm:
        converged = False
    if change["total"] >= max_energy_total_change:
        converged = False
    if change["1e"] >= max_energy_1e_change:
        converged = False
    return ScfConvergence(error_norm, change, converged)


# --------------------------------------------------------------------------------------
# Iterate state + DIIS machinery
# --------------------------------------------------------------------------------------
@dataclass
class FockIterState:
    """types/ScfIterState.jl:10-27."""
    fock: np.ndarray
    error_pulay: np.ndarray
    energies: dict
    orbcoeff: object
    orben: object
    density: object


def update_iterate_matrix(it, fock):
    """ScfIterState.jl:32-35 — keeps the *old* energies/orbitals."""
    return FockIterState(fock, it.error_pulay, it.energies, it.orbcoeff, it.orben, it.density)


def spin(obj, s):
    """parts/misc.jl:8-10 — view of the last dimension (0-based here).  For the 2-D
    ROHF error matrix this is column `s`, as in the reference."""
    return obj[..., s]


class DiisState:
    """types/DiisState.jl:4-21."""

    def __init__(self, n_diis_size):
        self.iterate = CircularBuffer(n_diis_size)
        self.error = CircularBuffer(n_diis_size)
        self.iterationstate = CircularBuffer(n_diis_size)
        self.density = CircularBuffer(n_diis_size)
        self.energies = CircularBuffer(n_diis_size)
        self.n_diis_size = n_diis_size


class _DiisBase:
    needs_error = False
    needs_density = False

    def __init__(self, problem=None, n_diis_size=5, sync_spins=True,
                 conditioning_threshold=1e-14, coefficient_threshold=1e-6, **kwargs):
        self.state = (DiisState(n_diis_size), DiisState(n_diis_size))
        self.sync_spins = sync_spins
        self.conditioning_threshold = conditioning_threshold
        self.coefficient_threshold = coefficient_threshold

    # DiisState.jl:26-34
    def push_iterate(self, state, iterate, error, density, energies):
        state.iterate.pushfirst(iterate)
        state.error.pushfirst(error if error is not None else iterate - state.iterate[0])
        if self.needs_density:
            state.density.pushfirst(density)
        if energies is not None:
            state.energies.pushfirst(energies)

    # DiisState.jl:40-50
    def purge_from_state(self, state, count):
        if count == 0:
            return
        for _ in range(count + 1):
            state.iterate.pop()
            state.iterationstate.pop()
            if self.needs_error:
                state.error.pop()
            if self.needs_density:
                state.density.pop()

    def _build_rows(self, state, entry, bordered):
        """Common row-caching logic of cDIIS.jl:82-154 / EDIIS.jl:49-112."""
        assert state.n_diis_size > 0 and len(state.iterate) > 0
        m = min(state.n_diis_size, len(state.iterate))
        size = m + 1 if bordered else m
        a = np.zeros((size, size))
        new_values = CircularBuffer(state.n_diis_size)
        for j in range(m):
            new_values.push(entry(0, j))
        new_values.fill(0.0)
        state.iterationstate.pushfirst(new_values)
        if bordered:
            a[0, m] = 1
        for i in range(m):
            a[i, :m] = state.iterationstate[i].items[:m]
            if bordered:
                a[i, m] = 1
            state.iterationstate[i].pushfirst(0.0)
        # Symmetric(A): mirror the upper triangle
        return np.triu(a) + np.triu(a, 1).T


class cDIIS(_DiisBase):
    """algorithms/cDIIS.jl."""
    needs_error = True

    def build_matrix(self, state):  # cDIIS.jl:82-154
        def entry(i, j):
            ei, ej = state.error[i], state.error[j]
            return float(np.sum(ei * ej))  # tr(e_i' e_j); also right for the ROHF column vectors
        return self._build_rows(state, entry, bordered=True)

    def merge(self, a1, a2):  # cDIIS.jl:69-72 — the sum is computed and DISCARDED
        return a1

    def solve_coefficients(self, a, **kwargs):  # cDIIS.jl:33-63
        rhs = np.zeros(a.shape[0])
        rhs[-1] = 1
        lam, u = np.linalg.eigh(a)
        mask = np.abs(lam) > self.conditioning_threshold
        if not mask.any():
            # reference returns a bare vector here (latent crash);
            # handled gracefully: no extrapolation, nothing purged
            c = np.zeros(a.shape[0] - 1)
            c[0] = 1
            return c, 0
        c = u[:, mask] @ np.diag(1.0 / lam[mask]) @ u[:, mask].T @ rhs
        return c[:-1], int((~mask).sum())


class EDIIS(_DiisBase):
    """algorithms/EDIIS.jl."""
    needs_density = True

    def build_matrix(self, state):  # EDIIS.jl:49-112
        def entry(i, j):
            return float(np.trace((state.iterate[i] - state.iterate[j])
                                  @ (state.density[i] - state.density[j])))
        return self._build_rows(state, entry, bordered=False)

    def merge(self, b1, b2):  # EDIIS.jl:126-128
        return b1 + b2

    def solve_coefficients(self, b, energybuffer=None, **kwargs):  # EDIIS.jl:20-44
        m = b.shape[0]
        e = np.array([en["total"] for en in energybuffer.items[:m]])

        def f(x):
            c = x ** 2 / np.sum(x ** 2)
            return e @ c - 0.5 * c @ b @ c

        guess = np.ones(m) / m
        guess[0] = 1
        res = scipy.optimize.minimize(f, guess, method="BFGS", options={"xrtol": 1e-4})
        x = res.x
        c = x ** 2 / np.sum(x ** 2)
        if res.nit == 0:
            c = np.zeros(m)
            c[0] = 1
        return c, 0


def compute_next_iterate_matrix(state, c, out, coefficient_threshold):
    """parts/diis.jl:42-55 — mutates `c` in place exactly like the reference."""
    mask = np.abs(c) >= coefficient_threshold
    mask[0] = True
    c[int(np.argmax(c))] += c[~mask].sum()
    for i in np.nonzero(mask)[0]:
        out += c[i] * state.iterate[i]


def compute_next_iterate(acc, it):
    """parts/diis.jl:4-36."""
    if acc is None:
        return it
    iterate = it.fock
    n_spin = iterate.shape[-1]
    for s in range(n_spin):
        acc.push_iterate(acc.state[s], spin(it.fock, s), spin(it.error_pulay, s),
                         spin(it.density, s), it.energies)
    new_iterate = np.zeros(iterate.shape)

    def coefficients(a):
        return acc.solve_coefficients(a, energybuffer=acc.state[0].energies)

    if acc.sync_spins and n_spin == 2:
        a = acc.merge(acc.build_matrix(acc.state[0]), acc.build_matrix(acc.state[1]))
        c, purge = coefficients(a)
        for s in range(n_spin):
            compute_next_iterate_matrix(acc.state[s], c, new_iterate[..., s],
                                        acc.coefficient_threshold)
        for s in range(n_spin):
            acc.purge_from_state(acc.state[s], purge)
    else:
        for s in range(n_spin):
            c, purge = coefficients(acc.build_matrix(acc.state[s]))
            compute_next_iterate_matrix(acc.state[s], c, new_iterate[..., s],
                                        acc.coefficient_threshold)
            acc.purge_from_state(acc.state[s], purge)
    return update_iterate_matrix(it, new_iterate)


# --------------------------------------------------------------------------------------
# Guess + SCF loop
# --------------------------------------------------------------------------------------
def compute_guess_hcore(problem):
    """util/guess.jl:11-22."""
    n = problem.h_core.shape[0]
    _, c = compute_orbitals(problem, problem.h_core.reshape(n, n, 1))
    return compute_density(problem, c)


def break_spin_symmetry(density):
    """util/break_spin_symmetry.jl:5-18 — beta density <- its tridiagonal part, in place."""
    if density.shape[2] == 1:
        return density
    b = density[:, :, 1]
    density[:, :, 1] = np.triu(np.tril(b, 1), -1)
    return density


def roothan_step(problem, it, **kwargs):
    """algorithms/Roothaan.jl:4-9."""
    orben, orbcoeff = compute_orbitals(problem, it.fock)
    density = compute_density(problem, orbcoeff)
    fock, err, energies = compute_fock_matrix(problem, density)
    return FockIterState(fock, err, energies, orbcoeff, orben, density)


def run_scf(problem, guess_density, max_iter=100, ediis_max_error_norm=0.1, verbose=False,
            start_with_ediis=True, **kwargs):
    """run_scf.jl:4-65, including the stale-result-on-convergence quirk (:40 vs :48):
    the returned energies/fock/orbitals are those of the last *extrapolated* iterate.
    Extra keys (not in the reference): "energies_newest" = energies of the newest
    Fock build, for same-iterate comparisons at the 1e-10 Eh level."""
    damping = EDIIS(problem, **kwargs) if start_with_ediis else cDIIS(problem, **kwargs)
    switched = not start_with_ediis
    scfconv = None
    n_iter = 0
    fock, err, energies = compute_fock_matrix(problem, guess_density)
    it = FockIterState(fock, err, energies, None, None, guess_density)
    newit = it
    for i in range(1, max_iter + 1):
        n_iter = i
        it = compute_next_iterate(damping, it)
        newit = roothan_step(problem, it)
        scfconv = check_convergence(it, newit, **kwargs)
        if verbose:
            e = newit.energies
            print(f" {i:4d} {e['1e']:14.8f} {e['2e']:14.8f} {e['total']:14.8f} "
                  f"{scfconv.error_norm:16.9g}")
        if scfconv.is_converged:
            break
        if scfconv.error_norm < ediis_max_error_norm and not switched:
            damping = cDIIS(problem, **kwargs)
            switched = True
        it = newit
    return {
        "n_iter": n_iter, "n_applies": math.nan, "problem": problem,
        "orben": it.orben, "orbcoeff": it.orbcoeff,
        "density": compute_density(problem, it.orbcoeff) if it.orbcoeff is not None else None,
        "converged": scfconv.is_converged, "fock": it.fock, "energies": it.energies,
        "error_norm": scfconv.error_norm, "energy_change": scfconv.energy_change,
        "energies_newest": newit.energies,
    }


# --------------------------------------------------------------------------------------
# Post-SCF energy split (examples/HartreeFock.jl:67-98)
# --------------------------------------------------------------------------------------
def termwise_jk_energies(eri, density):
    n_spin = density.shape[2]
    da, db = density[:, :, 0], density[:, :, n_spin - 1]
    ej = 0.5 * sum(np.einsum("mn,abmn,ab->", x, eri, y) for x in (da, db) for y in (da, db))
    ek = -0.5 * (np.einsum("mn,mban,ab->", da, eri, da) + np.einsum("mn,mban,ab->", db, eri, db))
    return float(ej), float(ek)


# --------------------------------------------------------------------------------------
# The reference's CPU execution strategy for the hot path (for the CPU baseline):
# J as a GEMV on the no-copy (ab)x(mu nu) matrix view, K as permute-copy + GEMV
# (how TensorOperations' BLAS backend lowers the two contractions).
# --------------------------------------------------------------------------------------
def add_2e_term_blas(out, eri_f, density, total_density):
    """`eri_f` must be Fortran-ordered (Julia memory layout).  Same results as
    add_2e_term up to summation order."""
    n = density.shape[0]
    n2 = n * n
    m = eri_f.reshape(n2, n2, order="F")                # rows (a,b), cols (mu,nu): a view
    j = (m.T @ total_density.reshape(n2, order="F")).reshape(n, n, order="F")
    # K: bring (a=3rd index, b) to the front: P[a,b,mu,nu] = eri[mu,b,a,nu]  (full n^4 copy)
    p = np.asfortranarray(eri_f.transpose(2, 1, 0, 3)).reshape(n2, n2, order="F")
    for s in range(density.shape[2]):
        k = (p.T @ density[:, :, s].reshape(n2, order="F")).reshape(n, n, order="F")
        out[:, :, s] += j - k
    return out
