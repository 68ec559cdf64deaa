"""The small-matrix eigensolver of the device-resident SCF step (csrc/eig_jacobi.cu, exported as
scfb_syev_jacobi) against LAPACK (numpy.linalg.eigh — what compute_orbitals.jl:14 calls after the
reduction to standard form): eigenvalues, residuals, orthonormality, ordering; cold and warm starts;
every shape class of the kernels (odd n -> phantom column / row, n <= 64 -> one CTA with the matrix in
shared memory, n >= 65 -> a cluster of 8 CTAs exchanging column blocks over distributed shared memory,
block widths with and without a bye in the intra-block tournament, n > 320 -> a cluster of 16, n > 448 -> three block buffers per CTA
and a two-phase exchange, n = 512 -> the size limit)."""
import numpy as np
import pytest

import scf_b200

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def builder():
    b = scf_b200.JKBuilderCUDA.synthetic(4, "hash", 1)
    yield b
    b.close()


def sym(rng, n, scale=1.0):
    a = rng.standard_normal((n, n)) * scale
    return (a + a.T) / 2


def check(a, w, z, tol=1e-12):
    n = a.shape[0]
    nrm = max(np.linalg.norm(a, 2), 1e-300)
    w_ref = np.linalg.eigvalsh(a)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - w_ref).max() <= tol * nrm
    assert np.abs(a @ z - z * w).max() <= 4 * tol * nrm
    assert np.abs(z.T @ z - np.eye(n)).max() <= tol


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 16, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 167, 168, 169, 200, 255, 256, 320, 321, 383, 384, 448, 449, 500, 511, 512])
def test_cold_start_matches_lapack(builder, n):
    rng = np.random.default_rng(n)
    a = sym(rng, n)
    w, z, sweeps = builder.syev_jacobi(a)
    check(a, w, z)
    assert 1 <= sweeps <= 16


def test_batch_of_two_and_fock_like_spectrum(builder):
    """Two matrices in one launch (the UHF spin blocks); spectrum like a Fock matrix in an orthonormal
    basis: a few strongly negative core levels, a gap, a band of virtuals."""
    rng = np.random.default_rng(7)
    n = 128
    a = np.empty((2, n, n))
    for b in range(2):
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.concatenate([-20 + rng.standard_normal(4), -1.5 + 0.5 * rng.standard_normal(28),
                              0.3 + 2.0 * rng.random(n - 32)])
        a[b] = (q * lam) @ q.T
        a[b] = (a[b] + a[b].T) / 2
    w, z, sweeps = builder.syev_jacobi(a)
    for b in range(2):
        check(a[b], w[b], z[b])


@pytest.mark.parametrize("n", [24, 127, 128, 168, 256, 384, 512])
def test_warm_start_takes_fewer_sweeps(builder, n):
    rng = np.random.default_rng(100 + n)
    a = sym(rng, n)
    w0, z0, cold = builder.syev_jacobi(a)
    a2 = a + 1e-3 * sym(rng, n)          # the next SCF iterate: a small change
    w, z, warm = builder.syev_jacobi(a2, v0=z0)
    check(a2, w, z)
    assert warm < cold and warm <= 4
    w1, z1, again = builder.syev_jacobi(a2, v0=z)   # converged SCF: the iterate no longer changes
    check(a2, w1, z1)
    assert again <= 2


def test_degenerate_and_trivial_matrices(builder):
    n = 40
    w, z, _ = builder.syev_jacobi(np.zeros((n, n)))
    assert np.all(w == 0) and np.abs(z.T @ z - np.eye(n)).max() < 1e-14
    w, z, _ = builder.syev_jacobi(np.eye(n) * -3.0)
    assert np.abs(w + 3.0).max() < 1e-13
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.repeat([-2.0, -2.0, 0.5, 0.5, 0.5, 1.0, 4.0, 4.0], 5)   # eight 5-fold degenerate levels
    a = (q * lam) @ q.T
    a = (a + a.T) / 2
    w, z, _ = builder.syev_jacobi(a)
    check(a, w, z)
    d = np.diag(np.arange(n, dtype=float) - 7.5)                      # already diagonal, unsorted input order
    w, z, sweeps = builder.syev_jacobi(d[::-1, ::-1].copy())
    check(d[::-1, ::-1], w, z)
    assert sweeps == 1


@pytest.mark.parametrize("scale", [1e-30, 1e-8, 1e8, 1e30])
def test_scale_invariance(builder, scale):
    """The rotation angles come from single-precision units: the working matrix is normalised first,
    so the magnitude of A must not matter."""
    rng = np.random.default_rng(11)
    a = sym(rng, 64, scale)
    w, z, _ = builder.syev_jacobi(a)
    check(a, w, z)


def test_rejects_matrices_beyond_the_size_limit(builder):
    with pytest.raises(scf_b200.ScfbError):
        builder.syev_jacobi(np.eye(513))
