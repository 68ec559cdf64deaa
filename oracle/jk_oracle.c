/* CPU ORACLE (test infrastructure only) — plain C restatement of the reference's J/K
 * contraction and of the synthetic tensor families.  Never linked into the product.
 *
 * jk_literal_ld follows /root/reference/src/algorithms/JKBuilderFromTensor.jl:36-38,43-47
 * literally (valid for non-symmetric D and tensors without symmetry), accumulating in long
 * double and rounding once: the arbiter for the 1e-12 relative J/K gate at sizes where the
 * NumPy longdouble einsum is too slow.  jk_onepass_omp is a best-effort single-pass
 * OpenMP CPU kernel (what a CPU would do with the GPU kernel's idea) used only as an extra,
 * clearly labelled CPU reference point in bench.py.
 *
 * Layout: column-major (Julia): eri[a + n*(b + n*(c + n*d))], density[m + n*(v + n*s)].
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; no -march so the .so travels).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* J[m,v] = sum_ab eri[a,b,m,v] Dtot[a,b];  K[m,v,s] = sum_ab eri[m,b,a,v] D[a,b,s] */
void jk_literal_ld(const double* eri, int n, int n_spin, const double* density, const double* dtot,
                   double* J, double* K) {
  const size_t N = (size_t)n;
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < n; ++v)
    for (int m = 0; m < n; ++m) {
      long double j = 0.0L, k0 = 0.0L, k1 = 0.0L;
      for (size_t b = 0; b < N; ++b)
        for (size_t a = 0; a < N; ++a) {
          j += (long double)eri[a + N * (b + N * (m + N * v))] * (long double)dtot[a + N * b];
          const long double x = (long double)eri[m + N * (b + N * (a + N * v))];
          k0 += x * (long double)density[a + N * b];
          if (n_spin == 2) k1 += x * (long double)density[a + N * b + N * N];
        }
      J[m + N * v] = (double)j;
      K[m + N * v] = (double)k0;
      if (n_spin == 2) K[m + N * v + N * N] = (double)k1;
    }
}

/* The same literal contraction for a SAMPLE of nu-slabs: `eri_slabs` holds n_slabs slabs
 * X[a,b,c] = eri[a,b,c,nu] back to back; column s of the outputs is J[:,nu_s] (n doubles) and
 * K[:,nu_s,spin] ((n, n_slabs, n_spin) column-major).  J[m,nu] and K[m,nu] need only slab nu
 * (JKBuilderFromTensor.jl:37: the last tensor index is the open index nu in both terms), which is
 * what lets the tests check dense densities at sizes whose full tensor no host can hold. */
void jk_literal_ld_slabs(const double* eri_slabs, int n, int n_spin, int n_slabs, const double* density,
                         const double* dtot, double* Jcols, double* Kcols) {
  const size_t N = (size_t)n;
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < n_slabs; ++v)
    for (int m = 0; m < n; ++m) {
      const double* X = eri_slabs + (size_t)v * N * N * N;
      long double j = 0.0L, k0 = 0.0L, k1 = 0.0L;
      for (size_t b = 0; b < N; ++b)
        for (size_t a = 0; a < N; ++a) {
          j += (long double)X[a + N * (b + N * m)] * (long double)dtot[a + N * b];
          const long double x = (long double)X[m + N * (b + N * a)];
          k0 += x * (long double)density[a + N * b];
          if (n_spin == 2) k1 += x * (long double)density[a + N * b + N * N];
        }
      Jcols[m + N * v] = (double)j;
      Kcols[m + N * v] = (double)k0;
      if (n_spin == 2) Kcols[m + N * v + N * (size_t)n_slabs] = (double)k1;
    }
}

/* One pass over nu-slabs [nu_lo, nu_hi) of a tensor stored slab-contiguously (`eri` points at
 * slab nu_lo): every element feeds J and every K.  fp64, one thread per slab. */
void jk_onepass_omp(const double* eri, int n, int n_spin, const double* density, const double* dtot,
                    int nu_lo, int nu_hi, double* J, double* K) {
  const size_t N = (size_t)n, N2 = N * N;
#pragma omp parallel for schedule(dynamic, 1)
  for (int v = nu_lo; v < nu_hi; ++v) {
    const double* X = eri + (size_t)(v - nu_lo) * N2 * N;
    double* Kv0 = K + N * v;
    double* Kv1 = K + N * v + N2;
    for (size_t a = 0; a < N; ++a) Kv0[a] = 0.0;
    if (n_spin == 2)
      for (size_t a = 0; a < N; ++a) Kv1[a] = 0.0;
    for (size_t c = 0; c < N; ++c) {
      double j = 0.0;
      for (size_t b = 0; b < N; ++b) {
        const double* x = X + N * (b + N * c);
        const double* dt = dtot + N * b;
        const double d0 = density[c + N * b];
        const double d1 = n_spin == 2 ? density[c + N * b + N2] : 0.0;
        double jj = 0.0;
        if (n_spin == 2) {
          for (size_t a = 0; a < N; ++a) {
            jj += x[a] * dt[a];
            Kv0[a] += x[a] * d0;
            Kv1[a] += x[a] * d1;
          }
        } else {
          for (size_t a = 0; a < N; ++a) {
            jj += x[a] * dt[a];
            Kv0[a] += x[a] * d0;
          }
        }
        j += jj;
      }
      J[c + N * v] = j;
    }
  }
}

/* The permuted temporary the reference's K contraction materialises on a CPU
 * (JKBuilderFromTensor.jl:37,45-46 `eri[mu,beta,alpha,nu] * D[alpha,beta]`: the open indices (1,4)
 * cannot be fused, so TensorOperations copies the tensor into contraction order and calls BLAS;
 * SURVEY.md 8a): P[a,b,m,v] = X[m,b,a,v] for the nu-slabs in `eri_slabs`, as a cache-blocked,
 * all-core transpose of the (first, third) index pair — what a tuned permute does, so that the
 * CPU reference arm is not handicapped by a strided NumPy copy. */
void permute_mban_slabs(const double* eri_slabs, int n, int n_slabs, double* out) {
  const size_t N = (size_t)n, N2 = N * N, N3 = N2 * N;
  const int T = 32;
#pragma omp parallel for collapse(2) schedule(static)
  for (int v = 0; v < n_slabs; ++v)
    for (int b = 0; b < n; ++b) {
      const double* X = eri_slabs + (size_t)v * N3 + N * (size_t)b; /* X[m + N2*a] */
      double* P = out + (size_t)v * N3 + N * (size_t)b;             /* P[a + N2*m] */
      for (int a0 = 0; a0 < n; a0 += T)
        for (int m0 = 0; m0 < n; m0 += T) {
          const int a1 = a0 + T < n ? a0 + T : n, m1 = m0 + T < n ? m0 + T : n;
          for (int m = m0; m < m1; ++m)
            for (int a = a0; a < a1; ++a) P[(size_t)a + N2 * (size_t)m] = X[(size_t)m + N2 * (size_t)a];
        }
    }
}

/* ---- synthetic families (DESIGN.md "synthetic inputs"; twins of csrc/eri_gen.cu) ---------- */
static inline uint64_t pair_index(uint64_t i, uint64_t j) {
  return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i;
}

static inline double hash_value(uint64_t p1, uint64_t p2, uint64_t seed) {
  const uint64_t p = p1 > p2 ? p1 : p2, q = p1 > p2 ? p2 : p1;
  uint64_t z = p * (p + 1) / 2 + q + (seed + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * 0x1.0p-52 - 1.0;
}

/* slabs [nu_lo, nu_hi) of the HASH family, slab-contiguous */
void hash_eri_slabs(double* out, int n, uint64_t seed, int nu_lo, int nu_hi) {
  const size_t N = (size_t)n;
#pragma omp parallel for collapse(2) schedule(static)
  for (int d = nu_lo; d < nu_hi; ++d)
    for (int c = 0; c < n; ++c) {
      const uint64_t p2 = pair_index((uint64_t)c, (uint64_t)d);
      double* dst = out + N * N * (c + N * (size_t)(d - nu_lo));
      for (size_t b = 0; b < N; ++b)
        for (size_t a = 0; a < N; ++a) dst[a + N * b] = hash_value(pair_index(a, b), p2, seed);
    }
}

/* slabs of the CHAIN family; aux = [soft, x[n], S[n*n]] */
void chain_eri_slabs(double* out, int n, const double* aux, int nu_lo, int nu_hi) {
  const size_t N = (size_t)n;
  const double soft = aux[0];
  const double* x = aux + 1;
  const double* S = aux + 1 + n;
#pragma omp parallel for collapse(2) schedule(static)
  for (int d = nu_lo; d < nu_hi; ++d)
    for (int c = 0; c < n; ++c) {
      const double s_cd = S[c + N * d], r_cd = 0.5 * (x[c] + x[d]);
      double* dst = out + N * N * (c + N * (size_t)(d - nu_lo));
      for (size_t b = 0; b < N; ++b)
        for (size_t a = 0; a < N; ++a) {
          const double r_ab = 0.5 * (x[a] + x[b]);
          const double dd = r_ab - r_cd;
          double v = (S[a + N * b] * s_cd) / sqrt(soft + dd * dd);
          if (fabs(v) < 1e-200) v = 0.0;
          dst[a + N * b] = v;
        }
    }
}
