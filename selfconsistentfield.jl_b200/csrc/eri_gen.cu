// K3 — closed-form synthetic ERI families generated directly in HBM (a 34 GB / 550 GB
// tensor cannot be staged through the host).  Formulas are frozen in DESIGN.md
// ("synthetic inputs") and mirrored bit-for-bit by oracle/synth.py and oracle/jk_oracle.c;
// tests compare device vs host twins at small n.
//
// Tensor element eri[a,b,c,d] = (ab|cd) in chemists' notation, the index order the
// reference contracts in (/root/reference/src/algorithms/JKBuilderFromTensor.jl:37).
#include "scfb_internal.h"

namespace {

__host__ __device__ __forceinline__ uint64_t pair_index(uint64_t i, uint64_t j) {
  return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i;
}

// HASH family: value depends only on the canonical 8-fold orbit of (a,b,c,d).
__device__ __forceinline__ double hash_value(uint64_t p1, uint64_t p2, uint64_t seed) {
  const uint64_t p = p1 > p2 ? p1 : p2, q = p1 > p2 ? p2 : p1;
  uint64_t z = p * (p + 1) / 2 + q + (seed + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return __dadd_rn(__dmul_rn((double)(z >> 11), 0x1.0p-52), -1.0);  // uniform in [-1, 1)
}

// CHAIN family: (ab|cd) = S_ab S_cd / sqrt(soft + (R_ab - R_cd)^2), R_ab = (x_a + x_b)/2,
// |v| < 1e-200 flushed to 0.  Explicit _rn intrinsics: no FMA contraction, so the host
// twin (compiled with -ffp-contract=off) matches bit for bit.
__device__ __forceinline__ double chain_value(double s_ab, double r_ab, double s_cd, double r_cd,
                                              double soft) {
  const double d = __dadd_rn(r_ab, -r_cd);
  const double den = __dsqrt_rn(__dadd_rn(soft, __dmul_rn(d, d)));
  const double v = __ddiv_rn(__dmul_rn(s_ab, s_cd), den);
  return fabs(v) < 1e-200 ? 0.0 : v;
}

// one CTA per (b, c, d) row at a time, threads along a
__global__ void generate_full_kernel(double* __restrict__ eri, int n, int nu_lo, int n_slabs,
                                     int family, uint64_t seed, const double* __restrict__ aux) {
  const uint64_t rows = (uint64_t)n * n * n_slabs;
  const double soft = family == SCFB_ERI_CHAIN ? aux[0] : 0.0;
  const double* x = aux + 1;
  const double* S = aux + 1 + n;
  for (uint64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row % n);
    const int c = (int)((row / n) % n);
    const int d = nu_lo + (int)(row / ((uint64_t)n * n));
    double* dst = eri + row * n;
    if (family == SCFB_ERI_HASH) {
      const uint64_t p2 = pair_index(c, d);
      for (int a = threadIdx.x; a < n; a += blockDim.x)
        dst[a] = hash_value(pair_index(a, b), p2, seed);
    } else {
      const double s_cd = S[c + (size_t)n * d];
      const double r_cd = __dmul_rn(0.5, __dadd_rn(x[c], x[d]));
      for (int a = threadIdx.x; a < n; a += blockDim.x) {
        const double r_ab = __dmul_rn(0.5, __dadd_rn(x[a], x[b]));
        dst[a] = chain_value(S[a + (size_t)n * b], r_ab, s_cd, r_cd, soft);
      }
    }
  }
}

}  // namespace

int scfb_launch_generate_full(scfb_handle_s* h, int family, uint64_t seed, const double* d_aux) {
  const int n_slabs = h->nu_hi - h->nu_lo;
  if (n_slabs <= 0) return SCFB_OK;
  const uint64_t rows = (uint64_t)h->n * h->n * n_slabs;
  const int threads = h->n >= 256 ? 256 : ((h->n + 31) / 32) * 32;
  uint64_t blocks = rows < 148ull * 64 ? rows : 148ull * 64;
  generate_full_kernel<<<(unsigned)blocks, threads, 0, h->stream>>>(h->eri, h->n, h->nu_lo, n_slabs,
                                                                   family, seed, d_aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "generate_full_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}
