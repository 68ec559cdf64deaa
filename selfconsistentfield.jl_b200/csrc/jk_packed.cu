// K2 — single-pass J+K over the 8-fold-symmetry-packed ERI (sm_100a).
//
// Storage (this library's own HBM layout, "packed8"): pair index pr(i,j) = i(i+1)/2 + j for
// i >= j, P = n(n+1)/2 pairs; element (p,q), p >= q, at p(p+1)/2 + q (row p contiguous,
// p+1 entries).  Stored value f = (ij|kl) * 0.5^{[i=j] + [k=l] + [p=q]} (exact power-of-two
// pre-scaling), so that applying all 8 index permutations to every stored element gives each
// distinct tensor element weight 1.  Only valid for 8-fold symmetric
// tensors and SYMMETRIC densities (precondition documented at the ABI); under it the
// reference's literal contraction (/root/reference/src/algorithms/JKBuilderFromTensor.jl:37)
// equals  J_ab = sum_cd (ab|cd) Dtot_cd,  K_ac = sum_bd (ab|cd) D_bd.
//
// Per stored element (i,j,k,l; f), with Dq2 = 2*Dtot in pair space:
//   J~[p] += f Dq2[q]          J~[q] += f Dq2[p]       (J_ij = J_ji = J~[pr(i,j)], J_ii = 2 J~[pr(i,i)])
//   K'[i,k] += f D[j,l]   K'[i,l] += f D[j,k]   K'[j,k] += f D[i,l]   K'[j,l] += f D[i,k]
// and K = K' + K'^T.  2 + 4 n_spin FMAs per 8 bytes: still HBM-bound (1.5-2.5 flop/B).
//
// Mapping: work item = (i, chunk of KC consecutive k <= i); lanes run along l (coalesced
// 64-bit loads, warps with l > k exit), the CTA loops over the rows j = 0..i, each of which
// contributes one contiguous run of the packed row p = pr(i,j).  Sums whose output does not
// depend on j (K'[i,k], K'[i,l], J~[q]) stay in registers for the whole item; per-row outputs
// use a warp transpose-reduce (K'[j,k], J~[p]) or one RED per lane (K'[j,l]).  Cross-warp /
// cross-CTA combination is fp64 RED.ADD into zeroed accumulators, so the packed path is
// deterministic only up to summation order (well inside the 1e-12 relative gate).
#include "scfb_internal.h"

namespace {

constexpr int KC = 8;

__host__ __device__ __forceinline__ uint64_t tri(uint64_t i) { return i * (i + 1) / 2; }

__device__ __forceinline__ double ldg_stream(const double* p) {
  double r;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// Sum 8 per-lane values over the 32 lanes of a warp with 9 (instead of 40) shuffles.
// On return lane L with (L & 3) == 0 holds the total of value index (L >> 2) in `r`.
__device__ __forceinline__ double warp_transpose_reduce8(const double (&v)[8], int lane) {
  double a[4], b[2], c;
  const bool hi16 = lane & 16;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double keep = hi16 ? v[t + 4] : v[t];
    const double send = hi16 ? v[t] : v[t + 4];
    a[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool hi8 = lane & 8;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double keep = hi8 ? a[t + 2] : a[t];
    const double send = hi8 ? a[t] : a[t + 2];
    b[t] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const bool hi4 = lane & 4;
  {
    const double keep = hi4 ? b[1] : b[0];
    const double send = hi4 ? b[0] : b[1];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;  // value index = 4*bit4 + 2*bit3 + bit2 of lane
}
__device__ __forceinline__ int transpose_reduce8_index(int lane) {
  return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// grid = (n_kchunks, n_i_owned); block = roundup32(n) threads (<= 512).
// D: n x n x NSPIN symmetric; dq2: packed 2*Dtot (P entries).
// Kp: n x n x NSPIN accumulators K'[a + n*b]; Jp: P accumulators.
template <int NSPIN>
__global__ void __launch_bounds__(512, 1)
jk_packed_kernel(const double* __restrict__ eri, uint64_t shard_offset,
                 const double* __restrict__ D, const double* __restrict__ dq2,
                 double* __restrict__ Kp, double* __restrict__ Jp, int n, int i_lo) {
  const int i = i_lo + blockIdx.y;
  const int k0 = blockIdx.x * KC;
  if (k0 > i) return;
  const int l = threadIdx.x;
  const int lane = l & 31;
  const int kmax = min(k0 + KC - 1, i);
  if ((l & ~31) > kmax) return;  // whole warp beyond the triangle: nothing to do
  const bool l_ok = l <= kmax;   // lanes past kmax only take part in the shuffles
  const size_t n2 = (size_t)n * n;
  const int lc = l_ok ? l : 0;   // clamped index for loads of idle lanes

  // per-item constants
  double di_l[NSPIN], di_k[NSPIN][KC], dq[KC];
  uint64_t off[KC];
  bool ok[KC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) di_l[s] = __ldg(D + s * n2 + (size_t)i * n + lc);
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    const int k = k0 + t;
    ok[t] = l_ok && k <= i && l <= k;
    const int kc = k <= i ? k : i;
    off[t] = tri(kc) + lc;
    dq[t] = ok[t] ? __ldg(dq2 + off[t]) : 0.0;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) di_k[s][t] = __ldg(D + s * n2 + (size_t)i * n + kc);
  }

  double a1[NSPIN][KC], a2[NSPIN], jq[KC];
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    jq[t] = 0.0;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) a1[s][t] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) a2[s] = 0.0;

  const uint64_t p0 = tri(i);
  for (int j = 0; j <= i; ++j) {
    const uint64_t p = p0 + j;
    const double* row = eri + (tri(p) - shard_offset);
    double w[KC];
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      // q <= p: inside the diagonal block (k == i) only l <= j is stored
      const bool valid = ok[t] && (k0 + t < i || l <= j);
      w[t] = valid ? ldg_stream(row + off[t]) : 0.0;
    }
    const double dp2 = __ldg(dq2 + p);
    double dj_l[NSPIN], dj_k[NSPIN][KC];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      dj_l[s] = __ldg(D + s * n2 + (size_t)j * n + lc);
#pragma unroll
      for (int t = 0; t < KC; ++t)
        dj_k[s][t] = __ldg(D + s * n2 + (size_t)j * n + min(k0 + t, i));
    }
    double jp = 0.0;
    double y[NSPIN][KC], z[NSPIN];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) z[s] = 0.0;
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      jp = fma(w[t], dq[t], jp);
      jq[t] = fma(w[t], dp2, jq[t]);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        a1[s][t] = fma(w[t], dj_l[s], a1[s][t]);
        a2[s] = fma(w[t], dj_k[s][t], a2[s]);
        y[s][t] = w[t] * di_l[s];
        z[s] = fma(w[t], di_k[s][t], z[s]);
      }
    }
    // per-row outputs
    jp = warp_sum(jp);
    if (lane == 0) atomicAdd(Jp + p, jp);
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      if (l_ok) atomicAdd(Kp + s * n2 + j + (size_t)n * l, z[s]);  // K'[j,l]
      const double r = warp_transpose_reduce8(y[s], lane);
      const int t = transpose_reduce8_index(lane);
      if ((lane & 3) == 0 && k0 + t <= i) atomicAdd(Kp + s * n2 + j + (size_t)n * (k0 + t), r);  // K'[j,k]
    }
  }

  // per-item outputs
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    if (l_ok) atomicAdd(Kp + s * n2 + i + (size_t)n * l, a2[s]);  // K'[i,l]
    const double r = warp_transpose_reduce8(a1[s], lane);
    const int t = transpose_reduce8_index(lane);
    if ((lane & 3) == 0 && k0 + t <= i) atomicAdd(Kp + s * n2 + i + (size_t)n * (k0 + t), r);  // K'[i,k]
  }
#pragma unroll
  for (int t = 0; t < KC; ++t)
    if (ok[t]) atomicAdd(Jp + off[t], jq[t]);  // J~[q]
}

// dq2[pr(k,l)] = 2 * Dtot[k,l]
__global__ void pack_density_kernel(const double* __restrict__ dtot, double* __restrict__ dq2, int n) {
  const uint64_t P = tri(n);
  for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < P;
       q += (uint64_t)gridDim.x * blockDim.x) {
    // k = floor((sqrt(8q+1)-1)/2), fixed up for rounding
    uint64_t k = (uint64_t)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
    while (tri(k + 1) <= q) ++k;
    while (tri(k) > q) --k;
    const uint64_t l = q - tri(k);
    dq2[q] = 2.0 * dtot[k + (size_t)n * l];
  }
}

// jk = [J | K_0 | K_1] from the accumulators: J[a,b] = J~[pr(a,b)], K = K' + K'^T
__global__ void unpack_jk_kernel(const double* __restrict__ Jp, const double* __restrict__ Kp,
                                 double* __restrict__ jk, int n, int n_spin) {
  const size_t n2 = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t a = e % n, b = e / n;
    const size_t hi = a > b ? a : b, lo = a > b ? b : a;
    // a diagonal pair (a,a) is the target of both the (ab|..) and (ba|..) permutations
    jk[e] = (a == b ? 2.0 : 1.0) * Jp[tri(hi) + lo];
    for (int s = 0; s < n_spin; ++s) jk[(1 + s) * n2 + e] = Kp[s * n2 + e] + Kp[s * n2 + b + n * a];
  }
}

// synthetic families written directly in packed, pre-scaled form
__device__ __forceinline__ double hash_value(uint64_t p, uint64_t q, uint64_t seed) {
  uint64_t z = tri(p) + q + (seed + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return __dadd_rn(__dmul_rn((double)(z >> 11), 0x1.0p-52), -1.0);
}

__global__ void generate_packed_kernel(double* __restrict__ eri, uint64_t shard_offset, int n,
                                       int i_lo, int i_hi, int family, uint64_t seed,
                                       const double* __restrict__ aux) {
  const double soft = family == SCFB_ERI_CHAIN ? aux[0] : 0.0;
  const double* x = aux + 1;
  const double* S = aux + 1 + n;
  const uint64_t p_lo = tri(i_lo), p_hi = tri(i_hi);
  for (uint64_t p = p_lo + blockIdx.x; p < p_hi; p += gridDim.x) {
    uint64_t i = (uint64_t)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tri(i + 1) <= p) ++i;
    while (tri(i) > p) --i;
    const uint64_t j = p - tri(i);
    double* row = eri + (tri(p) - shard_offset);
    uint64_t k = 0, kbase = 0;  // walk (k,l) incrementally per thread stride
    for (uint64_t q = threadIdx.x; q <= p; q += blockDim.x) {
      while (kbase + k + 1 <= q) {
        kbase += k + 1;
        ++k;
      }
      const uint64_t l = q - kbase;
      double v;
      if (family == SCFB_ERI_HASH) {
        v = hash_value(p, q, seed);
      } else {
        const double r_ab = __dmul_rn(0.5, __dadd_rn(x[i], x[j]));
        const double r_cd = __dmul_rn(0.5, __dadd_rn(x[k], x[l]));
        const double d = __dadd_rn(r_ab, -r_cd);
        const double den = __dsqrt_rn(__dadd_rn(soft, __dmul_rn(d, d)));
        v = __ddiv_rn(__dmul_rn(S[i + (size_t)n * j], S[k + (size_t)n * l]), den);
        if (fabs(v) < 1e-200) v = 0.0;
      }
      double scale = 1.0;
      if (i == j) scale *= 0.5;
      if (k == l) scale *= 0.5;
      if (p == q) scale *= 0.5;
      row[q] = v * scale;
    }
  }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------
// i-blocks owned by `rank`: contiguous ranges balanced by stored-element count (~ i^4 / 8).
void scfb_packed_split(int n, int rank, int n_ranks, int* i_lo, int* i_hi) {
  auto elems_below = [](uint64_t i) {  // stored elements of all rows with first index < i
    const uint64_t P = tri(i);
    return tri(P);
  };
  const uint64_t total = elems_below(n);
  auto bound = [&](int g) -> int {
    if (g <= 0) return 0;
    if (g >= n_ranks) return n;
    const long double target = (long double)total * g / n_ranks;
    int lo = 0, hi = n;
    while (lo < hi) {  // smallest i with elems_below(i) >= target
      const int mid = (lo + hi) / 2;
      if ((long double)elems_below(mid) >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  *i_lo = bound(rank);
  *i_hi = bound(rank + 1);
}

uint64_t scfb_packed_offset(int i) { return tri(tri((uint64_t)i)); }

int scfb_launch_jk_packed(scfb_handle_s* h, int n_spin, const double* d_density,
                          const double* d_total) {
  const int n = h->n;
  const size_t n2 = (size_t)n * n;
  const uint64_t P = tri(n);
  cudaError_t e;
  // zero the accumulators: [Jp (P) | Kp (2 n^2)]
  e = cudaMemsetAsync(h->pk_acc, 0, (P + 2 * n2) * sizeof(double), h->stream);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  pack_density_kernel<<<148 * 2, 256, 0, h->stream>>>(d_total, h->pk_dq2, n);
  h->launches += 1;
  if (h->timed) cudaEventRecord(h->ev[3], h->stream);
  const int n_i = h->i_hi - h->i_lo;
  if (n_i > 0) {
    dim3 grid((n + KC - 1) / KC, n_i);
    const int threads = ((n + 31) / 32) * 32;
    double* Jp = h->pk_acc;
    double* Kp = h->pk_acc + P;
    if (n_spin == 1)
      jk_packed_kernel<1><<<grid, threads, 0, h->stream>>>(h->eri, h->pk_offset, d_density, h->pk_dq2,
                                                         Kp, Jp, n, h->i_lo);
    else
      jk_packed_kernel<2><<<grid, threads, 0, h->stream>>>(h->eri, h->pk_offset, d_density, h->pk_dq2,
                                                         Kp, Jp, n, h->i_lo);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return scfb_fail(h, SCFB_ERR_CUDA, "jk_packed_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
  }
  if (h->timed) cudaEventRecord(h->ev[1], h->stream);
  unpack_jk_kernel<<<148 * 2, 256, 0, h->stream>>>(h->pk_acc, h->pk_acc + P, h->jk, n, n_spin);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "unpack_jk_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

int scfb_launch_generate_packed(scfb_handle_s* h, int family, uint64_t seed, const double* d_aux) {
  if (h->i_hi <= h->i_lo) return SCFB_OK;
  generate_packed_kernel<<<148 * 16, 256, 0, h->stream>>>(h->eri, h->pk_offset, h->n, h->i_lo, h->i_hi,
                                                         family, seed, d_aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "generate_packed_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

// Host-side packing of this rank's rows from a full column-major tensor (small n: the full
// tensor exists on the host anyway).  Representative of each orbit: eri[i,j,k,l] with
// i>=j, k>=l, pr(i,j) >= pr(k,l).
void scfb_pack_host(const double* full, int n, int i_lo, int i_hi, double* out) {
  const uint64_t N = n;
  uint64_t pos = 0;
  for (uint64_t i = i_lo; i < (uint64_t)i_hi; ++i)
    for (uint64_t j = 0; j <= i; ++j) {
      const uint64_t p = tri(i) + j;
      for (uint64_t k = 0; k <= i; ++k)
        for (uint64_t l = 0; l <= k; ++l) {
          const uint64_t q = tri(k) + l;
          if (q > p) break;
          double v = full[i + N * (j + N * (k + N * l))];
          if (i == j) v *= 0.5;
          if (k == l) v *= 0.5;
          if (p == q) v *= 0.5;
          out[pos++] = v;
        }
    }
}
