// K2 — single-pass J+K over the 8-fold-symmetry-packed ERI (sm_100a).
//
// Storage (this library's own HBM layout, "packed8"): pair index pr(i,j) = i(i+1)/2 + j,
// i >= j.  Row p = pr(i,j) holds the canonical elements (ij|kl) with pr(k,l) <= p, grouped
// in k-segments (l = 0..k; the last segment k = i stops at l = j).  Every segment is padded
// with one zero to EVEN length so that all segment and row starts are 16-byte aligned:
//   E(k)   = k(k+1)/2 + (k+1)/2        padded elements of segments 0..k-1
//   R(i,j) = B[i] + j*E(i) + E(j)      offset of row (i,j);  B[i+1] = B[i] + (i+1)E(i) + E(i+1)
// (padding overhead ~0.5 % at n = 384; roofline figures use the UNPADDED byte count).
// Stored value f = (ij|kl) * 0.5^{[i=j] + [k=l] + [p=q]} — exact power-of-two pre-scaling, so
// that applying all 8 index permutations to every stored element gives each distinct tensor
// element weight 1.  Valid only for 8-fold symmetric tensors and
// SYMMETRIC densities (precondition documented at the ABI); under it the reference's literal
// contraction (/root/reference/src/algorithms/JKBuilderFromTensor.jl:37) equals
//   J_ab = sum_cd (ab|cd) Dtot_cd,   K_ac = sum_bd (ab|cd) D_bd.
//
// Per stored element (i,j,k,l; f), with Dq2 = 2*Dtot in pair space:
//   J~[p] += f Dq2[q]      J~[q] += f Dq2[p]      (J_ij = J_ji = J~[pr(i,j)], J_ii = 2 J~[pr(i,i)])
//   K'[i,k] += f D[j,l]   K'[i,l] += f D[j,k]   K'[j,k] += f D[i,l]   K'[j,l] += f D[i,k]
// and K = K' + K'^T.  2 + 4 n_spin FMAs per 8 bytes: still HBM-bound (1.5-2.5 flop/B).
//
// Mapping: a warp owns (i, chunk of KC=8 consecutive k <= i, 64 consecutive l); each lane
// owns an l-PAIR (one 128-bit load per k and row) and the warp loops over the rows j = 0..i,
// each of which contributes one contiguous run of the packed row.  Sums whose output does not
// depend on j (K'[i,k], K'[i,l], J~[q]) stay in registers for the whole item; per-row outputs
// use a warp transpose-reduce (K'[j,k], J~[p]) or one RED per lane (K'[j,l]).  Warps are
// fully independent (no CTA barrier); cross-warp combination is fp64 RED.ADD into zeroed
// accumulators, so the packed path is deterministic only up to summation order (well inside
// the 1e-12 relative gate).
#include "scfb_internal.h"

#include <vector>

namespace {

constexpr int KC = 8;    // k values per work item
constexpr int LB = 128;  // threads per CTA = 4 independent warps, 256 l values

__host__ __device__ __forceinline__ uint64_t tri(uint64_t i) { return i * (i + 1) / 2; }
__host__ __device__ __forceinline__ uint64_t seg_off(uint64_t k) { return tri(k) + (k + 1) / 2; }

__device__ __forceinline__ double2 ldg_stream2(const double* p) {
  double2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// Sum 8 per-lane values over the 32 lanes of a warp with 9 (instead of 40) shuffles.
// On return lanes with (lane & 3) == 0 hold the total of value index
// 4*bit4 + 2*bit3 + bit2 of their lane id.
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
  return c;
}
__device__ __forceinline__ int transpose_reduce8_index(int lane) {
  return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// One work item of one warp.  D: n x n x NSPIN symmetric; dq2: 2*Dtot in pair space, stored
// with the SAME padded segment offsets seg_off(k)+l (pads zero).  Kp: n x n x NSPIN
// accumulators; Jp: accumulators indexed like dq2.
template <int NSPIN, bool DIAG>
__device__ __forceinline__ void packed_item(const double* __restrict__ eri_i,  // row (i,0)
                                            const double* __restrict__ D,
                                            const double* __restrict__ dq2,
                                            double* __restrict__ Kp, double* __restrict__ Jp, int n,
                                            int i, int k0, int l0, int lane, uint32_t rot) {
  const size_t n2 = (size_t)n * n;
  const int kmax = min(k0 + KC - 1, i);
  // lanes past the triangle edge (l > kmax, possibly l >= n) only feed zeros to the shuffles
  const bool l0_in = l0 <= kmax, l1_in = l0 + 1 <= kmax;
  const int lc = l0_in ? l0 : 0;
  const uint64_t Ei = seg_off(i);  // elements of the full segments 0..i-1 of every row (i,*)

  // per-item constants
  double di_l[NSPIN][2], di_k[NSPIN][KC], dq[KC][2];
  uint32_t off[KC];
  bool ok[KC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    const double* di = D + s * n2 + (size_t)i * n;
    di_l[s][0] = __ldg(di + lc);
    di_l[s][1] = l1_in ? __ldg(di + lc + 1) : 0.0;
  }
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    const int k = k0 + t;
    ok[t] = k <= i && l0 <= k;
    const int kc = k <= i ? k : i;
    off[t] = (uint32_t)seg_off(kc) + (uint32_t)lc;
    const double2 q2 =
        ok[t] ? __ldg(reinterpret_cast<const double2*>(dq2 + off[t])) : make_double2(0.0, 0.0);
    dq[t][0] = q2.x;
    dq[t][1] = q2.y;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) di_k[s][t] = __ldg(D + s * n2 + (size_t)i * n + kc);
  }

  double a1[NSPIN][KC], a2[NSPIN][2], jq[KC][2];
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    jq[t][0] = jq[t][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) a1[s][t] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) a2[s][0] = a2[s][1] = 0.0;

  const int rows = i + 1;
  auto load_row = [&](const double* row, int j, double2 (&w)[KC]) {
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      // the diagonal segment k == i of row (i,j) stops at l = j (padded to even)
      const bool valid = DIAG ? (ok[t] && (k0 + t < i || l0 <= j)) : ok[t];
      w[t] = valid ? ldg_stream2(row + off[t]) : make_double2(0.0, 0.0);
    }
  };

  // Rows are visited in a rotated order that differs per work item, so that concurrently
  // running warps do not all RED into the same K'[j,:] row at the same time.
  int j = (int)(rot % (uint32_t)rows);
  const double* row = eri_i + (uint64_t)j * Ei + seg_off(j);
  double2 w[KC], wn[KC];
  load_row(row, j, w);
  for (int step = 0; step < rows; ++step) {
    const int jn = j + 1 == rows ? 0 : j + 1;
    // row (i,j+1) starts E(i) + ceil2(j+1) elements after row (i,j)
    const double* row_n = jn ? row + Ei + ((j + 2) & ~1) : eri_i;
    if (step + 1 < rows) load_row(row_n, jn, wn);  // next row in flight while this one is consumed

    const double dp2 = __ldg(dq2 + Ei + j);
    double dj_l[NSPIN][2], dj_k[NSPIN][KC];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const double* dj = D + s * n2 + (size_t)j * n;
      dj_l[s][0] = __ldg(dj + lc);
      dj_l[s][1] = l1_in ? __ldg(dj + lc + 1) : 0.0;
#pragma unroll
      for (int t = 0; t < KC; ++t) dj_k[s][t] = __ldg(dj + min(k0 + t, i));
    }
    double jp = 0.0;
    double y[NSPIN][KC], z[NSPIN][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) z[s][0] = z[s][1] = 0.0;
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const double w0 = w[t].x, w1 = w[t].y;
      jp = fma(w0, dq[t][0], jp);
      jp = fma(w1, dq[t][1], jp);
      jq[t][0] = fma(w0, dp2, jq[t][0]);
      jq[t][1] = fma(w1, dp2, jq[t][1]);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        a1[s][t] = fma(w0, dj_l[s][0], a1[s][t]);
        a1[s][t] = fma(w1, dj_l[s][1], a1[s][t]);
        a2[s][0] = fma(w0, dj_k[s][t], a2[s][0]);
        a2[s][1] = fma(w1, dj_k[s][t], a2[s][1]);
        y[s][t] = fma(w1, di_l[s][1], w0 * di_l[s][0]);
        z[s][0] = fma(w0, di_k[s][t], z[s][0]);
        z[s][1] = fma(w1, di_k[s][t], z[s][1]);
      }
    }
    // per-row outputs (K' entries are accumulated at whichever of (a,b)/(b,a) coalesces:
    // K = K' + K'^T does not care)
    jp = warp_sum(jp);
    if (lane == 0) atomicAdd(Jp + Ei + j, jp);
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      double* kj = Kp + s * n2 + (size_t)n * j;
      if (l0_in) atomicAdd(kj + l0, z[s][0]);  // K'[j,l]
      if (l1_in) atomicAdd(kj + l0 + 1, z[s][1]);
      const double r = warp_transpose_reduce8(y[s], lane);
      const int t = transpose_reduce8_index(lane);
      if ((lane & 3) == 0 && k0 + t <= i) atomicAdd(kj + k0 + t, r);  // K'[j,k]
    }
#pragma unroll
    for (int t = 0; t < KC; ++t) w[t] = wn[t];
    row = row_n;
    j = jn;
  }

  // per-item outputs
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    double* ki = Kp + s * n2 + (size_t)n * i;
    if (l0_in) atomicAdd(ki + l0, a2[s][0]);  // K'[i,l]
    if (l1_in) atomicAdd(ki + l0 + 1, a2[s][1]);
    const double r = warp_transpose_reduce8(a1[s], lane);
    const int t = transpose_reduce8_index(lane);
    if ((lane & 3) == 0 && k0 + t <= i) atomicAdd(ki + k0 + t, r);  // K'[i,k]
  }
#pragma unroll
  for (int t = 0; t < KC; ++t)
    if (ok[t]) {  // J~[q]; the pad slot only ever receives zeros
      atomicAdd(Jp + off[t], jq[t][0]);
      atomicAdd(Jp + off[t] + 1, jq[t][1]);
    }
}

// grid = (n_kchunks, n_i_owned, n_lblocks), block = LB threads = 4 independent warps.
template <int NSPIN>
__global__ void __launch_bounds__(LB, 2)
jk_packed_kernel(const double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                 uint64_t shard_offset, const double* __restrict__ D,
                 const double* __restrict__ dq2, double* __restrict__ Kp, double* __restrict__ Jp,
                 int n, int i_lo) {
  const int i = i_lo + blockIdx.y;
  const int k0 = blockIdx.x * KC;
  if (k0 > i) return;
  const int m = blockIdx.z * LB + threadIdx.x;  // l-pair index
  const int lane = threadIdx.x & 31;
  const int kmax = min(k0 + KC - 1, i);
  if (2 * (m & ~31) > kmax) return;  // whole warp beyond the triangle
  const int l0 = 2 * m;
  const double* eri_i = eri + (Btab[i] - shard_offset);
  const uint32_t rot =
      blockIdx.x * 40503u + blockIdx.z * 9973u + blockIdx.y * 17u + (threadIdx.x >> 5) * 7u;
  if (kmax < i)
    packed_item<NSPIN, false>(eri_i, D, dq2, Kp, Jp, n, i, k0, l0, lane, rot);
  else
    packed_item<NSPIN, true>(eri_i, D, dq2, Kp, Jp, n, i, k0, l0, lane, rot);
}

// dq2[seg_off(k) + l] = 2 * Dtot[k,l] for l <= k, pads = 0
__global__ void pack_density_kernel(const double* __restrict__ dtot, double* __restrict__ dq2, int n) {
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    const uint64_t base = seg_off(k);
    const int len = (k + 2) & ~1;
    for (int l = threadIdx.x; l < len; l += blockDim.x)
      dq2[base + l] = l <= k ? 2.0 * dtot[k + (size_t)n * l] : 0.0;
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
    jk[e] = (a == b ? 2.0 : 1.0) * Jp[seg_off(hi) + lo];
    for (int s = 0; s < n_spin; ++s) jk[(1 + s) * n2 + e] = Kp[s * n2 + e] + Kp[s * n2 + b + n * a];
  }
}

// synthetic families written directly in packed, pre-scaled, padded form
__device__ __forceinline__ double hash_value(uint64_t p, uint64_t q, uint64_t seed) {
  uint64_t z = tri(p) + q + (seed + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return __dadd_rn(__dmul_rn((double)(z >> 11), 0x1.0p-52), -1.0);
}

// one CTA per row (i,j) at a time; threads walk the padded row
__global__ void generate_packed_kernel(double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                                       uint64_t shard_offset, int n, int i_lo, int i_hi, int family,
                                       uint64_t seed, const double* __restrict__ aux) {
  const double soft = family == SCFB_ERI_CHAIN ? aux[0] : 0.0;
  const double* x = aux + 1;
  const double* S = aux + 1 + n;
  const uint64_t p_lo = tri(i_lo), p_hi = tri(i_hi);
  for (uint64_t p = p_lo + blockIdx.x; p < p_hi; p += gridDim.x) {
    uint64_t i = (uint64_t)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tri(i + 1) <= p) ++i;
    while (tri(i) > p) --i;
    const uint64_t j = p - tri(i);
    double* row = eri + (Btab[i] + j * seg_off(i) + seg_off(j) - shard_offset);
    const uint64_t len = seg_off(i) + ((j + 2) & ~1ull);
    uint64_t k = 0;  // segment containing the thread's position (walks forward only)
    for (uint64_t e = threadIdx.x; e < len; e += blockDim.x) {
      while (seg_off(k + 1) <= e) ++k;
      const uint64_t l = e - seg_off(k);
      const uint64_t lmax = k < i ? k : j;
      double v = 0.0;
      if (l <= lmax) {
        const uint64_t q = tri(k) + l;
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
        if (i == j) v *= 0.5;
        if (k == l) v *= 0.5;
        if (p == q) v *= 0.5;
      }
      row[e] = v;
    }
  }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------
uint64_t scfb_packed_offset(int i) {  // B[i]: padded elements of all rows with first index < i
  uint64_t b = 0;
  for (uint64_t ii = 0; ii < (uint64_t)i; ++ii) b += (ii + 1) * seg_off(ii) + seg_off(ii + 1);
  return b;
}

uint64_t scfb_packed_dq_elems(int n) { return seg_off(n); }

// i-blocks owned by `rank`: contiguous ranges balanced by stored-element count (~ i^4 / 8).
void scfb_packed_split(int n, int rank, int n_ranks, int* i_lo, int* i_hi) {
  std::vector<uint64_t> B(n + 1, 0);
  for (int i = 0; i < n; ++i) B[i + 1] = B[i] + (uint64_t)(i + 1) * seg_off(i) + seg_off(i + 1);
  auto bound = [&](int g) -> int {
    if (g <= 0) return 0;
    if (g >= n_ranks) return n;
    const long double target = (long double)B[n] * g / n_ranks;
    int lo = 0, hi = n;
    while (lo < hi) {  // smallest i with B[i] >= target
      const int mid = (lo + hi) / 2;
      if ((long double)B[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  *i_lo = bound(rank);
  *i_hi = bound(rank + 1);
}

int scfb_packed_upload_table(scfb_handle_s* h) {
  std::vector<uint64_t> B(h->n + 1, 0);
  for (int i = 0; i < h->n; ++i) B[i + 1] = B[i] + (uint64_t)(i + 1) * seg_off(i) + seg_off(i + 1);
  cudaError_t e = cudaMalloc(&h->pk_btab, (h->n + 1) * sizeof(uint64_t));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->pk_btab, B.data(), (h->n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess)
    return scfb_fail(h, e == cudaErrorMemoryAllocation ? SCFB_ERR_OOM : SCFB_ERR_CUDA,
                     "packed row table: %s", cudaGetErrorString(e));
  return SCFB_OK;
}

int scfb_launch_jk_packed(scfb_handle_s* h, int n_spin, const double* d_density,
                          const double* d_total) {
  const int n = h->n;
  const size_t n2 = (size_t)n * n;
  const uint64_t PQ = seg_off(n);
  cudaError_t e;
  // zero the accumulators: [J~ (PQ) | K' (2 n^2)]
  e = cudaMemsetAsync(h->pk_acc, 0, (PQ + 2 * n2) * sizeof(double), h->stream);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  pack_density_kernel<<<n < 296 ? n : 296, 128, 0, h->stream>>>(d_total, h->pk_dq2, n);
  h->launches += 1;
  if (h->timed) cudaEventRecord(h->ev[3], h->stream);
  const int n_i = h->i_hi - h->i_lo;
  if (n_i > 0) {
    dim3 grid((n + KC - 1) / KC, n_i, (n + 2 * LB - 1) / (2 * LB));
    double* Jp = h->pk_acc;
    double* Kp = h->pk_acc + PQ;
    if (n_spin == 1)
      jk_packed_kernel<1><<<grid, LB, 0, h->stream>>>(h->eri, h->pk_btab, h->pk_offset, d_density,
                                                    h->pk_dq2, Kp, Jp, n, h->i_lo);
    else
      jk_packed_kernel<2><<<grid, LB, 0, h->stream>>>(h->eri, h->pk_btab, h->pk_offset, d_density,
                                                    h->pk_dq2, Kp, Jp, n, h->i_lo);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return scfb_fail(h, SCFB_ERR_CUDA, "jk_packed_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
  }
  if (h->timed) cudaEventRecord(h->ev[1], h->stream);
  unpack_jk_kernel<<<148 * 2, 256, 0, h->stream>>>(h->pk_acc, h->pk_acc + PQ, h->jk, n, n_spin);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "unpack_jk_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

int scfb_launch_generate_packed(scfb_handle_s* h, int family, uint64_t seed, const double* d_aux) {
  if (h->i_hi <= h->i_lo) return SCFB_OK;
  generate_packed_kernel<<<148 * 16, 256, 0, h->stream>>>(h->eri, h->pk_btab, h->pk_offset, h->n,
                                                         h->i_lo, h->i_hi, family, seed, d_aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "generate_packed_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

// Host-side packing of this rank's rows from a full column-major tensor (small n: the full
// tensor exists on the host anyway).  Representative of each orbit: eri[i,j,k,l] with
// i>=j, k>=l, pr(i,j) >= pr(k,l).  `out` must hold B[i_hi] - B[i_lo] doubles.
void scfb_pack_host(const double* full, int n, int i_lo, int i_hi, double* out) {
  const uint64_t N = n;
  uint64_t pos = 0;
  for (uint64_t i = i_lo; i < (uint64_t)i_hi; ++i)
    for (uint64_t j = 0; j <= i; ++j) {
      const uint64_t p = tri(i) + j;
      for (uint64_t k = 0; k <= i; ++k) {
        const uint64_t lmax = k < i ? k : j;
        for (uint64_t l = 0; l <= lmax; ++l) {
          const uint64_t q = tri(k) + l;
          double v = full[i + N * (j + N * (k + N * l))];
          if (i == j) v *= 0.5;
          if (k == l) v *= 0.5;
          if (p == q) v *= 0.5;
          out[pos++] = v;
        }
        if ((lmax + 1) & 1) out[pos++] = 0.0;  // pad the segment to even length
      }
    }
}
