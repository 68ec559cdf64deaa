// K1 — single-pass J+K stream over the full-layout ERI shard (sm_100a).
//
// Replaces the two @tensor contractions of
//   /root/reference/src/algorithms/JKBuilderFromTensor.jl:36-38 (RHF) and :43-47 (UHF)
// with ONE read of every ERI element for J and for K of every spin block.
//
// Index algebra (Julia layout, eri[a,b,c,v] at a + n b + n^2 c + n^3 v):
//   J[c,v]   = sum_{a,b} eri[a,b,c,v] * Dtot[a,b]        (mu=c)
//   K_s[a,v] = sum_{b,c} eri[a,b,c,v] * D_s[c,b]         (mu=a, beta=b, alpha=c)
// Work item = (nu-slab v, chunk of TC consecutive c).  Threads lie along a (coalesced,
// 128-bit loads when n is even) x G groups along b.  Per loaded element: 1 FMA into the
// J accumulator of its c, n_spin FMAs into the thread-private K accumulators of its a.
// J[c,v] is completed inside the CTA (warp shuffles + one smem hop); K rows are written
// as per-chunk partials and summed in fixed order by `k_reduce_kernel`, so results are
// bit-reproducible run to run.  HBM-bound: 8*n^4 algorithmic bytes per build; D, Dtot
// (8 n^2 B each) are L2/L1 resident and re-read through the read-only path.
#include "scfb_internal.h"

namespace {

constexpr int TC = 8;          // c values per work item
constexpr int MAX_THREADS = 256;

__device__ __forceinline__ double2 ldg_stream_v2(const double* p) {
  double2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ldg_stream_v1(const double* p) {
  double r;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

template <int VEC>
struct Vec;
template <>
struct Vec<2> {
  double v[2];
  __device__ __forceinline__ void load_stream(const double* p) {
    double2 t = ldg_stream_v2(p);
    v[0] = t.x;
    v[1] = t.y;
  }
  __device__ __forceinline__ void load_cached(const double* p) {
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  }
};
template <>
struct Vec<1> {
  double v[1];
  __device__ __forceinline__ void load_stream(const double* p) { v[0] = ldg_stream_v1(p); }
  __device__ __forceinline__ void load_cached(const double* p) { v[0] = __ldg(p); }
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// One b-row worth of work for one thread: TC streamed vectors against Dtot[a..,b] and
// D_s[c0..c0+TC, b].
template <int VEC, int NSPIN, bool FULL>
__device__ __forceinline__ void consume_row(const Vec<VEC> (&x)[TC], const Vec<VEC>& dt,
                                            const double* __restrict__ dens, size_t n2,
                                            size_t dsoff, int tcn, double (&jacc)[TC],
                                            double (&kacc)[NSPIN][VEC]) {
  double ds[NSPIN][TC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    const double* dp = dens + s * n2 + dsoff;
    if (VEC == 2 && FULL) {
#pragma unroll
      for (int t = 0; t < TC; t += 2) {
        double2 d2 = __ldg(reinterpret_cast<const double2*>(dp + t));
        ds[s][t] = d2.x;
        ds[s][t + 1] = d2.y;
      }
    } else {
#pragma unroll
      for (int t = 0; t < TC; ++t) ds[s][t] = (FULL || t < tcn) ? __ldg(dp + t) : 0.0;
    }
  }
#pragma unroll
  for (int t = 0; t < TC; ++t) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      jacc[t] = fma(x[t].v[v], dt.v[v], jacc[t]);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) kacc[s][v] = fma(x[t].v[v], ds[s][t], kacc[s][v]);
    }
  }
}

template <int VEC, bool FULL>
__device__ __forceinline__ void load_row(Vec<VEC> (&x)[TC], const double* __restrict__ xp,
                                         size_t n2, int tcn) {
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    if (FULL || t < tcn) {
      x[t].load_stream(xp + t * n2);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[t].v[v] = 0.0;
    }
  }
}

template <int VEC, int NSPIN, bool FULL>
__device__ __forceinline__ void stream_rows(const double* __restrict__ xbase,
                                            const double* __restrict__ dtot,
                                            const double* __restrict__ dens, int n, size_t n2,
                                            int a0, int c0, int tcn, int g, int G, int b_lo, int b_hi,
                                            double (&jacc)[TC], double (&kacc)[NSPIN][VEC]) {
  Vec<VEC> xc[TC], xn[TC];
  Vec<VEC> dtc, dtn;
  int b = b_lo + g;
  if (b >= b_hi) return;
  load_row<VEC, FULL>(xc, xbase + (size_t)b * n, n2, tcn);
  dtc.load_cached(dtot + (size_t)b * n + a0);
  for (; b < b_hi; b += G) {
    const int bn = b + G;
    if (bn < b_hi) {  // prefetch the next row this group owns while the FMAs below run
      load_row<VEC, FULL>(xn, xbase + (size_t)bn * n, n2, tcn);
      dtn.load_cached(dtot + (size_t)bn * n + a0);
    }
    consume_row<VEC, NSPIN, FULL>(xc, dtc, dens, n2, (size_t)b * n + c0, tcn, jacc, kacc);
#pragma unroll
    for (int t = 0; t < TC; ++t) xc[t] = xn[t];
    dtc = dtn;
  }
}

// grid = (n_chunks, n_slabs, BS); block = roundup32(R*G) threads; dynamic smem = kred.
// MAXT = 256 (2 CTAs/SM) covers rows of up to 512 doubles; MAXT = 512 (1 CTA/SM) rows up to 1024.
// BS = gridDim.z > 1 splits the b range of a work item over BS CTAs (used when a shard has too
// few equal-size items to fill the SMs evenly): J then also becomes a per-CTA partial (jpart).
template <int VEC, int NSPIN, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 2 : 1)
jk_full_stream_kernel(const double* __restrict__ eri, const double* __restrict__ dtot,
                      const double* __restrict__ dens, double* __restrict__ jk,
                      double* __restrict__ kpart, double* __restrict__ jpart, int n, int nu_lo,
                      int R, int G) {
  extern __shared__ double kred[];  // [G][NSPIN][n] when G > 1
  __shared__ double jred[MAXT / 32][TC];

  const int tid = threadIdx.x;
  const int g = tid / R;
  const int r = tid - g * R;
  const bool active = g < G;
  const int a0 = r * VEC;
  const int chunk = blockIdx.x;
  const int n_chunks = gridDim.x;
  const int slab = blockIdx.y;
  const int c0 = chunk * TC;
  const int tcn = min(TC, n - c0);
  const size_t n2 = (size_t)n * n;
  const int BS = gridDim.z, bz = blockIdx.z;
  const int nb = (n + BS - 1) / BS;
  const int b_lo = bz * nb, b_hi = min(n, b_lo + nb);

  double jacc[TC];
  double kacc[NSPIN][VEC];
#pragma unroll
  for (int t = 0; t < TC; ++t) jacc[t] = 0.0;
#pragma unroll
  for (int s = 0; s < NSPIN; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v) kacc[s][v] = 0.0;

  if (active) {
    const double* xbase = eri + ((size_t)slab * n + c0) * n2 + a0;
    if (tcn == TC)
      stream_rows<VEC, NSPIN, true>(xbase, dtot, dens, n, n2, a0, c0, tcn, g, G, b_lo, b_hi, jacc, kacc);
    else
      stream_rows<VEC, NSPIN, false>(xbase, dtot, dens, n, n2, a0, c0, tcn, g, G, b_lo, b_hi, jacc, kacc);
  }

  // ---- J[c0+t, nu]: reduce over every thread of the CTA (all a, all b) ----
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    double v = warp_sum(jacc[t]);
    if (lane == 0) jred[warp][t] = v;
  }
  // ---- K partial rows: reduce over the b-groups ----
  if (G > 1 && active) {
#pragma unroll
    for (int s = 0; s < NSPIN; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) kred[((size_t)g * NSPIN + s) * n + a0 + v] = kacc[s][v];
  }
  __syncthreads();
  const int n_warps = (blockDim.x + 31) >> 5;
  const size_t item = ((size_t)slab * n_chunks + chunk) * BS + bz;
  if (tid < tcn) {
    double v = 0.0;
    for (int w = 0; w < n_warps; ++w) v += jred[w][tid];
    if (BS == 1)
      jk[(size_t)(nu_lo + slab) * n + c0 + tid] = v;
    else
      jpart[item * TC + tid] = v;
  }
  double* kp = kpart + item * NSPIN * n;
  if (G > 1) {
    for (int i = tid; i < NSPIN * n; i += blockDim.x) {
      double v = 0.0;
      for (int gg = 0; gg < G; ++gg) v += kred[(size_t)gg * NSPIN * n + i];
      kp[i] = v;
    }
  } else if (active) {
#pragma unroll
    for (int s = 0; s < NSPIN; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) kp[(size_t)s * n + a0 + v] = kacc[s][v];
  }
}

// K[a, nu, s] = sum over (chunk, b-split) of the partial rows; with b-splits also
// J[c, nu] = sum over the b-splits.  Fixed order -> deterministic.
__global__ void k_reduce_kernel(const double* __restrict__ kpart, const double* __restrict__ jpart,
                                double* __restrict__ jk, int n, int nu_lo, int n_slabs,
                                int n_chunks, int n_spin, int bs) {
  const size_t total = (size_t)n * n_slabs * n_spin;
  const size_t parts = (size_t)n_chunks * bs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(i % n);
    const size_t rest = i / n;
    const int slab = (int)(rest % n_slabs);
    const int s = (int)(rest / n_slabs);
    const double* p = kpart + (((size_t)slab * parts) * n_spin + s) * n + a;
    double v = 0.0;
    for (size_t c = 0; c < parts; ++c) v += p[c * n_spin * n];
    jk[(size_t)(1 + s) * n * n + (size_t)(nu_lo + slab) * n + a] = v;
    if (bs > 1 && s == 0) {  // this thread also owns J[c = a, nu]
      const double* q = jpart + (((size_t)slab * n_chunks + a / TC) * bs) * TC + a % TC;
      double j = 0.0;
      for (int z = 0; z < bs; ++z) j += q[(size_t)z * TC];
      jk[(size_t)(nu_lo + slab) * n + a] = j;
    }
  }
}

// out[:,:,s] += J - K_s
__global__ void accumulate_kernel(const double* __restrict__ jk, double* __restrict__ out, int n,
                                  int n_spin) {
  const size_t n2 = (size_t)n * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * n_spin;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / n2, e = i - s * n2;
    out[i] += jk[e] - jk[(1 + s) * n2 + e];
  }
}

template <int VEC, int NSPIN, int MAXT>
cudaError_t launch_stream(scfb_handle_s* h, const double* d_density, const double* d_total) {
  const int n = h->n;
  const int n_slabs = h->nu_hi - h->nu_lo;
  const int n_chunks = (n + TC - 1) / TC;
  const int R = (n + VEC - 1) / VEC;
  int G = MAXT / R;
  if (G < 1) G = 1;
  if (G > n) G = n;
  const int threads = ((R * G + 31) / 32) * 32;
  const size_t smem = G > 1 ? (size_t)G * NSPIN * n * sizeof(double) : 0;
  dim3 grid(n_chunks, n_slabs, h->b_splits);
  jk_full_stream_kernel<VEC, NSPIN, MAXT><<<grid, threads, smem, h->stream>>>(
      h->eri, d_total, d_density, h->jk, h->kpart, h->jpart, n, h->nu_lo, R, G);
  return cudaGetLastError();
}

}  // namespace

// b-splits: equal-size work items fill 2 x 148 resident CTA slots in waves; below ~4 waves the
// last, partly filled wave costs > 10 % (n_bf = 256 over 8 GPUs: 1024 items = 3.46 -> 4 waves).
int scfb_full_b_splits(int n, int n_slabs) {
  const long items = (long)((n + TC - 1) / TC) * n_slabs;
  if (items <= 0) return 1;
  int bs = 1;
  while (bs < 4 && items * bs < 4L * 2 * 148 && n / (2 * bs) >= 16) bs *= 2;
  return bs;
}

uint64_t scfb_kpart_elems_full(int n, int n_slabs) {
  const uint64_t n_chunks = (n + TC - 1) / TC;
  return (uint64_t)n_slabs * n_chunks * scfb_full_b_splits(n, n_slabs) * 2 * n;
}

uint64_t scfb_jpart_elems_full(int n, int n_slabs) {
  const uint64_t n_chunks = (n + TC - 1) / TC;
  return (uint64_t)n_slabs * n_chunks * scfb_full_b_splits(n, n_slabs) * TC;
}

int scfb_launch_jk_full(scfb_handle_s* h, int n_spin, const double* d_density,
                        const double* d_total) {
  const int n = h->n;
  const int n_slabs = h->nu_hi - h->nu_lo;
  if (n_slabs <= 0) {  // this rank owns nothing: its partial stays zero
    if (h->timed) cudaEventRecord(h->ev[3], h->stream);
    if (h->timed) cudaEventRecord(h->ev[1], h->stream);
    return SCFB_OK;
  }
  const bool even = (n % 2 == 0);
  // VEC=2 needs R = n/2 <= MAX_THREADS; wider rows fall back to more rows per thread? no:
  // rows longer than 2*MAX_THREADS are rejected at create time.
  cudaError_t e;
  if (h->timed) cudaEventRecord(h->ev[3], h->stream);
  const bool wide = (even ? n / 2 : n) > MAX_THREADS;  // rows longer than one 256-thread pass
  if (even && !wide)
    e = n_spin == 1 ? launch_stream<2, 1, 256>(h, d_density, d_total)
                    : launch_stream<2, 2, 256>(h, d_density, d_total);
  else if (even)
    e = n_spin == 1 ? launch_stream<2, 1, 512>(h, d_density, d_total)
                    : launch_stream<2, 2, 512>(h, d_density, d_total);
  else if (!wide)
    e = n_spin == 1 ? launch_stream<1, 1, 256>(h, d_density, d_total)
                    : launch_stream<1, 2, 256>(h, d_density, d_total);
  else
    e = n_spin == 1 ? launch_stream<1, 1, 512>(h, d_density, d_total)
                    : launch_stream<1, 2, 512>(h, d_density, d_total);
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "jk_full_stream_kernel launch: %s", cudaGetErrorString(e));
  if (h->timed) cudaEventRecord(h->ev[1], h->stream);
  const int n_chunks = (n + TC - 1) / TC;
  const size_t total = (size_t)n * n_slabs * n_spin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_reduce_kernel<<<blocks, 256, 0, h->stream>>>(h->kpart, h->jpart, h->jk, n, h->nu_lo, n_slabs,
                                                 n_chunks, n_spin, h->b_splits);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "k_reduce_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 2;
  return SCFB_OK;
}

int scfb_launch_accumulate(scfb_handle_s* h, int n_spin, const double* d_jk, double* d_out) {
  const size_t total = (size_t)h->n * h->n * n_spin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  accumulate_kernel<<<blocks, 256, 0, h->stream>>>(d_jk, d_out, h->n, n_spin);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "accumulate_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}
