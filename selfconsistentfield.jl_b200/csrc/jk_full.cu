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

#include <cstdio>
#include <cstdlib>

namespace {

#ifndef SCFB_TMA_WAIT_HINT
#define SCFB_TMA_WAIT_HINT 20000u  // mbarrier.try_wait suspend-time hint, ns
#endif

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

// ------------------------------------------------------------------------------------------
// K1 over the TMA engine (even n <= 512; the default there).  Same work item, same thread
// mapping and the same epilogue as `jk_full_stream_kernel`, but nothing is prefetched through
// registers.  A stage holds the G consecutive b-rows the CTA consumes together: per c-plane one
// contiguous run of G*n doubles, the matching Dtot rows, and D_s[c0..c0+8, b] per row and spin.
// Each of these is ONE 1-D bulk copy (cp.async.bulk, SASS UBLKCP) completing on the stage's
// "full" mbarrier, posted by lane 0 of four producer warps (a bulk copy costs its issuer ~35
// instructions / ~70 cycles, see benchmarks/micro/bulk_copy_rate.cu; 8 consumer warps + 1..4
// producers are allocated as 12 warps anyway).  The consumer warps read their operands with
// LDS.128 (four planes at a time, D_s pair by pair: 80 registers, no spills), do the FMAs and
// release the stage through the "empty" mbarrier.  Bytes in flight are bounded by shared memory
// (2 CTAs x S=3 stages x ~36 KB per SM) instead of by the register file (64 KB per SM for the
// LDG version), which is what a 7+ TB/s read stream needs.  CTAs are persistent and draw work
// items from a global counter, so the ring keeps running across item boundaries.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(SCFB_TMA_WAIT_HINT)  // suspend-time hint (ns): fewer spins
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

constexpr int TMA_MAX_STAGES = 4;

// Timeline probe of the persistent kernel (tuning builds only: -DSCFB_K1_TRACE, see
// benchmarks/k1_trace.py).  Off by default: the shipped kernel carries none of it.
#ifdef SCFB_K1_TRACE
__device__ unsigned long long g_k1_trace[1024][8];
__device__ __forceinline__ unsigned long long k1_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K1_TRACE(slot, val) do { if (tid == 0 && blockIdx.x < 1024) g_k1_trace[blockIdx.x][slot] = (val); } while (0)
#else
#define K1_TRACE(slot, val) do { } while (0)
#endif
__host__ __device__ inline size_t tma_stage_doubles(int n, int G, int n_spin) {
  return (size_t)(TC + 1) * G * n + (size_t)G * n_spin * TC;
}

constexpr int TMA_PRODUCERS = 4;  // producer warps per CTA (9..12 warps are allocated as 12 anyway)

// Persistent: grid = min(items, 2 * SMs) CTAs drawing work items from a global counter, where
// item = chunk + n_chunks * (slab + n_slabs * bz).  The stage ring runs on ACROSS items: while
// the consumers finish an item (J/K reductions, partial-row stores) the producers are already S-1
// stages into the next one, so the ring's fill latency and the epilogue are paid once per CTA, not
// once per item (measured per-item bubble of the one-item-per-CTA version: ~5 us).
// block = MAX_THREADS consumer threads + 32 * TMA_PRODUCERS; dynamic smem = S stages.
template <int NSPIN>
__global__ void __launch_bounds__(MAX_THREADS + 32 * TMA_PRODUCERS, 2)
jk_full_tma_kernel(const double* __restrict__ eri, const double* __restrict__ dtot,
                   const double* __restrict__ dens, double* __restrict__ jk,
                   double* __restrict__ kpart, double* __restrict__ jpart, int n, int nu_lo, int R,
                   int G, int S, int n_chunks, int n_slabs, int whole_slabs, int BSB, int* __restrict__ sched,
                   int* __restrict__ slab_done, int* __restrict__ grp_done, unsigned int* __restrict__ grid_done,
                   ScfbFuse fuse) {
  extern __shared__ __align__(128) double ring[];  // [S][stage]
  __shared__ double jred[MAX_THREADS / 32][TC];
  __shared__ int slab_last, grp_last;
  __shared__ uint64_t full[TMA_MAX_STAGES], empty[TMA_MAX_STAGES];
  // Work items are handed out dynamically (SMs do not all stream at the same rate: a static
  // round-robin was 6 % slower than one CTA per item).  Lane 0 of producer warp 0 draws the CTA's
  // next item from a global counter and publishes it in a two-deep queue; the other producers and
  // the consumer warps pick it up from there.  -1 ends the CTA.
  __shared__ int item_q[2];
  __shared__ uint64_t item_full[2], item_empty[2];

  const int tid = threadIdx.x;
  constexpr int CW = MAX_THREADS / 32;  // consumer warps (those beyond R*G threads only keep the barriers going)
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = warp >= CW;
  const size_t n2 = (size_t)n * n;
  const size_t stage_doubles = tma_stage_doubles(n, G, NSPIN);
  const int Gn = G * n;
  // Items [0, n_whole): (chunk, slab) of the first `whole_slabs` slabs, all b.  Items beyond: the
  // remaining slabs cut BSB ways along b, (bz, chunk, slab) with bz fastest — the END of the
  // counter hands out pieces 1/BSB the size, so the CTAs run dry within a fraction of an item of
  // each other instead of within a whole one.  n % BSB == 0 (host), so no piece is empty.
  const int n_whole = whole_slabs * n_chunks;
  const int n_items = n_whole + (n_slabs - whole_slabs) * n_chunks * BSB;
  const int nb_split = n / BSB;
#ifdef SCFB_K1_TRACE
  K1_TRACE(0, k1_now());  // CTA start
  unsigned long long tr_epi = 0, tr_fin = 0, tr_last_start = 0;
  unsigned int tr_items = 0, tr_smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(tr_smid));
#endif

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, TMA_PRODUCERS);
      mbar_init(empty + s, CW);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(item_full + q, 1);
      mbar_init(item_empty + q, CW + TMA_PRODUCERS - 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // k-th item of this CTA: slot k & 1, phase (k >> 1) & 1
  auto publish_item = [&](int k) {  // scheduler only
    if (k >= 2) mbar_wait(item_empty + (k & 1), ((k >> 1) & 1) ^ 1);
    int item = k == 0 ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(sched, 1);
    if (item >= n_items) item = -1;
    item_q[k & 1] = item;
    mbar_arrive(item_full + (k & 1));  // release: the store above is visible to the waiters
    return item;
  };
  auto fetch_item = [&](int k) {  // everyone else; one arrival per warp on item_empty
    mbar_wait(item_full + (k & 1), (k >> 1) & 1);
    const int item = item_q[k & 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(item_empty + (k & 1));
    return item;
  };

  // Register budget: 80 per thread (2 CTAs x 384 threads).  The consumers fit because D_s is read
  // from the stage pair by pair inside the FMA loop instead of being preloaded (a setmaxnreg
  // split 32/104 measured the same, but left the 32-register producers spilling).
  if (producer) {
    // Copy jobs of one stage, dealt round-robin to the producer warps (job J -> warp J % 4):
    //   J < tcn       plane c0+J, rows*n doubles                         -> slot J of the stage
    //   J == tcn      the Dtot rows, rows*n doubles                      -> slot TC
    //   J = tcn+1+q   D_s[c0..c0+tcn, b+gg], q = gg*NSPIN+sp, tcn doubles -> the stage's tail
    //                 (n even => tcn even => a multiple of 16 bytes)
    // Lane 0 posts its jobs, then announces their byte count on the stage's "full" barrier (4
    // arrivals + the bytes complete a phase; the phase cannot complete before all four have
    // arrived, so copies that land before the announcement are accounted for).  Every source
    // address advances by G*n doubles per stage: the pointers of a warp's first MAXMINE jobs
    // live in registers and a copy costs an add plus the UBLKCP sequence (~35 instructions) —
    // kept out of the consumers' instruction stream, which is what made the first, consumer-posted
    // version instruction-bound (8.5 ms).
    if (lane == 0) {
      const int pw = warp - CW;
      int s = 0;
      uint32_t parity = 1;  // a fresh "empty" barrier passes a wait on the opposite parity
      double* dst = ring;
      for (int k = 0;; ++k) {
        const int item = pw == 0 ? publish_item(k) : fetch_item(k);
        if (item < 0) break;
        int chunk, slab, b_lo, b_hi;
        if (item < n_whole) {
          chunk = item % n_chunks, slab = item / n_chunks, b_lo = 0, b_hi = n;
        } else {
          const int j = item - n_whole, bz = j % BSB, cs = j / BSB;
          chunk = cs % n_chunks, slab = whole_slabs + cs / n_chunks, b_lo = bz * nb_split, b_hi = b_lo + nb_split;
        }
        const int c0 = chunk * TC;
        const int tcn = min(TC, n - c0);
        const int n_iter = (b_hi - b_lo + G - 1) / G;
        const double* xbase = eri + ((size_t)slab * n + c0) * n2 + (size_t)b_lo * n;  // plane c0, row b_lo
        const int n_jobs = tcn + 1 + G * NSPIN;
        auto job_src = [&](int J) -> const double* {
          if (J < tcn) return xbase + (size_t)J * n2;
          if (J == tcn) return dtot + (size_t)b_lo * n;
          const int q = J - tcn - 1, gg = q / NSPIN, sp = q - gg * NSPIN;
          return dens + sp * n2 + (size_t)(b_lo + gg) * n + c0;
        };
        auto job_dst = [&](int J) -> int {
          return J < tcn ? J * Gn : (J == tcn ? TC * Gn : (TC + 1) * Gn + (J - tcn - 1) * TC);
        };
        constexpr int MAXMINE = 4;
        const double* jsrc[MAXMINE];
        int jdst[MAXMINE], jrow[MAXMINE];  // jrow: -1 = row-sized job, else the row gg of a small job
#pragma unroll
        for (int k = 0; k < MAXMINE; ++k) {
          const int J = pw + k * TMA_PRODUCERS;
          jsrc[k] = J < n_jobs ? job_src(J) : nullptr;
          jdst[k] = J < n_jobs ? job_dst(J) : 0;
          jrow[k] = J <= tcn ? -1 : (J - tcn - 1) / NSPIN;
        }
        const uint32_t small_bytes = (uint32_t)tcn * 8u;
        for (int it = 0; it < n_iter; ++it) {
          mbar_wait(empty + s, parity);
          const int rows = min(G, b_hi - (b_lo + it * G));
          const uint32_t row_bytes = (uint32_t)(rows * n) * 8u;
          uint64_t* bar = full + s;
          uint32_t bytes = 0;
#pragma unroll
          for (int k = 0; k < MAXMINE; ++k) {
            if (jsrc[k] != nullptr) {
              if (jrow[k] < 0) {
                bulk_g2s(dst + jdst[k], jsrc[k], row_bytes, bar);
                bytes += row_bytes;
              } else if (jrow[k] < rows) {
                bulk_g2s(dst + jdst[k], jsrc[k], small_bytes, bar);
                bytes += small_bytes;
              }
              jsrc[k] += Gn;
            }
          }
          for (int J = pw + MAXMINE * TMA_PRODUCERS; J < n_jobs; J += TMA_PRODUCERS) {  // small n only
            const bool big = J <= tcn;
            if (big || (J - tcn - 1) / NSPIN < rows) {
              bulk_g2s(dst + job_dst(J), job_src(J) + (size_t)it * Gn, big ? row_bytes : small_bytes, bar);
              bytes += big ? row_bytes : small_bytes;
            }
          }
          mbar_expect_tx(bar, bytes);
          dst += stage_doubles;
          if (++s == S) {
            s = 0;
            parity ^= 1;
            dst = ring;
          }
        }
      }
    }
    return;  // the consumers synchronise among themselves (named barrier 1)
  }

  const int g = tid / R;
  const int r = tid - g * R;
  const bool active = g < G;  // (tid / R >= G for the consumer threads beyond R*G)
  const int a0 = r * 2;
  // per-thread operand offsets inside a stage (doubles)
  const int x_off = g * n + a0;
  const int ds_off = (TC + 1) * Gn + g * NSPIN * TC;
  int s = 0;
  uint32_t parity = 0;
  const double* st = ring;

  for (int k = 0;; ++k) {
    const int item = fetch_item(k);
    if (item < 0) break;
    int chunk, slab, b_lo, b_hi;
    const bool split = item >= n_whole;
    if (!split) {
      chunk = item % n_chunks, slab = item / n_chunks, b_lo = 0, b_hi = n;
    } else {
      const int j = item - n_whole, bz = j % BSB, cs = j / BSB;
      chunk = cs % n_chunks, slab = whole_slabs + cs / n_chunks, b_lo = bz * nb_split, b_hi = b_lo + nb_split;
    }
    const int c0 = chunk * TC;
    const int tcn = min(TC, n - c0);
    const int n_iter = (b_hi - b_lo + G - 1) / G;

    double jacc[TC];
    double kacc[NSPIN][2];
#pragma unroll
    for (int t = 0; t < TC; ++t) jacc[t] = 0.0;
#pragma unroll
    for (int sp = 0; sp < NSPIN; ++sp) kacc[sp][0] = kacc[sp][1] = 0.0;

    int b = b_lo + g;
    double* last_stage = nullptr;  // the item's last stage doubles as kred before it is released
    int last_slot = 0;
#ifdef SCFB_K1_TRACE
    tr_last_start = k1_now();
    ++tr_items;
#endif
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(full + s, parity);
#ifdef SCFB_K1_TRACE
      if (k == 0 && it == 0) K1_TRACE(1, k1_now());  // first stage landed
#endif
      const bool live = active && b < b_hi;
      if (live) {
        // two half-chunks of four planes: 16 operand registers live instead of 32
        const double2 dt = *reinterpret_cast<const double2*>(st + x_off + TC * Gn);  // slot TC: Dtot rows
        const double* dsp = st + ds_off;                                             // broadcast reads
        const double* xp = st + x_off;
#pragma unroll
        for (int h = 0; h < TC; h += 4) {
          double2 x[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            x[t] = h + t < tcn ? *reinterpret_cast<const double2*>(xp) : make_double2(0.0, 0.0);
            xp += Gn;
          }
#pragma unroll
          for (int t = 0; t < 4; t += 2) {
            jacc[h + t] = fma(x[t].x, dt.x, jacc[h + t]);
            jacc[h + t] = fma(x[t].y, dt.y, jacc[h + t]);
            jacc[h + t + 1] = fma(x[t + 1].x, dt.x, jacc[h + t + 1]);
            jacc[h + t + 1] = fma(x[t + 1].y, dt.y, jacc[h + t + 1]);
#pragma unroll
            for (int sp = 0; sp < NSPIN; ++sp) {
              double2 d2 = *reinterpret_cast<const double2*>(dsp + sp * TC + h + t);
              // columns of D beyond the last chunk's tcn are never copied: mask instead of multiply
              if (h + t >= tcn) d2.x = 0.0;
              if (h + t + 1 >= tcn) d2.y = 0.0;
              kacc[sp][0] = fma(x[t].x, d2.x, kacc[sp][0]);
              kacc[sp][1] = fma(x[t].y, d2.x, kacc[sp][1]);
              kacc[sp][0] = fma(x[t + 1].x, d2.y, kacc[sp][0]);
              kacc[sp][1] = fma(x[t + 1].y, d2.y, kacc[sp][1]);
            }
          }
        }
      }
      __syncwarp();
      if (it + 1 < n_iter) {
        if (lane == 0) mbar_arrive(empty + s);  // all reads of the stage are done: hand it back
      } else {
        last_stage = const_cast<double*>(st);
        last_slot = s;
      }
      b += G;
      st += stage_doubles;
      if (++s == S) {
        s = 0;
        parity ^= 1;
        st = ring;
      }
    }

    // ---- item epilogue (as in jk_full_stream_kernel; kred lives in the item's last stage) ----
#ifdef SCFB_K1_TRACE
    const unsigned long long tr_e0 = k1_now();
#endif
    double* kred = last_stage;  // [G][NSPIN][n]
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      double v = warp_sum(jacc[t]);
      if (lane == 0) jred[warp][t] = v;
    }
    // every consumer has copied its operands out of the last stage and finished its FMAs
    asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
    if (G > 1 && active) {
#pragma unroll
      for (int sp = 0; sp < NSPIN; ++sp) {
        kred[((size_t)g * NSPIN + sp) * n + a0] = kacc[sp][0];
        kred[((size_t)g * NSPIN + sp) * n + a0 + 1] = kacc[sp][1];
      }
    }
    if (tid < tcn) {
      double v = 0.0;
      for (int w = 0; w < CW; ++w) v += jred[w][tid];
      if (!split)
        jk[(size_t)(nu_lo + slab) * n + c0 + tid] = v;
      else
        jpart[(size_t)(item - n_whole) * TC + tid] = v;
    }
    double* kp = kpart + (size_t)item * NSPIN * n;  // one partial row pair per item, in item order
    if (G > 1) {
      asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
      for (int i = tid; i < NSPIN * n; i += MAX_THREADS) {
        double v = 0.0;
        for (int gg = 0; gg < G; ++gg) v += kred[(size_t)gg * NSPIN * n + i];
        kp[i] = v;
      }
    } else if (active) {
#pragma unroll
      for (int sp = 0; sp < NSPIN; ++sp) {
        kp[(size_t)sp * n + a0] = kacc[sp][0];
        kp[(size_t)sp * n + a0 + 1] = kacc[sp][1];
      }
    }
    // jred and kred are free again: release the last stage to the producers.  (The generic-proxy
    // writes to kred are ordered before the async-proxy refill by the fence + barrier arrival.)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
    if (lane == 0) mbar_arrive(empty + last_slot);

    // ---- fused tail: the CTA that completes a nu-slab reduces it, accumulates / ships it ----
    // (the producers are already streaming the next item; "last block" pattern: every writer
    // fences, one thread counts, the CTA that sees the last count reads the partials back from L2.
    // A split item first completes its (slab, chunk) group: the last of the BSB pieces folds the
    // group's rows into the first one and its J partials into J, then counts for the slab like a
    // whole item.  Every sum runs in a fixed order: the result does not depend on who finishes.)
    const size_t stride = (size_t)NSPIN * n;
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
    if (split) {
      const int grp = (item - n_whole) / BSB;
      if (tid == 0) {
        const int last = atomicAdd(grp_done + grp, 1) == BSB - 1;
        if (last) grp_done[grp] = 0;  // ready for the next build
        grp_last = last;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
      if (grp_last) {
        __threadfence();
        double* row0 = kpart + (size_t)(n_whole + grp * BSB) * stride;
        for (int i = tid; i < NSPIN * n; i += MAX_THREADS) {
          double v = 0.0;
          for (int z = 0; z < BSB; ++z) v += __ldcg(row0 + (size_t)z * stride + i);
          row0[i] = v;
        }
        if (tid < tcn) {
          double v = 0.0;
          for (int z = 0; z < BSB; ++z) v += __ldcg(jpart + (size_t)(grp * BSB + z) * TC + tid);
          jk[(size_t)(nu_lo + slab) * n + c0 + tid] = v;
        }
        __threadfence();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
    }
    if (tid == 0) {
      int last = 0;
      if (!split || grp_last) {
        last = atomicAdd(slab_done + slab, 1) == n_chunks - 1;
        if (last) slab_done[slab] = 0;
      }
      slab_last = last;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
#ifdef SCFB_K1_TRACE
    const unsigned long long tr_f0 = k1_now();
    tr_epi += tr_f0 - tr_e0;
#endif
    if (slab_last) {
      __threadfence();
      const size_t col = (size_t)(nu_lo + slab) * n;
      // the slab's n_chunks rows: consecutive for whole items, every BSB-th (the folded ones) otherwise
      const size_t row0 = split ? (size_t)n_whole + (size_t)(slab - whole_slabs) * n_chunks * BSB : (size_t)slab * n_chunks;
      const size_t rstep = (split ? (size_t)BSB : 1) * stride;
      for (int i2 = tid; i2 < NSPIN * n / 2; i2 += MAX_THREADS) {  // n is even: a pair never straddles the spins
        const int i = 2 * i2, sp = i >= n ? 1 : 0, a = i - sp * n;
        const double* p = kpart + row0 * stride + i;
        double2 v = make_double2(0.0, 0.0);
        int c = 0;
        for (; c + 8 <= n_chunks; c += 8) {  // eight 16-byte loads in flight, added in the same fixed order
          double2 t[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = __ldcg(reinterpret_cast<const double2*>(p + (size_t)(c + u) * rstep));
#pragma unroll
          for (int u = 0; u < 8; ++u) v.x += t[u].x, v.y += t[u].y;
        }
        for (; c < n_chunks; ++c) {
          const double2 t = __ldcg(reinterpret_cast<const double2*>(p + (size_t)c * rstep));
          v.x += t.x, v.y += t.y;
        }
        const double2 jv = __ldcg(reinterpret_cast<const double2*>(jk + col + a));  // J[a, nu]: written by this slab's items
        *reinterpret_cast<double2*>(jk + (size_t)(1 + sp) * n2 + col + a) = v;
        if (fuse.out) {
          double* o = fuse.out + (size_t)sp * n2 + col + a;  // (caller's pointer: no alignment assumed)
          o[0] += jv.x - v.x;
          o[1] += jv.y - v.y;
        }
        for (int r = 0; r < fuse.n_peers; ++r) {  // plain stores into every rank's landing zone (NVLink)
          *reinterpret_cast<double2*>(fuse.peer_jk[r] + (size_t)(1 + sp) * n2 + col + a) = v;
          if (sp == 0) *reinterpret_cast<double2*>(fuse.peer_jk[r] + col + a) = jv;
        }
      }
      if (fuse.n_peers) __threadfence_system();
#ifdef SCFB_K1_TRACE
      tr_fin += k1_now() - tr_f0;
#endif
    }
  }
#ifdef SCFB_K1_TRACE
  K1_TRACE(2, k1_now());  // left the item loop
  K1_TRACE(4, (unsigned long long)tr_items | ((unsigned long long)tr_smid << 32));
  K1_TRACE(5, tr_epi);
  K1_TRACE(6, tr_last_start);
  K1_TRACE(7, tr_fin);
#endif

  // ---- kernel tail: the last CTA re-arms the counters and publishes this rank's epoch ----
  asm volatile("bar.sync 1, %0;" ::"n"(MAX_THREADS) : "memory");
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(grid_done, 1u) == gridDim.x - 1) {  // every CTA has drawn its last item and fenced its stores
      *grid_done = 0;
      *sched = 0;
      if (fuse.n_peers) {
        __threadfence_system();
        for (int r = 0; r < fuse.n_peers; ++r) *reinterpret_cast<volatile unsigned long long*>(fuse.peer_flag[r]) = fuse.epoch;
        __threadfence_system();
      }
    }
  }
  K1_TRACE(3, k1_now());  // CTA end
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
    const size_t stride = (size_t)n_spin * n;
    double v = 0.0;
    size_t c = 0;
    for (; c + 8 <= parts; c += 8) {  // eight loads in flight, added in the same fixed order
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(p + (c + u) * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; c < parts; ++c) v += __ldg(p + c * stride);
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

template <int NSPIN>
cudaError_t launch_tma(scfb_handle_s* h, const double* d_density, const double* d_total, const ScfbFuse& fuse) {
  const int n = h->n;
  const int n_slabs = h->nu_hi - h->nu_lo;
  const int n_chunks = (n + TC - 1) / TC;
  const int R = n / 2;
  int G = 256 / R;
  if (G < 1) G = 1;
  if (G > n) G = n;
  const size_t stage = tma_stage_doubles(n, G, NSPIN) * sizeof(double);
  // two CTAs per SM share 227 KB (1 KB per CTA is reserved, ~0.6 KB static)
  int S = (int)((size_t)(113 * 1024 - 2048) / stage);
  if (S > TMA_MAX_STAGES) S = TMA_MAX_STAGES;
  if (S < 2) S = 2;
  const size_t smem = (size_t)S * stage;
  if (!h->tma_attr[NSPIN]) {  // function attributes are per device, hence per handle
    cudaError_t e = cudaFuncSetAttribute(jk_full_tma_kernel<NSPIN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e == cudaSuccess)  // two CTAs only fit with the largest shared-memory carveout
      e = cudaFuncSetAttribute(jk_full_tma_kernel<NSPIN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    h->tma_attr[NSPIN] = true;
  }
  static const bool verbose = std::getenv("SCFB_VERBOSE") != nullptr;
  if (verbose) {
    int occ = 0, occ0 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jk_full_tma_kernel<NSPIN>, MAX_THREADS + 32 * TMA_PRODUCERS, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, jk_full_tma_kernel<NSPIN>, MAX_THREADS + 32 * TMA_PRODUCERS, 0);
    fprintf(stderr, "[scfb] jk_full_tma_kernel<%d>: n=%d R=%d G=%d stages=%d smem=%zu B threads=%d CTAs/SM=%d "
            "(%d without dynamic smem)\n", NSPIN, n, R, G, S, smem, MAX_THREADS + 32 * TMA_PRODUCERS, occ, occ0);
  }
  const ScfbFullPlan& pl = h->plan;
  const long items = (long)n_chunks * (pl.whole_slabs + (long)(n_slabs - pl.whole_slabs) * pl.tail_bs);
  const int n_sm = h->n_sm > 0 ? h->n_sm : 148;
  static const int per_sm = [] {  // SCFB_FULL_TMA_CTAS=0: one CTA per item (non-persistent), else CTAs per SM
    const char* v = std::getenv("SCFB_FULL_TMA_CTAS");
    return v ? std::atoi(v) : 2;
  }();
  const long cap = per_sm > 0 ? (long)per_sm * n_sm : items;
  const int grid = (int)(items < cap ? items : cap);
  // (the work-item counter and the completion counters are re-armed by the kernel's last CTA)
  jk_full_tma_kernel<NSPIN><<<grid, MAX_THREADS + 32 * TMA_PRODUCERS, smem, h->stream>>>(
      h->eri, d_total, d_density, h->jk, h->kpart, h->jpart, n, h->nu_lo, R, G, S, n_chunks, n_slabs,
      pl.whole_slabs, pl.tail_bs, h->sched, h->slab_done, h->grp_done, h->grid_done, fuse);
  return cudaGetLastError();
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
  dim3 grid(n_chunks, n_slabs, h->plan.b_splits);
  jk_full_stream_kernel<VEC, NSPIN, MAXT><<<grid, threads, smem, h->stream>>>(
      h->eri, d_total, d_density, h->jk, h->kpart, h->jpart, n, h->nu_lo, R, G);
  return cudaGetLastError();
}

}  // namespace

#ifdef SCFB_K1_TRACE
extern "C" int scfb_debug_k1_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_k1_trace, sizeof(g_k1_trace));
}
#endif

static bool tma_path(int n) {  // the selection rule of scfb_launch_jk_full
  const char* v = std::getenv("SCFB_FULL_TMA");
  return (v ? std::atoi(v) : 1) != 0 && n % 2 == 0 && n >= 8 && n / 2 <= MAX_THREADS;
}

// Work-item plan of a rank's shard.
//
// LDG kernel (one CTA per item): equal-size items fill 2 x n_sm resident CTA slots in waves, and
// below ~4 waves the last, partly filled wave costs > 10 %, so every (nu, c-chunk) is split over
// 2-4 CTAs along b (gridDim.z); J then also becomes a per-CTA partial and k_reduce_kernel sums.
//
// Persistent TMA kernel: items are drawn from a counter, so only the END of the build is ragged —
// with whole items (8 planes = 64 n^2 bytes, ~170 us at n_bf = 256) the CTAs run dry up to one item
// apart (traced on a 1/8 shard of n_bf = 256: 160 of 296 CTAs idle for the last ~95 of 622 us).  The
// last ~one wave of whole items (ceil(2 n_sm / n_chunks) slabs) is therefore cut tail_bs ways along b
// into pieces of >= 1 MB; the kernel folds a group's pieces before the slab finisher sees them.
// SCFB_FULL_BSPLIT=k forces k pieces for EVERY item (both kernels), SCFB_FULL_TAIL=0 keeps all items
// whole; both are read per handle creation (tests, tuning).
ScfbFullPlan scfb_full_plan(int n, int n_slabs, int n_sm) {
  ScfbFullPlan pl;
  pl.tma = tma_path(n);
  if (n_slabs <= 0) return pl;
  if (n_sm <= 0) n_sm = 148;
  const int n_chunks = (n + TC - 1) / TC;
  const long items = (long)n_chunks * n_slabs;
  int forced = 0;
  if (const char* v = std::getenv("SCFB_FULL_BSPLIT")) forced = std::atoi(v);
  if (!pl.tma) {
    int bs = 1;
    if (forced > 0)
      bs = n / (2 * forced) >= 1 ? forced : 1;
    else
      while (bs < 4 && items * bs < 4L * 2 * n_sm && n / (2 * bs) >= 16) bs *= 2;
    pl.b_splits = bs;
    pl.kpart_rows = (uint64_t)items * bs;
    pl.jpart_rows = bs > 1 ? (uint64_t)items * bs : 0;
    return pl;
  }
  int bs = 1, split_slabs = 0;
  if (forced > 0) {
    if (n % forced == 0 && n / forced >= 2) bs = forced, split_slabs = n_slabs;
  } else {
    const char* t = std::getenv("SCFB_FULL_TAIL");
    if (!t || std::atoi(t) != 0) {
      const size_t item_bytes = (size_t)TC * n * n * sizeof(double);
      while (bs < 8 && n % (2 * bs) == 0 && item_bytes / (2 * bs) >= ((size_t)1 << 20)) bs *= 2;
      if (bs > 1) {
        split_slabs = t && std::atoi(t) > 1 ? std::atoi(t) : (2 * n_sm + n_chunks - 1) / n_chunks;
        if (split_slabs > n_slabs) split_slabs = n_slabs;
      }
    }
  }
  if (bs == 1) split_slabs = 0;
  pl.whole_slabs = n_slabs - split_slabs;
  pl.tail_bs = bs;
  pl.n_groups = split_slabs * n_chunks;
  pl.jpart_rows = (uint64_t)pl.n_groups * bs;
  pl.kpart_rows = (uint64_t)pl.whole_slabs * n_chunks + pl.jpart_rows;
  return pl;
}

// `fuse` (may be null): what the TMA kernel's slab finisher should do beyond filling h->jk;
// *fused tells the caller whether it was honoured (TMA path, unsplit items) — otherwise the caller
// runs the separate accumulate / push launches.
int scfb_launch_jk_full(scfb_handle_s* h, int n_spin, const double* d_density,
                        const double* d_total, const ScfbFuse* fuse, bool* fused) {
  const int n = h->n;
  const int n_slabs = h->nu_hi - h->nu_lo;
  if (fused) *fused = false;
  if (n_slabs <= 0) {  // this rank owns nothing: its partial stays zero
    if (h->timed) scfb_ev_alias(h, 3, 0);
    if (h->timed) scfb_ev_alias(h, 1, 0);
    return SCFB_OK;
  }
  const bool even = (n % 2 == 0);
  // VEC=2 needs R = n/2 <= MAX_THREADS; wider rows fall back to more rows per thread? no:
  // rows longer than 2*MAX_THREADS are rejected at create time.
  cudaError_t e;
  if (h->timed) scfb_ev_alias(h, 3, 0);  // the stream kernel is the build's first launch
  const bool wide = (even ? n / 2 : n) > MAX_THREADS;  // rows longer than one 256-thread pass
  const bool tma = h->plan.tma;  // SCFB_FULL_TMA=0 selects the register-prefetch (LDG) kernel
  const bool in_kernel_tail = tma;  // every slab is finished inside the kernel
  if (tma) {
    ScfbFuse f{};
    if (fuse && in_kernel_tail) {
      f = *fuse;
      if (fused) *fused = true;
    }
    e = n_spin == 1 ? launch_tma<1>(h, d_density, d_total, f) : launch_tma<2>(h, d_density, d_total, f);
  }
  else if (even && !wide)
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
  if (h->timed) scfb_ev_record(h, 1);
  if (in_kernel_tail) {  // K rows were reduced by the slab finishers
    h->launches += 1;
    return SCFB_OK;
  }
  const int n_chunks = (n + TC - 1) / TC;
  const size_t total = (size_t)n * n_slabs * n_spin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_reduce_kernel<<<blocks, 256, 0, h->stream>>>(h->kpart, h->jpart, h->jk, n, h->nu_lo, n_slabs,
                                                 n_chunks, n_spin, h->plan.b_splits);
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
