// K2 — single-pass J+K over the 8-fold-symmetry-packed ERI (sm_100a).
//
// Storage (this library's own HBM layout, "packed8"): pair index pr(i,j) = i(i+1)/2 + j,
// i >= j.  Row p = pr(i,j) holds the canonical elements (ij|kl) with pr(k,l) <= p, grouped
// in k-segments (l = 0..k; the last segment k = i stops at l = j).  Every segment is zero-padded
// to a multiple of 8 elements so that all segment and row starts are 64-byte aligned — every
// 512-byte run a warp reads then covers whole DRAM atoms (16-byte alignment made the stream
// fetch 1.13-1.25x its bytes):
//   E(k)   = sum_{m<k} ceil8(m+1)       padded elements of segments 0..k-1
//   R(i,j) = B[i] + j*E(i) + E(j)      offset of row (i,j);  B[i+1] = B[i] + (i+1)E(i) + E(i+1)
// (padding overhead 2.4 % at n = 384; roofline figures use the UNPADDED byte count).
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
// Mapping (details at `packed_item`): one warp owns a work item (two consecutive first indices
// iA, iA+1; a chunk of KC consecutive k; 64 consecutive l); each lane owns an l-PAIR (16 bytes per
// k, row and stream) and the warp loops over the rows j, consuming row (iA,j) and row (iA+1,j)
// together, each of which contributes contiguous 512-byte runs of the packed row.
//   * all per-row data (ERI pairs from HBM, the density row from L2) travels through a DEPTH-deep
//     cp.async (LDGSTS) ring in shared memory; the row loop is unrolled DEPTH times so every ring
//     and tile address is a compile-time offset;
//   * sums whose output does not depend on j (K'[i,k], K'[i,l], J~[q]) stay in registers for the
//     whole item; K'[j,l] is lane-private (one RED per lane and row PAIR); the lane-reductions
//     (K'[j,k], J~[p]) go through a conflict-free shared-memory tile flushed every DEPTH row pairs;
//   * warps never synchronise with each other: a persistent CTA is 8 independent warps (one per
//     SM sub-partition slot, 255 registers) that draw items from a global counter; cross-warp
//     combination is fp64 RED.ADD into zeroed accumulators, so the packed path is deterministic
//     only up to summation order (well inside the 1e-12 relative gate).
// The l-blocks of one (pair, k-chunk) are neighbours in the item list and walk the rows in the same
// order, so what two neighbours share of a DRAM page is fetched close in time.
#include "scfb_internal.h"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace {

__host__ __device__ __forceinline__ uint64_t tri(uint64_t i) { return i * (i + 1) / 2; }
// every k-segment is padded with zeros to a multiple of SEG_PAD elements (64 bytes)
constexpr uint64_t SEG_PAD = 8;
__host__ __device__ __forceinline__ uint64_t seg_len(uint64_t l_count) { return (l_count + SEG_PAD - 1) & ~(SEG_PAD - 1); }
// padded elements of the segments 0..k-1: sum_{m<k} seg_len(m+1) = 8 (q+1) (4q + r), k = 8q + r
__host__ __device__ __forceinline__ uint64_t seg_off(uint64_t k) {
  const uint64_t q = k / SEG_PAD, r = k % SEG_PAD;
  return SEG_PAD * (q + 1) * (SEG_PAD / 2 * q + r);
}

// Sum V (4 or 8) per-lane values over the 32 lanes of a warp with V+1 (instead of 5V)
// shuffles: each butterfly stage halves the number of values a lane still carries.
// On return lanes with (lane % (32/V)) == 0 hold the total of value index lane / (32/V)
// in bit-reversed-free order given by transpose_reduce_index<V>().
template <int V>
__device__ __forceinline__ double warp_transpose_reduce(const double (&v)[V], int lane) {
  static_assert(V == 8 || V == 4, "V must be 4 or 8");
  double a[4], b[2], c;
  if (V == 8) {
    const bool hi16 = lane & 16;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double keep = hi16 ? v[(t + 4) % V] : v[t];
      const double send = hi16 ? v[t] : v[(t + 4) % V];
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
    const double keep = hi4 ? b[1] : b[0];
    const double send = hi4 ? b[0] : b[1];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  } else {
    const bool hi16 = lane & 16;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const double keep = hi16 ? v[(t + 2) % V] : v[t];
      const double send = hi16 ? v[t] : v[(t + 2) % V];
      b[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const bool hi8 = lane & 8;
    const double keep = hi8 ? b[1] : b[0];
    const double send = hi8 ? b[0] : b[1];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
  }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}
template <int V>
__device__ __forceinline__ int transpose_reduce_index(int lane) {
  return V == 8 ? ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
                : ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
}
template <int V>
__device__ __forceinline__ bool transpose_reduce_owner(int lane) {
  return V == 8 ? (lane & 3) == 0 : (lane & 7) == 0;
}

// D: (n*n + DPAD) x NSPIN symmetric, padded so that the KC-wide uniform reads D[j, k0..k0+KC)
// and the pair reads D[j, l0..l0+1] never leave the allocation; dq2: 2*Dtot in pair space,
// stored with the SAME padded segment offsets seg_off(k)+l (pads zero).  Kp: (n*n + DPAD) x
// NSPIN accumulators; Jp: accumulators indexed like dq2.
constexpr int DPAD = 64;
constexpr int YS = 34;    // padded lane stride of the row-sum tile (doubles)

// shared memory of one warp, in doubles
// DEPTH: cp.async ring depth (row pairs) == row pairs parked between flushes == unroll
template <int NSPIN, int KC, int DEPTH>
struct PkSmem {
  static constexpr int NK = NSPIN * KC;
  static constexpr int NV = NK + 2;  // lane-reduced values per row pair: K'[j,k] per spin and k, J~[(iA,j)], J~[(iB,j)]
  static_assert(DEPTH * NV <= 32, "one flush pass: one lane per (row, value)");
  static constexpr int DKN = (NK + 3) & ~1;  // D[j, k-chunk] per spin + 2Dtot[(iA,j)] + 2Dtot[(iB,j)], padded even
  static constexpr int TILE = 0;                                // [DEPTH][NV][YS]
  static constexpr int RING = TILE + DEPTH * NV * YS;           // [DEPTH][2][KC][32] double2
  static constexpr int RING_DL = RING + DEPTH * 2 * KC * 64;    // [DEPTH][NSPIN][32] double2
  static constexpr int RING_DK = RING_DL + DEPTH * NSPIN * 64;  // [DEPTH][DKN]
  static constexpr int TOTAL = RING_DK + DEPTH * DKN;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// One work item of one warp: TWO consecutive first indices iA, iB = iA + 1 (stream A = rows
// (iA, j), stream B = rows (iB, j); `has_b` false for a shard's unpaired last index), one chunk of
// KC consecutive k and 64 consecutive l.  Row j of both streams is consumed together: every output
// indexed by j — K'[j,k] (a 32-lane reduction per k), K'[j,l] (a RED per lane) — and every per-row
// load (D[j,l], D[j,k]) is then paid once per TWO packed rows, which is what bounds this kernel
// (LSU wavefronts, instruction issue and L2 atomics, not DRAM).
// MODE 0: interior warp — every lane's pair lies inside every segment of the chunk of both streams
// MODE 1: edge warp — per-(t, lane) validity fixed for the item
// MODE 2: the chunk reaches iA — validity also depends on the stream and on the row j
//
// Row sums over the lanes (K'[j,k], J~[pr(iA,j)], J~[pr(iB,j)]) do not use shuffles: every row
// pair each lane parks its NV partial sums in the warp's shared-memory tile (conflict-free STS),
// and once DEPTH row pairs are parked lane r*NV+v adds the 32 partials of (row r, value v) with
// 16 conflict-free LDS.128 (row stride 34 doubles) in a fixed order and issues the RED.
template <int NSPIN, int KC, int DEPTH, int MODE, bool VECD>
__device__ __forceinline__ void packed_item(const double* __restrict__ eri_a,  // row (iA,0)
                                            const double* __restrict__ eri_b,  // row (iB,0)
                                            const double* __restrict__ D, size_t dstride,
                                            const double* __restrict__ dq2,
                                            double* __restrict__ Kp, double* __restrict__ Jp,
                                            double* __restrict__ sm,  // this warp's PkSmem block
                                            int n, int iA, bool has_b, int k0, int l0, int lane,
                                            uint32_t rot, int j_lo, int j_cnt) {
  using S = PkSmem<NSPIN, KC, DEPTH>;
  constexpr int NK = S::NK, NV = S::NV, DKN = S::DKN;
  const int iB = iA + 1;
  const int iT = has_b ? iB : iA;  // largest first index of the item
  const int kmax = min(k0 + KC - 1, iT);
  // lanes past the triangle edge (l > kmax, possibly l >= n) only feed zeros
  const bool l0_in = MODE == 0 || l0 <= kmax, l1_in = MODE == 0 || l0 + 1 <= kmax;
  const int lc = l0_in ? l0 : 0;
  const uint64_t EA = seg_off(iA), EB = seg_off(iB);  // elements of the full segments 0..i-1 of every row (i,*)

  // shared-memory addresses of this lane's slots (stage / value offsets are immediates below)
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t ring_a = sm_base + (S::RING + 2 * lane) * 8;        // + ((stage*2 + stream)*KC + t) * 512
  const uint32_t ring_dl_a = sm_base + (S::RING_DL + 2 * lane) * 8;  // + (stage*NSPIN + s) * 512
  const uint32_t ring_dk_a = sm_base + (S::RING_DK + lane) * 8;      // + stage*DKN*8   (lanes < NK + 2)
  const double2* ring = reinterpret_cast<const double2*>(sm + S::RING) + lane;
  const double2* ring_dl = reinterpret_cast<const double2*>(sm + S::RING_DL) + lane;
  const double* ring_dk = sm + S::RING_DK;
  double* ys = sm + S::TILE;

  // per-item constants
  double dA_l[NSPIN][2], dB_l[NSPIN][2], dA_k[NSPIN][KC], dB_k[NSPIN][KC], dq[KC][2];
  uint32_t off[KC];  // BYTE offset of this lane's pair inside a row, per k of the chunk (same for both streams)
  bool okA[KC], okB[KC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    const double* da = D + s * dstride + (size_t)iA * n;
    const double* db = D + s * dstride + (size_t)iT * n;
    dA_l[s][0] = __ldg(da + lc);
    dA_l[s][1] = __ldg(da + lc + 1);  // beyond-the-edge values only ever meet zero ERI entries
    dB_l[s][0] = __ldg(db + lc);
    dB_l[s][1] = __ldg(db + lc + 1);
  }
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    const int k = k0 + t;
    okA[t] = MODE == 0 || (k <= iA && l0 <= k);
    okB[t] = MODE == 0 || (has_b && k <= iB && l0 <= k);
    const int kc = k <= iT ? k : iT;
    const uint32_t o = (uint32_t)seg_off(kc) + (uint32_t)lc;
    off[t] = o * 8u;
    const double2 q2 =
        (okA[t] || okB[t]) ? __ldg(reinterpret_cast<const double2*>(dq2 + o)) : make_double2(0.0, 0.0);
    dq[t][0] = q2.x;
    dq[t][1] = q2.y;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      dA_k[s][t] = __ldg(D + s * dstride + (size_t)iA * n + kc);
      dB_k[s][t] = __ldg(D + s * dstride + (size_t)iT * n + kc);
    }
  }

  double a1A[NSPIN][KC], a1B[NSPIN][KC], a2A[NSPIN][2], a2B[NSPIN][2], jq[KC][2];
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    jq[t][0] = jq[t][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) a1A[s][t] = a1B[s][t] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) a2A[s][0] = a2A[s][1] = a2B[s][0] = a2B[s][1] = 0.0;

  // The item covers the rows j in [j_lo, j_hi) of both streams (stream A has no row j = iB).
  // Rows are visited in a rotated order that differs per (i, k-chunk), so that concurrently
  // running warps do not all RED into the same K'[j,:] row at the same time; the l-blocks of one
  // (i, k-chunk, row range) share the order (their reads of a row are neighbours in memory).
  const int rows = j_cnt, j_hi = j_lo + j_cnt;
  const int j_first = j_lo + (int)(rot % (uint32_t)rows);

  // ---- load side: row pair jl goes to the ring DEPTH-1 pairs ahead of its use (warp-uniform state) ----
  int jl = j_first;
  // row (i,j+1) starts E(i) + seg_len(j+1) elements after row (i,j); stream A's pointer is only
  // dereferenced while jl <= iA
  const char* rowA = reinterpret_cast<const char*>(eri_a + (uint64_t)jl * EA + seg_off(jl));
  const char* rowB = reinterpret_cast<const char*>(eri_b + (uint64_t)jl * EB + seg_off(jl));
  const uint32_t stepA = (uint32_t)EA * 8u, stepB = (uint32_t)EB * 8u;
  // per-lane source of the warp-uniform values: lanes < NK fetch D[j, k0 + t] of spin s, lane NK
  // 2Dtot[(iA,j)], lane NK+1 2Dtot[(iB,j)]:  src = dk_base + j * dk_step
  const double* dk_base;
  uint32_t dk_step;
  {
    const int s = lane / KC, t = lane - s * KC;
    dk_base = lane < NK ? D + s * dstride + k0 + t : (lane == NK ? dq2 + EA : dq2 + (has_b ? EB : EA));
    dk_step = lane < NK ? (uint32_t)n : 1u;
  }
  const double* dl_base = D + lc;
  // Everything a row pair needs travels global -> shared with cp.async (LDGSTS), so neither the
  // HBM latency of the ERI pairs nor the L2 latency of the density row (L1 cannot hold D) sits on
  // the critical path and no registers are tied up: ERI pairs and D[j, l-pair] go to slots only the
  // issuing thread reads back; the warp-uniform values are fetched by the first lanes and read
  // back as shared-memory broadcasts after a __syncwarp().
  auto issue_row = [&](const int stage) {  // `stage` is a compile-time constant after unrolling
    const bool rowA_in = jl <= iA;
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const uint32_t dstA = ring_a + ((stage * 2 + 0) * KC + t) * 512;
      const uint32_t dstB = ring_a + ((stage * 2 + 1) * KC + t) * 512;
      bool vA = rowA_in, vB = has_b;
      if (MODE != 0) {
        vA = vA && okA[t] && (MODE != 2 || k0 + t < iA || l0 <= jl);
        vB = vB && okB[t] && (MODE != 2 || k0 + t < iB || l0 <= jl);
      }
      cp_async16(dstA, rowA + off[t], vA ? 16u : 0u);  // 0 -> the 16 bytes are zero-filled
      cp_async16(dstB, rowB + off[t], vB ? 16u : 0u);
    }
    const double* dj_l = dl_base + (size_t)jl * n;  // D[j, l-pair]   (+ s*dstride per spin)
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const uint32_t dst = ring_dl_a + (stage * NSPIN + s) * 512;
      if (VECD) {  // n even: (j*n + even) is 16-byte aligned
        cp_async16(dst, dj_l + s * dstride);
      } else {
        cp_async8(dst, dj_l + s * dstride);
        cp_async8(dst + 8, dj_l + s * dstride + 1);
      }
    }
    if (lane < NK + 2) cp_async8(ring_dk_a + stage * DKN * 8, dk_base + (size_t)jl * dk_step);
    // advance to the next row in the rotated order
    if (jl + 1 == j_hi) {
      jl = j_lo;
      rowA = reinterpret_cast<const char*>(eri_a + (uint64_t)j_lo * EA + seg_off(j_lo));
      rowB = reinterpret_cast<const char*>(eri_b + (uint64_t)j_lo * EB + seg_off(j_lo));
    } else {
      const uint32_t c2 = (uint32_t)seg_len(jl + 1) * 8u;
      rowA += stepA + c2;
      rowB += stepB + c2;
      ++jl;
    }
  };

  // ---- consume side ----
  int jc = j_first;  // j of the row pair being consumed
  const int red_r = lane / NV, red_v = lane - red_r * NV;  // lane r*NV + v reduces (row r, value v) of the tile
  int j_tile = jc;                                         // j of the first parked row pair

  auto flush_tile = [&](int n_rows) {
    __syncwarp();
    if (red_r < n_rows) {  // (red_r < DEPTH always holds for lanes < DEPTH*NV)
      const double2* src = reinterpret_cast<const double2*>(ys + lane * YS);
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 16; q += 2) {
        const double2 u = src[q], v = src[q + 1];
        acc[0] += u.x;
        acc[1] += u.y;
        acc[2] += v.x;
        acc[3] += v.y;
      }
      const double tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      int jr = j_tile + red_r;
      if (jr >= j_hi) jr -= rows;
      if (red_v == NK) {
        if (jr <= iA) atomicAdd(Jp + EA + jr, tot);  // J~[pr(iA,j)]
      } else if (red_v == NK + 1) {
        if (has_b) atomicAdd(Jp + EB + jr, tot);  // J~[pr(iB,j)]
      } else {
        const int sp = red_v / KC, t = red_v - sp * KC;
        if (MODE == 0 || k0 + t <= iT) atomicAdd(Kp + sp * dstride + (size_t)jr * n + k0 + t, tot);  // K'[j,k]
      }
    }
    __syncwarp();
  };

  auto consume = [&](const int stage) {  // compile-time constant; the row pair is parked in tile row `stage`
    const double* dk = ring_dk + stage * DKN;  // same address in every lane: broadcast reads
    const double2 dp2 = *reinterpret_cast<const double2*>(dk + NK);  // 2Dtot[(iA,j)], 2Dtot[(iB,j)]
    double dj_l[NSPIN][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const double2 v = ring_dl[(stage * NSPIN + s) * 32];
      dj_l[s][0] = v.x;
      dj_l[s][1] = v.y;
    }
    double jpA = 0.0, jpB = 0.0;
    double z[NSPIN][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) z[s][0] = z[s][1] = 0.0;
    double* park = ys + stage * NV * YS + lane;
#pragma unroll
    for (int t = 0; t < KC; t += 2) {
      double2 wA[2], wB[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        wA[u] = ring[((stage * 2 + 0) * KC + t + u) * 32];
        wB[u] = ring[((stage * 2 + 1) * KC + t + u) * 32];
      }
      double2 dk2[NSPIN];
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) dk2[s] = *reinterpret_cast<const double2*>(dk + s * KC + t);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const double x0 = wA[u].x, x1 = wA[u].y, y0 = wB[u].x, y1 = wB[u].y;
        jpA = fma(x0, dq[t + u][0], jpA);
        jpA = fma(x1, dq[t + u][1], jpA);
        jpB = fma(y0, dq[t + u][0], jpB);
        jpB = fma(y1, dq[t + u][1], jpB);
        jq[t + u][0] = fma(x0, dp2.x, jq[t + u][0]);
        jq[t + u][0] = fma(y0, dp2.y, jq[t + u][0]);
        jq[t + u][1] = fma(x1, dp2.x, jq[t + u][1]);
        jq[t + u][1] = fma(y1, dp2.y, jq[t + u][1]);
#pragma unroll
        for (int s = 0; s < NSPIN; ++s) {
          const double djk = u ? dk2[s].y : dk2[s].x;
          a1A[s][t + u] = fma(x0, dj_l[s][0], a1A[s][t + u]);
          a1A[s][t + u] = fma(x1, dj_l[s][1], a1A[s][t + u]);
          a1B[s][t + u] = fma(y0, dj_l[s][0], a1B[s][t + u]);
          a1B[s][t + u] = fma(y1, dj_l[s][1], a1B[s][t + u]);
          a2A[s][0] = fma(x0, djk, a2A[s][0]);
          a2A[s][1] = fma(x1, djk, a2A[s][1]);
          a2B[s][0] = fma(y0, djk, a2B[s][0]);
          a2B[s][1] = fma(y1, djk, a2B[s][1]);
          // -> K'[j,k]: both streams feed the same output
          park[(s * KC + t + u) * YS] =
              fma(y1, dB_l[s][1], fma(y0, dB_l[s][0], fma(x1, dA_l[s][1], x0 * dA_l[s][0])));
          z[s][0] = fma(x0, dA_k[s][t + u], z[s][0]);
          z[s][0] = fma(y0, dB_k[s][t + u], z[s][0]);
          z[s][1] = fma(x1, dA_k[s][t + u], z[s][1]);
          z[s][1] = fma(y1, dB_k[s][t + u], z[s][1]);
        }
      }
    }
    park[NK * YS] = jpA;        // -> J~[pr(iA,j)]
    park[(NK + 1) * YS] = jpB;  // -> J~[pr(iB,j)]
    // K'[j,l]: lane-private, one RED per lane and row pair (K' entries are accumulated at
    // whichever of (a,b)/(b,a) coalesces: K = K' + K'^T does not care)
    double* kj = Kp + (size_t)jc * n + l0;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      if (l0_in) atomicAdd(kj + s * dstride, z[s][0]);
      if (l1_in) atomicAdd(kj + s * dstride + 1, z[s][1]);
    }
    jc = jc + 1 == j_hi ? j_lo : jc + 1;
  };

  auto commit = [] { asm volatile("cp.async.commit_group;" ::: "memory"); };
  auto wait_row = [] {
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();  // the uniform values were fetched by other lanes
  };

  // prologue: DEPTH-1 row pairs in flight (empty groups keep the group count uniform)
  int to_issue = rows;
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) {
    if (to_issue > 0) {
      issue_row(d);
      --to_issue;
    }
    commit();
  }
  int remaining = rows;
  while (remaining >= DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (to_issue > 0) {
        issue_row((d + DEPTH - 1) % DEPTH);
        --to_issue;
      }
      commit();
      wait_row();
      consume(d);
    }
    flush_tile(DEPTH);
    j_tile = jc;
    remaining -= DEPTH;
  }
  if (remaining > 0) {  // every remaining row pair is already in flight
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) {
      if (d < remaining) {
        commit();
        wait_row();
        consume(d);
      }
    }
    flush_tile(remaining);
  }

  // per-item outputs
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    double* ka = Kp + s * dstride + (size_t)n * iA;
    if (l0_in) atomicAdd(ka + l0, a2A[s][0]);  // K'[iA,l]
    if (l1_in) atomicAdd(ka + l0 + 1, a2A[s][1]);
    const double ra = warp_transpose_reduce<KC>(a1A[s], lane);
    const int t = transpose_reduce_index<KC>(lane);
    const bool owner = transpose_reduce_owner<KC>(lane);
    if (owner && k0 + t <= iA) atomicAdd(ka + k0 + t, ra);  // K'[iA,k]
    const double rb = warp_transpose_reduce<KC>(a1B[s], lane);
    if (has_b) {
      double* kb = ka + n;
      if (l0_in) atomicAdd(kb + l0, a2B[s][0]);  // K'[iB,l]
      if (l1_in) atomicAdd(kb + l0 + 1, a2B[s][1]);
      if (owner && k0 + t <= iB) atomicAdd(kb + k0 + t, rb);  // K'[iB,k]
    }
  }
#pragma unroll
  for (int t = 0; t < KC; ++t)
    if (okA[t] || okB[t]) {  // J~[q]; the pad slot only ever receives zeros
      double* jo = reinterpret_cast<double*>(reinterpret_cast<char*>(Jp) + off[t]);
      atomicAdd(jo, jq[t][0]);
      atomicAdd(jo + 1, jq[t][1]);
    }
}

// Persistent: one CTA of WARPS independent warps per SM; every warp draws work items from a
// global counter over a host-built list of the LIVE (pair, k-chunk, l-block) triples, heaviest
// pair first and the l-blocks of one (pair, k-chunk) adjacent.  (One-warp CTAs launched per triple
// left 20-35 % of the warp slots empty: dead triples beyond the triangle, block-launch latency and
// warps that could not be placed on a sub-partition with enough free registers.)
// item = (pair << 16 | k-chunk << 6 | l-block, j_lo << 16 | row count); pair z covers iA = i_hi - 2 - 2z,
// iB = iA + 1; the last pair of an odd-sized shard is the single index i_lo (stream A only).  The
// rows of a (pair, k-chunk, l-block) are cut into ranges so that every warp slot sees a few dozen
// items (one item over all <= n rows takes up to ~0.8 ms: too coarse for a 1/8 shard).
template <int NSPIN, int KC, int DEPTH, int WARPS, bool VECD>
__global__ void __launch_bounds__(32 * WARPS, 1)
jk_packed_kernel(const double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                 uint64_t shard_offset, const double* __restrict__ D, size_t dstride,
                 const double* __restrict__ dq2, double* __restrict__ Kp, double* __restrict__ Jp,
                 int n, int i_lo, int i_hi, const uint2* __restrict__ items, int n_items,
                 int* __restrict__ sched) {
  extern __shared__ __align__(16) double pk_smem[];
  const int lane = threadIdx.x & 31;
  double* sm = pk_smem + (threadIdx.x >> 5) * PkSmem<NSPIN, KC, DEPTH>::TOTAL;
  for (;;) {
    int idx = lane == 0 ? atomicAdd(sched, 1) : 0;
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n_items) break;
    const uint2 it2 = __ldg(items + idx);
    const uint32_t it = it2.x;
    const int j_lo = (int)(it2.y >> 16), j_cnt = (int)(it2.y & 0xffffu);
    int iA = i_hi - 2 - 2 * (int)(it >> 16);
    const bool has_b = iA >= i_lo;
    if (!has_b) iA = i_lo;
    const int iT = has_b ? iA + 1 : iA;
    const int kc = (it >> 6) & 0x3ff, lb = it & 0x3f;
    const int k0 = kc * KC;
    const int wl = 64 * lb;  // first l of this warp
    const int l0 = wl + 2 * lane;
    const double* eri_a = eri + (Btab[iA] - shard_offset);
    const double* eri_b = eri + (Btab[iT] - shard_offset);
    const uint32_t rot = (uint32_t)kc * 40503u + (uint32_t)iA * 17u;
    if (k0 + KC - 1 >= iA)
      packed_item<NSPIN, KC, DEPTH, 2, VECD>(eri_a, eri_b, D, dstride, dq2, Kp, Jp, sm, n, iA, has_b, k0, l0, lane, rot, j_lo, j_cnt);
    else if (wl + 63 <= k0)
      packed_item<NSPIN, KC, DEPTH, 0, VECD>(eri_a, eri_b, D, dstride, dq2, Kp, Jp, sm, n, iA, has_b, k0, l0, lane, rot, j_lo, j_cnt);
    else
      packed_item<NSPIN, KC, DEPTH, 1, VECD>(eri_a, eri_b, D, dstride, dq2, Kp, Jp, sm, n, iA, has_b, k0, l0, lane, rot, j_lo, j_cnt);
    __syncwarp();
  }
}

// dq2[seg_off(k) + l] = 2 * Dtot[k,l] for l <= k, pads = 0;  dpad = D with DPAD zeros appended
// to every spin block (lets the kernel read D[j, k0..k0+KC) without clamping)
__global__ void pack_density_kernel(const double* __restrict__ dtot, const double* __restrict__ dens,
                                    double* __restrict__ dq2, double* __restrict__ dpad, int n,
                                    int n_spin) {
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    const uint64_t base = seg_off(k);
    const int len = (int)seg_len(k + 1);
    for (int l = threadIdx.x; l < len; l += blockDim.x)
      dq2[base + l] = l <= k ? 2.0 * dtot[k + (size_t)n * l] : 0.0;
    for (int s = 0; s < n_spin; ++s) {
      const size_t src = (size_t)s * n * n + (size_t)k * n, dst = (size_t)s * ((size_t)n * n + DPAD) + (size_t)k * n;
      for (int l = threadIdx.x; l < n; l += blockDim.x) dpad[dst + l] = dens[src + l];
      if (k == n - 1)
        for (int l = threadIdx.x; l < DPAD; l += blockDim.x) dpad[dst + n + l] = 0.0;
    }
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
    for (int s = 0; s < n_spin; ++s)
      jk[(1 + s) * n2 + e] = Kp[s * (n2 + DPAD) + e] + Kp[s * (n2 + DPAD) + b + n * a];
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
    const uint64_t len = seg_off(i) + seg_len(j + 1);
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

// K3b — pack uploaded nu-slabs into the packed8 shard (load_integral_hdf5.jl:9-10: the tensor never
// has to exist as a whole, neither on the host nor on the device).  Slab i of the Julia-ordered tensor
// is X[a,b,c] = eri[a,b,c,i] = (ab|ci); by the 8-fold symmetry (ij|kl) = (kl|ji) = X[k,l,j], so every
// packed row (i,j), j <= i, comes from slab nu = i alone: out[R(i,j) + E(k) + l] = f * X[k + n l + n^2 j].
// One CTA = one 32x32 (k,l) tile of one (slab, j): read coalesced along k, transposed through shared
// memory, written coalesced along l — pre-scaled and with the segment pads zeroed.
__global__ void __launch_bounds__(256) pack_slabs_kernel(const double* __restrict__ slabs, int i_first,
                                                        double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                                                        uint64_t shard_offset, int n, int n_tiles) {
  __shared__ double tile[32][33];
  const int i = i_first + blockIdx.z, j = blockIdx.y;
  if (j > i) return;
  const int kt = blockIdx.x / n_tiles, lt = blockIdx.x % n_tiles;
  const int k0 = kt * 32, l0 = lt * 32;
  if (k0 > i || l0 > k0 + 31) return;  // beyond the last segment / above the triangle (incl. its pad)
  const double* X = slabs + (size_t)blockIdx.z * n * n * n + (size_t)j * n * n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {  // tile[l - l0][k - k0]
    const int k = k0 + tx, l = l0 + r;
    tile[r][tx] = (k < n && l < n) ? X[k + (size_t)n * l] : 0.0;
  }
  __syncthreads();
  double* row = eri + (Btab[i] + (uint64_t)j * seg_off(i) + seg_off(j) - shard_offset);
  const uint64_t p = tri(i) + j;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, l = l0 + tx;
    if (k > i) continue;
    const int lmax = k < i ? k : j;
    if (l >= (int)seg_len(lmax + 1)) continue;
    double v = 0.0;
    if (l <= lmax) {
      v = tile[tx][r];
      if (i == j) v *= 0.5;
      if (k == l) v *= 0.5;
      if (p == tri(k) + l) v *= 0.5;
    }
    row[seg_off(k) + l] = v;
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

// i-blocks owned by `rank`: contiguous ranges balanced by WORK, not bytes.  A warp spends about the
// same time on a row of its (k-chunk, l-block) tile whether its 64 l-lanes are all inside the
// triangle or not, so the cost of first index i is its number of (tile, row) steps,
// (i+1) * sum over k-chunks of the l-blocks they touch; balancing stored elements instead made the
// small-i shard 40 % slower than the large-i one (0.72 vs 0.51 ms at n_bf = 384 over 8 ranks).
void scfb_packed_split(int n, int rank, int n_ranks, int* i_lo, int* i_hi) {
  std::vector<long double> W(n + 1, 0.0L);
  for (int i = 0; i < n; ++i) {
    long double tiles = 0;
    for (int kc = 0; kc <= i / 8; ++kc) tiles += std::min(8 * kc + 7, i) / 64 + 1;
    W[i + 1] = W[i] + tiles * (i + 1);
  }
  auto bound = [&](int g) -> int {
    if (g <= 0) return 0;
    if (g >= n_ranks) return n;
    const long double target = W[n] * g / n_ranks;
    int lo = 0, hi = n;
    while (lo < hi) {  // smallest i with W[i] >= target
      const int mid = (lo + hi) / 2;
      if (W[mid] >= target) hi = mid; else lo = mid + 1;
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

// Live work items of this shard for k-chunk width kc_w, in the order the warps draw them: heaviest
// pair first, the l-blocks of one (pair, k-chunk, row range) adjacent.  Row ranges are sized so that
// every warp slot of the device draws >= ~16 items (a dynamic schedule's tail is one item long).
static int build_items(scfb_handle_s* h, int kc_w, int warps_per_sm, int slot) {
  std::vector<uint2> items;
  const int n_pairs = (h->i_hi - h->i_lo + 1) / 2;
  uint64_t warp_rows = 0;  // total (warp, row pair) steps of the shard
  for (int z = 0; z < n_pairs; ++z) {
    int iA = h->i_hi - 2 - 2 * z;
    const bool has_b = iA >= h->i_lo;
    if (!has_b) iA = h->i_lo;
    const int iT = has_b ? iA + 1 : iA;
    for (int kc = iT / kc_w; kc >= 0; --kc) warp_rows += (uint64_t)(std::min(kc * kc_w + kc_w - 1, iT) / 64 + 1) * (iT + 1);
  }
  const uint64_t slots = (uint64_t)(h->n_sm > 0 ? h->n_sm : 148) * warps_per_sm;
  // (an item costs ~5 us of prologue / ring fill / epilogue next to ~1.5 us per row pair: ranges
  // shorter than 64 rows were measured to cost a small shard 40 %)
  int j_max = (int)(warp_rows / (16 * slots));
  j_max = std::max(64, std::min(j_max, 4096));
  for (int z = 0; z < n_pairs; ++z) {
    int iA = h->i_hi - 2 - 2 * z;
    const bool has_b = iA >= h->i_lo;
    if (!has_b) iA = h->i_lo;
    const int iT = has_b ? iA + 1 : iA;
    const int rows = iT + 1, n_ranges = (rows + j_max - 1) / j_max;
    for (int r = 0; r < n_ranges; ++r) {
      const int j_lo = (int)((int64_t)rows * r / n_ranges), j_hi = (int)((int64_t)rows * (r + 1) / n_ranges);
      for (int kc = iT / kc_w; kc >= 0; --kc) {  // longest segments first
        const int kmax = std::min(kc * kc_w + kc_w - 1, iT);
        for (int lb = 0; 64 * lb <= kmax; ++lb)
          items.push_back(make_uint2((uint32_t)z << 16 | (uint32_t)kc << 6 | (uint32_t)lb,
                                     (uint32_t)j_lo << 16 | (uint32_t)(j_hi - j_lo)));
      }
    }
  }
  h->pk_n_items[slot] = (int)items.size();
  if (items.empty()) return SCFB_OK;
  cudaError_t e = cudaMalloc(&h->pk_items[slot], items.size() * sizeof(uint2));
  if (e == cudaSuccess)
    e = cudaMemcpy(h->pk_items[slot], items.data(), items.size() * sizeof(uint2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess)
    return scfb_fail(h, e == cudaErrorMemoryAllocation ? SCFB_ERR_OOM : SCFB_ERR_CUDA, "packed item list: %s",
                     cudaGetErrorString(e));
  return SCFB_OK;
}

// One tuning point: k-chunk width, ring depth and warps per SM.  The register file is 16 K
// registers per SM sub-partition: 255 registers per thread at 8 warps per SM, 168 at 12, 128 at 16.
template <int NSPIN, int KC, int DEPTH, int WARPS, bool VECD>
static int launch_packed_cfg(scfb_handle_s* h, double* Jp, double* Kp, size_t dstride, int slot) {
  const size_t smem = (size_t)WARPS * PkSmem<NSPIN, KC, DEPTH>::TOTAL * sizeof(double);
  auto kern = jk_packed_kernel<NSPIN, KC, DEPTH, WARPS, VECD>;
  if (h->n_sm == 0) {
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
    if (h->n_sm <= 0) h->n_sm = 148;
  }
  if (!h->pk_items[slot] && h->pk_n_items[slot] < 0) {  // first launch of this instantiation on this handle
    if (int rc = build_items(h, KC, WARPS, slot)) return rc;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "jk_packed_kernel attributes: %s", cudaGetErrorString(e));
  }
  if (h->n_sm == 0) {
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
    if (h->n_sm <= 0) h->n_sm = 148;
  }
  const int n_items = h->pk_n_items[slot];
  if (n_items == 0) return SCFB_OK;
  cudaError_t e = cudaMemsetAsync(h->sched, 0, sizeof(int), h->stream);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  const int grid = std::min(h->n_sm, (n_items + WARPS - 1) / WARPS);
  kern<<<grid, 32 * WARPS, smem, h->stream>>>(h->eri, h->pk_btab, h->pk_offset, h->pk_dpad, dstride, h->pk_dq2,
                                            Kp, Jp, h->n, h->i_lo, h->i_hi, h->pk_items[slot], n_items, h->sched);
  e = cudaGetLastError();
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "jk_packed_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

// Tuning points (same-box A/B at n_bf = 384, profiles/README.md): RHF KC = 8 at 8 warps per SM
// (255 registers) beats KC = 4 at 12 or 16 warps; UHF needs KC = 4 to keep its 2x K state in
// registers.
template <int NSPIN, bool VECD>
static int launch_packed(scfb_handle_s* h, double* Jp, double* Kp, size_t dstride) {
  if (NSPIN == 1) return launch_packed_cfg<1, 8, 2, 8, VECD>(h, Jp, Kp, dstride, 0);
  return launch_packed_cfg<2, 4, 2, 8, VECD>(h, Jp, Kp, dstride, 1);
}

int scfb_launch_jk_packed(scfb_handle_s* h, int n_spin, const double* d_density,
                          const double* d_total) {
  const int n = h->n;
  const size_t n2 = (size_t)n * n;
  const uint64_t PQ = seg_off(n);
  cudaError_t e;
  // zero the accumulators: [J~ (PQ) | K' (2 n^2)]
  e = cudaMemsetAsync(h->pk_acc, 0, (PQ + 2 * (n2 + DPAD)) * sizeof(double), h->stream);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  pack_density_kernel<<<n < 296 ? n : 296, 128, 0, h->stream>>>(d_total, d_density, h->pk_dq2, h->pk_dpad, n, n_spin);
  const size_t dstride = n2 + DPAD;
  h->launches += 1;
  if (h->timed) scfb_ev_record(h, 3);
  if (h->i_hi > h->i_lo) {
    double* Jp = h->pk_acc;
    double* Kp = h->pk_acc + PQ;
    int rc;
    if (n % 2 == 0)
      rc = n_spin == 1 ? launch_packed<1, true>(h, Jp, Kp, dstride) : launch_packed<2, true>(h, Jp, Kp, dstride);
    else
      rc = n_spin == 1 ? launch_packed<1, false>(h, Jp, Kp, dstride) : launch_packed<2, false>(h, Jp, Kp, dstride);
    if (rc) return rc;
  }
  if (h->timed) scfb_ev_record(h, 1);
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

// Pack `count` consecutive nu-slabs (device memory, slab i_first first) into the shard.
int scfb_launch_pack_slabs(scfb_handle_s* h, const double* d_slabs, int i_first, int count) {
  const int n_tiles = (h->n + 31) / 32;
  dim3 grid(n_tiles * n_tiles, i_first + count, count);  // (tile, j, slab); j > i exits at once
  pack_slabs_kernel<<<grid, 256, 0, h->stream>>>(d_slabs, i_first, h->eri, h->pk_btab, h->pk_offset, h->n, n_tiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "pack_slabs_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}
