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
// Mapping (details at `packed_item`): one warp = one CTA owns (i, chunk of KC = 8 consecutive
// k <= i, 64 consecutive l); each lane owns an l-PAIR (16 bytes per k and row) and the warp loops
// over the rows j = 0..i, each of which contributes contiguous runs of the packed row.
//   * all per-row data (ERI pairs from HBM, the density row from L2) travels through a 4-row
//     cp.async (LDGSTS) ring in shared memory, three rows ahead of its use;
//   * sums whose output does not depend on j (K'[i,k], K'[i,l], J~[q]) stay in registers for the
//     whole item; K'[j,l] is lane-private (one RED per lane and row); the lane-reductions
//     (K'[j,k], J~[p]) go through a conflict-free shared-memory tile flushed every few rows;
//   * warps never synchronise with each other; cross-warp combination is fp64 RED.ADD into
//     zeroed accumulators, so the packed path is deterministic only up to summation order
//     (well inside the 1e-12 relative gate).
// Compile-time knobs SCFB_EXP_NO_Z / SCFB_EXP_NO_FLUSH remove one cost centre each (wrong
// results!) and exist only to reproduce the ablation in profiles/README.md.
#include "scfb_internal.h"

#include <cstdlib>
#include <vector>

namespace {

// One warp per CTA: the warps of an item never cooperate, and in the triangular iteration space
// multi-warp CTAs keep dead warps' slots (and their shared memory) pinned until the longest
// warp finishes; with LB = 32 every resident CTA is a live warp (measured 5.6 -> 4.9 ms).
#ifndef SCFB_LB
#define SCFB_LB 32
#endif
constexpr int LB = SCFB_LB;  // threads per CTA = LB/32 independent warps, 2*LB l values

__host__ __device__ __forceinline__ uint64_t tri(uint64_t i) { return i * (i + 1) / 2; }
__host__ __device__ __forceinline__ uint64_t seg_off(uint64_t k) { return tri(k) + (k + 1) / 2; }

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

// One work item of one warp.  D: (n*n + DPAD) x NSPIN symmetric, padded so that the KC-wide
// uniform reads D[j, k0..k0+KC) and the pair reads D[j, l0..l0+1] never leave the allocation;
// dq2: 2*Dtot in pair space, stored with the SAME padded segment offsets seg_off(k)+l (pads
// zero).  Kp: (n*n + DPAD) x NSPIN accumulators; Jp: accumulators indexed like dq2.
// MODE 0: interior warp — every lane's pair lies inside every segment of the chunk: no predicates
// MODE 1: edge warp — per-(t, lane) validity fixed for the item
// MODE 2: diagonal chunk (k == i inside) — validity also depends on the row j
//
// Row sums over the lanes (K'[j,k] and J~[pr(i,j)]) do not use shuffles: every row each lane
// parks its NV = NSPIN*KC + 1 partial sums in a per-warp shared-memory tile (conflict-free
// STS), and once RB = 32/NV rows are parked lane r*NV+v adds the 32 partials of (row r,
// value v) with 16 conflict-free LDS.128 (row stride 34 doubles) in a fixed order and issues
// the RED.  That is ~13 instructions per row instead of ~56 for shuffle butterflies.
constexpr int DPAD = 64;
constexpr int YS = 34;  // padded lane stride of the row-sum tile (doubles)
constexpr int DEPTH = 4;  // cp.async ring depth (rows)
// rows parked in the row-sum tile between flushes, and doubles per row of the uniform-D ring
__host__ __device__ constexpr int packed_rb(int nspin, int kc) {
  return 32 / (nspin * kc + 1) < 4 ? 32 / (nspin * kc + 1) : 4;
}
__host__ __device__ constexpr int packed_dkn(int nspin, int kc) { return (nspin * kc + 2) & ~1; }
__host__ __device__ constexpr size_t packed_smem_bytes(int nspin, int kc) {
  return (size_t)(LB / 32) * packed_rb(nspin, kc) * (nspin * kc + 1) * YS * sizeof(double)  // row-sum tiles
         + (size_t)DEPTH * kc * LB * 16                                                    // ERI ring
         + (size_t)DEPTH * nspin * LB * 16                                                 // D[j,l] ring
         + (size_t)DEPTH * (LB / 32) * packed_dkn(nspin, kc) * sizeof(double);             // D[j,k], 2Dtot ring
}

template <int NSPIN, int KC, int MODE, bool VECD>
__device__ __forceinline__ void packed_item(const double* __restrict__ eri_i,  // row (i,0)
                                            const double* __restrict__ D, size_t dstride,
                                            const double* __restrict__ dq2,
                                            double* __restrict__ Kp, double* __restrict__ Jp,
                                            double* __restrict__ ys,  // this warp's 32*YS tile
                                            double2* __restrict__ ring,     // ERI ring, this thread's slot 0
                                            double2* __restrict__ ring_dl,  // D[j,l-pair] ring, this thread's slot 0
                                            double* __restrict__ ring_dk,   // this warp's [DEPTH][DKN] block
                                            int n, int i, int k0, int l0, int lane, uint32_t rot) {
  constexpr int NV = NSPIN * KC + 1;
  constexpr int RB = packed_rb(NSPIN, KC);
  constexpr int DKN = packed_dkn(NSPIN, KC);
  const int kmax = min(k0 + KC - 1, i);
  // lanes past the triangle edge (l > kmax, possibly l >= n) only feed zeros
  const bool l0_in = MODE == 0 || l0 <= kmax, l1_in = MODE == 0 || l0 + 1 <= kmax;
  const int lc = l0_in ? l0 : 0;
  const uint64_t Ei = seg_off(i);  // elements of the full segments 0..i-1 of every row (i,*)

  // per-item constants
  double di_l[NSPIN][2], di_k[NSPIN][KC], dq[KC][2];
  uint32_t off[KC];  // element offset of this lane's pair inside a row, per k of the chunk
  bool ok[KC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    const double* di = D + s * dstride + (size_t)i * n;
    di_l[s][0] = __ldg(di + lc);
    di_l[s][1] = __ldg(di + lc + 1);  // beyond-the-edge values only ever meet zero ERI entries
  }
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    const int k = k0 + t;
    ok[t] = MODE == 0 || (k <= i && l0 <= k);
    const int kc = k <= i ? k : i;
    off[t] = (uint32_t)seg_off(kc) + (uint32_t)lc;
    const double2 q2 =
        ok[t] ? __ldg(reinterpret_cast<const double2*>(dq2 + off[t])) : make_double2(0.0, 0.0);
    dq[t][0] = q2.x;
    dq[t][1] = q2.y;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) di_k[s][t] = __ldg(D + s * dstride + (size_t)i * n + kc);
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
  // Everything a row needs travels global -> shared with cp.async (LDGSTS) into a DEPTH-deep
  // ring, DEPTH-1 rows ahead of its use, so neither the HBM latency of the ERI pairs nor the
  // L2 latency of the density row (L1 cannot hold D) sits on a row's critical path and no
  // registers are tied up:  ERI pairs and D[j, l-pair] go to slots only the issuing thread reads
  // back; the warp-uniform D[j, k0..k0+KC) and 2Dtot[(i,j)] are fetched by the first lanes and
  // read back as shared-memory broadcasts after a __syncwarp().
  auto issue_row = [&](int stage, const double* row, const double* dj0, const double* dqp, int j) {
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const bool valid = MODE == 0 ? true : (MODE == 2 ? (ok[t] && (k0 + t < i || l0 <= j)) : ok[t]);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (stage * KC + t) * LB);
      const uint32_t nbytes = valid ? 16u : 0u;  // 0 -> the 16 bytes are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(row + off[t]),
                   "r"(nbytes)
                   : "memory");
    }
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const double* src = dj0 + s * dstride + lc;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring_dl + (stage * NSPIN + s) * LB);
      if (VECD) {  // n even: (j*n + even) is 16-byte aligned
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8), "l"(src + 1) : "memory");
      }
    }
    if (lane <= NSPIN * KC) {
      const int s = lane / KC, t = lane - s * KC;
      const double* src = lane == NSPIN * KC ? dqp : dj0 + s * dstride + k0 + t;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring_dk + stage * DKN + lane);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
  };

  // Rows are visited in a rotated order that differs per work item, so that concurrently
  // running warps do not all RED into the same K'[j,:] row at the same time.
  const int j_first = (int)(rot % (uint32_t)rows);
  // load side (DEPTH-1 rows ahead), warp-uniform
  int j = j_first;
  const double* row = eri_i + (uint64_t)j * Ei + seg_off(j);
  const double* dj0 = D + (size_t)j * n;  // D[j, :]   (+ s*dstride per spin)
  const double* dqp = dq2 + Ei + j;       // 2 Dtot[(i,j)]
  // consume side
  int jc = j_first;                       // j of the row being consumed
  double* kj0 = Kp + (size_t)jc * n;      // K'[j, :]  (+ s*dstride per spin)
  int parked = 0;                         // rows parked in the tile
  int j_tile = jc;                        // j of the first parked row

  // lane r*NV + v reduces (row r, value v) of the tile
  const int red_r = lane / NV, red_v = lane - red_r * NV;
  auto flush_tile = [&](int n_rows) {
    __syncwarp();
    if (red_r < n_rows) {  // (red_r < RB always holds for lanes < RB*NV)
      const double2* src = reinterpret_cast<const double2*>(ys + lane * YS);
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double2 v = src[q];
        acc0 += v.x;
        acc1 += v.y;
      }
      const double tot = acc0 + acc1;
      int jr = j_tile + red_r;
      if (jr >= rows) jr -= rows;
      if (red_v == NV - 1) {
        atomicAdd(Jp + Ei + jr, tot);  // J~[pr(i,j)]
      } else {
        const int sp = red_v / KC, t = red_v - sp * KC;
        if (MODE == 0 || k0 + t <= i) atomicAdd(Kp + sp * dstride + (size_t)jr * n + k0 + t, tot);  // K'[j,k]
      }
    }
    __syncwarp();
  };

  auto consume = [&](int stage) {
    double2 w[KC];
#pragma unroll
    for (int t = 0; t < KC; ++t) w[t] = ring[(stage * KC + t) * LB];
    struct {
      double dj_l[NSPIN][2], dj_k[NSPIN][KC], dp2;
    } r;
    const double* dk = ring_dk + stage * DKN;  // same address in every lane: broadcast reads
    r.dp2 = dk[NSPIN * KC];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const double2 v = ring_dl[(stage * NSPIN + s) * LB];
      r.dj_l[s][0] = v.x;
      r.dj_l[s][1] = v.y;
#pragma unroll
      for (int t = 0; t < KC; t += 2) {
        const double2 u = *reinterpret_cast<const double2*>(dk + s * KC + t);
        r.dj_k[s][t] = u.x;
        r.dj_k[s][t + 1] = u.y;
      }
    }
    double jp0 = 0.0, jp1 = 0.0;
    double z[NSPIN][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) z[s][0] = z[s][1] = 0.0;
    double* park = ys + parked * NV * YS + lane;
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const double w0 = w[t].x, w1 = w[t].y;
      jp0 = fma(w0, dq[t][0], jp0);
      jp1 = fma(w1, dq[t][1], jp1);
      jq[t][0] = fma(w0, r.dp2, jq[t][0]);
      jq[t][1] = fma(w1, r.dp2, jq[t][1]);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        a1[s][t] = fma(w0, r.dj_l[s][0], a1[s][t]);
        a1[s][t] = fma(w1, r.dj_l[s][1], a1[s][t]);
        a2[s][0] = fma(w0, r.dj_k[s][t], a2[s][0]);
        a2[s][1] = fma(w1, r.dj_k[s][t], a2[s][1]);
        park[(s * KC + t) * YS] = fma(w1, di_l[s][1], w0 * di_l[s][0]);  // -> K'[j,k]
        z[s][0] = fma(w0, di_k[s][t], z[s][0]);
        z[s][1] = fma(w1, di_k[s][t], z[s][1]);
      }
    }
    park[(NV - 1) * YS] = jp0 + jp1;  // -> J~[pr(i,j)]
    // K'[j,l]: lane-private, one RED per lane and row (K' entries are accumulated at
    // whichever of (a,b)/(b,a) coalesces: K = K' + K'^T does not care)
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      double* kj = kj0 + s * dstride;
#ifndef SCFB_EXP_NO_Z
      if (l0_in) atomicAdd(kj + l0, z[s][0]);
      if (l1_in) atomicAdd(kj + l0 + 1, z[s][1]);
#else
      a2[s][0] += z[s][0] + z[s][1];
      (void)kj;
#endif
    }
#ifdef SCFB_EXP_NO_FLUSH
    parked = 0;
#endif
    if (++parked == RB) {
      flush_tile(RB);
      parked = 0;
      j_tile = jc + 1 == rows ? 0 : jc + 1;
    }
    // next consumed row
    if (jc + 1 == rows) {
      jc = 0;
      kj0 = Kp;
    } else {
      ++jc;
      kj0 += n;
    }
  };
  // row (i,j+1) starts E(i) + ceil2(j+1) elements after row (i,j); wraps to row (i,0)
  auto advance = [&]() {
    if (j + 1 == rows) {
      j = 0;
      row = eri_i;
      dj0 = D;
      dqp = dq2 + Ei;
    } else {
      row += Ei + ((j + 2) & ~1);
      ++j;
      dj0 += n;
      ++dqp;
    }
  };

  // prologue: DEPTH-1 rows in flight (empty groups keep the group count uniform)
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) {
    if (d < rows) {
      issue_row(d, row, dj0, dqp, j);
      advance();
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int step = 0; step < rows; ++step) {
    if (step + DEPTH - 1 < rows) {
      issue_row((step + DEPTH - 1) % DEPTH, row, dj0, dqp, j);
      advance();
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();  // the uniform values were fetched by other lanes
    consume(step % DEPTH);
  }
  if (parked) flush_tile(parked);

  // per-item outputs
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    double* ki = Kp + s * dstride + (size_t)n * i;
    if (l0_in) atomicAdd(ki + l0, a2[s][0]);  // K'[i,l]
    if (l1_in) atomicAdd(ki + l0 + 1, a2[s][1]);
    const double r = warp_transpose_reduce<KC>(a1[s], lane);
    const int t = transpose_reduce_index<KC>(lane);
    if (transpose_reduce_owner<KC>(lane) && k0 + t <= i) atomicAdd(ki + k0 + t, r);  // K'[i,k]
  }
#pragma unroll
  for (int t = 0; t < KC; ++t)
    if (ok[t]) {  // J~[q]; the pad slot only ever receives zeros
      atomicAdd(Jp + off[t], jq[t][0]);
      atomicAdd(Jp + off[t] + 1, jq[t][1]);
    }
}

// grid = (n_kchunks, n_i_owned, n_lblocks), block = LB threads = 4 independent warps.
template <int NSPIN, int KC, int MINB, bool VECD>
__global__ void __launch_bounds__(LB, MINB)
jk_packed_kernel(const double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                 uint64_t shard_offset, const double* __restrict__ D, size_t dstride,
                 const double* __restrict__ dq2, double* __restrict__ Kp, double* __restrict__ Jp,
                 int n, int i_lo) {
  extern __shared__ __align__(16) unsigned char pk_smem[];
  constexpr int TILE = packed_rb(NSPIN, KC) * (NSPIN * KC + 1) * YS;  // doubles per warp
  constexpr int DKN = packed_dkn(NSPIN, KC);
  double* ys_all = reinterpret_cast<double*>(pk_smem);                                // [warps][TILE]
  double2* ring_all = reinterpret_cast<double2*>(ys_all + (LB / 32) * TILE);          // [DEPTH][KC][LB]
  double2* ring_dl_all = ring_all + DEPTH * KC * LB;                                  // [DEPTH][NSPIN][LB]
  double* ring_dk_all = reinterpret_cast<double*>(ring_dl_all + DEPTH * NSPIN * LB);  // [warps][DEPTH][DKN]
  const int i = i_lo + blockIdx.y;
  const int k0 = blockIdx.x * KC;
  if (k0 > i) return;
  const int m = blockIdx.z * LB + threadIdx.x;  // l-pair index
  const int lane = threadIdx.x & 31;
  const int kmax = min(k0 + KC - 1, i);
  const int wl = 2 * (m & ~31);  // first l of this warp
  if (wl > kmax) return;         // whole warp beyond the triangle
  const int l0 = 2 * m;
  const double* eri_i = eri + (Btab[i] - shard_offset);
  double* ys = ys_all + (threadIdx.x >> 5) * TILE;
  double2* ring = ring_all + threadIdx.x;
  double2* ring_dl = ring_dl_all + threadIdx.x;
  double* ring_dk = ring_dk_all + (threadIdx.x >> 5) * DEPTH * DKN;
  const uint32_t rot =
      blockIdx.x * 40503u + blockIdx.z * 9973u + blockIdx.y * 17u + (threadIdx.x >> 5) * 7u;
  if (kmax == i)
    packed_item<NSPIN, KC, 2, VECD>(eri_i, D, dstride, dq2, Kp, Jp, ys, ring, ring_dl, ring_dk, n, i, k0, l0, lane, rot);
  else if (wl + 63 <= k0)
    packed_item<NSPIN, KC, 0, VECD>(eri_i, D, dstride, dq2, Kp, Jp, ys, ring, ring_dl, ring_dk, n, i, k0, l0, lane, rot);
  else
    packed_item<NSPIN, KC, 1, VECD>(eri_i, D, dstride, dq2, Kp, Jp, ys, ring, ring_dl, ring_dk, n, i, k0, l0, lane, rot);
}

// ---- mbarrier / bulk-copy (TMA) helpers ----------------------------------------------------
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
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------
// v5 (SCFB_PACKED_TMA=1): the one-warp work item of `packed_item`, fed by the TMA engine from a
// DIFFERENT warp.  ncu on v2 shows the LSU data pipe at 66 % with two warps per scheduler: per
// row and warp ~160 LSU cycles, 80 of them the LDGSTS writes of the ERI pairs and of the density
// row.  v4 let lane 0 of the consuming warp post bulk copies instead (cp.async.bulk, SASS UBLKCP)
// and lost: a bulk copy costs its issuer ~35 instructions / ~70 cycles
// (benchmarks/micro/bulk_copy_rate.cu), ten of them per row came out of the consumer's own row
// budget (7.0 vs 5.5 ms).  Here a CTA is WS_PAIRS independent (producer warp, consumer warp)
// pairs: the producer draws work items from a global counter (heaviest i first), publishes them
// to its consumer through a two-deep queue and posts, up to DEPTH rows ahead, one bulk copy per
// k-segment (<= 512 B: the consumer's 64 l values), one per spin for D[j, wl..wl+63] and one per
// spin for D[j, k0..k0+KC), all completing on the stage's "full" mbarrier; the consumer reads its
// operands out of the stage, hands it back ("empty" mbarrier) and does the math of `packed_item`.
// Pairs never synchronise with each other; the ring runs on across items.  setmaxnreg moves
// registers from the producer warpgroups to the consumer warpgroups.  Requires n even.
// ------------------------------------------------------------------------------------------
constexpr int WS_PAIRS = 8;
constexpr int WS_PRODUCER_REGS = 48, WS_CONSUMER_REGS = 208;  // 256 * (48 + 208) = 65536
__host__ __device__ constexpr int ws_stage_doubles(int nspin, int kc) {
  return kc * 64 + nspin * 64 + nspin * kc;
}
__host__ __device__ constexpr int ws_pair_doubles(int nspin, int kc) {
  return packed_rb(nspin, kc) * (nspin * kc + 1) * YS + DEPTH * ws_stage_doubles(nspin, kc);
}

struct WsItem {
  int i, k0, wl;
  uint32_t rot;
  bool valid;
};
// linear index -> item, heaviest rows first; kc fastest so that neighbours share D rows in L2
template <int KC>
__device__ __forceinline__ WsItem ws_decode(int idx, int n, int i_hi) {
  const int nk = (n + KC - 1) / KC, nl = (n + 63) / 64;
  const int per_i = nk * nl;
  WsItem it;
  it.i = i_hi - 1 - idx / per_i;
  const int r = idx % per_i;
  const int kc = r % nk, lw = r / nk;
  it.k0 = kc * KC;
  it.wl = lw * 64;
  it.rot = (uint32_t)kc * 40503u + (uint32_t)lw * 9973u + (uint32_t)it.i * 17u;
  it.valid = it.k0 <= it.i && it.wl <= min(it.k0 + KC - 1, it.i);
  return it;
}

// consumer side of one work item; `sc` is the pair's running stage counter
template <int NSPIN, int KC, int MODE>
__device__ __forceinline__ void ws_consume_item(const double* __restrict__ D, size_t dstride,
                                                const double* __restrict__ dq2,
                                                double* __restrict__ Kp, double* __restrict__ Jp,
                                                double* __restrict__ ys,
                                                const double* __restrict__ stages,
                                                uint64_t* __restrict__ full, uint64_t* __restrict__ empty,
                                                int n, int i, int k0, int wl, int lane, uint32_t rot,
                                                uint32_t& sc) {
  constexpr int NV = NSPIN * KC + 1;
  constexpr int RB = packed_rb(NSPIN, KC);
  constexpr int STG = ws_stage_doubles(NSPIN, KC);
  const int l0 = wl + 2 * lane;
  const int kmax = min(k0 + KC - 1, i);
  const bool l0_in = MODE == 0 || l0 <= kmax, l1_in = MODE == 0 || l0 + 1 <= kmax;
  const int lc = l0_in ? l0 : 0;
  const uint64_t Ei = seg_off(i);

  double di_l[NSPIN][2], di_k[NSPIN][KC], dq[KC][2];
  bool ok[KC];
#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    const double* di = D + s * dstride + (size_t)i * n;
    di_l[s][0] = __ldg(di + lc);
    di_l[s][1] = __ldg(di + lc + 1);  // beyond-the-edge values only ever meet zero ERI entries
  }
#pragma unroll
  for (int t = 0; t < KC; ++t) {
    const int k = k0 + t;
    ok[t] = MODE == 0 || (k <= i && l0 <= k);
    const int kc = k <= i ? k : i;
    const double2 q2 = ok[t] ? __ldg(reinterpret_cast<const double2*>(dq2 + seg_off(kc) + lc))
                             : make_double2(0.0, 0.0);
    dq[t][0] = q2.x;
    dq[t][1] = q2.y;
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) di_k[s][t] = __ldg(D + s * dstride + (size_t)i * n + kc);
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
  int jc = (int)(rot % (uint32_t)rows);  // same rotated row order as the producer
  double* kj0 = Kp + (size_t)jc * n;
  int parked = 0, j_tile = jc;
  double dp2 = __ldg(dq2 + Ei + jc);

  const int red_r = lane / NV, red_v = lane - red_r * NV;
  auto flush_tile = [&](int n_rows) {
    __syncwarp();
    if (red_r < n_rows) {
      const double2* src = reinterpret_cast<const double2*>(ys + lane * YS);
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double2 v = src[q];
        acc0 += v.x;
        acc1 += v.y;
      }
      const double tot = acc0 + acc1;
      int jr = j_tile + red_r;
      if (jr >= rows) jr -= rows;
      if (red_v == NV - 1) {
        atomicAdd(Jp + Ei + jr, tot);  // J~[pr(i,j)]
      } else {
        const int sp = red_v / KC, t = red_v - sp * KC;
        if (MODE == 0 || k0 + t <= i) atomicAdd(Kp + sp * dstride + (size_t)jr * n + k0 + t, tot);  // K'[j,k]
      }
    }
    __syncwarp();
  };

  for (int step = 0; step < rows; ++step) {
    const int jn = jc + 1 == rows ? 0 : jc + 1;
    const double dp2_next = __ldg(dq2 + Ei + jn);
    const uint32_t slot = sc % DEPTH;
    mbar_wait(full + slot, (sc / DEPTH) & 1);
    const double* st = stages + slot * STG;
    double2 w[KC];
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const double2 v = reinterpret_cast<const double2*>(st + t * 64)[lane];
      // bytes beyond the copied length are stale: select, never multiply
      const bool valid = MODE == 0 ? true : (MODE == 2 ? (ok[t] && (k0 + t < i || l0 <= jc)) : ok[t]);
      w[t] = valid ? v : make_double2(0.0, 0.0);
    }
    double dj_l[NSPIN][2], dj_k[NSPIN][KC];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      const double2 v = reinterpret_cast<const double2*>(st + KC * 64 + s * 64)[lane];
      dj_l[s][0] = v.x;
      dj_l[s][1] = v.y;
      const double* dk = st + KC * 64 + NSPIN * 64 + s * KC;  // broadcast reads
#pragma unroll
      for (int t = 0; t < KC; t += 2) {
        const double2 u = *reinterpret_cast<const double2*>(dk + t);
        dj_k[s][t] = u.x;
        dj_k[s][t + 1] = u.y;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);  // operands are in registers: hand the stage back
    ++sc;

    double jp0 = 0.0, jp1 = 0.0;
    double z[NSPIN][2];
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) z[s][0] = z[s][1] = 0.0;
    double* park = ys + parked * NV * YS + lane;
#pragma unroll
    for (int t = 0; t < KC; ++t) {
      const double w0 = w[t].x, w1 = w[t].y;
      jp0 = fma(w0, dq[t][0], jp0);
      jp1 = fma(w1, dq[t][1], jp1);
      jq[t][0] = fma(w0, dp2, jq[t][0]);
      jq[t][1] = fma(w1, dp2, jq[t][1]);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        a1[s][t] = fma(w0, dj_l[s][0], a1[s][t]);
        a1[s][t] = fma(w1, dj_l[s][1], a1[s][t]);
        a2[s][0] = fma(w0, dj_k[s][t], a2[s][0]);
        a2[s][1] = fma(w1, dj_k[s][t], a2[s][1]);
        park[(s * KC + t) * YS] = fma(w1, di_l[s][1], w0 * di_l[s][0]);  // -> K'[j,k]
        z[s][0] = fma(w0, di_k[s][t], z[s][0]);
        z[s][1] = fma(w1, di_k[s][t], z[s][1]);
      }
    }
    park[(NV - 1) * YS] = jp0 + jp1;  // -> J~[pr(i,j)]
#pragma unroll
    for (int s = 0; s < NSPIN; ++s) {
      double* kj = kj0 + s * dstride;
      if (l0_in) atomicAdd(kj + l0, z[s][0]);  // K'[j,l]
      if (l1_in) atomicAdd(kj + l0 + 1, z[s][1]);
    }
    if (++parked == RB) {
      flush_tile(RB);
      parked = 0;
      j_tile = jn;
    }
    jc = jn;
    kj0 = jn ? kj0 + n : Kp;
    dp2 = dp2_next;
  }
  if (parked) flush_tile(parked);

#pragma unroll
  for (int s = 0; s < NSPIN; ++s) {
    double* ki = Kp + s * dstride + (size_t)n * i;
    if (l0_in) atomicAdd(ki + l0, a2[s][0]);  // K'[i,l]
    if (l1_in) atomicAdd(ki + l0 + 1, a2[s][1]);
    const double r = warp_transpose_reduce<KC>(a1[s], lane);
    const int t = transpose_reduce_index<KC>(lane);
    if (transpose_reduce_owner<KC>(lane) && k0 + t <= i) atomicAdd(ki + k0 + t, r);  // K'[i,k]
  }
#pragma unroll
  for (int t = 0; t < KC; ++t)
    if (ok[t]) {  // J~[q]
      double* jo = Jp + seg_off(min(k0 + t, i)) + lc;
      atomicAdd(jo, jq[t][0]);
      atomicAdd(jo + 1, jq[t][1]);
    }
}

// grid = one CTA per SM; block = 2 * WS_PAIRS warps (consumers first); dynamic smem = per-pair
// [row-sum tile | DEPTH stages].
template <int NSPIN, int KC>
__global__ void __launch_bounds__(64 * WS_PAIRS, 1)
jk_packed_ws_kernel(const double* __restrict__ eri, const uint64_t* __restrict__ Btab,
                    uint64_t shard_offset, const double* __restrict__ D, size_t dstride,
                    const double* __restrict__ dq2, double* __restrict__ Kp,
                    double* __restrict__ Jp, int n, int i_lo, int i_hi, int* __restrict__ sched) {
  extern __shared__ __align__(128) double ws_smem[];
  __shared__ uint64_t full[WS_PAIRS][DEPTH], empty[WS_PAIRS][DEPTH];
  __shared__ uint64_t item_full[WS_PAIRS][2], item_empty[WS_PAIRS][2];
  __shared__ int item_q[WS_PAIRS][2];
  constexpr int TILE = packed_rb(NSPIN, KC) * (NSPIN * KC + 1) * YS;
  constexpr int STG = ws_stage_doubles(NSPIN, KC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp >= WS_PAIRS;
  const int p = producer ? warp - WS_PAIRS : warp;
  double* ys = ws_smem + (size_t)p * ws_pair_doubles(NSPIN, KC);
  double* stages = ys + TILE;
  const int per_i = ((n + KC - 1) / KC) * ((n + 63) / 64);
  const int total = (i_hi - i_lo) * per_i;

  if (threadIdx.x < WS_PAIRS) {
    for (int d = 0; d < DEPTH; ++d) {
      mbar_init(&full[threadIdx.x][d], 1);
      mbar_init(&empty[threadIdx.x][d], 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&item_full[threadIdx.x][q], 1);
      mbar_init(&item_empty[threadIdx.x][q], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (producer) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_PRODUCER_REGS));
    // All 32 lanes stay converged and each owns at most ONE copy job of a row (lane t < KC: the
    // k-segment k0+t; then per spin the D[j, l-window] piece and the D[j, k-chunk] piece): the
    // address/length arithmetic runs once, SIMT-parallel, and only the UBLKCP issue itself is
    // serialised over the active lanes (~8 instructions each) — a single posting lane needed
    // ~450 instructions per row and was the bottleneck (6.25 ms).
    static_assert(KC + 2 * NSPIN <= 32, "one copy job per producer lane");
    uint32_t sc = 0;
    for (int k = 0;; ++k) {
      // draw the next valid item (lane 0 draws, everybody decodes)
      int idx;
      WsItem it;
      it.valid = false;
      do {
        idx = lane == 0 ? atomicAdd(sched, 1) : 0;
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx < total) it = ws_decode<KC>(idx, n, i_hi);
      } while (idx < total && !it.valid);
      if (idx >= total) idx = -1;
      if (lane == 0) {
        if (k >= 2) mbar_wait(&item_empty[p][k & 1], ((k >> 1) & 1) ^ 1);
        item_q[p][k & 1] = idx;
        mbar_arrive(&item_full[p][k & 1]);  // release: the store above is visible to the waiter
      }
      __syncwarp();
      if (idx < 0) break;

      const int i = it.i, k0 = it.k0, wl = it.wl;
      const int rows = i + 1;
      const uint64_t Ei = seg_off(i);
      const double* eri_i = eri + (Btab[i] - shard_offset);
      const bool interior = k0 + KC - 1 < i && wl + 63 <= k0;  // MODE 0: every ERI copy is 512 bytes
      int j = (int)(it.rot % (uint32_t)rows);
      const double* row = eri_i + (uint64_t)j * Ei + seg_off(j);
      // this lane's job: ERI segment (lane < KC) or a density piece
      const int kk = k0 + lane;                       // segment index if lane < KC
      const int q = lane - KC, sp = q >> 1;           // density job if 0 <= q < 2*NSPIN
      const bool seg_job = lane < KC, dl_job = q >= 0 && q < 2 * NSPIN && !(q & 1),
                 dk_job = q >= 0 && q < 2 * NSPIN && (q & 1);
      const uint32_t soff = seg_job ? (uint32_t)seg_off(min(kk, i)) + (uint32_t)wl : 0u;
      const int my_dst = seg_job ? lane * 64 : (dl_job ? KC * 64 + sp * 64 : KC * 64 + NSPIN * 64 + sp * KC);
      const double* dsrc = D + (dl_job || dk_job ? sp * dstride : 0) + (dl_job ? wl : k0);  // + j*n per row
      for (int step = 0; step < rows; ++step) {
        const uint32_t slot = sc % DEPTH;
        mbar_wait(&empty[p][slot], ((sc / DEPTH) & 1) ^ 1);
        double* dst = stages + slot * STG + my_dst;
        uint64_t* bar = &full[p][slot];
        uint32_t bytes = 0;
        const double* src = row + soff;
        if (seg_job) {
          uint32_t len = 64;
          if (!interior) {
            // stored (padded) length of segment kk of row (i,j), clipped to this warp's l window
            int seg = kk > i ? 0 : ((kk < i ? kk : j) + 2) & ~1;
            seg -= wl;
            len = seg <= 0 ? 0u : (uint32_t)(seg < 64 ? seg : 64);
          }
          bytes = len * 8u;
        } else if (dl_job || dk_job) {
          src = dsrc + (size_t)j * n;
          bytes = dl_job ? 512u : KC * 8u;
        }
        if (bytes) bulk_g2s(dst, src, bytes, bar);
        const uint32_t all = __reduce_add_sync(0xffffffffu, bytes);
        // the one arrival of the phase; copies that landed earlier are already accounted for
        if (lane == 0) mbar_expect_tx(bar, all);
        ++sc;
        if (j + 1 == rows) {
          j = 0;
          row = eri_i;
        } else {
          row += Ei + ((j + 2) & ~1);
          ++j;
        }
      }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_CONSUMER_REGS));
  uint32_t sc = 0;
  for (int k = 0;; ++k) {
    mbar_wait(&item_full[p][k & 1], (k >> 1) & 1);
    const int idx = item_q[p][k & 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(&item_empty[p][k & 1]);
    if (idx < 0) break;
    const WsItem it = ws_decode<KC>(idx, n, i_hi);
    const int kmax = min(it.k0 + KC - 1, it.i);
    if (kmax == it.i)
      ws_consume_item<NSPIN, KC, 2>(D, dstride, dq2, Kp, Jp, ys, stages, full[p], empty[p], n, it.i, it.k0,
                                    it.wl, lane, it.rot, sc);
    else if (it.wl + 63 <= it.k0)
      ws_consume_item<NSPIN, KC, 0>(D, dstride, dq2, Kp, Jp, ys, stages, full[p], empty[p], n, it.i, it.k0,
                                    it.wl, lane, it.rot, sc);
    else
      ws_consume_item<NSPIN, KC, 1>(D, dstride, dq2, Kp, Jp, ys, stages, full[p], empty[p], n, it.i, it.k0,
                                    it.wl, lane, it.rot, sc);
  }
}

// dq2[seg_off(k) + l] = 2 * Dtot[k,l] for l <= k, pads = 0;  dpad = D with DPAD zeros appended
// to every spin block (lets the kernel read D[j, k0..k0+KC) without clamping)
__global__ void pack_density_kernel(const double* __restrict__ dtot, const double* __restrict__ dens,
                                    double* __restrict__ dq2, double* __restrict__ dpad, int n,
                                    int n_spin) {
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    const uint64_t base = seg_off(k);
    const int len = (k + 2) & ~1;
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
  e = cudaMemsetAsync(h->pk_acc, 0, (PQ + 2 * (n2 + DPAD)) * sizeof(double), h->stream);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  pack_density_kernel<<<n < 296 ? n : 296, 128, 0, h->stream>>>(d_total, d_density, h->pk_dq2, h->pk_dpad, n, n_spin);
  const size_t dstride = n2 + DPAD;
  h->launches += 1;
  if (h->timed) cudaEventRecord(h->ev[3], h->stream);
  const int n_i = h->i_hi - h->i_lo;
  if (n_i > 0) {
    double* Jp = h->pk_acc;
    double* Kp = h->pk_acc + PQ;
    // SCFB_PACKED_TMA=1 selects the warp-specialised bulk-copy (v5) kernel (RHF, even n); read per
    // launch so that one test process can cover both kernels
    const char* tma_env = std::getenv("SCFB_PACKED_TMA");
    const int use_tma = tma_env ? std::atoi(tma_env) : 0;
    static const int minb_sel = [] {  // tuning knob: SCFB_PACKED_MINB=3|4 (CTAs per SM, KC=4 RHF)
      const char* v = std::getenv("SCFB_PACKED_MINB");
      return v ? std::atoi(v) : 4;
    }();
    static const int kc_sel = [] {  // tuning knob: SCFB_PACKED_KC=4|8
      const char* v = std::getenv("SCFB_PACKED_KC");
      return v ? std::atoi(v) : 0;
    }();
    const int kc = kc_sel == 4 || kc_sel == 8 ? kc_sel : 8;
    if (use_tma && n % 2 == 0 && n_spin == 1 && kc == 8) {
      if (h->n_sm == 0) {
        cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
        if (h->n_sm <= 0) h->n_sm = 148;
      }
      const int n_sm = h->n_sm;
      const size_t sm_ = (size_t)WS_PAIRS * ws_pair_doubles(1, 8) * sizeof(double);
      cudaFuncSetAttribute(jk_packed_ws_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_);
      e = cudaMemsetAsync(h->sched, 0, sizeof(int), h->stream);
      if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
      jk_packed_ws_kernel<1, 8><<<n_sm, 64 * WS_PAIRS, sm_, h->stream>>>(
          h->eri, h->pk_btab, h->pk_offset, h->pk_dpad, dstride, h->pk_dq2, Kp, Jp, n, h->i_lo, h->i_hi,
          h->sched);
      e = cudaGetLastError();
      if (e != cudaSuccess)
        return scfb_fail(h, SCFB_ERR_CUDA, "jk_packed_ws_kernel launch: %s", cudaGetErrorString(e));
      h->launches += 1;
      if (h->timed) cudaEventRecord(h->ev[1], h->stream);
      unpack_jk_kernel<<<148 * 2, 256, 0, h->stream>>>(h->pk_acc, h->pk_acc + PQ, h->jk, n, n_spin);
      e = cudaGetLastError();
      if (e != cudaSuccess)
        return scfb_fail(h, SCFB_ERR_CUDA, "unpack_jk_kernel launch: %s", cudaGetErrorString(e));
      h->launches += 1;
      return SCFB_OK;
    }
    dim3 grid((n + kc - 1) / kc, n_i, (n + 2 * LB - 1) / (2 * LB));
#define SCFB_PK_LAUNCH(NS, KCV, MINB)                                                          \
  do {                                                                                          \
    const size_t sm_ = packed_smem_bytes(NS, KCV);                                              \
    if (n % 2 == 0) {                                                                           \
      cudaFuncSetAttribute(jk_packed_kernel<NS, KCV, MINB, true>,                               \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_);              \
      jk_packed_kernel<NS, KCV, MINB, true><<<grid, LB, sm_, h->stream>>>(                      \
          h->eri, h->pk_btab, h->pk_offset, h->pk_dpad, dstride, h->pk_dq2, Kp, Jp, n, h->i_lo); \
    } else {                                                                                    \
      cudaFuncSetAttribute(jk_packed_kernel<NS, KCV, MINB, false>,                              \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_);              \
      jk_packed_kernel<NS, KCV, MINB, false><<<grid, LB, sm_, h->stream>>>(                     \
          h->eri, h->pk_btab, h->pk_offset, h->pk_dpad, dstride, h->pk_dq2, Kp, Jp, n, h->i_lo); \
    }                                                                                           \
  } while (0)
    // resident CTAs per SM targeted by __launch_bounds__ (register cap = 65536 / (LB * MINB))
    constexpr int W = 128 / LB;  // CTAs per 128 threads
    (void)minb_sel;
    if (n_spin == 1 && kc == 8) SCFB_PK_LAUNCH(1, 8, 2 * W);
    else if (n_spin == 1 && minb_sel == 4) SCFB_PK_LAUNCH(1, 4, 4 * W);
    else if (n_spin == 1) SCFB_PK_LAUNCH(1, 4, 3 * W);
    else if (kc == 8) SCFB_PK_LAUNCH(2, 8, 2 * W);
    else SCFB_PK_LAUNCH(2, 4, 3 * W);
#undef SCFB_PK_LAUNCH
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
