// K5b — symmetric eigensolver for the small matrices of the device-resident SCF step
// (compute_orbitals.jl:14 asks LAPACK for ALL orbitals of every iterate; at n_bf <= 256 cuSOLVER's
// syevd / syevj spend 1-2.5 ms in ~200 tiny launches, 3-5x the Fock build of the same step at n_bf = 128).
//
// One-sided (Hestenes) Jacobi with the matrix resident in shared memory — of one CTA (n <= 32) or
// spread over a cluster of 8 (n <= 320) or 16 (n <= 512) CTAs that exchange column blocks through
// distributed shared memory:
//
//   U = (A / sigma + I) V0        V0 = eigenvectors of the previous SCF step (or I): U's columns are
//                                 then nearly orthogonal already and 2-4 sweeps suffice
//   sweeps of plane rotations of column pairs (p, q); a pair is rotated while |u_p . u_q| >
//   tol |u_p| |u_q| (tol ~ the rounding noise of the dot product), and the iteration stops after the
//   first sweep that met no pair above kQuad = 1e-8: Jacobi converges quadratically, so what that
//   sweep's own rotations leave behind is O(kQuad^2 / relative gap) — below tol
//   => U = (A / sigma + I) V with orthonormal V:  u_i = (lambda_i / sigma + 1) v_i
//   lambda_i = sigma (|u_i| - 1),  v_i = u_i / |u_i|,  sorted ascending (LAPACK order)
//
// sigma = 1.5 |A|_F makes A / sigma + I positive definite with condition number <= 5, so the column
// norms ARE the shifted eigenvalues (no sign ambiguity) and V is orthonormal to ~tol.  A round
// pairs all columns at once (round-robin tournament: m/2 disjoint pairs, m - 1 rounds per sweep), a
// pair is handled by TPP lanes of one warp (dot products by shuffles), and rounds are separated by
// ONE __syncthreads: no global memory, no launches inside the iteration.
#include <cooperative_groups.h>

#include "scfb_internal.h"

#include <cstdlib>

namespace {

constexpr int kMaxSweeps = 40;
constexpr double kQuad = 1e-8;

__device__ __forceinline__ double group_sum(double v, int tpp) {
  for (int o = tpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid = matrices; block = pairs * tpp threads (rounded up to a warp); dynamic smem = (m*m + 3m) doubles.
// U is [m][m] column-major, m = n rounded up to even (odd n: one phantom zero column and row — never
// rotated, its dot products are 0).  A lane owns the row pairs 2 (l + k tpp) + {0, 1}, k < RPT2, of its
// two columns; with RPT2 > 0 (n <= 128) it keeps them in registers between the dot product and the
// rotation, RPT2 = 0 re-reads shared memory in loops of any length.
// Squared column norms are cached in shared memory: recomputed from the data in round 0 of every sweep,
// updated by the Jacobi identities |x'|^2 = |x|^2 - t g, |y'|^2 = |y|^2 + t g in between, so a round
// costs ONE dot product (g = x . y) and one shuffle reduction.
template <int RPT2>
__global__ void __launch_bounds__(1024, 1)
jacobi_eig_kernel(const double* __restrict__ AV, const double* __restrict__ V0, double* __restrict__ Z,
                  double* __restrict__ W, int* __restrict__ info, int* __restrict__ sweeps_out, int n, int tpp,
                  double tol) {
  extern __shared__ __align__(16) double sm[];
  const int m = n + (n & 1);
  double* U = sm;
  double* nrm2 = sm + (size_t)m * m;  // [m] squared column norms (during the sweeps), then norms
  double* lam = nrm2 + m;
  __shared__ double red[32];
  __shared__ double sigma_s;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const size_t n2 = (size_t)n * n;
  const double* src = AV + blockIdx.x * n2;
  const double* v0 = V0 ? V0 + blockIdx.x * n2 : nullptr;

  double ss = 0.0;
  for (int i = tid; i < m * m; i += nt) {
    const int col = i / m, row = i - col * m;
    const double x = (col < n && row < n) ? src[(size_t)col * n + row] : 0.0;
    U[i] = x;
    ss += x * x;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    const double fro = sqrt(t);
    sigma_s = fro > 0.0 ? 1.5 * fro : 1.0;
  }
  __syncthreads();
  const double sigma = sigma_s, inv_sigma = 1.0 / sigma;
  for (int i = tid; i < m * m; i += nt) {
    const int col = i / m, row = i - col * m;
    // scaled by 1 / sigma: column norms end up in [1/3, 5/3] whatever the magnitude of A (the fp32
    // rotation angles below then cannot under- or overflow)
    if (col < n && row < n) U[i] = fma(U[i], inv_sigma, v0 ? v0[(size_t)col * n + row] : (row == col ? 1.0 : 0.0));
  }
  __syncthreads();

  const int pairs = m >> 1, half = m >> 1;  // half: double2 chunks per column
  const int g = tid / tpp, l = tid - g * tpp;
  const bool active = g < pairs;
  const double tol2 = tol * tol;
  int sweep = 0;
  bool converged = false;
  for (; sweep < kMaxSweeps; ++sweep) {
    int big = 0;
    for (int r = 0; r < m - 1; ++r) {
      // round-robin tournament: column m-1 stays, the others rotate
      int p = m - 1, q = r;
      if (g > 0) {
        p = r + g;
        if (p >= m - 1) p -= m - 1;
        q = r - g;
        if (q < 0) q += m - 1;
      }
      if (!active) p = q = 0;
      double2* up = reinterpret_cast<double2*>(U + (size_t)p * m);
      double2* uq = reinterpret_cast<double2*>(U + (size_t)q * m);
      double2 x[RPT2 > 0 ? RPT2 : 1], y[RPT2 > 0 ? RPT2 : 1];
      double a = 0.0, b = 0.0, c = 0.0;
      if (RPT2 > 0) {
#pragma unroll
        for (int k = 0; k < RPT2; ++k) {
          const int ch = l + k * tpp;
          const bool in = active && ch < half;
          x[k] = in ? up[ch] : make_double2(0.0, 0.0);
          y[k] = in ? uq[ch] : make_double2(0.0, 0.0);
          c = fma(x[k].x, y[k].x, c);
          c = fma(x[k].y, y[k].y, c);
          if (r == 0) {
            a = fma(x[k].x, x[k].x, a), a = fma(x[k].y, x[k].y, a);
            b = fma(y[k].x, y[k].x, b), b = fma(y[k].y, y[k].y, b);
          }
        }
      } else if (active) {
        for (int ch = l; ch < half; ch += tpp) {
          const double2 xv = up[ch], yv = uq[ch];
          c = fma(xv.x, yv.x, c);
          c = fma(xv.y, yv.y, c);
          if (r == 0) {
            a = fma(xv.x, xv.x, a), a = fma(xv.y, xv.y, a);
            b = fma(yv.x, yv.x, b), b = fma(yv.y, yv.y, b);
          }
        }
      }
      c = group_sum(c, tpp);
      if (r == 0) {  // (uniform) fresh norms once per sweep
        a = group_sum(a, tpp);
        b = group_sum(b, tpp);
      } else {
        a = nrm2[p];
        b = nrm2[q];
      }
      const double c2 = c * c, ab = a * b;
      const bool rot = active && c2 > tol2 * ab;
      if (rot) {
        big |= c2 > kQuad * kQuad * ab;
        // tan of the rotation angle: the smaller root of t^2 + 2 rho t - 1 = 0, rho = (b - a) / 2c.
        // The fp64 pipe is what bounds this kernel (divide + two square roots cost ~100 DFMA-class
        // instructions per warp and round), and the angle needs no more than single precision — a
        // rotation that leaves 1e-7 of g behind converges just the same; what must hold to fp64
        // accuracy is cs^2 + sn^2 = 1, so t comes from the fp32 units and cs = (1 + t^2)^(-1/2) from
        // an fp32 seed plus two Newton steps in fp64.
        const float rho = __fdividef((float)(b - a), 2.0f * (float)c);
        const float t32 = copysignf(1.0f, rho) * __frcp_rn(fabsf(rho) + sqrtf(fmaf(rho, rho, 1.0f)));
        const double t = (double)t32;
        const double xx = fma(t, t, 1.0);
        double cs = (double)rsqrtf((float)xx);
        cs = cs * fma(-0.5 * xx, cs * cs, 1.5);
        cs = cs * fma(-0.5 * xx, cs * cs, 1.5);
        const double sn = cs * t;
        if (RPT2 > 0) {
#pragma unroll
          for (int k = 0; k < RPT2; ++k) {
            const int ch = l + k * tpp;
            if (ch < half) {
              up[ch] = make_double2(fma(-sn, y[k].x, cs * x[k].x), fma(-sn, y[k].y, cs * x[k].y));
              uq[ch] = make_double2(fma(sn, x[k].x, cs * y[k].x), fma(sn, x[k].y, cs * y[k].y));
            }
          }
        } else {
          for (int ch = l; ch < half; ch += tpp) {
            const double2 xv = up[ch], yv = uq[ch];
            up[ch] = make_double2(fma(-sn, yv.x, cs * xv.x), fma(-sn, yv.y, cs * xv.y));
            uq[ch] = make_double2(fma(sn, xv.x, cs * yv.x), fma(sn, xv.y, cs * yv.y));
          }
        }
        a = fma(-t, c, a);
        b = fma(t, c, b);
      }
      __syncwarp();  // every lane has read the cached norms
      if (active && l == 0 && (rot || r == 0)) {
        nrm2[p] = a;
        nrm2[q] = b;
      }
      __syncthreads();
    }
    if (!__syncthreads_or(big)) {
      converged = true;
      ++sweep;
      break;
    }
  }

  // eigenvalues from the column norms, ascending order by ranking, normalised columns out
  for (int j = warp; j < n; j += nwarps) {
    double s = 0.0;
    for (int row = lane; row < n; row += 32) s = fma(U[(size_t)j * m + row], U[(size_t)j * m + row], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const double nj = sqrt(s);
      nrm2[j] = nj;
      lam[j] = sigma * (nj - 1.0);
    }
  }
  __syncthreads();
  double* Zb = Z + blockIdx.x * n2;
  double* Wb = W + (size_t)blockIdx.x * n;
  for (int j = warp; j < n; j += nwarps) {
    const double lj = lam[j];
    int rank = 0;
    for (int k = lane; k < n; k += 32) {
      const double lk = lam[k];
      rank += (lk < lj || (lk == lj && k < j)) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    const double inv = nrm2[j] > 0.0 ? 1.0 / nrm2[j] : 0.0;
    for (int row = lane; row < n; row += 32) Zb[(size_t)rank * n + row] = U[(size_t)j * m + row] * inv;
    if (lane == 0) Wb[rank] = lj;
  }
  if (tid == 0) {
    if (!converged) atomicExch(info, 1000 + sweep);
    if (sweeps_out) sweeps_out[blockIdx.x] = sweep;
  }
}

// ------------------------------------------------------------------------------------------
// The same iteration spread over a thread-block cluster of 8 CTAs (Brent-Luk ring).  One CTA reads
// and writes the whole matrix once per round, and at n = 128 that shared-memory traffic (256 KB a
// round, 127 rounds a sweep) is what a sweep costs (225 us measured).  Here the m columns are cut
// into 16 blocks of bw columns; CTA k holds two of them (slots L, R: 16 KB at n = 128).  A sweep is
//   * an intra-block tournament of both resident blocks (all pairs inside a block), then
//   * 15 steps of the block round-robin: all bw^2 cross pairs (L_i, R_j) in bw rounds of bw disjoint
//     pairs, then every CTA PUSHES its blocks (and their cached squared norms) into the neighbour's
//     spare buffer over distributed shared memory — L_k -> L_k+1, R_k -> R_k-1, R_0 -> L_1,
//     L_7 -> R_7, L_0 stays — and one cluster barrier flips the buffers.
// Every pair of columns meets exactly once per sweep, as in the one-CTA version; a round touches
// 2 bw columns per CTA, the 8 CTAs run their rounds concurrently.
// ------------------------------------------------------------------------------------------
// kCluster = 16 (non-portable cluster size, opt-in per launch) halves the block width again and
// carries the scheme to n = 448 (4 bw LD doubles per CTA must fit 226 KB).  TRIPLE keeps THREE block
// buffers per CTA instead of four (n <= 512: 3 x 64 KB): an exchange is then two phases — (1) L
// blocks (R of rank 0) go into the right neighbour's spare buffer, barrier, (2) R blocks go into the
// left neighbour's just-vacated L buffer (rank 0: its vacated R buffer), barrier — after which the
// roles rotate (L <- spare, R <- old L, spare <- old R) identically in every CTA but rank 0.
namespace cg = cooperative_groups;

template <int kCluster, int RPT2, int MAXT, bool TRIPLE>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_eig_cluster_kernel(const double* __restrict__ AV, const double* __restrict__ V0, double* __restrict__ Z,
                          double* __restrict__ W, int* __restrict__ info, int* __restrict__ sweeps_out, int n,
                          int bw, double tol) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int mat = blockIdx.x / kCluster;
  const int LD = n + (n & 1), half = LD >> 1;
  const int bsz = bw * LD;  // doubles per block
  extern __shared__ __align__(16) double sm[];
  constexpr int NBUF = TRIPLE ? 3 : 4;
  double* blk = sm;                         // [NBUF block buffers][bw][LD]  (double-buffered: [2][2 slots])
  double* nrm = sm + NBUF * (size_t)bsz;    // [NBUF][bw] squared norms travelling with the blocks
  double* lam_all = nrm + NBUF * bw;        // [2 kCluster bw] gathered column norms (end)
  __shared__ double red[32];
  __shared__ double part[kCluster];
  __shared__ int flags[kCluster];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const size_t n2 = (size_t)n * n;
  const double* src = AV + mat * n2;
  const double* v0 = V0 ? V0 + mat * n2 : nullptr;

  // ---- load this CTA's two blocks (block rank -> L, block 8 + rank -> R), Frobenius norm ----
  double ss = 0.0;
  for (int i = tid; i < 2 * bsz; i += nt) {
    const int slot = i / bsz, rem = i - slot * bsz;
    const int j = rem / LD, row = rem - j * LD;
    const int col = (slot * kCluster + rank) * bw + j;
    const double x = (col < n && row < n) ? src[(size_t)col * n + row] : 0.0;
    blk[i] = x;
    ss += x * x;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (tid < kCluster) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    cl.map_shared_rank(part, tid)[rank] = t;  // all-gather of the partial sums
  }
  cl.sync();
  double fro2 = 0.0;
  for (int r = 0; r < kCluster; ++r) fro2 += part[r];  // same order everywhere: one sigma
  const double fro = sqrt(fro2);
  const double sigma = fro > 0.0 ? 1.5 * fro : 1.0, inv_sigma = 1.0 / sigma;
  for (int i = tid; i < 2 * bsz; i += nt) {
    const int slot = i / bsz, rem = i - slot * bsz;
    const int j = rem / LD, row = rem - j * LD;
    const int col = (slot * kCluster + rank) * bw + j;
    if (col < n && row < n) blk[i] = fma(blk[i], inv_sigma, v0 ? v0[(size_t)col * n + row] : (row == col ? 1.0 : 0.0));
  }
  __syncthreads();

  const int g = warp;                 // one warp per column pair
  const int mt = bw + (bw & 1);       // players of the intra-block tournament (odd bw: one bye)
  const double tol2 = tol * tol;
  // physical buffers of the resident L and R blocks (and, TRIPLE, of the spare) — of the ranks >= 1;
  // rank 0 of the TRIPLE scheme keeps L in buffer 0 and receives every R in buffer 1
  int iL = 0, iR = 1, iS = 2;

  // rotate columns x, y (local shared memory) with cached squared norms *na, *nb
  auto do_pair = [&](double* xc, double* yc, double* na, double* nb, int& big) {
    double2* up = reinterpret_cast<double2*>(xc);
    double2* uq = reinterpret_cast<double2*>(yc);
    double2 x[RPT2], y[RPT2];
    double c = 0.0;
#pragma unroll
    for (int k = 0; k < RPT2; ++k) {
      const int ch = lane + k * 32;
      const bool in = ch < half;
      x[k] = in ? up[ch] : make_double2(0.0, 0.0);
      y[k] = in ? uq[ch] : make_double2(0.0, 0.0);
      c = fma(x[k].x, y[k].x, c);
      c = fma(x[k].y, y[k].y, c);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    double a = *na, b = *nb;
    const double c2 = c * c, ab = a * b;
    if (c2 > tol2 * ab) {
      big |= c2 > kQuad * kQuad * ab;
      const float rho = __fdividef((float)(b - a), 2.0f * (float)c);
      const float t32 = copysignf(1.0f, rho) * __frcp_rn(fabsf(rho) + sqrtf(fmaf(rho, rho, 1.0f)));
      const double t = (double)t32;
      const double xx = fma(t, t, 1.0);
      double cs = (double)rsqrtf((float)xx);
      cs = cs * fma(-0.5 * xx, cs * cs, 1.5);
      cs = cs * fma(-0.5 * xx, cs * cs, 1.5);
      const double sn = cs * t;
#pragma unroll
      for (int k = 0; k < RPT2; ++k) {
        const int ch = lane + k * 32;
        if (ch < half) {
          up[ch] = make_double2(fma(-sn, y[k].x, cs * x[k].x), fma(-sn, y[k].y, cs * x[k].y));
          uq[ch] = make_double2(fma(sn, x[k].x, cs * y[k].x), fma(sn, x[k].y, cs * y[k].y));
        }
      }
      __syncwarp();  // every lane has read the cached norms
      if (lane == 0) {
        *na = fma(-t, c, a);
        *nb = fma(t, c, b);
      }
    }
  };
  // block copy into (possibly remote) shared memory, with the block's cached norms
  auto push = [&](int src_buf, int dst_rank, int dst_buf) {
    double2* d = reinterpret_cast<double2*>(cl.map_shared_rank(blk, dst_rank) + (size_t)dst_buf * bsz);
    const double2* sp = reinterpret_cast<const double2*>(blk + (size_t)src_buf * bsz);
    for (int i = tid; i < bsz / 2; i += nt) d[i] = sp[i];
    if (tid < bw) cl.map_shared_rank(nrm, dst_rank)[dst_buf * bw + tid] = nrm[src_buf * bw + tid];
  };

  int sweep = 0;
  bool converged = false;
  for (; sweep < kMaxSweeps; ++sweep) {
    int big = 0;
    const bool fixed0 = TRIPLE && rank == 0;
    double* BL = blk + (size_t)(fixed0 ? 0 : iL) * bsz;
    double* BR = blk + (size_t)(fixed0 ? 1 : iR) * bsz;
    double* NL = nrm + (fixed0 ? 0 : iL) * bw;
    double* NR = nrm + (fixed0 ? 1 : iR) * bw;
    // fresh squared norms of the 2 bw resident columns
    for (int j = warp; j < 2 * bw; j += nwarps) {
      const double* col = (j < bw ? BL + (size_t)j * LD : BR + (size_t)(j - bw) * LD);
      double s = 0.0;
      for (int row = lane; row < LD; row += 32) s = fma(col[row], col[row], s);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) (j < bw ? NL[j] : NR[j - bw]) = s;
    }
    __syncthreads();
    // intra-block tournament: warps [0, mt/2) work on L, [mt/2, mt) on R
    {
      const int slot = g / (mt >> 1), pi = g - slot * (mt >> 1);
      double* B = slot ? BR : BL;
      double* N = slot ? NR : NL;
      for (int r = 0; r < mt - 1; ++r) {
        if (g < mt) {
          int p = mt - 1, q = r;
          if (pi > 0) {
            p = r + pi;
            if (p >= mt - 1) p -= mt - 1;
            q = r - pi;
            if (q < 0) q += mt - 1;
          }
          if (p < bw && q < bw) do_pair(B + (size_t)p * LD, B + (size_t)q * LD, N + p, N + q, big);
        }
        __syncthreads();
      }
    }
    for (int step = 0; step < 2 * kCluster - 1; ++step) {
      BL = blk + (size_t)(fixed0 ? 0 : iL) * bsz, BR = blk + (size_t)(fixed0 ? 1 : iR) * bsz;
      NL = nrm + (fixed0 ? 0 : iL) * bw, NR = nrm + (fixed0 ? 1 : iR) * bw;
      for (int r = 0; r < bw; ++r) {
        if (g < bw) {
          int q = g + r;
          if (q >= bw) q -= bw;
          do_pair(BL + (size_t)g * LD, BR + (size_t)q * LD, NL + g, NR + q, big);
        }
        __syncthreads();
      }
      if (step == 2 * kCluster - 2) {  // the sweep's verdict travels with the last exchange
        const int any = __syncthreads_or(big);
        if (tid < kCluster) cl.map_shared_rank(flags, tid)[rank] = any;
      }
      // round-robin of the blocks: L_0 stays, R_0 -> L_1, L_k -> L_k+1, R_k -> R_k-1, L_last -> R_last
      if (!TRIPLE) {  // into the other half of the double buffer, one barrier
        const int nL = iL ^ 2, nR = iR ^ 2;
        if (rank == 0) {
          push(iL, 0, nL);
          push(iR, 1, nL);
        } else {
          push(iL, rank == kCluster - 1 ? rank : rank + 1, rank == kCluster - 1 ? nR : nL);
          push(iR, rank - 1, nR);
        }
        cl.sync();
        iL = nL, iR = nR;
      } else {
        if (rank == 0)
          push(1, 1, iS);
        else if (rank < kCluster - 1)
          push(iL, rank + 1, iS);
        cl.sync();
        if (rank == 1)
          push(iR, 0, 1);
        else if (rank > 1)
          push(iR, rank - 1, iL);
        cl.sync();
        const int oL = iL, oR = iR;
        iL = iS, iR = oL, iS = oR;
      }
    }
    int any = 0;
    for (int r = 0; r < kCluster; ++r) any |= flags[r];
    if (!any) {
      converged = true;
      ++sweep;
      break;
    }
    // (the next verdict is written 15+ cluster barriers from here: nobody can still be reading this one)
  }

  // ---- norms of the resident columns, all-gathered; rank among all real columns; write out ----
  const bool fixed0 = TRIPLE && rank == 0;
  const double* BL = blk + (size_t)(fixed0 ? 0 : iL) * bsz;
  const double* BR = blk + (size_t)(fixed0 ? 1 : iR) * bsz;
  for (int j = warp; j < 2 * bw; j += nwarps) {
    const double* col = (j < bw ? BL + (size_t)j * LD : BR + (size_t)(j - bw) * LD);
    double s = 0.0;
    for (int row = lane; row < LD; row += 32) s = fma(col[row], col[row], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double nj = sqrt(s);  // 0 exactly for the phantom columns (real ones are >= 1/3)
    for (int r = lane; r < kCluster; r += 32) cl.map_shared_rank(lam_all, r)[rank * 2 * bw + j] = nj;
  }
  cl.sync();
  double* Zb = Z + mat * n2;
  double* Wb = W + (size_t)mat * n;
  const int total = 2 * kCluster * bw;
  for (int j = warp; j < 2 * bw; j += nwarps) {
    const int me = rank * 2 * bw + j;
    const double nj = lam_all[me];
    if (nj == 0.0) continue;  // (warp-uniform)
    int rk = 0;
    for (int k = lane; k < total; k += 32) {
      const double nk = lam_all[k];
      rk += (nk != 0.0 && (nk < nj || (nk == nj && k < me))) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    const double inv = 1.0 / nj;
    const double* col = (j < bw ? BL + (size_t)j * LD : BR + (size_t)(j - bw) * LD);
    for (int row = lane; row < n; row += 32) Zb[(size_t)rk * n + row] = col[row] * inv;
    if (lane == 0) Wb[rk] = sigma * (nj - 1.0);
  }
  if (tid == 0 && rank == 0) {
    if (!converged) atomicExch(info, 1000 + sweep);
    if (sweeps_out) sweeps_out[mat] = sweep;
  }
}

template <int CL, int RPT2, int MAXT, bool TRIPLE>
cudaError_t launch_jacobi_cluster(scfb_handle_s* h, int batch, int threads, size_t smem, const double* AV,
                                  const double* V0, double* Z, double* W, int* info, int* sweeps_out, int n, int bw,
                                  double tol) {
  auto kernel = jacobi_eig_cluster_kernel<CL, RPT2, MAXT, TRIPLE>;
  static thread_local int attr_dev = -1;
  if (attr_dev != h->device) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e == cudaSuccess && CL > 8) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    attr_dev = h->device;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * CL);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, AV, V0, Z, W, info, sweeps_out, n, bw, tol);
}

template <int RPT2>
cudaError_t launch_jacobi(scfb_handle_s* h, int batch, int threads, size_t smem, const double* AV, const double* V0,
                          double* Z, double* W, int* info, int* sweeps_out, int n, int tpp, double tol) {
  static thread_local int attr_dev = -1;
  if (attr_dev != h->device) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_eig_kernel<RPT2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024 - 1024);
    if (e != cudaSuccess) return e;
    attr_dev = h->device;
  }
  jacobi_eig_kernel<RPT2><<<batch, threads, smem, h->stream>>>(AV, V0, Z, W, info, sweeps_out, n, tpp, tol);
  return cudaGetLastError();
}

}  // namespace

constexpr int kMaxN8 = 320;   // 8-CTA cluster: 4 bw LD doubles per CTA, bw = ceil(m / 16) <= 20 warps of pairs
constexpr int kMaxN16 = 448;  // 16-CTA cluster, double-buffered: bw = ceil(m / 32) <= 14
int scfb_jacobi_eig_max_n() { return 512; }  // 16-CTA cluster, three block buffers: bw <= 16, 3 x 64 KB
int scfb_jacobi_eig_default_max_n() { return 512; }  // up to here it replaces cuSOLVER by default (dense.cu)

// Eigen-decomposition of `batch` symmetric n x n matrices A_b (column-major, consecutive).
//   AV : A_b V0_b when a warm start V0 is given, else A_b            (not modified)
//   V0 : orthonormal warm-start bases, or nullptr
//   Z  : eigenvectors (columns, ascending eigenvalue), W: eigenvalues; info: device int, set non-zero
//        when a matrix did not converge within the sweep limit (the caller zeroes it)
int scfb_launch_jacobi_eig(scfb_handle_s* h, int n, int batch, const double* AV, const double* V0, double* Z, double* W,
                           int* info, int* sweeps_out) {
  if (n < 1 || n > scfb_jacobi_eig_max_n())
    return scfb_fail(h, SCFB_ERR_ARG, "jacobi eigensolver: n=%d outside [1, %d]", n, scfb_jacobi_eig_max_n());
  const double tol = 2.0 * 2.220446049250313e-16 * sqrt((double)n);
  static const int single_max = [] {  // SCFB_EIG_CLUSTER_MIN=k: matrices of k or more rows go to the cluster kernel
    const char* v = std::getenv("SCFB_EIG_CLUSTER_MIN");
    return (v ? std::atoi(v) : 33) - 1;
  }();
  if (n > single_max || n > 168) {
    const int LD = n + (n & 1);
    const int cl = n <= kMaxN8 ? 8 : 16;
    const bool triple = n > kMaxN16;
    const int bw = (LD + 2 * cl - 1) / (2 * cl);
    const int mt = bw + (bw & 1);
    const int threads = 32 * (mt > bw ? mt : bw);
    const int nbuf = triple ? 3 : 4;
    const size_t smem = (nbuf * (size_t)bw * LD + nbuf * bw + 2 * cl * bw) * sizeof(double);
    const int rpt2 = (LD / 2 + 31) / 32;
    cudaError_t e;
#define SCFB_JAC(CL, R, T, TR) \
  launch_jacobi_cluster<CL, R, T, TR>(h, batch, threads, smem, AV, V0, Z, W, info, sweeps_out, n, bw, tol)
    if (cl == 8)
      e = rpt2 <= 1   ? SCFB_JAC(8, 1, 768, false)
          : rpt2 <= 2 ? SCFB_JAC(8, 2, 768, false)
          : rpt2 <= 3 ? SCFB_JAC(8, 3, 768, false)
                      : SCFB_JAC(8, 5, 768, false);
    else if (!triple)
      e = SCFB_JAC(16, 7, 512, false);
    else
      e = SCFB_JAC(16, 8, 512, true);
#undef SCFB_JAC
    if (e != cudaSuccess)
      return scfb_fail(h, SCFB_ERR_CUDA, "jacobi_eig_cluster_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return SCFB_OK;
  }
  const int m = n + (n & 1), pairs = m / 2;
  int tpp = 32;
  while (tpp > 1 && pairs * tpp > 1024) tpp >>= 1;
  const int threads = ((pairs * tpp + 31) / 32) * 32;
  const size_t smem = ((size_t)m * m + 3 * (size_t)m) * sizeof(double);
  const int rpt2 = (m / 2 + tpp - 1) / tpp;  // double2 chunks per lane and column
  cudaError_t e;
  if (rpt2 <= 1)
    e = launch_jacobi<1>(h, batch, threads, smem, AV, V0, Z, W, info, sweeps_out, n, tpp, tol);
  else if (rpt2 <= 2)
    e = launch_jacobi<2>(h, batch, threads, smem, AV, V0, Z, W, info, sweeps_out, n, tpp, tol);
  else if (rpt2 <= 4)
    e = launch_jacobi<4>(h, batch, threads, smem, AV, V0, Z, W, info, sweeps_out, n, tpp, tol);
  else
    e = launch_jacobi<0>(h, batch, threads, smem, AV, V0, Z, W, info, sweeps_out, n, tpp, tol);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "jacobi_eig_kernel launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}
