// K4/K5 — the small dense steps that surround the Fock build in one SCF iteration, kept on
// the device so that only scalars and <= 6x6 systems cross PCIe per iteration:
//   Fock assembly + energies   /root/reference/src/parts/compute_fock_matrix.jl:8-58
//   Pulay error S D F - F D S  /root/reference/src/parts/compute_pulay_error.jl:17-25
//   orbitals  eigen(F, S)      /root/reference/src/parts/compute_orbitals.jl:4-19  (cuSOLVER Dsygvd)
//   density   C_occ C_occ'     /root/reference/src/parts/compute_density.jl:4-35
//   DIIS rows / extrapolation  /root/reference/src/algorithms/cDIIS.jl:123-127,
//                              /root/reference/src/algorithms/EDIIS.jl:87,
//                              /root/reference/src/parts/diis.jl:42-55
// These are O(n^3) / O(n^2) — <1 % of a build's bytes — so cuBLAS GEMMs are used for the
// products and tiny fixed-order (deterministic) reduction kernels for traces and dots.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cstdlib>
#include <vector>

#include "scfb_internal.h"

struct scfb_scf_state {
  int n = 0, n_orb = 0, n_elec[2] = {0, 0};
  int n_spin = 1;  // spin blocks of the Fock iterate / orbitals (1 = RHF and ROHF, 2 = UHF)
  int n_dens = 1;  // spin blocks of the density (compute_density.jl:24-27): 2 for UHF and ROHF
  bool restricted = true;
  bool rohf = false;  // restricted open shell: 2 densities, Roothaan effective 1-block Fock
  double e0 = 0.0;
  int cap = 5;  // n_diis_size

  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;

  double *S = nullptr, *H = nullptr;                       // n^2
  double *F = nullptr, *Fnew = nullptr, *D = nullptr;      // 2 n^2 each
  double *Dtot = nullptr, *G = nullptr, *E = nullptr;      // n^2, 2 n^2, 2 n^2
  double *C = nullptr, *eps = nullptr;                     // 2 n^2, 2 n
  double *Cprev = nullptr, *epsprev = nullptr, *Dprev = nullptr;  // previous Roothaan step
  double *T1 = nullptr, *T2 = nullptr, *Bw = nullptr;      // n^2 scratch
  double* R = nullptr;       // ROHF scratch: [Pc | Po | Pv | Fc | X | Feff], 6 n^2
  double* Lchol = nullptr;   // Cholesky factor of S (lower), computed once (ScfProblem.jl:48-49 TODO)
  int eig_mode = 0;          // 0 = Dsygvd per block (LAPACK path), 1 = cached Cholesky + Dsyevd,
                             // 2 = cached Cholesky + cuSOLVER Jacobi warm-started in the previous step's eigenbasis
                             // 3 = cached Cholesky + the shared-memory / cluster Jacobi of eig_jacobi.cu (n <= 512)
  double* Uj = nullptr;      // 2 n^2: (L^-1 F L^-T) Yprev, the input of mode 3
  int* jsweeps = nullptr;    // 2 ints: sweeps the last mode-3 solve took per spin block
  double* work_ev = nullptr;  // Dsyevd / Dsyevj workspace
  int lwork_ev = 0;
  double* Yprev = nullptr;   // 2 n^2: orthonormal eigenvectors of L^-1 F L^-T of the previous step
  bool have_y[2] = {false, false};
  int warm_steps[2] = {0, 0};
  syevjInfo_t jw = nullptr;
  double *work = nullptr;
  int lwork = 0;
  syevjInfo_t jacobi = nullptr;  // SCFB_EIG=jacobi: cusolverDnDsygvj instead of Dsygvd
  double* work_j = nullptr;
  int lwork_j = 0;
  int* dev_info = nullptr;
  double* scal = nullptr;    // device scalars
  double* h_scal = nullptr;  // pinned mirror
  double* coef = nullptr;    // device copy of DIIS coefficients

  // DIIS history: per spin, `cap` slots of (F, e, D); order[s] lists slot ids newest first
  double *Fh = nullptr, *Eh = nullptr, *Dh = nullptr;  // [2][cap][n^2]
  std::vector<int> order[2];
  bool have_density = false, have_fock = false;
};

namespace {

constexpr int kScal = 64;


#define SCFB_BLAS(h, call)                                                                  \
  do {                                                                                      \
    cublasStatus_t s_ = (call);                                                             \
    if (s_ != CUBLAS_STATUS_SUCCESS)                                                        \
      return scfb_fail((h), SCFB_ERR_CUSOLVER, "%s:%d %s -> cublas status %d", __FILE__, __LINE__, \
                       #call, (int)s_);                                                     \
  } while (0)
#define SCFB_SOLVER(h, call)                                                                \
  do {                                                                                      \
    cusolverStatus_t s_ = (call);                                                           \
    if (s_ != CUSOLVER_STATUS_SUCCESS)                                                      \
      return scfb_fail((h), SCFB_ERR_CUSOLVER, "%s:%d %s -> cusolver status %d", __FILE__,  \
                       __LINE__, #call, (int)s_);                                           \
  } while (0)

// out[k] = sum_ij A_k[i,j] * B_k[j,i]  (= tr(A_k B_k)); one CTA per k, fixed-order tree.
struct TracePair {
  const double* a;
  const double* b;
  int transpose_b;  // 1: tr(A B) = sum A[i,j] B[j,i];  0: <A,B> = sum A[i,j] B[i,j]
};
constexpr int kMaxPairs = 16;
struct TraceBatch {
  TracePair p[kMaxPairs];
};

__global__ void __launch_bounds__(1024) trace_kernel(TraceBatch batch, int n, double* out) {
  __shared__ double red[32];
  const TracePair pr = batch.p[blockIdx.x];
  const size_t n2 = (size_t)n * n;
  double acc = 0.0;
  // (i, j) of element e = i + j n advance with e (no division inside the loop); four elements per
  // trip so that the strided loads of the transposed operand are in flight together — the sum
  // still runs in element order
  unsigned i = threadIdx.x % n, j = threadIdx.x / n;
  const unsigned di = blockDim.x % n, dj = blockDim.x / n;
  for (size_t e = threadIdx.x; e < n2; e += 4 * (size_t)blockDim.x) {
    double av[4], bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t eu = e + u * (size_t)blockDim.x;
      const bool in = eu < n2;
      av[u] = in ? pr.a[eu] : 0.0;
      bv[u] = in ? (pr.transpose_b ? pr.b[j + (size_t)i * n] : pr.b[eu]) : 0.0;
      i += di, j += dj;
      if (i >= (unsigned)n) i -= n, ++j;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = fma(av[u], bv[u], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// EDIIS row entry: tr((F1 - Fj)(D1 - Dj)) = sum_ik (F1-Fj)[i,k] (D1-Dj)[k,i]
struct EdiisBatch {
  const double* f1;
  const double* d1;
  const double* fj[kMaxPairs];
  const double* dj[kMaxPairs];
};
__global__ void __launch_bounds__(1024) ediis_row_kernel(EdiisBatch b, int n, double* out) {
  __shared__ double red[32];
  const double* fj = b.fj[blockIdx.x];
  const double* dj = b.dj[blockIdx.x];
  const size_t n2 = (size_t)n * n;
  double acc = 0.0;
  unsigned i = threadIdx.x % n, k = threadIdx.x / n;
  const unsigned di = blockDim.x % n, dk = blockDim.x / n;
  for (size_t e = threadIdx.x; e < n2; e += 4 * (size_t)blockDim.x) {
    double fv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t eu = e + u * (size_t)blockDim.x;
      const size_t et = k + (size_t)i * n;
      const bool in = eu < n2;
      fv[u] = in ? b.f1[eu] - fj[eu] : 0.0;
      dv[u] = in ? b.d1[et] - dj[et] : 0.0;
      i += di, k += dk;
      if (i >= (unsigned)n) i -= n, ++k;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = fma(fv[u], dv[u], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// Dtot = 2 Da (one spin block) or Da + Db; G = 0
__global__ void prepare_build_kernel(const double* __restrict__ D, double* __restrict__ Dtot,
                                     double* __restrict__ G, size_t n2, int n_spin) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2;
       i += (size_t)gridDim.x * blockDim.x) {
    Dtot[i] = n_spin == 1 ? 2.0 * D[i] : D[i] + D[n2 + i];
    G[i] = 0.0;
    if (n_spin == 2) G[n2 + i] = 0.0;
  }
}

// F[:,:,s] = G[:,:,s] + h_core   (compute_fock_matrix.jl:58)
__global__ void fock_add_kernel(const double* __restrict__ G, const double* __restrict__ H,
                                double* __restrict__ F, size_t n2, int n_spin) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * n_spin;
       i += (size_t)gridDim.x * blockDim.x)
    F[i] = G[i] + H[i % n2];
}

// ROHF projector pieces (compute_fock_matrix.jl:105-110): Fc = (Fa + Fb)/2, Dab = Da - Db
__global__ void rohf_prepare_kernel(const double* __restrict__ F2, const double* __restrict__ D2,
                                    double* __restrict__ Fc, double* __restrict__ Dab, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    Fc[i] = (F2[i] + F2[n2 + i]) / 2;
    Dab[i] = D2[i] - D2[n2 + i];
  }
}
// Pv = I - Da S  given  Pv = Da S
__global__ void one_minus_kernel(double* __restrict__ P, int n) {
  const size_t n2 = (size_t)n * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
    P[i] = (i % n == i / n ? 1.0 : 0.0) - P[i];
}
// out = A + A'   (compute_fock_matrix.jl:122)
__global__ void add_transpose_kernel(const double* __restrict__ A, double* __restrict__ out, int n) {
  const size_t n2 = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e % n, j = e / n;
    out[e] = A[e] + A[j + i * n];
  }
}

struct LinComb {
  const double* m[kMaxPairs];
  int count;
};
// out = sum_i c[i] * m_i   in index order (parts/diis.jl:52-54 starts from zeros)
__global__ void lincomb_kernel(LinComb lc, const double* __restrict__ c, double* __restrict__ out,
                               size_t n2) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2;
       e += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int i = 0; i < lc.count; ++i) v += c[i] * lc.m[i][e];
    out[e] = v;
  }
}

inline int blocks_for(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b > 148 * 4 ? 148 * 4 : (b ? b : 1));
}

int gemm(scfb_handle_s* h, scfb_scf_state* st, cublasOperation_t ta, cublasOperation_t tb, int m,
         int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
         double* Cm, int ldc) {
  SCFB_BLAS(h, cublasDgemm(st->blas, ta, tb, m, n, k, &alpha, A, lda, B, ldb, &beta, Cm, ldc));
  h->launches += 1;
  return SCFB_OK;
}

// e = S D F - F D S   for one spin block (compute_pulay_error.jl:24)
int pulay_block(scfb_handle_s* h, scfb_scf_state* st, const double* F, const double* D, double* E) {
  const int n = st->n;
  if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, D, n, F, n, 0.0, st->T1, n)) return rc;
  if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, st->S, n, st->T1, n, 0.0, E, n)) return rc;
  if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, D, n, st->S, n, 0.0, st->T2, n)) return rc;
  return gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, -1.0, F, n, st->T2, n, 1.0, E, n);
}

// Roothaan effective Fock (compute_fock_matrix.jl:82-123) on resident matrices: F2 = (Fa, Fb),
// D2 = (Da, Db) -> Feff (one block).  Same association as the reference's expression:
//   Feff = Pc' Fc Pc / 2 + Po' Fb Pc + Po' Fc Po / 2 + Pv' Fc Pc + Pv' Fa Po + Pv' Fc Pv / 2;  result Feff + Feff'
int rohf_effective_fock(scfb_handle_s* h, scfb_scf_state* st, const double* F2, const double* D2, double* out) {
  const int n = st->n;
  const size_t n2 = (size_t)n * n;
  double *Pc = st->R, *Po = st->R + n2, *Pv = st->R + 2 * n2, *Fc = st->R + 3 * n2, *X = st->R + 4 * n2,
         *Fe = st->R + 5 * n2;
  const double *Fa = F2, *Fb = F2 + n2, *Da = D2, *Db = D2 + n2;
  rohf_prepare_kernel<<<blocks_for(n2), 256, 0, h->stream>>>(F2, D2, Fc, st->T1, n2);  // T1 = Da - Db
  h->launches += 1;
  auto mm = [&](cublasOperation_t ta, double alpha, const double* A, const double* B, double beta, double* Cm) {
    return gemm(h, st, ta, CUBLAS_OP_N, n, n, n, alpha, A, n, B, n, beta, Cm, n);
  };
  if (int rc = mm(CUBLAS_OP_N, 1.0, Db, st->S, 0.0, Pc)) return rc;      // Pc = Db S
  if (int rc = mm(CUBLAS_OP_N, 1.0, st->T1, st->S, 0.0, Po)) return rc;  // Po = (Da - Db) S
  if (int rc = mm(CUBLAS_OP_N, 1.0, Da, st->S, 0.0, Pv)) return rc;      // Pv = I - Da S
  one_minus_kernel<<<blocks_for(n2), 256, 0, h->stream>>>(Pv, n);
  h->launches += 1;
  if (int rc = mm(CUBLAS_OP_N, 1.0, Fc, Pc, 0.0, X)) return rc;   // X = Fc Pc
  if (int rc = mm(CUBLAS_OP_T, 0.5, Pc, X, 0.0, Fe)) return rc;   //   Pc' Fc Pc / 2
  if (int rc = mm(CUBLAS_OP_T, 1.0, Pv, X, 1.0, Fe)) return rc;   // + Pv' Fc Pc
  if (int rc = mm(CUBLAS_OP_N, 1.0, Fb, Pc, 0.0, X)) return rc;   // X = Fb Pc
  if (int rc = mm(CUBLAS_OP_T, 1.0, Po, X, 1.0, Fe)) return rc;   // + Po' Fb Pc
  if (int rc = mm(CUBLAS_OP_N, 1.0, Fc, Po, 0.0, X)) return rc;   // X = Fc Po
  if (int rc = mm(CUBLAS_OP_T, 0.5, Po, X, 1.0, Fe)) return rc;   // + Po' Fc Po / 2
  if (int rc = mm(CUBLAS_OP_N, 1.0, Fa, Po, 0.0, X)) return rc;   // X = Fa Po
  if (int rc = mm(CUBLAS_OP_T, 1.0, Pv, X, 1.0, Fe)) return rc;   // + Pv' Fa Po
  if (int rc = mm(CUBLAS_OP_N, 1.0, Fc, Pv, 0.0, X)) return rc;   // X = Fc Pv
  if (int rc = mm(CUBLAS_OP_T, 0.5, Pv, X, 1.0, Fe)) return rc;   // + Pv' Fc Pv / 2
  add_transpose_kernel<<<blocks_for(n2), 256, 0, h->stream>>>(Fe, out, n);
  h->launches += 1;
  return SCFB_OK;
}

// Orbitals of the iterate F (n_spin blocks) into st->C / st->eps (compute_orbitals.jl:14): LAPACK
// sygvd semantics (C' S C = I, ascending eps).  eig_mode 0: cusolverDnDsygvd per block.  eig_mode
// 1/2/3: S = L L' was factored once at setup, so a block costs A = L^-1 F L^-T (two Dtrsm), a
// standard symmetric eigensolve (Dsyevd per block; cuSOLVER's Jacobi in the previous step's
// eigenbasis; or, mode 3, ONE launch of eig_jacobi.cu for both spin blocks, warm-started the same
// way) and C = L^-T Y (one Dtrsm).
int solve_orbitals(scfb_handle_s* h, scfb_scf_state* st) {
  const int n = st->n, ns = st->n_spin;
  const size_t n2 = (size_t)n * n;
  if (st->eig_mode == 0) {
    for (int s = 0; s < ns; ++s) {
      double* Cs = st->C + s * n2;
      SCFB_CUDA(h, cudaMemcpyAsync(Cs, st->F + s * n2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      SCFB_CUDA(h, cudaMemcpyAsync(st->Bw, st->S, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      if (st->jacobi)
        SCFB_SOLVER(h, cusolverDnDsygvj(st->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                        CUBLAS_FILL_MODE_LOWER, n, Cs, n, st->Bw, n, st->eps + s * n,
                                        st->work_j, st->lwork_j, st->dev_info, st->jacobi));
      else
        SCFB_SOLVER(h, cusolverDnDsygvd(st->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                        CUBLAS_FILL_MODE_LOWER, n, Cs, n, st->Bw, n, st->eps + s * n,
                                        st->work, st->lwork, st->dev_info));
      h->launches += 1;
    }
    return SCFB_OK;
  }
  const double one = 1.0;
  SCFB_CUDA(h, cudaMemcpyAsync(st->C, st->F, n2 * ns * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  // (the spin blocks sit side by side: one left-solve covers both)
  SCFB_BLAS(h, cublasDtrsm(st->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT,
                           n, n * ns, &one, st->Lchol, n, st->C, n));  // A <- L^-1 F
  h->launches += 1;
  for (int s = 0; s < ns; ++s) {
    double* A = st->C + s * n2;
    SCFB_BLAS(h, cublasDtrsm(st->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                             n, n, &one, st->Lchol, n, A, n));  // A <- A L^-T
    h->launches += 1;
    if (st->eig_mode == 3) continue;  // both blocks are solved by one launch below
    if (st->eig_mode == 1) {
      SCFB_SOLVER(h, cusolverDnDsyevd(st->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n,
                                      st->eps + s * n, st->work_ev, st->lwork_ev, st->dev_info));
      h->launches += 1;
    } else {
      // The iterate changes little from one SCF step to the next: in the previous step's eigenbasis
      // it is nearly diagonal and the cyclic Jacobi method converges in a few sweeps.
      double* Yp = st->Yprev + s * n2;
      if (++st->warm_steps[s] % 16 == 0) st->have_y[s] = false;  // periodic cold solve: Y = Yp V V' ... stays orthonormal
      if (st->have_y[s]) {
        if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, A, n, Yp, n, 0.0, st->T1, n)) return rc;
        if (int rc = gemm(h, st, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, 1.0, Yp, n, st->T1, n, 0.0, st->T2, n)) return rc;
      } else {
        SCFB_CUDA(h, cudaMemcpyAsync(st->T2, A, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      }
      SCFB_SOLVER(h, cusolverDnDsyevj(st->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, st->T2, n,
                                      st->eps + s * n, st->work_ev, st->lwork_ev, st->dev_info, st->jw));
      h->launches += 1;
      if (st->have_y[s]) {
        if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, Yp, n, st->T2, n, 0.0, A, n)) return rc;  // Y = Yp V
      } else {
        SCFB_CUDA(h, cudaMemcpyAsync(A, st->T2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      }
      SCFB_CUDA(h, cudaMemcpyAsync(Yp, A, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      st->have_y[s] = true;
    }
    SCFB_BLAS(h, cublasDtrsm(st->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n,
                             n, &one, st->Lchol, n, A, n));  // C <- L^-T Y
    h->launches += 1;
  }
  if (st->eig_mode == 3) {
    // U = A Yprev: in the previous step's eigenbasis the columns are nearly orthogonal already
    const bool warm = st->have_y[0];
    if (warm) {
      for (int s = 0; s < ns; ++s)
        if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, 1.0, st->C + s * n2, n, st->Yprev + s * n2, n, 0.0,
                          st->Uj + s * n2, n))
          return rc;
    } else {
      SCFB_CUDA(h, cudaMemcpyAsync(st->Uj, st->C, n2 * ns * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    SCFB_CUDA(h, cudaMemsetAsync(st->dev_info, 0, sizeof(int), h->stream));
    if (int rc = scfb_launch_jacobi_eig(h, n, ns, st->Uj, warm ? st->Yprev : nullptr, st->C, st->eps, st->dev_info,
                                        st->jsweeps))
      return rc;
    SCFB_CUDA(h, cudaMemcpyAsync(st->Yprev, st->C, n2 * ns * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    st->have_y[0] = st->have_y[1] = true;
    SCFB_BLAS(h, cublasDtrsm(st->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n,
                             n * ns, &one, st->Lchol, n, st->C, n));  // C <- L^-T Y, both blocks
    h->launches += 1;
  }
  return SCFB_OK;
}

int ensure_dense_handles(scfb_handle_s* h) {
  if (!h->blas) {
    cublasHandle_t b = nullptr;
    SCFB_BLAS(h, cublasCreate(&b));
    h->blas = b;
  }
  if (!h->solver) {
    cusolverDnHandle_t sv = nullptr;
    SCFB_SOLVER(h, cusolverDnCreate(&sv));
    h->solver = sv;
  }
  cublasSetStream(static_cast<cublasHandle_t>(h->blas), h->stream);
  cusolverDnSetStream(static_cast<cusolverDnHandle_t>(h->solver), h->stream);
  return SCFB_OK;
}

int require_state(scfb_handle_s* h) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  if (!h->scf) return scfb_fail(h, SCFB_ERR_STATE, "scfb_scf_setup has not been called");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cublasSetStream(h->scf->blas, h->stream);
  cusolverDnSetStream(h->scf->solver, h->stream);
  return SCFB_OK;
}

}  // namespace

void scfb_scf_free(scfb_handle_s* h) {
  scfb_scf_state* st = h->scf;
  if (!st) return;
  for (double* p : {st->S, st->H, st->F, st->Fnew, st->D, st->Dtot, st->G, st->E, st->C, st->eps,
                    st->T1, st->T2, st->Bw, st->work, st->scal, st->coef, st->Fh, st->Eh, st->Dh,
                    st->Cprev, st->epsprev, st->Dprev, st->R, st->Lchol, st->work_ev, st->Yprev, st->Uj})
    cudaFree(p);
  cudaFree(st->jsweeps);
  if (st->jw) cusolverDnDestroySyevjInfo(st->jw);
  cudaFree(st->dev_info);
  cudaFree(st->work_j);
  if (st->jacobi) cusolverDnDestroySyevjInfo(st->jacobi);
  if (st->h_scal) cudaFreeHost(st->h_scal);
  delete st;
  h->scf = nullptr;
}

// cuBLAS / cuSOLVER handle creation costs tens of milliseconds: done once per scfb handle
void scfb_dense_handles_destroy(scfb_handle_s* h) {
  if (h->blas) cublasDestroy(static_cast<cublasHandle_t>(h->blas));
  if (h->solver) cusolverDnDestroy(static_cast<cusolverDnHandle_t>(h->solver));
  h->blas = h->solver = nullptr;
}

extern "C" {

int scfb_scf_setup(scfb_handle_t h, const double* overlap, const double* h_core, double e_0e,
                   int n_elec_a, int n_elec_b, int n_orb, int restricted, int n_diis_size) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, overlap && h_core, "null matrix pointer");
  const int n = h->n;
  SCFB_REQUIRE(h, n_orb >= 1 && n_orb <= n, "n_orb=%d outside [1, %d]", n_orb, n);
  // compute_density.jl:8-10 asserts strictly fewer electrons than orbitals
  SCFB_REQUIRE(h, n_elec_a >= 0 && n_elec_b >= 0 && n_elec_a < n_orb && n_elec_b < n_orb,
               "n_elec of alpha and beta should be less than n_orb");
  SCFB_REQUIRE(h, n_diis_size >= 1 && n_diis_size <= kMaxPairs, "n_diis_size must be in [1, %d]",
               kMaxPairs);
  SCFB_CUDA(h, cudaSetDevice(h->device));
  // A second SCF on the same handle (same sizes, e.g. RHF after the guess or a restart) reuses the
  // resident buffers, cuSOLVER workspaces and the overlap's Cholesky workspace: ~40 cudaMalloc /
  // cudaFree calls (each a device synchronisation) would otherwise cost more than a small SCF itself.
  const bool want_rohf = restricted && n_elec_a != n_elec_b;
  int want_eig;
  {
    const char* eig = std::getenv("SCFB_EIG");
    // default (same-box timings in profiles/README.md): the warm-started Jacobi kernels of eig_jacobi.cu up to
    // n_bf = 512, Dsyevd on the cached factor beyond.  RHF iteration at n_bf = 128: Dsygvd 3.5 ms, Dsyevd on the
    // cached factor 3.6, cuSOLVER's Jacobi warm-started 2.4, eig_jacobi.cu 1.3; at n_bf = 256: Dsyevd 8.3, eig_jacobi.cu 7.7;
    // at n_bf = 384 (packed8): Dsyevd 9.6, eig_jacobi.cu 7.8; at n_bf = 512 (packed8): 19.9 / 19.1.
    const int native_ok = n <= scfb_jacobi_eig_max_n();
    want_eig = !eig ? (n <= scfb_jacobi_eig_default_max_n() ? 3 : 1)
                    : (std::strcmp(eig, "syevd") == 0 ? 1
                       : (std::strcmp(eig, "warm") == 0 ? 2 : (std::strcmp(eig, "native") == 0 && native_ok ? 3 : 0)));
  }
  const bool want_jacobi = [] { const char* e = std::getenv("SCFB_EIG"); return e && std::strcmp(e, "jacobi") == 0; }();
  scfb_scf_state* st = h->scf;
  const bool reuse = st && st->n == n && st->cap == n_diis_size && st->eig_mode == want_eig &&
                     (st->R != nullptr || !want_rohf) && ((st->jacobi != nullptr) == want_jacobi);
  if (!reuse) {
    scfb_scf_free(h);
    st = new scfb_scf_state();
    h->scf = st;
  }
  st->n = n;
  st->n_orb = n_orb;
  st->n_elec[0] = n_elec_a;
  st->n_elec[1] = n_elec_b;
  st->restricted = restricted != 0;
  st->n_spin = st->restricted ? 1 : 2;
  st->rohf = st->restricted && n_elec_a != n_elec_b;
  st->n_dens = (!st->restricted || st->rohf) ? 2 : 1;  // compute_density.jl:24-27
  st->e0 = e_0e;
  st->cap = n_diis_size;
  const size_t n2 = (size_t)n * n;
  if (int rc = ensure_dense_handles(h)) return rc;
  st->blas = static_cast<cublasHandle_t>(h->blas);
  st->solver = static_cast<cusolverDnHandle_t>(h->solver);
  cublasSetStream(st->blas, h->stream);
  cusolverDnSetStream(st->solver, h->stream);
  if (!reuse) {
    auto alloc = [&](double** p, size_t count) { return cudaMalloc(p, count * sizeof(double)); };
    SCFB_CUDA(h, alloc(&st->S, n2));
    SCFB_CUDA(h, alloc(&st->H, n2));
    SCFB_CUDA(h, alloc(&st->F, 2 * n2));
    SCFB_CUDA(h, alloc(&st->Fnew, 2 * n2));
    SCFB_CUDA(h, alloc(&st->D, 2 * n2));
    SCFB_CUDA(h, alloc(&st->Dtot, n2));
    SCFB_CUDA(h, alloc(&st->G, 2 * n2));
    SCFB_CUDA(h, alloc(&st->E, 2 * n2));
    SCFB_CUDA(h, alloc(&st->C, 2 * n2));
    SCFB_CUDA(h, alloc(&st->eps, 2 * (size_t)n));
    SCFB_CUDA(h, alloc(&st->Cprev, 2 * n2));
    SCFB_CUDA(h, alloc(&st->epsprev, 2 * (size_t)n));
    SCFB_CUDA(h, alloc(&st->Dprev, 2 * n2));
    SCFB_CUDA(h, alloc(&st->T1, n2));
    SCFB_CUDA(h, alloc(&st->T2, n2));
    SCFB_CUDA(h, alloc(&st->Bw, n2));
    SCFB_CUDA(h, alloc(&st->scal, kScal));
    SCFB_CUDA(h, alloc(&st->coef, kMaxPairs));
    SCFB_CUDA(h, alloc(&st->Fh, 2 * (size_t)st->cap * n2));
    SCFB_CUDA(h, alloc(&st->Eh, 2 * (size_t)st->cap * n2));
    SCFB_CUDA(h, alloc(&st->Dh, 2 * (size_t)st->cap * n2));
    SCFB_CUDA(h, cudaMalloc(&st->dev_info, sizeof(int)));
    SCFB_CUDA(h, cudaMallocHost(&st->h_scal, kScal * sizeof(double)));
    SCFB_SOLVER(h, cusolverDnDsygvd_bufferSize(st->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                               CUBLAS_FILL_MODE_LOWER, n, st->T1, n, st->Bw, n, st->eps,
                                               &st->lwork));
    SCFB_CUDA(h, alloc(&st->work, st->lwork > 0 ? st->lwork : 1));
    if (want_rohf) SCFB_CUDA(h, alloc(&st->R, 6 * n2));
    st->eig_mode = want_eig;  // SCFB_EIG = sygvd | syevd | warm | jacobi (see solve_orbitals)
    if (st->eig_mode != 0) {
      SCFB_CUDA(h, alloc(&st->Lchol, n2));
      if (st->eig_mode == 3) {
        SCFB_CUDA(h, alloc(&st->Yprev, 2 * n2));
        SCFB_CUDA(h, alloc(&st->Uj, 2 * n2));
        SCFB_CUDA(h, cudaMalloc(&st->jsweeps, 2 * sizeof(int)));
      } else if (st->eig_mode == 1) {
        SCFB_SOLVER(h, cusolverDnDsyevd_bufferSize(st->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                                   st->T1, n, st->eps, &st->lwork_ev));
      } else {
        SCFB_CUDA(h, alloc(&st->Yprev, 2 * n2));
        SCFB_SOLVER(h, cusolverDnCreateSyevjInfo(&st->jw));
        SCFB_SOLVER(h, cusolverDnXsyevjSetTolerance(st->jw, 1e-15));
        SCFB_SOLVER(h, cusolverDnXsyevjSetMaxSweeps(st->jw, 100));
        SCFB_SOLVER(h, cusolverDnXsyevjSetSortEig(st->jw, 1));
        SCFB_SOLVER(h, cusolverDnDsyevj_bufferSize(st->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                                   st->T1, n, st->eps, &st->lwork_ev, st->jw));
      }
      SCFB_CUDA(h, alloc(&st->work_ev, st->lwork_ev > 0 ? st->lwork_ev : 1));
    }
  }
  st->order[0].clear();
  st->order[1].clear();
  st->have_density = st->have_fock = false;
  st->have_y[0] = st->have_y[1] = false;
  st->warm_steps[0] = st->warm_steps[1] = 0;
  SCFB_CUDA(h, cudaMemcpyAsync(st->S, overlap, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(st->H, h_core, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (st->eig_mode != 0) {  // S = L L' once: the overlap is constant over the SCF (ScfProblem.jl:48-49)
    SCFB_CUDA(h, cudaMemcpyAsync(st->Lchol, st->S, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    int lw = 0;
    SCFB_SOLVER(h, cusolverDnDpotrf_bufferSize(st->solver, CUBLAS_FILL_MODE_LOWER, n, st->Lchol, n, &lw));
    double* w = st->T2;  // scratch: n^2 doubles cover potrf's workspace at every size seen; else allocate
    if ((size_t)lw > n2) SCFB_CUDA(h, cudaMalloc(&w, (size_t)lw * sizeof(double)));
    cusolverStatus_t cs = cusolverDnDpotrf(st->solver, CUBLAS_FILL_MODE_LOWER, n, st->Lchol, n, w, lw, st->dev_info);
    int info = 0;
    cudaMemcpyAsync(&info, st->dev_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    if (w != st->T2) cudaFree(w);
    if (cs != CUSOLVER_STATUS_SUCCESS || ce != cudaSuccess || info != 0)
      return scfb_fail(h, SCFB_ERR_CUSOLVER, "Cholesky factorisation of the overlap failed (info=%d): S must be "
                       "symmetric positive definite", info);
  }
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  return SCFB_OK;
}

int scfb_scf_n_spin(scfb_handle_t h, int* n_spin) {
  if (int rc = require_state(h)) return rc;
  if (n_spin) *n_spin = h->scf->n_spin;
  return SCFB_OK;
}

int scfb_scf_n_densities(scfb_handle_t h, int* n_densities) {
  if (int rc = require_state(h)) return rc;
  if (n_densities) *n_densities = h->scf->n_dens;
  return SCFB_OK;
}

int scfb_scf_set_density(scfb_handle_t h, int n_spin, const double* density) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, density, "null density");
  SCFB_REQUIRE(h, n_spin == st->n_dens, "density has %d spin blocks, problem needs %d", n_spin,
               st->n_dens);
  const size_t n2 = (size_t)st->n * st->n;
  SCFB_CUDA(h, cudaMemcpyAsync(st->D, density, n2 * n_spin * sizeof(double), cudaMemcpyHostToDevice,
                               h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  st->have_density = true;
  return SCFB_OK;
}

int scfb_scf_set_fock(scfb_handle_t h, int n_spin, const double* fock) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, fock && n_spin == st->n_spin, "bad fock argument");
  const size_t n2 = (size_t)st->n * st->n;
  SCFB_CUDA(h, cudaMemcpyAsync(st->F, fock, n2 * n_spin * sizeof(double), cudaMemcpyHostToDevice,
                               h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  st->have_fock = true;
  return SCFB_OK;
}

// compute_fock_matrix(problem, density) on the device-resident density:
// Dtot, G = sum builders (here: this handle's J-K), energies, Fnew = G + h_core, E = SDF - FDS.
int scfb_scf_fock(scfb_handle_t h, double* energies4, double* error_norm) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  if (!st->have_density) return scfb_fail(h, SCFB_ERR_STATE, "no density: call scfb_scf_set_density");
  const int n = st->n, ns = st->n_spin, nd = st->n_dens;
  const size_t n2 = (size_t)n * n;
  prepare_build_kernel<<<blocks_for(n2), 256, 0, h->stream>>>(st->D, st->Dtot, st->G, n2, nd);
  h->launches += 1;
  if (int rc = scfb_fock_jk_dev(h, nd, st->D, st->Dtot, st->G)) return rc;
  if (st->rohf) {
    // compute_fock_matrix.jl:58-65: the two-block G + h_core goes through the Roothaan effective
    // Fock; the Pulay error is that of the effective Fock with the TOTAL density (one block)
    fock_add_kernel<<<blocks_for(n2 * 2), 256, 0, h->stream>>>(st->G, st->H, st->E, n2, 2);  // E = scratch (Fa, Fb)
    h->launches += 1;
    if (int rc = rohf_effective_fock(h, st, st->E, st->D, st->Fnew)) return rc;
    if (int rc = pulay_block(h, st, st->Fnew, st->Dtot, st->E)) return rc;
  } else {
    fock_add_kernel<<<blocks_for(n2 * ns), 256, 0, h->stream>>>(st->G, st->H, st->Fnew, n2, ns);
    h->launches += 1;
    for (int s = 0; s < ns; ++s)
      if (int rc = pulay_block(h, st, st->Fnew + s * n2, st->D + s * n2, st->E + s * n2)) return rc;
  }
  // scalars: [0] tr(h Dtot)  [1],[2] tr(G_s D_s)  [3],[4] <E_s, E_s>
  TraceBatch tb{};
  int k = 0;
  tb.p[k++] = {st->H, st->Dtot, 1};
  for (int s = 0; s < 2; ++s) {
    const int ss = s < nd ? s : 0;
    tb.p[k++] = {st->G + ss * n2, st->D + ss * n2, 1};
  }
  for (int s = 0; s < 2; ++s) {
    const int ss = s < ns ? s : 0;
    tb.p[k++] = {st->E + ss * n2, st->E + ss * n2, 0};
  }
  trace_kernel<<<k, 1024, 0, h->stream>>>(tb, n, st->scal);
  h->launches += 1;
  SCFB_CUDA(h, cudaGetLastError());
  SCFB_CUDA(h, cudaMemcpyAsync(st->h_scal, st->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost,
                               h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  const double* v = st->h_scal;
  const double e1 = v[0];
  const double e2 = nd == 1 ? v[1] : 0.5 * (v[1] + v[2]);  // compute_fock_matrix.jl:44-53
  if (energies4) {
    energies4[0] = st->e0;
    energies4[1] = e1;
    energies4[2] = e2;
    energies4[3] = st->e0 + e1 + e2;
  }
  if (error_norm) *error_norm = sqrt(ns == 1 ? v[3] : v[3] + v[4]);  // check_convergence.jl:18
  return SCFB_OK;
}

// roothan_step's first two lines: orbitals from the iterate F, then the density.
int scfb_scf_roothaan(scfb_handle_t h) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  if (!st->have_fock) return scfb_fail(h, SCFB_ERR_STATE, "no Fock iterate on the device");
  const int n = st->n, ns = st->n_spin, nd = st->n_dens;
  const size_t n2 = (size_t)n * n;
  // keep the previous step's orbitals/density: run_scf.jl:52-58 returns those on convergence
  SCFB_CUDA(h, cudaMemcpyAsync(st->Cprev, st->C, 2 * n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(st->epsprev, st->eps, 2 * (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(st->Dprev, st->D, 2 * n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (int rc = solve_orbitals(h, st)) return rc;
  int info = 0;
  SCFB_CUDA(h, cudaMemcpyAsync(&info, st->dev_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  static const bool verbose = std::getenv("SCFB_VERBOSE") != nullptr;
  int sweeps[2] = {0, 0};
  if (verbose && st->eig_mode == 3)
    SCFB_CUDA(h, cudaMemcpyAsync(sweeps, st->jsweeps, ns * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  if (info != 0) return scfb_fail(h, SCFB_ERR_CUSOLVER, "eigensolver failed: info=%d", info);
  if (verbose && st->eig_mode == 3) fprintf(stderr, "[scfb] jacobi_eig_kernel: %d / %d sweeps\n", sweeps[0], sweeps[1]);
  // D_s = C[:, 1:n_elec[s], s'] C[...]'   (compute_density.jl:16-32; ROHF: both densities from the one orbital block)
  for (int s = 0; s < nd; ++s) {
    const int nocc = st->n_elec[s];
    double* Ds = st->D + s * n2;
    if (nocc == 0) {
      SCFB_CUDA(h, cudaMemsetAsync(Ds, 0, n2 * sizeof(double), h->stream));
      continue;
    }
    const double* Cs = st->C + (s < ns ? s : ns - 1) * n2;
    if (int rc = gemm(h, st, CUBLAS_OP_N, CUBLAS_OP_T, n, n, nocc, 1.0, Cs, n, Cs, n, 0.0, Ds, n)) return rc;
  }
  st->have_density = true;
  return SCFB_OK;
}

// Push the newest (Fnew, E, D) of `spin` into the history (newest first, capacity n_diis_size:
// DiisState.jl:26-34) and return the new first rows of both DIIS matrices:
//   row_cdiis[j] = tr(e_1' e_j)                 (cDIIS.jl:126-127)
//   row_ediis[j] = tr((F_1 - F_j)(D_1 - D_j))   (EDIIS.jl:87)
int scfb_scf_diis_push(scfb_handle_t h, int spin, int* m_out, double* row_cdiis, double* row_ediis) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, spin >= 0 && spin < st->n_spin, "spin %d out of range", spin);
  const size_t n2 = (size_t)st->n * st->n;
  std::vector<int>& ord = st->order[spin];
  int slot;
  if ((int)ord.size() == st->cap) {  // full: the oldest entry is dropped, its slot reused
    slot = ord.back();
    ord.pop_back();
  } else {
    std::vector<bool> used(st->cap, false);
    for (int s : ord) used[s] = true;
    slot = 0;
    while (used[slot]) ++slot;
  }
  ord.insert(ord.begin(), slot);
  const size_t off = ((size_t)spin * st->cap + slot) * n2;
  SCFB_CUDA(h, cudaMemcpyAsync(st->Fh + off, st->Fnew + spin * n2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(st->Eh + off, st->E + spin * n2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(st->Dh + off, st->D + spin * n2, n2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  const int m = (int)ord.size();
  TraceBatch tb{};
  EdiisBatch eb{};
  eb.f1 = st->Fh + off;
  eb.d1 = st->Dh + off;
  for (int j = 0; j < m; ++j) {
    const size_t oj = ((size_t)spin * st->cap + ord[j]) * n2;
    tb.p[j] = {st->Eh + off, st->Eh + oj, 0};  // tr(e1' ej) = <e1, ej>
    eb.fj[j] = st->Fh + oj;
    eb.dj[j] = st->Dh + oj;
  }
  trace_kernel<<<m, 1024, 0, h->stream>>>(tb, st->n, st->scal);
  ediis_row_kernel<<<m, 1024, 0, h->stream>>>(eb, st->n, st->scal + kMaxPairs);
  h->launches += 2;
  SCFB_CUDA(h, cudaGetLastError());
  SCFB_CUDA(h, cudaMemcpyAsync(st->h_scal, st->scal, 2 * kMaxPairs * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  if (m_out) *m_out = m;
  for (int j = 0; j < m; ++j) {
    if (row_cdiis) row_cdiis[j] = st->h_scal[j];
    if (row_ediis) row_ediis[j] = st->h_scal[kMaxPairs + j];
  }
  return SCFB_OK;
}

// F[:,:,spin] <- sum_{i<m} c[i] * Fhist_i  (newest first; zero coefficients are skipped, as the
// masked entries are in parts/diis.jl:52-54)
int scfb_scf_diis_extrapolate(scfb_handle_t h, int spin, int m, const double* c) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, spin >= 0 && spin < st->n_spin, "spin %d out of range", spin);
  SCFB_REQUIRE(h, c && m >= 1 && m <= (int)st->order[spin].size(), "m=%d outside the history (%d)", m,
               (int)st->order[spin].size());
  const size_t n2 = (size_t)st->n * st->n;
  LinComb lc{};
  double cc[kMaxPairs];
  int k = 0;
  for (int i = 0; i < m; ++i) {
    if (c[i] == 0.0) continue;
    lc.m[k] = st->Fh + ((size_t)spin * st->cap + st->order[spin][i]) * n2;
    cc[k++] = c[i];
  }
  lc.count = k;
  std::memcpy(st->h_scal + 2 * kMaxPairs, cc, k * sizeof(double));
  SCFB_CUDA(h, cudaMemcpyAsync(st->coef, st->h_scal + 2 * kMaxPairs, kMaxPairs * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream));
  lincomb_kernel<<<blocks_for(n2), 256, 0, h->stream>>>(lc, st->coef, st->F + spin * n2, n2);
  h->launches += 1;
  SCFB_CUDA(h, cudaGetLastError());
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;  // h_scal staging is reused by the next call
  st->have_fock = true;
  return SCFB_OK;
}

// DiisState.jl:40-50 — pops count+1 OLDEST entries when count > 0
int scfb_scf_diis_purge(scfb_handle_t h, int spin, int count) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, spin >= 0 && spin < st->n_spin && count >= 0, "bad purge arguments");
  if (count == 0) return SCFB_OK;
  std::vector<int>& ord = st->order[spin];
  for (int i = 0; i < count + 1 && !ord.empty(); ++i) ord.pop_back();
  return SCFB_OK;
}

int scfb_scf_diis_reset(scfb_handle_t h) {
  if (int rc = require_state(h)) return rc;
  h->scf->order[0].clear();
  h->scf->order[1].clear();
  return SCFB_OK;
}

// Use the newest Fock build as the iterate without extrapolation (accelerator == nothing).
int scfb_scf_accept_fock(scfb_handle_t h) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  const size_t n2 = (size_t)st->n * st->n;
  SCFB_CUDA(h, cudaMemcpyAsync(st->F, st->Fnew, n2 * st->n_spin * sizeof(double),
                               cudaMemcpyDeviceToDevice, h->stream));
  st->have_fock = true;
  return SCFB_OK;
}

int scfb_scf_get(scfb_handle_t h, int what, double* host_out) {
  if (int rc = require_state(h)) return rc;
  scfb_scf_state* st = h->scf;
  SCFB_REQUIRE(h, host_out, "null output");
  const size_t n = st->n, n2 = n * n, ns = st->n_spin;
  const double* src = nullptr;
  size_t count = 0;
  switch (what) {
    case SCFB_GET_FOCK: src = st->F; count = n2 * ns; break;
    case SCFB_GET_FOCK_NEW: src = st->Fnew; count = n2 * ns; break;
    case SCFB_GET_DENSITY: src = st->D; count = n2 * st->n_dens; break;
    case SCFB_GET_ERROR: src = st->E; count = n2 * ns; break;
    case SCFB_GET_ORBCOEFF: src = st->C; count = n2 * ns; break;  // all n columns per spin
    case SCFB_GET_ORBEN: src = st->eps; count = n * ns; break;
    case SCFB_GET_ORBCOEFF_PREV: src = st->Cprev; count = n2 * ns; break;
    case SCFB_GET_ORBEN_PREV: src = st->epsprev; count = n * ns; break;
    case SCFB_GET_DENSITY_PREV: src = st->Dprev; count = n2 * st->n_dens; break;
    default: return scfb_fail(h, SCFB_ERR_ARG, "unknown quantity %d", what);
  }
  SCFB_CUDA(h, cudaMemcpyAsync(host_out, src, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (int rc_ = scfb_stream_sync_checked(h)) return rc_;
  return SCFB_OK;
}

// ---- stateless host-pointer twins of the reference's plain functions ---------------------
// compute_pulay_error(fock, density, overlap) -> error   (compute_pulay_error.jl:1-25)
int scfb_pulay_error(scfb_handle_t h, int n_spin, const double* fock, const double* density,
                     const double* overlap, double* error_out) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, (n_spin == 1 || n_spin == 2) && fock && density && overlap && error_out,
               "bad arguments");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  const int n = h->n;
  const size_t n2 = (size_t)n * n;
  scfb_scf_state tmp;  // borrows a cuBLAS handle + scratch just for this call
  double* buf = nullptr;
  SCFB_CUDA(h, cudaMalloc(&buf, (3 + 3 * (size_t)n_spin) * n2 * sizeof(double)));
  tmp.n = n;
  tmp.S = buf;
  tmp.T1 = buf + n2;
  tmp.T2 = buf + 2 * n2;
  double* F = buf + 3 * n2;
  double* D = F + n_spin * n2;
  double* E = D + n_spin * n2;
  int rc = ensure_dense_handles(h);
  if (rc) {
    cudaFree(buf);
    return rc;
  }
  tmp.blas = static_cast<cublasHandle_t>(h->blas);
  cudaMemcpyAsync(tmp.S, overlap, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(F, fock, n_spin * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(D, density, n_spin * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  for (int s = 0; s < n_spin && rc == SCFB_OK; ++s)
    rc = pulay_block(h, &tmp, F + s * n2, D + s * n2, E + s * n2);
  cudaMemcpyAsync(error_out, E, n_spin * n2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "pulay_error: %s", cudaGetErrorString(e));
  return SCFB_OK;
}

// compute_orbitals(problem, fock) (compute_orbitals.jl:4-19): per spin F C = S C eps with
// LAPACK sygvd semantics (C' S C = I, ascending eps); first n_orb columns are returned.
int scfb_sygvd(scfb_handle_t h, int n_spin, int n_orb, const double* fock, const double* overlap,
               double* orben_out, double* orbcoeff_out) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  const int n = h->n;
  SCFB_REQUIRE(h, (n_spin == 1 || n_spin == 2) && fock && overlap && orben_out && orbcoeff_out &&
                      n_orb >= 1 && n_orb <= n, "bad arguments");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  const size_t n2 = (size_t)n * n;
  if (int rc0 = ensure_dense_handles(h)) return rc0;
  cusolverDnHandle_t solver = static_cast<cusolverDnHandle_t>(h->solver);
  double* buf = nullptr;
  int* info_d = nullptr;
  int lwork = 0, rc = SCFB_OK, info = 0;
  cudaError_t e = cudaMalloc(&buf, (2 * n2 + n) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&info_d, sizeof(int));
  double *A = buf, *B = buf + n2, *W = buf + 2 * n2, *work = nullptr;
  if (e == cudaSuccess &&
      cusolverDnDsygvd_bufferSize(solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                  CUBLAS_FILL_MODE_LOWER, n, A, n, B, n, W, &lwork) != CUSOLVER_STATUS_SUCCESS)
    rc = scfb_fail(h, SCFB_ERR_CUSOLVER, "Dsygvd_bufferSize failed");
  if (e == cudaSuccess && rc == SCFB_OK) e = cudaMalloc(&work, (lwork > 0 ? lwork : 1) * sizeof(double));
  for (int s = 0; s < n_spin && rc == SCFB_OK && e == cudaSuccess; ++s) {
    cudaMemcpyAsync(A, fock + s * n2, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(B, overlap, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (cusolverDnDsygvd(solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                         n, A, n, B, n, W, work, lwork, info_d) != CUSOLVER_STATUS_SUCCESS)
      rc = scfb_fail(h, SCFB_ERR_CUSOLVER, "Dsygvd failed to launch");
    h->launches += 1;
    cudaMemcpyAsync(&info, info_d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(orben_out + (size_t)s * n_orb, W, n_orb * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(orbcoeff_out + (size_t)s * n * n_orb, A, (size_t)n * n_orb * sizeof(double),
                    cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
    if (rc == SCFB_OK && info != 0) rc = scfb_fail(h, SCFB_ERR_CUSOLVER, "Dsygvd: info=%d", info);
  }
  cudaFree(work);
  cudaFree(info_d);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return scfb_fail(h, e == cudaErrorMemoryAllocation ? SCFB_ERR_OOM : SCFB_ERR_CUDA,
                                         "sygvd: %s", cudaGetErrorString(e));
  return SCFB_OK;
}

// Stateless twin of the small-matrix eigensolver the device-resident SCF step uses (eig_jacobi.cu):
// `batch` symmetric n x n matrices (column-major, consecutive), optional orthonormal warm-start
// bases v0 (the eigenvectors of a nearby matrix).  Eigenvalues ascending, eigenvectors in columns.
int scfb_syev_jacobi(scfb_handle_t h, int n, int batch, const double* a, const double* v0, double* w_out,
                     double* z_out, int* sweeps_out) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, a && w_out && z_out && batch >= 1 && batch <= 2, "bad arguments");
  SCFB_REQUIRE(h, n >= 1 && n <= scfb_jacobi_eig_max_n(), "n=%d outside [1, %d]", n, scfb_jacobi_eig_max_n());
  SCFB_CUDA(h, cudaSetDevice(h->device));
  if (int rc0 = ensure_dense_handles(h)) return rc0;
  const size_t n2 = (size_t)n * n;
  double* buf = nullptr;
  int* idev = nullptr;
  SCFB_CUDA(h, cudaMalloc(&buf, (4 * n2 + n) * batch * sizeof(double)));
  cudaError_t e = cudaMalloc(&idev, 3 * sizeof(int));
  double *A = buf, *V = A + batch * n2, *U = V + batch * n2, *Z = U + batch * n2, *W = Z + batch * n2;
  int rc = SCFB_OK, host_i[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemsetAsync(idev, 0, 3 * sizeof(int), h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(A, a, batch * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && v0) {
    e = cudaMemcpyAsync(V, v0, batch * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const double one = 1.0, zero = 0.0;
    for (int b = 0; b < batch && e == cudaSuccess && rc == SCFB_OK; ++b)
      if (cublasDgemm(static_cast<cublasHandle_t>(h->blas), CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A + b * n2, n,
                      V + b * n2, n, &zero, U + b * n2, n) != CUBLAS_STATUS_SUCCESS)
        rc = scfb_fail(h, SCFB_ERR_CUSOLVER, "Dgemm failed");
  }
  if (e == cudaSuccess && rc == SCFB_OK)
    rc = scfb_launch_jacobi_eig(h, n, batch, v0 ? U : A, v0 ? V : nullptr, Z, W, idev, idev + 1);
  if (e == cudaSuccess && rc == SCFB_OK) {
    cudaMemcpyAsync(host_i, idev, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(w_out, W, (size_t)batch * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(z_out, Z, batch * n2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(idev);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "syev_jacobi: %s", cudaGetErrorString(e));
  if (host_i[0] != 0) return scfb_fail(h, SCFB_ERR_CUSOLVER, "Jacobi eigensolver did not converge (info=%d)", host_i[0]);
  if (sweeps_out)
    for (int b = 0; b < batch; ++b) sweeps_out[b] = host_i[1 + b];
  return SCFB_OK;
}

// compute_density(problem, orbcoeff) (compute_density.jl:4-35): D_s = C_occ,s C_occ,s' with the
// reference's block rule — alpha from spin block 0, beta from the LAST block of orbcoeff
// (n_spin_c blocks of n x n_orb), n_densities blocks written.
int scfb_density(scfb_handle_t h, int n_spin_c, int n_densities, int n_occ_a, int n_occ_b, int n_orb,
                 const double* orbcoeff, double* density_out) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  const int n = h->n;
  SCFB_REQUIRE(h, (n_spin_c == 1 || n_spin_c == 2) && (n_densities == 1 || n_densities == 2) &&
                      orbcoeff && density_out && n_orb >= 1 && n_orb <= n, "bad arguments");
  // compute_density.jl:8-10
  SCFB_REQUIRE(h, n_occ_a >= 0 && n_occ_b >= 0 && n_occ_a < n_orb && n_occ_b < n_orb,
               "n_elec of alpha and beta should be less than n_orb");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  if (int rc = ensure_dense_handles(h)) return rc;
  const size_t cn = (size_t)n * n_orb, n2 = (size_t)n * n;
  double* buf = nullptr;
  SCFB_CUDA(h, cudaMalloc(&buf, (n_spin_c * cn + n_densities * n2) * sizeof(double)));
  double* C = buf;
  double* D = buf + n_spin_c * cn;
  cudaMemcpyAsync(C, orbcoeff, n_spin_c * cn * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  int rc = SCFB_OK;
  const int nocc[2] = {n_occ_a, n_occ_b};
  for (int s = 0; s < n_densities && rc == SCFB_OK; ++s) {
    const double* Cs = C + (s == 0 ? 0 : (size_t)(n_spin_c - 1) * cn);
    if (nocc[s] == 0) {
      cudaMemsetAsync(D + s * n2, 0, n2 * sizeof(double), h->stream);
      continue;
    }
    const double one = 1.0, zero = 0.0;
    if (cublasDgemm(static_cast<cublasHandle_t>(h->blas), CUBLAS_OP_N, CUBLAS_OP_T, n, n, nocc[s], &one,
                    Cs, n, Cs, n, &zero, D + s * n2, n) != CUBLAS_STATUS_SUCCESS)
      rc = scfb_fail(h, SCFB_ERR_CUSOLVER, "cublasDgemm failed in scfb_density");
    h->launches += 1;
  }
  cudaMemcpyAsync(density_out, D, n_densities * n2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "density: %s", cudaGetErrorString(e));
  return SCFB_OK;
}

// The assembly half of compute_fock_matrix (compute_fock_matrix.jl:38-58) on host arrays:
// e1e = tr(h_core Dtot), e2e = tr(G_1 D_1) (one block) or (tr(G_1 D_1) + tr(G_2 D_2))/2,
// fock_out[:,:,s] = G[:,:,s] + h_core.
int scfb_fock_assemble(scfb_handle_t h, int n_spin, const double* g, const double* h_core,
                       const double* density, const double* total_density, double* fock_out,
                       double* e1e, double* e2e) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, (n_spin == 1 || n_spin == 2) && g && h_core && density && total_density && fock_out,
               "bad arguments");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  const int n = h->n;
  const size_t n2 = (size_t)n * n, ns = n_spin;
  double* buf = nullptr;
  SCFB_CUDA(h, cudaMalloc(&buf, (3 * ns * n2 + 2 * n2 + 8) * sizeof(double)));
  double *G = buf, *D = G + ns * n2, *F = D + ns * n2, *H = F + ns * n2, *T = H + n2, *scal = T + n2;
  cudaMemcpyAsync(G, g, ns * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(D, density, ns * n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(H, h_core, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(T, total_density, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  fock_add_kernel<<<blocks_for(n2 * ns), 256, 0, h->stream>>>(G, H, F, n2, n_spin);
  TraceBatch tb{};
  tb.p[0] = {H, T, 1};
  tb.p[1] = {G, D, 1};
  tb.p[2] = {G + (ns - 1) * n2, D + (ns - 1) * n2, 1};
  trace_kernel<<<3, 1024, 0, h->stream>>>(tb, n, scal);
  h->launches += 2;
  double v[3] = {0, 0, 0};
  cudaMemcpyAsync(fock_out, F, ns * n2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaMemcpyAsync(v, scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(buf);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "fock_assemble: %s", cudaGetErrorString(e));
  if (e1e) *e1e = v[0];
  if (e2e) *e2e = n_spin == 1 ? v[1] : 0.5 * (v[1] + v[2]);
  return SCFB_OK;
}

}  // extern "C"
