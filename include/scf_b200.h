/* scf_b200.h — C ABI of libscf_b200.so: the B200-native Fock-build path that sits
 * behind SelfConsistentField.jl's `TwoElectronBuilder` / `add_2e_term!` boundary.
 *
 * Conventions
 *   - All matrices/tensors are dense, fp64, COLUMN-MAJOR (Julia layout):
 *       eri[a,b,c,d]   at  a + n*b + n^2*c + n^3*d          (n = n_bas)
 *       density[m,v,s] at  m + n*v + n^2*s                   (s = spin block)
 *   - Every function returns an int status: 0 = ok, negative = error class below.
 *     `scfb_last_error(h)` gives the message.  Nothing throws or aborts across the ABI.
 *   - Host entry points are synchronous: outputs are valid on return.
 *   - A handle is driven by ONE caller thread at a time (not re-entrant).
 *   - Multi-GPU, two ways.  (a) scfb_create_multi: ONE handle that owns the shards of G devices
 *     of this process — what a single-threaded caller such as the reference's run_scf needs; every
 *     entry point below works on it unchanged.  (b) One process (or thread) per GPU, each with its
 *     own handle created with (rank, n_ranks) plus an exchange (scfb_comm_init or scfb_peer_*).
 *     Either way rank / member g owns the contiguous nu-slabs (slowest ERI index; packed8: first
 *     indices i) `scfb_shard_range` reports.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   src/types/ScfProblem.jl:2-9                 abstract TwoElectronBuilder + add_2e_term!
 *   src/algorithms/JKBuilderFromTensor.jl:7-9   ERI storage           -> scfb_create / scfb_eri_*
 *   src/algorithms/JKBuilderFromTensor.jl:22-51 add_2e_term!          -> scfb_fock_jk
 *   src/parts/compute_fock_matrix.jl:8-71       Fock assembly/energies-> scfb_fock_assemble, scfb_scf_fock
 *   src/parts/compute_pulay_error.jl:1-25       S D F - F D S         -> scfb_pulay_error, scfb_scf_fock
 *   src/parts/compute_orbitals.jl:4-19          eigen(F, S)           -> scfb_sygvd, scfb_scf_roothaan
 *   src/parts/compute_density.jl:4-35           C_occ C_occ'          -> scfb_scf_roothaan
 *   src/parts/check_convergence.jl:18           ||e||_F               -> scfb_scf_fock
 *   src/algorithms/cDIIS.jl:123-127             B row tr(e1' ej)      -> scfb_scf_diis_push
 *   src/algorithms/EDIIS.jl:87                  tr((Fi-Fj)(Di-Dj))    -> scfb_scf_diis_push
 *   src/parts/diis.jl:42-55                     F <- sum c_i F_i      -> scfb_scf_diis_extrapolate
 *   src/types/DiisState.jl:40-50                purge                 -> scfb_scf_diis_purge
 */
#ifndef SCF_B200_H
#define SCF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCFB_OK 0
#define SCFB_ERR_ARG (-1)      /* bad argument / shape / precondition            */
#define SCFB_ERR_CUDA (-2)     /* CUDA runtime error (incl. no device)           */
#define SCFB_ERR_CUSOLVER (-3) /* cuSOLVER / cuBLAS failure                      */
#define SCFB_ERR_NCCL (-4)     /* NCCL failure or libnccl not loadable           */
#define SCFB_ERR_OOM (-5)      /* device allocation failed                       */
#define SCFB_ERR_STATE (-6)    /* call sequence error (e.g. ERI not loaded yet)  */

#define SCFB_LAYOUT_FULL 0    /* n^4 doubles, Julia order                                   */
#define SCFB_LAYOUT_PACKED8 1 /* ~P(P+1)/2 doubles, P=n(n+1)/2; ONLY for 8-fold-symmetric tensors AND
                                 symmetric density / total_density arguments (every SCF satisfies both) */

#define SCFB_ERI_HASH 0  /* dense splitmix64 family (DESIGN.md "synthetic inputs")  */
#define SCFB_ERI_CHAIN 1 /* soft-Coulomb dimerised chain (physical, SCF converges) */

typedef struct scfb_handle_s* scfb_handle_t;

/* ---- lifecycle --------------------------------------------------------------------- */
int scfb_version(void);
/* Creates a builder for an n_bas^4 ERI on CUDA device `device`.  (rank, n_ranks) select
 * the nu-slab shard this handle owns; use (0, 1) for a single GPU.  Allocates the shard. */
int scfb_create(scfb_handle_t* out, int n_bas, int layout, int device, int rank, int n_ranks);
/* One handle over n_dev devices of THIS process (device ordinals dev_ids[0..n_dev), <= 8, all
 * peer-accessible: one NVSwitch box).  Member g owns shard g of n_dev; a build fans the densities
 * out to every device, every device streams its shard, finished J/K columns travel over peer
 * memory (NVLink) and the result lands on dev_ids[0] — where the host-pointer calls read it back
 * and where the device-pointer / scfb_scf_* calls expect their device arrays.  The caller stays
 * single-threaded (src/run_scf.jl:4-65 unchanged).  n_dev == 1 is scfb_create(.., dev_ids[0], 0, 1). */
int scfb_create_multi(scfb_handle_t* out, int n_bas, int layout, int n_dev, const int* dev_ids);
int scfb_destroy(scfb_handle_t h); /* idempotent on NULL */
const char* scfb_last_error(scfb_handle_t h); /* h may be NULL: last error of this thread */
int scfb_sync(scfb_handle_t h);
/* Run the handle's work on the caller's CUDA stream (e.g. torch's current stream). */
int scfb_set_stream(scfb_handle_t h, void* cuda_stream);

/* ---- ERI storage (JKBuilderFromTensor.jl:7-9) -------------------------------------- */
int scfb_shard_range(scfb_handle_t h, int* nu_lo, int* nu_hi);
int scfb_eri_bytes(scfb_handle_t h, uint64_t* bytes);
/* `host_eri` is the FULL n^4 tensor; the handle copies (full layout) or packs on the device
 * (packed8: representatives eri[k,l,j,i], i >= j, k >= l, pr(i,j) >= pr(k,l)) what it owns. */
int scfb_eri_upload(scfb_handle_t h, const double* host_eri);
/* Chunked variant, both layouts: `host_slabs` holds nu_count consecutive nu-slabs eri[:,:,:,nu]
 * starting at global slab nu_first; slabs outside the shard are skipped.  packed8 packs them on the
 * device as they arrive (every packed row (i,j) needs nu-slab i only), so neither host nor device
 * ever holds n^4 doubles (load_integral_hdf5.jl:9-10). */
int scfb_eri_upload_slabs(scfb_handle_t h, const double* host_slabs, int nu_first, int nu_count);
/* Declare the shard complete after a sequence of partial scfb_eri_upload_slabs calls.  The handle
 * tracks which owned nu-slabs have arrived: SCFB_ERR_STATE if any is still missing. */
int scfb_eri_mark_ready(scfb_handle_t h);
/* Fill the shard on the device from a closed-form synthetic family ("synthetic
 * inputs").  `aux` is family specific: HASH ignores it (may be NULL); CHAIN expects
 * 1 + n + n*n doubles: [soft, x[0..n), S[0..n*n) column-major] (site positions and overlap). */
int scfb_eri_generate(scfb_handle_t h, int family, uint64_t seed, const double* aux);
/* Copy the shard back exactly as stored (full: owned nu-slabs; packed8: the pre-scaled, padded
 * packed rows — scfb_eri_bytes gives the size).  Debugging / parity at small n. */
int scfb_eri_download(scfb_handle_t h, double* host_shard);

/* ---- the hot path (JKBuilderFromTensor.jl:22-51) ----------------------------------- */
/* out[:,:,s] += J[total_density] - K[density[:,:,s]],  s < n_spin in {1,2}.
 *   J[m,v] = sum_ab eri[a,b,m,v] * total_density[a,b]
 *   K[m,v] = sum_ab eri[m,b,a,v] * density[a,b,s]
 * Host pointers; `out` is read, accumulated into and written back.  With n_ranks > 1 and
 * an initialised communicator the partial J/K are allreduced first, so every rank
 * returns the full result. */
int scfb_fock_jk(scfb_handle_t h, int n_spin, const double* density,
                 const double* total_density, double* out);
/* A handle created with n_ranks > 1 refuses to build (SCFB_ERR_STATE) until a cross-GPU exchange
 * is set up (scfb_comm_init or scfb_peer_import): without one the result would silently be this
 * rank's partial only.  scfb_allow_partial(h, 1) opts in to exactly that — the owned nu-columns
 * (full layout) or this shard's dense partial sums (packed8) — for shard tests and timing. */
int scfb_allow_partial(scfb_handle_t h, int allow);
/* Same with DEVICE pointers, asynchronous on the handle's stream.  d_density and
 * d_total_density must be 16-byte aligned (SCFB_ERR_ARG otherwise). */
int scfb_fock_jk_dev(scfb_handle_t h, int n_spin, const double* d_density,
                     const double* d_total_density, double* d_out);
/* J (n*n) and K (n*n*n_spin) separately, overwritten, host pointers. */
int scfb_jk_split(scfb_handle_t h, int n_spin, const double* density,
                  const double* total_density, double* J, double* K);
/* Device time of the last build on this handle, measured with CUDA events on the
 * handle's stream: `stream_ms` = the ERI-streaming kernel alone, `total_ms` = all
 * launches of the build (stream + reduce + collective + accumulate). */
int scfb_last_build_ms(scfb_handle_t h, double* stream_ms, double* total_ms);
/* The same two durations for the last `max_builds` builds (at most 128 are kept), oldest first,
 * WITHOUT having synchronised between them: what a back-to-back timing loop reads afterwards.
 * *n_builds = entries written; either array may be NULL. */
int scfb_build_ms_history(scfb_handle_t h, int max_builds, double* stream_ms, double* total_ms, int* n_builds);
/* Number of kernels this library launched on this handle since creation. */
int scfb_launch_count(scfb_handle_t h, uint64_t* launches);
/* Introspection (no device needed): how the full-layout kernel cuts a shard of n_slabs nu-slabs into work
 * items on a device with n_sm SMs.  plan6 = { persistent TMA kernel? , slabs handed out as whole (nu, c-chunk)
 * items, pieces the remaining slabs' items are cut into along b (LDG kernel: every item), partial K rows,
 * partial J rows, split (nu, c-chunk) groups }. */
int scfb_full_plan_query(int n_bas, int n_slabs, int n_sm, int* plan6);

/* ---- device-resident SCF step (rows a4-a9 of SURVEY.md 8) ---------------------------
 * The handle keeps S, h_core, the density, the Fock iterate, the Pulay error, the orbitals
 * and the DIIS history in HBM; per iteration only scalars and DIIS rows cross PCIe.  The
 * caller keeps the control flow of run_scf.jl:21-49 and the <= 6x6 coefficient solves
 * (cDIIS.jl:33-63, EDIIS.jl:20-44).  RHF (restricted, closed shell: 1 spin block), UHF (2 spin
 * blocks) and ROHF (restricted, n_elec_a != n_elec_b: 2 density blocks, ONE Fock / orbital block
 * — the Roothaan effective Fock of compute_fock_matrix.jl:82-123, built on the device). */
#define SCFB_GET_FOCK 0      /* the iterate F (after extrapolation)   (n,n,n_spin) */
#define SCFB_GET_FOCK_NEW 1  /* Fock matrix of the newest build       (n,n,n_spin) */
#define SCFB_GET_DENSITY 2   /*                                       (n,n,n_densities) */
#define SCFB_GET_ERROR 3     /* S D F - F D S of the newest build     (n,n,n_spin) */
#define SCFB_GET_ORBCOEFF 4  /* all n eigenvectors per spin           (n,n,n_spin) */
#define SCFB_GET_ORBEN 5     /* all n eigenvalues per spin            (n,n_spin)   */
#define SCFB_GET_ORBCOEFF_PREV 6 /* same three of the previous Roothaan step: what     */
#define SCFB_GET_ORBEN_PREV 7    /* run_scf.jl:52-58 returns when the loop breaks on   */
#define SCFB_GET_DENSITY_PREV 8  /* convergence (:40 precedes :48)                     */
int scfb_scf_setup(scfb_handle_t h, const double* overlap, const double* h_core, double e_0e,
                   int n_elec_a, int n_elec_b, int n_orb, int restricted, int n_diis_size);
int scfb_scf_n_spin(scfb_handle_t h, int* n_spin);           /* Fock / orbital / error blocks        */
int scfb_scf_n_densities(scfb_handle_t h, int* n_densities); /* density blocks (compute_density.jl:24-27) */
int scfb_scf_set_density(scfb_handle_t h, int n_spin, const double* density);
int scfb_scf_set_fock(scfb_handle_t h, int n_spin, const double* fock);
/* compute_fock_matrix.jl:8-71 on the resident density: energies4 = {0e, 1e, 2e, total},
 * error_norm = Frobenius norm of the Pulay error over all spins (check_convergence.jl:18). */
int scfb_scf_fock(scfb_handle_t h, double* energies4, double* error_norm);
/* Roothaan.jl:5-6: orbitals of the iterate F (cuSOLVER Dsygvd), then the density. */
int scfb_scf_roothaan(scfb_handle_t h);
/* Push the newest (F, e, D) of `spin` into the history (newest first, capacity
 * n_diis_size) and return the new first rows, length *m: tr(e1' ej) and
 * tr((F1-Fj)(D1-Dj)).  Either row pointer may be NULL. */
int scfb_scf_diis_push(scfb_handle_t h, int spin, int* m, double* row_cdiis, double* row_ediis);
/* F[:,:,spin] <- sum_{i<m} c[i] * F_i over the history, newest first (diis.jl:52-54). */
int scfb_scf_diis_extrapolate(scfb_handle_t h, int spin, int m, const double* c);
int scfb_scf_diis_purge(scfb_handle_t h, int spin, int count); /* DiisState.jl:40-50 */
int scfb_scf_diis_reset(scfb_handle_t h);
int scfb_scf_accept_fock(scfb_handle_t h); /* iterate <- newest Fock (no accelerator) */
int scfb_scf_get(scfb_handle_t h, int what, double* host_out);
/* Stateless host-pointer twins of the reference's plain functions. */
int scfb_pulay_error(scfb_handle_t h, int n_spin, const double* fock, const double* density,
                     const double* overlap, double* error_out);
int scfb_sygvd(scfb_handle_t h, int n_spin, int n_orb, const double* fock, const double* overlap,
               double* orben_out, double* orbcoeff_out);
/* The eigensolver scfb_scf_roothaan uses for n_bas <= 512 (one-sided Jacobi, the matrix resident in the shared
 * memory of one CTA or of a cluster of 8 / 16 CTAs; compute_orbitals.jl:14 after the reduction with the cached Cholesky factor of S):
 * `batch` (1 or 2) symmetric n x n matrices a (n <= 512), column-major and consecutive; v0 = NULL or orthonormal
 * warm-start bases (eigenvectors of a nearby matrix).  w_out (batch*n) ascending, z_out columns = eigenvectors,
 * sweeps_out[batch] = Jacobi sweeps taken.  n is independent of the handle's n_bas. */
int scfb_syev_jacobi(scfb_handle_t h, int n, int batch, const double* a, const double* v0, double* w_out,
                     double* z_out, int* sweeps_out);
/* compute_density.jl:4-35: density_out[:,:,s] = C_occ,s C_occ,s' (alpha from block 0, beta from the
 * last of the n_spin_c blocks of orbcoeff (n, n_orb, n_spin_c)); n_densities blocks are written. */
int scfb_density(scfb_handle_t h, int n_spin_c, int n_densities, int n_occ_a, int n_occ_b, int n_orb,
                 const double* orbcoeff, double* density_out);
/* compute_fock_matrix.jl:38-58: fock_out = G + h_core, *e1e = tr(h_core Dtot),
 * *e2e = tr(G_1 D_1) or (tr(G_1 D_1) + tr(G_2 D_2)) / 2. */
int scfb_fock_assemble(scfb_handle_t h, int n_spin, const double* g, const double* h_core,
                       const double* density, const double* total_density, double* fock_out,
                       double* e1e, double* e2e);

/* ---- multi-GPU exchange (one process per GPU) ------------------------- */
/* Rank 0 calls scfb_comm_unique_id and ships the 128 bytes to the other ranks by any
 * means (torch.distributed / MPI / a file); every rank then calls scfb_comm_init. */
int scfb_comm_unique_id(char* id128);
int scfb_comm_init(scfb_handle_t h, const char* id128);
/* Opt-in alternative exchange over peer memory (all ranks on one NVSwitch box, <= 8): every rank
 * calls scfb_peer_export (128 bytes of CUDA IPC handles), the host gathers the n_ranks x 128
 * bytes in rank order, every rank calls scfb_peer_import.  Builds then store their partial J/K
 * straight into every peer's receive area and sum the slots locally (deterministic order)
 * instead of calling ncclAllReduce.  A rank that never arrives makes the wait time out
 * (SCFB_ERR_NCCL at the next synchronising call) instead of hanging. */
int scfb_peer_export(scfb_handle_t h, char* handle128);
int scfb_peer_import(scfb_handle_t h, const char* all_handles);

#ifdef __cplusplus
}
#endif
#endif /* SCF_B200_H */
