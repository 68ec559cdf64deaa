// Internal state of a libscf_b200 handle.  Not part of the ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/scf_b200.h"

struct scfb_scf_state;  // dense.cu
struct scfb_multi_s;     // multi.cu

constexpr int SCFB_EV_RING = 128;

// What the full-layout stream kernel's slab finisher does with a finished nu-slab (columns
// J[:,nu], K_s[:,nu]) beyond storing it in the handle's own jk buffer.  Passed by value to the kernel.
struct ScfbFuse {
  double* out = nullptr;  // single rank: out[:,nu,s] += J[:,nu] - K_s[:,nu]  (no separate accumulate launch)
  int n_peers = 0;        // > 0: store the columns into every rank's landing zone (peer memory over NVLink) ...
  double* peer_jk[8] = {};                // ... [J | K_a | K_b] of rank r for this build's parity
  unsigned long long* peer_flag[8] = {};  // this rank's flag word inside rank r's flag array
  unsigned long long epoch = 0;           // published there by the kernel's last CTA
};

// How the full-layout kernels cut a rank's nu-slabs into work items (jk_full.cu: scfb_full_plan).
struct ScfbFullPlan {
  bool tma = false;          // persistent bulk-copy kernel (else the register-prefetch LDG kernel)
  int b_splits = 1;          // LDG kernel: CTAs sharing one (nu, c-chunk) along b (gridDim.z)
  int whole_slabs = 0;       // TMA kernel: the first slabs are handed out as whole (nu, c-chunk) items,
  int tail_bs = 1;           //   the remaining ones cut tail_bs ways along b (finer items at the end)
  uint64_t kpart_rows = 0;   // partial K rows ([2][n] doubles each) / partial J rows (TC doubles each)
  uint64_t jpart_rows = 0;
  int n_groups = 0;          // split (nu, c-chunk) groups (completion counters)
};

struct scfb_handle_s {
  int n = 0;             // n_bas
  int layout = SCFB_LAYOUT_FULL;
  int device = 0;
  int rank = 0, n_ranks = 1;
  int nu_lo = 0, nu_hi = 0;  // owned nu-slabs (full layout) — [lo, hi)
  // packed layout: owned first-index blocks i in [i_lo, i_hi) (all rows pr(i,j), j <= i)
  int i_lo = 0, i_hi = 0;
  uint64_t pk_offset = 0;       // packed elements stored before this shard
  double* pk_acc = nullptr;     // [J~ (P) | K' (2 n^2)] fp64 RED accumulators
  double* pk_dq2 = nullptr;     // 2*Dtot in (padded) pair space
  double* pk_dpad = nullptr;    // padded copy of D (2 x (n^2 + 64))
  uint64_t* pk_btab = nullptr;  // B[i]: offset of row (i,0), n+1 entries

  double* eri = nullptr;  // device shard
  uint64_t eri_elems = 0;
  bool eri_ready = false;
  bool allow_partial = false;   // n_ranks > 1 without an exchange: return this rank's partial instead of failing
  uint64_t* slab_seen = nullptr;  // host bitmap of the owned nu-slabs (packed8: first indices i) uploaded so far
  double* pack_stage = nullptr;   // packed8: device staging for the nu-slabs being packed
  size_t pack_stage_slabs = 0;

  // device scratch
  double* d_density = nullptr;  // n*n*2   (host-pointer API staging)
  double* d_total = nullptr;    // n*n
  double* d_out = nullptr;      // n*n*2
  double* jk = nullptr;         // [J | K_0 | K_1], 3*n*n, this rank's partial: zero outside owned columns
  double* jk_sum = nullptr;     // allreduce result (n_ranks > 1); == jk when single rank
  double* kpart = nullptr;      // partial K rows, one [n_spin][n] row per work item
  double* jpart = nullptr;      // partial J values of the b-split work items
  int* sched = nullptr;         // work-item counter of the persistent kernels
  int* slab_done = nullptr;     // per owned nu-slab: finished work items of the running build (K1 fused tail)
  int* grp_done = nullptr;      // per split (nu, c-chunk): finished pieces of the running build
  unsigned int* grid_done = nullptr;  // CTAs of the running K1 launch that have finished
  int n_sm = 0;                 // SM count of this handle's device (lazily queried)
  bool tma_attr[3] = {false, false, false};  // per NSPIN: smem attributes of jk_full_tma_kernel set on this device
  uint2* pk_items[2] = {};   // live work items of jk_packed_kernel per spin count (built at first launch)
  int pk_n_items[2] = {-1, -1};
  ScfbFullPlan plan;            // how K1 cuts the owned nu-slabs into work items
  double* h_stage = nullptr;    // pinned host staging, 5*n*n

  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t side = nullptr;  // host-pointer API: `out` travels here, overlapped with the ERI stream
  cudaEvent_t ev_side[2] = {nullptr, nullptr};  // [d_out free again, d_out uploaded]
  // CUDA-event timing of the last SCFB_EV_RING builds: [build start, stream-kernel end, build end, stream-kernel start]
  cudaEvent_t ev_ring[SCFB_EV_RING][4] = {};
  unsigned char ev_map[SCFB_EV_RING][4] = {};  // logical -> recorded event: neighbours with nothing enqueued between them share one record
  uint64_t build_no = 0;  // builds started on this handle; build b uses ring slot b % SCFB_EV_RING
  bool timed = false;
  uint64_t launches = 0;

  scfb_scf_state* scf = nullptr;  // device-resident SCF step (scfb_scf_setup)
  scfb_multi_s* multi = nullptr;  // set on member 0 of a single-process multi-GPU group (scfb_create_multi)
  void* blas = nullptr;           // cublasHandle_t, created on first use
  void* solver = nullptr;         // cusolverDnHandle_t, created on first use

  void* peer = nullptr;       // PeerState (peer.cu) when the peer-memory exchange is set up
  void* nccl_comm = nullptr;  // ncclComm_t when n_ranks > 1 and initialised

  char err[512] = {0};
};

// event i of the build in progress (or, between builds, of the last one)
inline int scfb_ev_slot(const scfb_handle_s* h) { return (int)((h->build_no + SCFB_EV_RING - 1) % SCFB_EV_RING); }
inline cudaEvent_t scfb_ev(scfb_handle_s* h, int i) {
  const int slot = scfb_ev_slot(h);
  return h->ev_ring[slot][h->ev_map[slot][i]];
}
inline cudaError_t scfb_ev_record(scfb_handle_s* h, int i) {
  const int slot = scfb_ev_slot(h);
  h->ev_map[slot][i] = (unsigned char)i;
  return cudaEventRecord(h->ev_ring[slot][i], h->stream);
}
// event i is the same point of the stream as event j (nothing was enqueued in between): an event
// record costs the GPU front end a few microseconds, which a sub-millisecond build notices
inline void scfb_ev_alias(scfb_handle_s* h, int i, int j) {
  const int slot = scfb_ev_slot(h);
  h->ev_map[slot][i] = h->ev_map[slot][j];
}

// thread-local message for calls that fail before a handle exists
extern thread_local char scfb_tls_err[512];

inline int scfb_fail(scfb_handle_s* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  std::snprintf(scfb_tls_err, sizeof scfb_tls_err, "%s", buf);
  if (h) std::snprintf(h->err, sizeof h->err, "%s", buf);
  return code;
}

#define SCFB_CUDA(h, call)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return scfb_fail((h), e_ == cudaErrorMemoryAllocation ? SCFB_ERR_OOM : SCFB_ERR_CUDA, \
                       "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define SCFB_REQUIRE(h, cond, ...) \
  do {                             \
    if (!(cond)) return scfb_fail((h), SCFB_ERR_ARG, __VA_ARGS__); \
  } while (0)

// contiguous nu-slab ownership; ragged remainder goes to the first ranks
inline void scfb_split_range(int n, int rank, int n_ranks, int* lo, int* hi) {
  int base = n / n_ranks, rem = n % n_ranks;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

// ---- kernels / launchers implemented in the .cu files --------------------------------
// jk_full.cu
int scfb_launch_jk_full(scfb_handle_s* h, int n_spin, const double* d_density, const double* d_total,
                        const ScfbFuse* fuse, bool* fused);  // fills h->jk for owned columns
int scfb_launch_accumulate(scfb_handle_s* h, int n_spin, const double* d_jk, double* d_out);  // out += J - K_s
ScfbFullPlan scfb_full_plan(int n, int n_slabs, int n_sm);
// jk_packed.cu
void scfb_packed_split(int n, int rank, int n_ranks, int* i_lo, int* i_hi);
uint64_t scfb_packed_offset(int i);
uint64_t scfb_packed_dq_elems(int n);
int scfb_packed_upload_table(scfb_handle_s* h);
int scfb_launch_jk_packed(scfb_handle_s* h, int n_spin, const double* d_density, const double* d_total);
int scfb_launch_generate_packed(scfb_handle_s* h, int family, uint64_t seed, const double* d_aux);
int scfb_launch_pack_slabs(scfb_handle_s* h, const double* d_slabs, int i_first, int count);
// eri_gen.cu
int scfb_launch_generate_full(scfb_handle_s* h, int family, uint64_t seed, const double* d_aux);
// scfb_api.cu
bool scfb_is_pinned(const void* p);
int scfb_stage_inputs(scfb_handle_s* h, int n_spin, const double* density, const double* total, const double* out);
int scfb_check_build_args(scfb_handle_s* h, int n_spin, const void* a, const void* b, const void* c);
int scfb_build_jk(scfb_handle_s* h, int n_spin, const double* d_density, const double* d_total, double* d_out,
                  cudaEvent_t out_event, const double** jk_final);
// multi.cu
void scfb_multi_destroy(scfb_handle_s* h0);
int scfb_multi_n_dev(scfb_handle_s* h0);
scfb_handle_s* scfb_multi_member(scfb_handle_s* h0, int g);
int scfb_multi_sync(scfb_handle_s* h0);
int scfb_multi_build_dev(scfb_handle_s* h0, int n_spin, const double* d_density, const double* d_total, double* d_out,
                         cudaEvent_t out_event, const double** jk_final, bool fan_out);
int scfb_multi_stage_inputs(scfb_handle_s* h0, int n_spin, const double* density, const double* total,
                            const double* out);
int scfb_multi_inputs_consumed(scfb_handle_s* h0);
// comm.cu
int scfb_comm_allreduce_jk(scfb_handle_s* h, int n_spin);
int scfb_comm_destroy(scfb_handle_s* h);
// eig_jacobi.cu
int scfb_jacobi_eig_max_n();
int scfb_jacobi_eig_default_max_n();
int scfb_launch_jacobi_eig(scfb_handle_s* h, int n, int batch, const double* AV, const double* V0, double* Z, double* W,
                           int* info, int* sweeps_out);
// dense.cu
void scfb_scf_free(scfb_handle_s* h);
void scfb_dense_handles_destroy(scfb_handle_s* h);
// peer.cu
bool scfb_peer_ready(scfb_handle_s* h);
void scfb_peer_destroy(scfb_handle_s* h);
void scfb_peer_begin(scfb_handle_s* h, ScfbFuse* fuse);
int scfb_peer_exchange(scfb_handle_s* h, int n_spin, double* d_out, bool pushed);  // d_out == nullptr: fill jk_sum
int scfb_peer_check(scfb_handle_s* h);
int scfb_peer_alloc_local(scfb_handle_s* h);
void scfb_peer_local_ptrs(scfb_handle_s* h, double** area, unsigned long long** flags);
int scfb_peer_wire(scfb_handle_s* h, double* const* areas, unsigned long long* const* flags);
inline int scfb_stream_sync_checked(scfb_handle_s* h) {  // every sync that may follow a build reports a peer timeout
  cudaError_t e_ = cudaStreamSynchronize(h->stream);
  if (e_ != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "cudaStreamSynchronize -> %s", cudaGetErrorString(e_));
  return scfb_peer_check(h);
}
