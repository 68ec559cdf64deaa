// C-ABI entry points of libscf_b200.so (declared in include/scf_b200.h): lifecycle, ERI
// storage and the host/device-pointer Fock-build calls.  No CPU fallback anywhere: every
// compute entry point needs a CUDA device and fails with SCFB_ERR_CUDA otherwise.
#include <cstdlib>
#include <new>

#include "scfb_internal.h"

thread_local char scfb_tls_err[512] = {0};

namespace {

constexpr int kVersion = 100;  // 0.1.0

int check_handle(scfb_handle_s* h) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
  return SCFB_OK;
}

// Members of the handle's group (a plain handle is its own only member); other members' errors
// are copied into the caller-visible handle.
int n_members(scfb_handle_s* h) { return h->multi ? scfb_multi_n_dev(h) : 1; }
scfb_handle_s* member(scfb_handle_s* h, int g) { return h->multi ? scfb_multi_member(h, g) : h; }
int lift(scfb_handle_s* h, scfb_handle_s* m, int rc) {
  if (rc != SCFB_OK && m != h) std::snprintf(h->err, sizeof h->err, "device %d: %.480s", m->device, m->err);
  return rc;
}

}  // namespace

bool scfb_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory on old drivers: clear the sticky error
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Push the caller's host arrays to the device staging block [density 2n^2 | total n^2 | out 2n^2].
// Page-locked caller buffers are copied straight from where they are; pageable ones go through
// the handle's pinned staging buffer (one memcpy + one H2D copy).  density / total_density travel on
// the build stream (the stream kernel needs them); `out` — only needed by the final accumulate —
// travels on the handle's side stream so that its PCIe time hides behind the ERI stream.
int scfb_stage_inputs(scfb_handle_s* h, int n_spin, const double* density, const double* total,
                      const double* out) {
  const size_t n2 = (size_t)h->n * h->n;
  const bool pinned = scfb_is_pinned(density) && scfb_is_pinned(total) && (!out || scfb_is_pinned(out));
  const double* src_d = density;
  const double* src_t = total;
  const double* src_o = out;
  if (!pinned) {
    std::memcpy(h->h_stage, density, n2 * n_spin * sizeof(double));
    std::memcpy(h->h_stage + 2 * n2, total, n2 * sizeof(double));
    if (out) std::memcpy(h->h_stage + 3 * n2, out, n2 * n_spin * sizeof(double));
    src_d = h->h_stage;
    src_t = h->h_stage + 2 * n2;
    src_o = h->h_stage + 3 * n2;
  }
  if (out) {
    // the side stream may only overwrite d_out once the previous build's reads of it are done
    SCFB_CUDA(h, cudaEventRecord(h->ev_side[0], h->stream));
    SCFB_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_side[0], 0));
    SCFB_CUDA(h, cudaMemcpyAsync(h->d_out, src_o, n2 * n_spin * sizeof(double), cudaMemcpyHostToDevice, h->side));
    SCFB_CUDA(h, cudaEventRecord(h->ev_side[1], h->side));
  }
  SCFB_CUDA(h, cudaMemcpyAsync(h->d_density, src_d, n2 * n_spin * sizeof(double), cudaMemcpyHostToDevice,
                               h->stream));
  SCFB_CUDA(h, cudaMemcpyAsync(h->d_total, src_t, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return SCFB_OK;
}

int scfb_check_build_args(scfb_handle_s* h, int n_spin, const void* a, const void* b, const void* c) {
  SCFB_REQUIRE(h, n_spin == 1 || n_spin == 2, "n_spin must be 1 or 2 (got %d)", n_spin);
  SCFB_REQUIRE(h, a && b && c, "null array pointer");
  for (int g = 0; g < n_members(h); ++g)
    if (!member(h, g)->eri_ready)
      return scfb_fail(h, SCFB_ERR_STATE, "ERI not loaded: call scfb_eri_upload/_generate first");
  if (h->n_ranks > 1 && !h->nccl_comm && !scfb_peer_ready(h) && !h->allow_partial)
    return scfb_fail(h, SCFB_ERR_STATE,
                     "rank %d of %d has no cross-GPU exchange (scfb_comm_init / scfb_peer_import): the result "
                     "would be this rank's partial only; scfb_allow_partial(h, 1) opts in to that",
                     h->rank, h->n_ranks);
  return SCFB_OK;
}

// One Fock build on the handle's stream: the J/K kernels of this rank's shard, the cross-GPU
// exchange when there is one, and (d_out != nullptr) out += J - K_s.  *jk_final = where the complete
// [J | K_a | K_b] lives afterwards.  `out_event`: d_out becomes valid only once this event (recorded on
// another stream) has fired — waited for right before the first kernel that touches d_out, and it
// rules out accumulating inside the stream kernel.
int scfb_build_jk(scfb_handle_s* h, int n_spin, const double* d_density, const double* d_total, double* d_out,
                  cudaEvent_t out_event, const double** jk_final) {
  const bool use_peer = h->n_ranks > 1 && scfb_peer_ready(h);
  const bool use_nccl = h->n_ranks > 1 && !use_peer && h->nccl_comm;
  ScfbFuse fuse{};
  if (use_peer)
    scfb_peer_begin(h, &fuse);
  else if (!use_nccl && !out_event)
    fuse.out = d_out;  // single rank: the slab finishers accumulate, no further launch
  bool fused = false;
  ++h->build_no;
  if (h->timed) SCFB_CUDA(h, scfb_ev_record(h, 0));
  if (int rc = h->layout == SCFB_LAYOUT_FULL ? scfb_launch_jk_full(h, n_spin, d_density, d_total, &fuse, &fused)
                                             : scfb_launch_jk_packed(h, n_spin, d_density, d_total))
    return rc;
  if (out_event) SCFB_CUDA(h, cudaStreamWaitEvent(h->stream, out_event, 0));
  const double* jk = h->jk;
  bool tail_work = out_event != nullptr;  // anything enqueued after the layout's own launches?
  if (use_peer) {
    if (int rc = scfb_peer_exchange(h, n_spin, d_out, fused)) return rc;
    jk = h->jk_sum;  // (valid when d_out == nullptr)
    tail_work = true;
  } else {
    if (use_nccl) {
      if (int rc = scfb_comm_allreduce_jk(h, n_spin)) return rc;
      jk = h->jk_sum;
      tail_work = true;
    }
    if (d_out && !(fused && fuse.out)) {
      if (int rc = scfb_launch_accumulate(h, n_spin, jk, d_out)) return rc;
      tail_work = true;
    }
  }
  if (h->timed) {
    // single launch (the fused K1 tail did everything): the build ends where the stream kernel does
    if (h->layout == SCFB_LAYOUT_FULL && h->plan.tma && !tail_work)
      scfb_ev_alias(h, 2, 1);
    else
      SCFB_CUDA(h, scfb_ev_record(h, 2));
  }
  if (jk_final) *jk_final = jk;
  return SCFB_OK;
}

extern "C" {

int scfb_version(void) { return kVersion; }

const char* scfb_last_error(scfb_handle_t h) { return h ? h->err : scfb_tls_err; }

int scfb_create(scfb_handle_t* out, int n_bas, int layout, int device, int rank, int n_ranks) {
  if (!out) return scfb_fail(nullptr, SCFB_ERR_ARG, "null output pointer");
  *out = nullptr;
  SCFB_REQUIRE(nullptr, n_bas >= 1, "n_bas must be >= 1 (got %d)", n_bas);
  SCFB_REQUIRE(nullptr, layout == SCFB_LAYOUT_FULL || layout == SCFB_LAYOUT_PACKED8,
               "unknown layout %d", layout);
  SCFB_REQUIRE(nullptr, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank %d of %d", rank,
               n_ranks);
  if (layout == SCFB_LAYOUT_FULL)
    SCFB_REQUIRE(nullptr, (n_bas % 2 == 0 && n_bas <= 1024) || (n_bas % 2 == 1 && n_bas <= 511),
                 "n_bas=%d unsupported by the full layout: even n <= 1024 or odd n <= 511", n_bas);
  else
    SCFB_REQUIRE(nullptr, n_bas <= 2048, "n_bas=%d unsupported by the packed8 layout: n <= 2048", n_bas);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return scfb_fail(nullptr, SCFB_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)",
                     cudaGetErrorString(e));
  SCFB_REQUIRE(nullptr, device >= 0 && device < n_dev, "device %d out of range (%d visible)",
               device, n_dev);

  scfb_handle_s* h = new (std::nothrow) scfb_handle_s();
  if (!h) return scfb_fail(nullptr, SCFB_ERR_OOM, "host allocation failed");
  h->n = n_bas;
  h->layout = layout;
  h->device = device;
  h->rank = rank;
  h->n_ranks = n_ranks;
  const uint64_t n = n_bas, n2 = n * n;
  if (layout == SCFB_LAYOUT_FULL) {
    scfb_split_range(n_bas, rank, n_ranks, &h->nu_lo, &h->nu_hi);
    h->eri_elems = n2 * n * (uint64_t)(h->nu_hi - h->nu_lo);
  } else {
    scfb_packed_split(n_bas, rank, n_ranks, &h->i_lo, &h->i_hi);
    h->pk_offset = scfb_packed_offset(h->i_lo);
    h->eri_elems = scfb_packed_offset(h->i_hi) - h->pk_offset;
  }

#define CREATE_CUDA(call)                                                                   \
  do {                                                                                      \
    cudaError_t e2_ = (call);                                                               \
    if (e2_ != cudaSuccess) {                                                               \
      int rc_ = scfb_fail(nullptr, e2_ == cudaErrorMemoryAllocation ? SCFB_ERR_OOM : SCFB_ERR_CUDA, \
                          "scfb_create: %s -> %s", #call, cudaGetErrorString(e2_));         \
      scfb_destroy(h);                                                                      \
      return rc_;                                                                           \
    }                                                                                       \
  } while (0)

  CREATE_CUDA(cudaSetDevice(device));
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& slot : h->ev_ring)
    for (auto& ev : slot) CREATE_CUDA(cudaEventCreate(&ev));
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  for (auto& ev : h->ev_side) CREATE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  h->timed = true;
  if (h->eri_elems) CREATE_CUDA(cudaMalloc(&h->eri, h->eri_elems * sizeof(double)));
  CREATE_CUDA(cudaMalloc(&h->d_density, 5 * n2 * sizeof(double)));
  h->d_total = h->d_density + 2 * n2;
  h->d_out = h->d_density + 3 * n2;
  CREATE_CUDA(cudaMalloc(&h->jk, 3 * n2 * sizeof(double)));
  CREATE_CUDA(cudaMemsetAsync(h->jk, 0, 3 * n2 * sizeof(double), h->stream));
  h->jk_sum = h->jk;
  if (n_ranks > 1) CREATE_CUDA(cudaMalloc(&h->jk_sum, 3 * n2 * sizeof(double)));
  CREATE_CUDA(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
  if (layout == SCFB_LAYOUT_FULL) {
    h->plan = scfb_full_plan(n_bas, h->nu_hi - h->nu_lo, h->n_sm);
    if (h->plan.kpart_rows) CREATE_CUDA(cudaMalloc(&h->kpart, h->plan.kpart_rows * 2 * n * sizeof(double)));
    if (h->plan.jpart_rows) CREATE_CUDA(cudaMalloc(&h->jpart, h->plan.jpart_rows * 8 * sizeof(double)));
    if (h->plan.n_groups) {
      CREATE_CUDA(cudaMalloc(&h->grp_done, h->plan.n_groups * sizeof(int)));
      CREATE_CUDA(cudaMemsetAsync(h->grp_done, 0, h->plan.n_groups * sizeof(int), h->stream));
    }
  }
  CREATE_CUDA(cudaMalloc(&h->sched, 2 * sizeof(int)));  // [work-item counter | finished-CTA counter]
  CREATE_CUDA(cudaMemsetAsync(h->sched, 0, 2 * sizeof(int), h->stream));
  h->grid_done = reinterpret_cast<unsigned int*>(h->sched + 1);
  if (layout == SCFB_LAYOUT_FULL && h->nu_hi > h->nu_lo) {
    CREATE_CUDA(cudaMalloc(&h->slab_done, (h->nu_hi - h->nu_lo) * sizeof(int)));
    CREATE_CUDA(cudaMemsetAsync(h->slab_done, 0, (h->nu_hi - h->nu_lo) * sizeof(int), h->stream));
  }
  if (layout == SCFB_LAYOUT_PACKED8) {
    const uint64_t PQ = scfb_packed_dq_elems(n_bas);
    CREATE_CUDA(cudaMalloc(&h->pk_acc, (PQ + 2 * (n2 + 64)) * sizeof(double)));
    CREATE_CUDA(cudaMalloc(&h->pk_dq2, PQ * sizeof(double)));
    CREATE_CUDA(cudaMalloc(&h->pk_dpad, 2 * (n2 + 64) * sizeof(double)));
    if (int rc_ = scfb_packed_upload_table(h)) {
      scfb_destroy(h);
      return rc_;
    }
  }
  CREATE_CUDA(cudaMallocHost(&h->h_stage, 5 * n2 * sizeof(double)));
  CREATE_CUDA(cudaStreamSynchronize(h->stream));
#undef CREATE_CUDA
  *out = h;
  return SCFB_OK;
}

int scfb_destroy(scfb_handle_t h) {
  if (!h) return SCFB_OK;
  if (h->multi) scfb_multi_destroy(h);  // the other members of a multi-GPU group go first
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  scfb_comm_destroy(h);
  scfb_peer_destroy(h);
  scfb_scf_free(h);
  scfb_dense_handles_destroy(h);
  cudaFree(h->eri);
  cudaFree(h->d_density);
  if (h->jk_sum != h->jk) cudaFree(h->jk_sum);
  cudaFree(h->jk);
  cudaFree(h->kpart);
  cudaFree(h->jpart);
  cudaFree(h->sched);
  cudaFree(h->slab_done);
  cudaFree(h->grp_done);
  cudaFree(h->pk_acc);
  cudaFree(h->pk_dq2);
  cudaFree(h->pk_dpad);
  cudaFree(h->pk_btab);
  cudaFree(h->pack_stage);
  for (auto* p : h->pk_items) cudaFree(p);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  std::free(h->slab_seen);
  for (auto& slot : h->ev_ring)
    for (auto& ev : slot)
      if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->ev_side)
    if (ev) cudaEventDestroy(ev);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return SCFB_OK;
}

int scfb_sync(scfb_handle_t h) {
  if (int rc = check_handle(h)) return rc;
  if (h->multi) {
    const int rc = scfb_multi_sync(h);
    cudaSetDevice(h->device);
    return rc;
  }
  SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return scfb_peer_check(h);
}

int scfb_set_stream(scfb_handle_t h, void* cuda_stream) {
  if (int rc = check_handle(h)) return rc;
  SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return SCFB_OK;
}

int scfb_shard_range(scfb_handle_t h, int* nu_lo, int* nu_hi) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  const bool full = h->layout == SCFB_LAYOUT_FULL;
  if (nu_lo) *nu_lo = h->multi ? 0 : (full ? h->nu_lo : h->i_lo);  // a multi-GPU group owns everything
  if (nu_hi) *nu_hi = h->multi ? h->n : (full ? h->nu_hi : h->i_hi);
  return SCFB_OK;
}

int scfb_eri_bytes(scfb_handle_t h, uint64_t* bytes) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  if (bytes) {
    *bytes = 0;
    for (int g = 0; g < n_members(h); ++g) *bytes += member(h, g)->eri_elems * sizeof(double);
  }
  return SCFB_OK;
}

static int one_eri_upload_slabs(scfb_handle_t h, const double* host_slabs, int nu_first, int nu_count) {
  if (int rc = check_handle(h)) return rc;
  SCFB_REQUIRE(h, host_slabs, "null host pointer");
  SCFB_REQUIRE(h, nu_first >= 0 && nu_count >= 0 && nu_first + nu_count <= h->n,
               "slab range [%d, %d) outside [0, %d)", nu_first, nu_first + nu_count, h->n);
  const bool full = h->layout == SCFB_LAYOUT_FULL;
  const int own_lo = full ? h->nu_lo : h->i_lo, own_hi = full ? h->nu_hi : h->i_hi;
  const int lo = nu_first > own_lo ? nu_first : own_lo;
  const int hi = nu_first + nu_count < own_hi ? nu_first + nu_count : own_hi;
  if (hi > lo) {
    if (!h->slab_seen) {
      h->slab_seen = static_cast<uint64_t*>(std::calloc((h->n + 63) / 64, sizeof(uint64_t)));
      if (!h->slab_seen) return scfb_fail(h, SCFB_ERR_OOM, "host allocation failed");
    }
    const uint64_t n3 = (uint64_t)h->n * h->n * h->n;
    if (full) {
      SCFB_CUDA(h, cudaMemcpyAsync(h->eri + (uint64_t)(lo - h->nu_lo) * n3,
                                   host_slabs + (uint64_t)(lo - nu_first) * n3,
                                   (uint64_t)(hi - lo) * n3 * sizeof(double), cudaMemcpyHostToDevice,
                                   h->stream));
    } else {
      // packed8: every row (i, j <= i) of the shard comes from nu-slab i alone (8-fold symmetry), so
      // the owned slabs pass through a bounded device staging buffer and are packed there
      size_t batch = (size_t)(256u << 20) / (n3 * sizeof(double));
      if (batch < 1) batch = 1;
      if (batch > (size_t)(hi - lo)) batch = hi - lo;
      if (h->pack_stage_slabs < batch) {
        cudaFree(h->pack_stage);
        h->pack_stage = nullptr;
        h->pack_stage_slabs = 0;
        SCFB_CUDA(h, cudaMalloc(&h->pack_stage, batch * n3 * sizeof(double)));
        h->pack_stage_slabs = batch;
      }
      for (int v = lo; v < hi; v += (int)batch) {
        const int cnt = v + (int)batch < hi ? (int)batch : hi - v;
        SCFB_CUDA(h, cudaMemcpyAsync(h->pack_stage, host_slabs + (uint64_t)(v - nu_first) * n3,
                                     (uint64_t)cnt * n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (int rc = scfb_launch_pack_slabs(h, h->pack_stage, v, cnt)) return rc;
      }
    }
    SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int v = lo; v < hi; ++v) h->slab_seen[v >> 6] |= 1ull << (v & 63);
  }
  if (lo <= own_lo && hi >= own_hi) h->eri_ready = true;  // whole shard arrived in one call
  return SCFB_OK;
}

static int one_eri_upload(scfb_handle_t h, const double* host_eri) {
  int rc = one_eri_upload_slabs(h, host_eri, 0, h ? h->n : 0);
  if (rc == SCFB_OK) h->eri_ready = true;
  return rc;
}

static int one_eri_mark_ready(scfb_handle_t h) {  // after a sequence of partial slab uploads
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  if (h->eri_ready) return SCFB_OK;
  const bool full = h->layout == SCFB_LAYOUT_FULL;
  const int own_lo = full ? h->nu_lo : h->i_lo, own_hi = full ? h->nu_hi : h->i_hi;
  for (int v = own_lo; v < own_hi; ++v)
    if (!h->slab_seen || !(h->slab_seen[v >> 6] >> (v & 63) & 1))
      return scfb_fail(h, SCFB_ERR_STATE, "nu-slab %d of the shard [%d, %d) was never uploaded", v, own_lo, own_hi);
  cudaFree(h->pack_stage);  // the staging buffer is only needed while loading
  h->pack_stage = nullptr;
  h->pack_stage_slabs = 0;
  h->eri_ready = true;
  return SCFB_OK;
}

int scfb_allow_partial(scfb_handle_t h, int allow) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  h->allow_partial = allow != 0;
  return SCFB_OK;
}

static int one_eri_generate(scfb_handle_t h, int family, uint64_t seed, const double* aux) {
  if (int rc = check_handle(h)) return rc;
  SCFB_REQUIRE(h, family == SCFB_ERI_HASH || family == SCFB_ERI_CHAIN, "unknown family %d", family);
  double* d_aux = nullptr;
  if (family == SCFB_ERI_CHAIN) {
    SCFB_REQUIRE(h, aux, "CHAIN family needs aux = [soft, x[n], S[n*n]]");
    const size_t cnt = 1 + (size_t)h->n + (size_t)h->n * h->n;
    SCFB_CUDA(h, cudaMalloc(&d_aux, cnt * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_aux, aux, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) {
      cudaFree(d_aux);
      return scfb_fail(h, SCFB_ERR_CUDA, "aux upload: %s", cudaGetErrorString(e));
    }
  }
  int rc = h->layout == SCFB_LAYOUT_FULL ? scfb_launch_generate_full(h, family, seed, d_aux)
                                         : scfb_launch_generate_packed(h, family, seed, d_aux);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (d_aux) cudaFree(d_aux);
  if (rc) return rc;
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "generate: %s", cudaGetErrorString(e));
  h->eri_ready = true;
  return SCFB_OK;
}

static int one_eri_download(scfb_handle_t h, double* host_shard) {
  if (int rc = check_handle(h)) return rc;
  SCFB_REQUIRE(h, host_shard, "null host pointer");  // packed: raw pre-scaled packed rows
  SCFB_CUDA(h, cudaMemcpyAsync(host_shard, h->eri, h->eri_elems * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
  SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SCFB_OK;
}

// ---- the same entry points over every member of a multi-GPU group ----
#define SCFB_EACH_MEMBER(h, call)                              \
  do {                                                         \
    if (!(h)) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle"); \
    for (int g_ = 0; g_ < n_members(h); ++g_) {                \
      scfb_handle_s* m = member(h, g_);                        \
      if (int rc_ = lift(h, m, call)) {                        \
        cudaSetDevice((h)->device);                            \
        return rc_;                                            \
      }                                                        \
    }                                                          \
    cudaSetDevice((h)->device);                                \
    return SCFB_OK;                                            \
  } while (0)

int scfb_eri_upload_slabs(scfb_handle_t h, const double* host_slabs, int nu_first, int nu_count) {
  SCFB_EACH_MEMBER(h, one_eri_upload_slabs(m, host_slabs, nu_first, nu_count));
}
int scfb_eri_upload(scfb_handle_t h, const double* host_eri) { SCFB_EACH_MEMBER(h, one_eri_upload(m, host_eri)); }
int scfb_eri_mark_ready(scfb_handle_t h) { SCFB_EACH_MEMBER(h, one_eri_mark_ready(m)); }
int scfb_eri_generate(scfb_handle_t h, int family, uint64_t seed, const double* aux) {
  SCFB_EACH_MEMBER(h, one_eri_generate(m, family, seed, aux));
}
int scfb_eri_download(scfb_handle_t h, double* host_shard) {  // members' shards back to back (rank order)
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, host_shard, "null host pointer");
  for (int g = 0; g < n_members(h); ++g) {
    scfb_handle_s* m = member(h, g);
    if (int rc = lift(h, m, one_eri_download(m, host_shard))) return rc;
    host_shard += m->eri_elems;
  }
  cudaSetDevice(h->device);
  return SCFB_OK;
}

int scfb_fock_jk_dev(scfb_handle_t h, int n_spin, const double* d_density,
                     const double* d_total_density, double* d_out) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = scfb_check_build_args(h, n_spin, d_density, d_total_density, d_out)) return rc;
  // the kernels read the densities with 128-bit loads / 16-byte-granular bulk copies: a misaligned
  // pointer would fault on the device (a sticky error), so it is refused here
  SCFB_REQUIRE(h, ((uintptr_t)d_density | (uintptr_t)d_total_density) % 16 == 0,
               "device density pointers must be 16-byte aligned");
  if (h->multi) return scfb_multi_build_dev(h, n_spin, d_density, d_total_density, d_out, nullptr, nullptr, true);
  return scfb_build_jk(h, n_spin, d_density, d_total_density, d_out, nullptr, nullptr);
}

int scfb_fock_jk(scfb_handle_t h, int n_spin, const double* density, const double* total_density,
                 double* out) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = scfb_check_build_args(h, n_spin, density, total_density, out)) return rc;
  if (h->multi) {
    if (int rc = scfb_multi_stage_inputs(h, n_spin, density, total_density, out)) return rc;
    if (int rc = scfb_multi_build_dev(h, n_spin, h->d_density, h->d_total, h->d_out, h->ev_side[1], nullptr, false))
      return rc;
  } else {
    if (int rc = scfb_stage_inputs(h, n_spin, density, total_density, out)) return rc;
    if (int rc = scfb_build_jk(h, n_spin, h->d_density, h->d_total, h->d_out, h->ev_side[1], nullptr)) return rc;
  }
  const size_t n2 = (size_t)h->n * h->n;
  const bool pinned = scfb_is_pinned(out);
  SCFB_CUDA(h, cudaMemcpyAsync(pinned ? out : h->h_stage + 3 * n2, h->d_out, n2 * n_spin * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
  SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->multi)
    if (int rc = scfb_multi_inputs_consumed(h)) return rc;
  if (!pinned) std::memcpy(out, h->h_stage + 3 * n2, n2 * n_spin * sizeof(double));
  return scfb_peer_check(h);
}

int scfb_jk_split(scfb_handle_t h, int n_spin, const double* density, const double* total_density,
                  double* J, double* K) {
  if (int rc = check_handle(h)) return rc;
  if (int rc = scfb_check_build_args(h, n_spin, density, total_density, J)) return rc;
  SCFB_REQUIRE(h, K, "null K pointer");
  const double* jk_final = nullptr;
  if (h->multi) {
    if (int rc = scfb_multi_stage_inputs(h, n_spin, density, total_density, nullptr)) return rc;
    if (int rc = scfb_multi_build_dev(h, n_spin, h->d_density, h->d_total, nullptr, nullptr, &jk_final, false)) return rc;
  } else {
    if (int rc = scfb_stage_inputs(h, n_spin, density, total_density, nullptr)) return rc;
    if (int rc = scfb_build_jk(h, n_spin, h->d_density, h->d_total, nullptr, nullptr, &jk_final)) return rc;
  }
  const size_t n2 = (size_t)h->n * h->n;
  SCFB_CUDA(h, cudaMemcpyAsync(h->h_stage, jk_final, (1 + n_spin) * n2 * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
  SCFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->multi)
    if (int rc = scfb_multi_inputs_consumed(h)) return rc;
  std::memcpy(J, h->h_stage, n2 * sizeof(double));
  std::memcpy(K, h->h_stage + n2, n2 * n_spin * sizeof(double));
  return scfb_peer_check(h);
}

int scfb_full_plan_query(int n_bas, int n_slabs, int n_sm, int* plan6) {
  if (!plan6 || n_bas < 1 || n_slabs < 0) return scfb_fail(nullptr, SCFB_ERR_ARG, "bad arguments");
  const ScfbFullPlan pl = scfb_full_plan(n_bas, n_slabs, n_sm);
  plan6[0] = pl.tma ? 1 : 0;
  plan6[1] = pl.tma ? pl.whole_slabs : n_slabs;
  plan6[2] = pl.tma ? pl.tail_bs : pl.b_splits;
  plan6[3] = (int)pl.kpart_rows;
  plan6[4] = (int)pl.jpart_rows;
  plan6[5] = pl.n_groups;
  return SCFB_OK;
}

int scfb_last_build_ms(scfb_handle_t h, double* stream_ms, double* total_ms) {
  if (int rc = check_handle(h)) return rc;
  if (h->build_no == 0) return scfb_fail(h, SCFB_ERR_STATE, "no build has run on this handle");
  SCFB_CUDA(h, cudaEventSynchronize(scfb_ev(h, 2)));
  float a = 0.f, b = 0.f;
  SCFB_CUDA(h, cudaEventElapsedTime(&a, scfb_ev(h, 3), scfb_ev(h, 1)));
  SCFB_CUDA(h, cudaEventElapsedTime(&b, scfb_ev(h, 0), scfb_ev(h, 2)));
  if (stream_ms) *stream_ms = a;
  if (total_ms) *total_ms = b;
  return SCFB_OK;
}

int scfb_build_ms_history(scfb_handle_t h, int max_builds, double* stream_ms, double* total_ms, int* n_builds) {
  if (int rc = check_handle(h)) return rc;
  SCFB_REQUIRE(h, max_builds >= 0 && n_builds, "bad arguments");
  int count = max_builds;
  if ((uint64_t)count > h->build_no) count = (int)h->build_no;
  if (count > SCFB_EV_RING) count = SCFB_EV_RING;
  *n_builds = count;
  if (count == 0) return SCFB_OK;
  SCFB_CUDA(h, cudaEventSynchronize(scfb_ev(h, 2)));
  for (int i = 0; i < count; ++i) {  // oldest first
    const int slot = (int)((h->build_no - count + i) % SCFB_EV_RING);
    cudaEvent_t* ev = h->ev_ring[slot];
    const unsigned char* m = h->ev_map[slot];
    float a = 0.f, b = 0.f;
    SCFB_CUDA(h, cudaEventElapsedTime(&a, ev[m[3]], ev[m[1]]));
    SCFB_CUDA(h, cudaEventElapsedTime(&b, ev[m[0]], ev[m[2]]));
    if (stream_ms) stream_ms[i] = a;
    if (total_ms) total_ms[i] = b;
  }
  return SCFB_OK;
}

int scfb_launch_count(scfb_handle_t h, uint64_t* launches) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  if (launches) {
    *launches = 0;
    for (int g = 0; g < n_members(h); ++g) *launches += member(h, g)->launches;
  }
  return SCFB_OK;
}

}  // extern "C"
