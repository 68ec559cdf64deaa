// One process, G devices, ONE handle (SURVEY.md 8b "Threading" / 8e "Process model"): what lets a
// single-threaded caller — the reference's run_scf (/root/reference/src/run_scf.jl:4-65) calling
// add_2e_term! once per Fock build (src/parts/compute_fock_matrix.jl:31-34) — use every GPU of the
// box, e.g. for the n_bf = 512 tensor (550 GB) that only exists sharded.
//
// The group is G ordinary per-rank handles (rank g on device dev_ids[g], its own stream, its own
// nu-slab / i-block shard) wired to each other with PEER ACCESS instead of CUDA IPC; member 0 is
// the handle the caller holds.  A build fans the densities out to every device (host-pointer API:
// one H2D copy per device, each over its own PCIe link; device-pointer API: peer copies from
// device 0), launches every member's J/K kernels from the calling thread and lets the peer-memory
// exchange of peer.cu (finished columns stored straight into every member's landing zone by the
// stream kernel itself; slots for the packed layout) deliver the sum to device 0, where
// out += J - K_s runs.  No host thread ever waits for another device: cross-device ordering is
// CUDA events plus the exchange's epoch flags.
#include "scfb_internal.h"

#include <new>

struct scfb_multi_s {
  int n_dev = 0;
  scfb_handle_s* member[8] = {};   // member[0] is the handle the caller holds
  cudaEvent_t ev_src = nullptr;    // device 0: the caller's device arrays are ready to be copied
  cudaEvent_t ev_h2d[8] = {};      // per member: its H2D copies of the caller's host arrays are done
};

namespace {

int set_dev(scfb_handle_s* m) {
  cudaError_t e = cudaSetDevice(m->device);
  if (e != cudaSuccess) return scfb_fail(m, SCFB_ERR_CUDA, "cudaSetDevice(%d): %s", m->device, cudaGetErrorString(e));
  return SCFB_OK;
}

// a member's error becomes the caller-visible handle's error
int lift(scfb_handle_s* h0, scfb_handle_s* m, int rc) {
  if (rc != SCFB_OK && m != h0) std::snprintf(h0->err, sizeof h0->err, "device %d: %.480s", m->device, m->err);
  return rc;
}

}  // namespace

extern "C" int scfb_create_multi(scfb_handle_t* out, int n_bas, int layout, int n_dev, const int* dev_ids) {
  if (!out) return scfb_fail(nullptr, SCFB_ERR_ARG, "null output pointer");
  *out = nullptr;
  SCFB_REQUIRE(nullptr, n_dev >= 1 && n_dev <= 8, "n_dev must be 1..8 (got %d)", n_dev);
  SCFB_REQUIRE(nullptr, dev_ids, "null device list");
  for (int a = 0; a < n_dev; ++a)
    for (int b = a + 1; b < n_dev; ++b)
      SCFB_REQUIRE(nullptr, dev_ids[a] != dev_ids[b], "device %d listed twice", dev_ids[a]);
  if (n_dev == 1) return scfb_create(out, n_bas, layout, dev_ids[0], 0, 1);

  scfb_multi_s* ms = new (std::nothrow) scfb_multi_s();
  if (!ms) return scfb_fail(nullptr, SCFB_ERR_OOM, "host allocation failed");
  ms->n_dev = n_dev;
  auto fail = [&](int rc) {
    for (int g = n_dev - 1; g >= 0; --g)
      if (ms->member[g]) {
        ms->member[g]->multi = nullptr;
        scfb_destroy(ms->member[g]);
      }
    delete ms;
    return rc;
  };
  for (int g = 0; g < n_dev; ++g)
    if (int rc = scfb_create(&ms->member[g], n_bas, layout, dev_ids[g], g, n_dev)) return fail(rc);
  // peer access in both directions between all members (one NVSwitch box: every pair is connected)
  for (int a = 0; a < n_dev; ++a) {
    cudaSetDevice(dev_ids[a]);
    for (int b = 0; b < n_dev; ++b) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, dev_ids[a], dev_ids[b]);
      if (!can)
        return fail(scfb_fail(nullptr, SCFB_ERR_CUDA, "device %d cannot access device %d's memory (no peer path)",
                              dev_ids[a], dev_ids[b]));
      cudaError_t e = cudaDeviceEnablePeerAccess(dev_ids[b], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess)
        return fail(scfb_fail(nullptr, SCFB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dev_ids[a],
                              dev_ids[b], cudaGetErrorString(e)));
    }
  }
  double* areas[8];
  unsigned long long* flags[8];
  for (int g = 0; g < n_dev; ++g) {
    cudaSetDevice(dev_ids[g]);
    if (int rc = scfb_peer_alloc_local(ms->member[g])) return fail(rc);
    scfb_peer_local_ptrs(ms->member[g], &areas[g], &flags[g]);
    if (cudaEventCreateWithFlags(&ms->ev_h2d[g], cudaEventDisableTiming) != cudaSuccess)
      return fail(scfb_fail(nullptr, SCFB_ERR_CUDA, "event creation failed"));
  }
  for (int g = 0; g < n_dev; ++g)
    if (int rc = scfb_peer_wire(ms->member[g], areas, flags)) return fail(rc);
  cudaSetDevice(dev_ids[0]);
  if (cudaEventCreateWithFlags(&ms->ev_src, cudaEventDisableTiming) != cudaSuccess)
    return fail(scfb_fail(nullptr, SCFB_ERR_CUDA, "event creation failed"));
  ms->member[0]->multi = ms;
  *out = ms->member[0];
  return SCFB_OK;
}

// ---- what the extern "C" entry points of scfb_api.cu dispatch to when h->multi is set ----------
void scfb_multi_destroy(scfb_handle_s* h0) {
  scfb_multi_s* ms = h0->multi;
  h0->multi = nullptr;
  for (int g = 0; g < ms->n_dev; ++g) {  // nobody may still be storing into a peer that is going away
    cudaSetDevice(ms->member[g]->device);
    cudaStreamSynchronize(ms->member[g]->stream);
  }
  for (int g = ms->n_dev - 1; g >= 1; --g) scfb_destroy(ms->member[g]);
  for (int g = 0; g < ms->n_dev; ++g)
    if (ms->ev_h2d[g]) cudaEventDestroy(ms->ev_h2d[g]);
  if (ms->ev_src) cudaEventDestroy(ms->ev_src);
  delete ms;
}

int scfb_multi_n_dev(scfb_handle_s* h0) { return h0->multi->n_dev; }
scfb_handle_s* scfb_multi_member(scfb_handle_s* h0, int g) { return h0->multi->member[g]; }

int scfb_multi_sync(scfb_handle_s* h0) {
  scfb_multi_s* ms = h0->multi;
  for (int g = ms->n_dev - 1; g >= 0; --g) {
    scfb_handle_s* m = ms->member[g];
    if (int rc = set_dev(m)) return lift(h0, m, rc);
    if (int rc = scfb_stream_sync_checked(m)) return lift(h0, m, rc);
  }
  return SCFB_OK;
}

// Densities already on device 0 (d_density, d_total on h0's stream): peer-copy them to the other
// members, launch every member's build; device 0 accumulates into d_out (nullptr: jk_sum only).
int scfb_multi_build_dev(scfb_handle_s* h0, int n_spin, const double* d_density, const double* d_total,
                         double* d_out, cudaEvent_t out_event, const double** jk_final, bool fan_out) {
  scfb_multi_s* ms = h0->multi;
  const size_t n2 = (size_t)h0->n * h0->n;
  if (fan_out) {
    if (int rc = set_dev(h0)) return rc;
    SCFB_CUDA(h0, cudaEventRecord(ms->ev_src, h0->stream));
  }
  for (int g = 1; g < ms->n_dev; ++g) {
    scfb_handle_s* m = ms->member[g];
    if (int rc = set_dev(m)) return lift(h0, m, rc);
    if (fan_out) {
      cudaError_t e = cudaStreamWaitEvent(m->stream, ms->ev_src, 0);
      if (e == cudaSuccess)
        e = cudaMemcpyPeerAsync(m->d_density, m->device, d_density, h0->device, n2 * n_spin * sizeof(double), m->stream);
      if (e == cudaSuccess)
        e = cudaMemcpyPeerAsync(m->d_total, m->device, d_total, h0->device, n2 * sizeof(double), m->stream);
      if (e != cudaSuccess) return scfb_fail(h0, SCFB_ERR_CUDA, "fan-out to device %d: %s", m->device, cudaGetErrorString(e));
    }
    if (int rc = scfb_build_jk(m, n_spin, m->d_density, m->d_total, nullptr, nullptr, nullptr)) return lift(h0, m, rc);
  }
  if (int rc = set_dev(h0)) return rc;
  return scfb_build_jk(h0, n_spin, d_density, d_total, d_out, out_event, jk_final);
}

// Host arrays: every member pulls the densities over its own PCIe link.
int scfb_multi_stage_inputs(scfb_handle_s* h0, int n_spin, const double* density, const double* total,
                            const double* out) {
  scfb_multi_s* ms = h0->multi;
  const size_t n2 = (size_t)h0->n * h0->n;
  if (int rc = set_dev(h0)) return rc;
  if (int rc = scfb_stage_inputs(h0, n_spin, density, total, out)) return rc;  // (pageable: copied to h0's pinned stage)
  const bool pinned = scfb_is_pinned(density) && scfb_is_pinned(total) && (!out || scfb_is_pinned(out));
  const double* src_d = pinned ? density : h0->h_stage;
  const double* src_t = pinned ? total : h0->h_stage + 2 * n2;
  for (int g = 1; g < ms->n_dev; ++g) {
    scfb_handle_s* m = ms->member[g];
    if (int rc = set_dev(m)) return lift(h0, m, rc);
    cudaError_t e = cudaMemcpyAsync(m->d_density, src_d, n2 * n_spin * sizeof(double), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_total, src_t, n2 * sizeof(double), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ms->ev_h2d[g], m->stream);
    if (e != cudaSuccess) return scfb_fail(h0, SCFB_ERR_CUDA, "H2D to device %d: %s", m->device, cudaGetErrorString(e));
  }
  return set_dev(h0);
}

// Before a host entry point returns: no member may still be reading the caller's (or the staging) arrays.
int scfb_multi_inputs_consumed(scfb_handle_s* h0) {
  scfb_multi_s* ms = h0->multi;
  for (int g = 1; g < ms->n_dev; ++g) {
    cudaError_t e = cudaEventSynchronize(ms->ev_h2d[g]);
    if (e != cudaSuccess) return scfb_fail(h0, SCFB_ERR_CUDA, "device %d: %s", ms->member[g]->device, cudaGetErrorString(e));
  }
  return SCFB_OK;
}
