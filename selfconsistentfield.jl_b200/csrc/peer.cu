// K6b — exchange of the partial [J | K_a | K_b] buffers over PEER MEMORY (NVLink / NVSwitch),
// the alternative to the NCCL allreduce of comm.cu.  One process per GPU wires it with CUDA IPC
// (scfb_peer_export / scfb_peer_import); the single-process multi-GPU handle (multi.cu) wires the
// same state directly with peer access.
//
// Full layout — FUSED with the stream kernel.  Rank g owns whole nu-slabs, so its J[:,nu] and
// K_s[:,nu] are final, disjoint columns.  The CTA of `jk_full_tma_kernel` that completes a slab
// stores its columns straight into the LANDING ZONE [2 parities][J | K_a | K_b] of EVERY rank
// (plain st.global on peer-mapped pointers) while the other slabs still stream; the kernel's last
// CTA fences system-wide and publishes the build's epoch in this rank's flag word of every peer.
// The exchange is therefore hidden behind the ERI stream tile by tile; what is left after the
// stream kernel is one small kernel that waits for all n_ranks flags and does out += J - K_s from
// its own, now complete, landing zone.  (Kernels without the fused tail — odd n, n > 512, split
// items — push their columns with `px_push_cols_kernel` after the K reduction instead.)
//
// Packed layout — every rank's partial is dense, so the ranks store whole partials into per-rank
// SLOTS [2 parities][n_ranks][3 n^2] of every peer and the receiver sums the slots in rank order
// (one-shot all-to-all + local, deterministic reduction).
//
// Double buffering by epoch parity is enough: a peer can only start build k+1's stores after it
// saw my flag k, which I publish after my own reads of build k-1 have completed (stream order).
// A rank that never arrives makes the wait give up after SCFB_PEER_TIMEOUT_MS (default 20 s):
// ONE CTA waits and publishes the verdict in a gate word the other CTAs follow, so a build is
// accumulated completely or not at all; the next synchronising call returns SCFB_ERR_NCCL.
#include "scfb_internal.h"

#include <cstdlib>

namespace {

constexpr int kFlagStride = 16;  // uint64 per flag slot (128 B apart)
constexpr int kMaxRanks = 8;

struct PeerPtrs {
  double* land[kMaxRanks];   // landing zones (2 x 3 n^2) of every rank
  double* slots[kMaxRanks];  // slot areas (packed layout), may be null
  unsigned long long* flags[kMaxRanks];
};

struct PeerState {
  double* area = nullptr;               // local [landing zones | slots]
  double* slots = nullptr;              // inside area (packed layout only)
  unsigned long long* flags = nullptr;  // local flag words
  unsigned long long* gate = nullptr;   // epoch << 2 | {1 waiting, 2 all flags seen, 3 timed out}
  unsigned int* counter = nullptr;
  int* timed_out = nullptr;  // device flag
  PeerPtrs peers{};
  void* opened[2 * kMaxRanks] = {nullptr};
  int n_opened = 0;
  unsigned long long epoch = 0;
  unsigned long long spin_limit = 0;  // x ~1 us
  bool ready = false;
};

PeerState* state(scfb_handle_s* h) { return static_cast<PeerState*>(h->peer); }

// Thread 0 of every CTA: the first one to claim the gate for `epoch` waits for all flags (bounded)
// and publishes the verdict; the others follow the gate.  Returns the same answer in every CTA.
__device__ __forceinline__ bool gate_wait(const unsigned long long* flags, int n_ranks, unsigned long long epoch,
                                          unsigned long long* gate, int* timed_out, unsigned long long spin_limit) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    volatile unsigned long long* g = gate;
    int res = -1;
    unsigned long long spins = 0;
    while (res < 0) {
      const unsigned long long v = *g;
      if ((v >> 2) == epoch && (v & 3) >= 2) {
        res = (v & 3) == 2;
      } else if ((v >> 2) < epoch && atomicCAS(gate, v, epoch << 2 | 1) == v) {
        int good = 1;
        for (int r = 0; r < n_ranks && good; ++r) {
          const volatile unsigned long long* f = flags + (size_t)r * kFlagStride;
          unsigned long long w = 0;
          while (*f < epoch) {
            __nanosleep(250);
            if (++w > 4 * spin_limit) {
              good = 0;
              break;
            }
          }
        }
        __threadfence_system();
        if (!good) *timed_out = 1;
        *g = epoch << 2 | (good ? 2 : 3);
        __threadfence();
        res = good;
      } else {
        __nanosleep(100);
        if (++spins > 20 * spin_limit * (unsigned long long)n_ranks) res = 0;  // the waiter itself never answered
      }
    }
    ok = res;
    __threadfence_system();
  }
  __syncthreads();
  return ok != 0;
}

__device__ __forceinline__ void publish_epoch(const PeerPtrs& peers, int n_ranks, int rank, unsigned long long epoch,
                                              unsigned int* done_counter) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {  // last CTA: every store of this launch is fenced
      *done_counter = 0;
      __threadfence_system();
      for (int p = 0; p < n_ranks; ++p)
        *reinterpret_cast<volatile unsigned long long*>(peers.flags[p] + (size_t)rank * kFlagStride) = epoch;
      __threadfence_system();
    }
  }
}

// full layout without the fused tail: push the owned columns [col_lo, col_lo + cols) of the n_mats
// matrices into every rank's landing zone, then publish the epoch
__global__ void px_push_cols_kernel(const double* __restrict__ jk, PeerPtrs peers, int n_ranks, int rank, int parity,
                                    int n, int n_mats, int col_lo, int cols, unsigned long long epoch,
                                    unsigned int* __restrict__ done_counter) {
  const size_t n2 = (size_t)n * n;
  const size_t per_mat = (size_t)cols * n;
  const size_t total = per_mat * n_mats;
  const size_t zone = (size_t)parity * 3 * n2;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t m = e / per_mat, r = e - m * per_mat;
    const size_t idx = m * n2 + (size_t)col_lo * n + r;
    const double v = jk[idx];
    for (int p = 0; p < n_ranks; ++p) peers.land[p][zone + idx] = v;
  }
  publish_epoch(peers, n_ranks, rank, epoch, done_counter);
}

// packed layout: the whole dense partial goes to slot[rank] of every peer
__global__ void px_send_slots_kernel(const double* __restrict__ jk, PeerPtrs peers, int n_ranks, int rank, int parity,
                                     int n, int n_mats, unsigned long long epoch,
                                     unsigned int* __restrict__ done_counter) {
  const size_t n2 = (size_t)n * n;
  const size_t total = n2 * n_mats;
  const size_t slot_off = ((size_t)parity * n_ranks + rank) * 3 * n2;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const double v = jk[e];
    for (int p = 0; p < n_ranks; ++p) peers.slots[p][slot_off + e] = v;
  }
  publish_epoch(peers, n_ranks, rank, epoch, done_counter);
}

// Wait for every rank's epoch, then (n_src = 1: landing zone, n_src = n_ranks: slots, summed in
// rank order)  out[:,:,s] += J - K_s  (out != null)  or  jk_sum = [J | K ...]  (out == null).
// The sources are written by peer GPUs while this kernel may already be resident: read through L2.
__global__ void px_wait_reduce_kernel(const double* src, int n_src, const unsigned long long* flags, int n_ranks,
                                      int n, int n_spin, unsigned long long epoch, unsigned long long* gate,
                                      double* __restrict__ out, double* __restrict__ jk_sum, int* timed_out,
                                      unsigned long long spin_limit) {
  if (!gate_wait(flags, n_ranks, epoch, gate, timed_out, spin_limit)) return;
  const size_t n2 = (size_t)n * n, stride = 3 * n2;
  if (out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * n_spin; i += (size_t)gridDim.x * blockDim.x) {
      const size_t s = i / n2, e = i - s * n2;
      double j = 0.0, k = 0.0;
      for (int r = 0; r < n_src; ++r) {
        j += __ldcg(src + r * stride + e);
        k += __ldcg(src + r * stride + (1 + s) * n2 + e);
      }
      out[i] += j - k;
    }
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * (1 + n_spin);
         i += (size_t)gridDim.x * blockDim.x) {
      double v = 0.0;
      for (int r = 0; r < n_src; ++r) v += __ldcg(src + r * stride + i);
      jk_sum[i] = v;
    }
  }
}

size_t area_doubles(const scfb_handle_s* h) {
  const size_t n2 = (size_t)h->n * h->n;
  return 2 * 3 * n2 + (h->layout == SCFB_LAYOUT_PACKED8 ? 2 * (size_t)h->n_ranks * 3 * n2 : 0);
}

int alloc_local(scfb_handle_s* h) {
  if (!h->peer) h->peer = new PeerState();
  PeerState* st = state(h);
  if (st->area) return SCFB_OK;
  const size_t n2 = (size_t)h->n * h->n;
  const size_t bytes = area_doubles(h) * sizeof(double);
  SCFB_CUDA(h, cudaMalloc(&st->area, bytes));
  SCFB_CUDA(h, cudaMemset(st->area, 0, bytes));
  st->slots = h->layout == SCFB_LAYOUT_PACKED8 ? st->area + 2 * 3 * n2 : nullptr;
  SCFB_CUDA(h, cudaMalloc(&st->flags, kMaxRanks * kFlagStride * sizeof(unsigned long long)));
  SCFB_CUDA(h, cudaMemset(st->flags, 0, kMaxRanks * kFlagStride * sizeof(unsigned long long)));
  SCFB_CUDA(h, cudaMalloc(&st->gate, sizeof(unsigned long long)));
  SCFB_CUDA(h, cudaMemset(st->gate, 0, sizeof(unsigned long long)));
  SCFB_CUDA(h, cudaMalloc(&st->counter, sizeof(unsigned int)));
  SCFB_CUDA(h, cudaMemset(st->counter, 0, sizeof(unsigned int)));
  SCFB_CUDA(h, cudaMalloc(&st->timed_out, sizeof(int)));
  SCFB_CUDA(h, cudaMemset(st->timed_out, 0, sizeof(int)));
  const char* t = std::getenv("SCFB_PEER_TIMEOUT_MS");
  const long ms = t ? std::atol(t) : 20000;
  st->spin_limit = (unsigned long long)(ms > 0 ? ms : 20000) * 1000ull;
  return SCFB_OK;
}

int blocks_for(scfb_handle_s* h, size_t total, int per_sm) {
  if (h->n_sm == 0) {
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
    if (h->n_sm <= 0) h->n_sm = 148;
  }
  size_t b = (total + 255) / 256;
  if (b > (size_t)h->n_sm * per_sm) b = (size_t)h->n_sm * per_sm;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

bool scfb_peer_ready(scfb_handle_s* h) { return h->peer && state(h)->ready; }

void scfb_peer_destroy(scfb_handle_s* h) {
  PeerState* st = state(h);
  if (!st) return;
  for (int i = 0; i < st->n_opened; ++i) cudaIpcCloseMemHandle(st->opened[i]);
  cudaFree(st->area);
  cudaFree(st->flags);
  cudaFree(st->gate);
  cudaFree(st->counter);
  cudaFree(st->timed_out);
  delete st;
  h->peer = nullptr;
}

// in-process wiring (multi.cu): every handle of the group has allocated its local area
int scfb_peer_alloc_local(scfb_handle_s* h) { return alloc_local(h); }
void scfb_peer_local_ptrs(scfb_handle_s* h, double** area, unsigned long long** flags) {
  *area = state(h)->area;
  *flags = state(h)->flags;
}
int scfb_peer_wire(scfb_handle_s* h, double* const* areas, unsigned long long* const* flags) {
  PeerState* st = state(h);
  if (!st || !st->area) return scfb_fail(h, SCFB_ERR_STATE, "peer area not allocated");
  const size_t n2 = (size_t)h->n * h->n;
  for (int r = 0; r < h->n_ranks; ++r) {
    st->peers.land[r] = areas[r];
    st->peers.slots[r] = h->layout == SCFB_LAYOUT_PACKED8 ? areas[r] + 2 * 3 * n2 : nullptr;
    st->peers.flags[r] = flags[r];
  }
  st->ready = true;
  return SCFB_OK;
}

// Start of a build on a rank with a peer exchange: the next epoch and, for the full layout, what
// the stream kernel's fused tail needs to ship finished columns itself.
void scfb_peer_begin(scfb_handle_s* h, ScfbFuse* fuse) {
  PeerState* st = state(h);
  const unsigned long long epoch = ++st->epoch;
  if (h->layout != SCFB_LAYOUT_FULL) return;
  const size_t zone = (size_t)(epoch & 1) * 3 * h->n * h->n;
  fuse->n_peers = h->n_ranks;
  for (int r = 0; r < h->n_ranks; ++r) {
    fuse->peer_jk[r] = st->peers.land[r] + zone;
    fuse->peer_flag[r] = st->peers.flags[r] + (size_t)h->rank * kFlagStride;
  }
  fuse->epoch = epoch;
}

// After the J/K kernels of the build begun with scfb_peer_begin: ship this rank's partial unless
// the stream kernel already did (`pushed`), then wait for everybody and accumulate into d_out
// (d_out == nullptr: fill jk_sum).
int scfb_peer_exchange(scfb_handle_s* h, int n_spin, double* d_out, bool pushed) {
  PeerState* st = state(h);
  const int n = h->n;
  const int n_mats = 1 + n_spin;
  const unsigned long long epoch = st->epoch;
  const int parity = (int)(epoch & 1);
  const bool full = h->layout == SCFB_LAYOUT_FULL;
  const size_t n2 = (size_t)n * n;
  if (!pushed) {
    if (full) {
      const int cols = h->nu_hi - h->nu_lo;
      px_push_cols_kernel<<<blocks_for(h, (size_t)cols * n * n_mats, 2), 256, 0, h->stream>>>(
          h->jk, st->peers, h->n_ranks, h->rank, parity, n, n_mats, h->nu_lo, cols, epoch, st->counter);
    } else {
      px_send_slots_kernel<<<blocks_for(h, n2 * n_mats, 2), 256, 0, h->stream>>>(
          h->jk, st->peers, h->n_ranks, h->rank, parity, n, n_mats, epoch, st->counter);
    }
    h->launches += 1;
  }
  const double* src = full ? st->area + (size_t)parity * 3 * n2 : st->slots + (size_t)parity * h->n_ranks * 3 * n2;
  px_wait_reduce_kernel<<<blocks_for(h, n2 * n_mats, 4), 256, 0, h->stream>>>(
      src, full ? 1 : h->n_ranks, st->flags, h->n_ranks, n, n_spin, epoch, st->gate, d_out, h->jk_sum, st->timed_out,
      st->spin_limit);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "peer exchange launch: %s", cudaGetErrorString(e));
  h->launches += 1;
  return SCFB_OK;
}

int scfb_peer_check(scfb_handle_s* h) {  // after a stream sync: did a wait give up?
  PeerState* st = state(h);
  if (!st || !st->ready) return SCFB_OK;
  int flag = 0;
  cudaError_t e = cudaMemcpy(&flag, st->timed_out, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "peer check: %s", cudaGetErrorString(e));
  if (flag) {
    cudaMemset(st->timed_out, 0, sizeof(int));  // reported once; the skipped build left `out` untouched
    return scfb_fail(h, SCFB_ERR_NCCL, "peer exchange timed out waiting for another rank (build skipped)");
  }
  return SCFB_OK;
}

extern "C" {

// 128 bytes: [cudaIpcMemHandle_t of the receive area | cudaIpcMemHandle_t of the flag words]
int scfb_peer_export(scfb_handle_t h, char* handle128) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, handle128, "null buffer");
  SCFB_REQUIRE(h, h->n_ranks >= 1 && h->n_ranks <= kMaxRanks, "peer exchange supports up to 8 ranks");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  if (int rc = alloc_local(h)) return rc;
  PeerState* st = state(h);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  cudaIpcMemHandle_t a, b;
  SCFB_CUDA(h, cudaIpcGetMemHandle(&a, st->area));
  SCFB_CUDA(h, cudaIpcGetMemHandle(&b, st->flags));
  std::memcpy(handle128, &a, 64);
  std::memcpy(handle128 + 64, &b, 64);
  return SCFB_OK;
}

// `all_handles`: n_ranks x 128 bytes in rank order (this rank's own entry is ignored)
int scfb_peer_import(scfb_handle_t h, const char* all_handles) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, all_handles, "null buffer");
  PeerState* st = state(h);
  if (!st || !st->area) return scfb_fail(h, SCFB_ERR_STATE, "call scfb_peer_export first");
  if (st->ready) return scfb_fail(h, SCFB_ERR_STATE, "peer handles were already imported on this handle");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  double* areas[kMaxRanks];
  unsigned long long* flags[kMaxRanks];
  for (int r = 0; r < h->n_ranks; ++r) {
    if (r == h->rank) {
      areas[r] = st->area;
      flags[r] = st->flags;
      continue;
    }
    cudaIpcMemHandle_t a, b;
    std::memcpy(&a, all_handles + (size_t)r * 128, 64);
    std::memcpy(&b, all_handles + (size_t)r * 128 + 64, 64);
    void *pa = nullptr, *pb = nullptr;
    SCFB_CUDA(h, cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess));
    st->opened[st->n_opened++] = pa;
    SCFB_CUDA(h, cudaIpcOpenMemHandle(&pb, b, cudaIpcMemLazyEnablePeerAccess));
    st->opened[st->n_opened++] = pb;
    areas[r] = static_cast<double*>(pa);
    flags[r] = static_cast<unsigned long long*>(pb);
  }
  return scfb_peer_wire(h, areas, flags);
}

}  // extern "C"
