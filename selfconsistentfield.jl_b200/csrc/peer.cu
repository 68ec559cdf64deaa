// K6b — peer-memory exchange of the partial [J | K_a | K_b] buffers (opt-in alternative to the
// NCCL allreduce of comm.cu; one process per GPU, all GPUs of one NVSwitch box).
//
// Every rank owns a receive area [2 parities][n_ranks slots][3 n^2] in its own HBM, exported
// with CUDA IPC.  After the J/K kernels of a build, `px_send_kernel` stores this rank's partial
// (full layout: only the owned nu-columns; packed: everything) straight into slot[rank] of EVERY
// peer over NVLink (plain st.global on IPC-mapped pointers), fences system-wide, and the last CTA
// publishes the build's epoch in each peer's flag word.  `px_recv_*` waits (bounded spin) until
// all n_ranks flags of ITS OWN flag array show the epoch, then sums the slots in rank order —
// a one-shot all-to-all + local reduction: deterministic, no ring, one NVLink hop.
// Double buffering by epoch parity is enough: a peer can only start build k+1's stores after it
// saw my flag k, which I publish after my own reads of build k-1 have completed (stream order).
#include "scfb_internal.h"

namespace {

constexpr int kFlagStride = 16;             // uint64 per flag slot (128 B apart)
constexpr unsigned long long kSpinLimit = 1ull << 22;  // x ~1 us  ->  give up after ~4 s

struct PeerPtrs {
  double* slots[8];
  unsigned long long* flags[8];
};

// region = (1 + n_spin) blocks of `cols` columns starting at column col_lo (n doubles each)
__global__ void px_send_kernel(const double* __restrict__ jk, PeerPtrs peers, int n_ranks, int rank,
                               int parity, int n, int n_mats, int col_lo, int cols,
                               unsigned long long epoch, unsigned int* __restrict__ done_counter) {
  const size_t n2 = (size_t)n * n;
  const size_t slot_elems = 3 * n2;
  const size_t per_mat = (size_t)cols * n;
  const size_t total = per_mat * n_mats;
  const size_t slot_off = ((size_t)parity * n_ranks + rank) * slot_elems;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t m = e / per_mat, r = e - m * per_mat;
    const size_t idx = m * n2 + (size_t)col_lo * n + r;
    const double v = jk[idx];
    for (int p = 0; p < n_ranks; ++p) peers.slots[p][slot_off + idx] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {  // last CTA: every store of this launch is fenced
      *done_counter = 0;
      __threadfence_system();
      for (int p = 0; p < n_ranks; ++p) {
        volatile unsigned long long* f = peers.flags[p] + (size_t)rank * kFlagStride;
        *f = epoch;
      }
      __threadfence_system();
    }
  }
}

__device__ __forceinline__ bool wait_flags(const unsigned long long* flags, int n_ranks,
                                           unsigned long long epoch, int* __restrict__ timed_out) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    for (int r = 0; r < n_ranks && ok; ++r) {
      const volatile unsigned long long* f = flags + (size_t)r * kFlagStride;
      unsigned long long spins = 0;
      while (*f < epoch) {
        __nanosleep(1000);
        if (++spins > kSpinLimit) {
          ok = 0;
          *timed_out = 1;
          break;
        }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  return ok != 0;
}

// out[:,:,s] += sum_r slot_r.J - sum_r slot_r.K_s   (rank order: deterministic)
__global__ void px_recv_accumulate_kernel(const double* __restrict__ slots,
                                          const unsigned long long* __restrict__ flags, int n_ranks,
                                          int parity, int n, int n_spin, unsigned long long epoch,
                                          double* __restrict__ out, int* __restrict__ timed_out) {
  if (!wait_flags(flags, n_ranks, epoch, timed_out)) return;
  const size_t n2 = (size_t)n * n, slot_elems = 3 * n2;
  const double* base = slots + (size_t)parity * n_ranks * slot_elems;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * n_spin;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / n2, e = i - s * n2;
    double j = 0.0, k = 0.0;
    for (int r = 0; r < n_ranks; ++r) {
      j += base[r * slot_elems + e];
      k += base[r * slot_elems + (1 + s) * n2 + e];
    }
    out[i] += j - k;
  }
}

// jk_sum = sum_r slot_r   (for scfb_jk_split)
__global__ void px_recv_sum_kernel(const double* __restrict__ slots,
                                   const unsigned long long* __restrict__ flags, int n_ranks,
                                   int parity, int n, int n_mats, unsigned long long epoch,
                                   double* __restrict__ jk_sum, int* __restrict__ timed_out) {
  if (!wait_flags(flags, n_ranks, epoch, timed_out)) return;
  const size_t n2 = (size_t)n * n, slot_elems = 3 * n2;
  const double* base = slots + (size_t)parity * n_ranks * slot_elems;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 * n_mats;
       i += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < n_ranks; ++r) v += base[r * slot_elems + i];
    jk_sum[i] = v;
  }
}

struct PeerState {
  double* slots = nullptr;              // local receive area
  unsigned long long* flags = nullptr;  // local flag words
  unsigned int* counter = nullptr;
  int* timed_out = nullptr;             // device flag
  PeerPtrs peers{};
  void* opened[16] = {nullptr};
  int n_opened = 0;
  unsigned long long epoch = 0;
  bool ready = false;
};

PeerState* state(scfb_handle_s* h) { return static_cast<PeerState*>(h->peer); }

}  // namespace

bool scfb_peer_ready(scfb_handle_s* h) { return h->peer && state(h)->ready; }

void scfb_peer_destroy(scfb_handle_s* h) {
  PeerState* st = state(h);
  if (!st) return;
  for (int i = 0; i < st->n_opened; ++i) cudaIpcCloseMemHandle(st->opened[i]);
  cudaFree(st->slots);
  cudaFree(st->flags);
  cudaFree(st->counter);
  cudaFree(st->timed_out);
  delete st;
  h->peer = nullptr;
}

// send this rank's partial to every peer; `to_out` != nullptr: accumulate, else fill jk_sum
int scfb_peer_exchange(scfb_handle_s* h, int n_spin, double* d_out) {
  PeerState* st = state(h);
  const int n = h->n;
  const int n_mats = 1 + n_spin;
  const unsigned long long epoch = ++st->epoch;
  const int parity = (int)(epoch & 1);
  const bool full = h->layout == SCFB_LAYOUT_FULL;
  const int col_lo = full ? h->nu_lo : 0;
  const int cols = full ? h->nu_hi - h->nu_lo : n;
  const size_t total = (size_t)cols * n * n_mats;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  px_send_kernel<<<blocks, 256, 0, h->stream>>>(h->jk, st->peers, h->n_ranks, h->rank, parity, n,
                                                n_mats, col_lo, cols, epoch, st->counter);
  const size_t n2 = (size_t)n * n;
  int rblocks = (int)((n2 * n_mats + 255) / 256);
  if (rblocks > 148 * 4) rblocks = 148 * 4;
  if (d_out)
    px_recv_accumulate_kernel<<<rblocks, 256, 0, h->stream>>>(st->slots, st->flags, h->n_ranks,
                                                             parity, n, n_spin, epoch, d_out,
                                                             st->timed_out);
  else
    px_recv_sum_kernel<<<rblocks, 256, 0, h->stream>>>(st->slots, st->flags, h->n_ranks, parity, n,
                                                      n_mats, epoch, h->jk_sum, st->timed_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return scfb_fail(h, SCFB_ERR_CUDA, "peer exchange launch: %s", cudaGetErrorString(e));
  h->launches += 2;
  return SCFB_OK;
}

int scfb_peer_check(scfb_handle_s* h) {  // after a stream sync: did a wait give up?
  PeerState* st = state(h);
  if (!st || !st->ready) return SCFB_OK;
  int flag = 0;
  cudaError_t e = cudaMemcpy(&flag, st->timed_out, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return scfb_fail(h, SCFB_ERR_CUDA, "peer check: %s", cudaGetErrorString(e));
  if (flag) return scfb_fail(h, SCFB_ERR_NCCL, "peer exchange timed out waiting for another rank");
  return SCFB_OK;
}

extern "C" {

// 128 bytes: [cudaIpcMemHandle_t of the receive area | cudaIpcMemHandle_t of the flag words]
int scfb_peer_export(scfb_handle_t h, char* handle128) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, handle128, "null buffer");
  SCFB_REQUIRE(h, h->n_ranks >= 1 && h->n_ranks <= 8, "peer exchange supports up to 8 ranks");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  if (!h->peer) h->peer = new PeerState();
  PeerState* st = state(h);
  const size_t n2 = (size_t)h->n * h->n;
  const size_t slot_bytes = 2 * (size_t)h->n_ranks * 3 * n2 * sizeof(double);
  if (!st->slots) {
    SCFB_CUDA(h, cudaMalloc(&st->slots, slot_bytes));
    SCFB_CUDA(h, cudaMemset(st->slots, 0, slot_bytes));
    SCFB_CUDA(h, cudaMalloc(&st->flags, 8 * kFlagStride * sizeof(unsigned long long)));
    SCFB_CUDA(h, cudaMemset(st->flags, 0, 8 * kFlagStride * sizeof(unsigned long long)));
    SCFB_CUDA(h, cudaMalloc(&st->counter, sizeof(unsigned int)));
    SCFB_CUDA(h, cudaMemset(st->counter, 0, sizeof(unsigned int)));
    SCFB_CUDA(h, cudaMalloc(&st->timed_out, sizeof(int)));
    SCFB_CUDA(h, cudaMemset(st->timed_out, 0, sizeof(int)));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  cudaIpcMemHandle_t a, b;
  SCFB_CUDA(h, cudaIpcGetMemHandle(&a, st->slots));
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
  if (!st || !st->slots) return scfb_fail(h, SCFB_ERR_STATE, "call scfb_peer_export first");
  SCFB_CUDA(h, cudaSetDevice(h->device));
  for (int r = 0; r < h->n_ranks; ++r) {
    if (r == h->rank) {
      st->peers.slots[r] = st->slots;
      st->peers.flags[r] = st->flags;
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
    st->peers.slots[r] = static_cast<double*>(pa);
    st->peers.flags[r] = static_cast<unsigned long long*>(pb);
  }
  st->ready = true;
  return SCFB_OK;
}

}  // extern "C"
