// Microbenchmark: how fast can one-warp CTAs pull HBM into shared memory with 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) as a function of the copy size?  Design input for the packed8
// Fock kernel (csrc/jk_packed.cu): its per-(row, k-segment) runs are <= 512 B per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_rate bulk_copy_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int DEPTH = 4;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every CTA (one warp) streams `bytes_per_cta` starting at its own offset, in stages of
// `per_stage` copies of `copy_bytes` each (copies of one stage are `stride` bytes apart)
__global__ void __launch_bounds__(32) bulk_rate(const char* __restrict__ src, size_t bytes_per_cta,
                                                int copy_bytes, int per_stage, size_t stride,
                                                double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem + 128;
  const int lane = threadIdx.x;
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  const size_t stage_bytes = (size_t)copy_bytes * per_stage;
  const int n_stage = (int)(bytes_per_cta / stage_bytes);
  if (lane == 0) {
    for (int d = 0; d < DEPTH; ++d)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int st) {
    const int slot = st % DEPTH;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + slot)),
                 "r"((uint32_t)stage_bytes) : "memory");
    for (int c = 0; c < per_stage; ++c)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(s32(ring + (size_t)slot * stage_bytes + (size_t)c * copy_bytes)),
                   "l"(base + (size_t)st * stage_bytes + (size_t)c * copy_bytes), "r"((uint32_t)copy_bytes),
                   "r"(s32(bars + slot)) : "memory");
    (void)stride;
  };
  if (lane == 0)
    for (int d = 0; d < DEPTH - 1 && d < n_stage; ++d) issue(d);
  double acc = 0.0;
  for (int st = 0; st < n_stage; ++st) {
    __syncwarp();
    if (lane == 0 && st + DEPTH - 1 < n_stage) issue(st + DEPTH - 1);
    const uint32_t parity = (st / DEPTH) & 1;
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::
                 "r"(s32(bars + st % DEPTH)), "r"(parity) : "memory");
    acc += reinterpret_cast<const double*>(ring + (size_t)(st % DEPTH) * stage_bytes)[lane % (stage_bytes / 8)];
  }
  if (acc == 123.456) sink[0] = acc;
}

int main() {
  const size_t total = 16ull << 30;
  char* src;
  double* sink;
  cudaMalloc(&src, total);
  cudaMalloc(&sink, 8);
  cudaMemset(src, 0, total);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int stage_kb = 4;  // like one row of the packed kernel: 8 x 512 B
  printf("copy_bytes per_stage ctas_per_sm  GB/s  copies/us/SM\n");
  for (int ctas_per_sm : {4, 8, 16}) {
    for (int copy_bytes : {128, 256, 512, 1024, 2048, 4096}) {
      const int per_stage = stage_kb * 1024 / copy_bytes;
      const int n_cta = 148 * ctas_per_sm * 4;
      const size_t per_cta = (total / n_cta) / (stage_kb * 1024) * (stage_kb * 1024);
      const size_t smem = 128 + (size_t)DEPTH * stage_kb * 1024;
      cudaFuncSetAttribute(bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      // occupancy is limited through the shared-memory request: pad it up to 227 KB / ctas_per_sm
      size_t pad = (227 * 1024 / ctas_per_sm) - 1024;
      if (pad < smem) pad = smem;
      cudaFuncSetAttribute(bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        bulk_rate<<<n_cta, 32, pad>>>(src, per_cta, copy_bytes, per_stage, 0, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      const double bytes = (double)per_cta * n_cta;
      printf("%9d %9d %11d %7.0f %9.1f\n", copy_bytes, per_stage, ctas_per_sm, bytes / best / 1e6,
             bytes / copy_bytes / (best * 1e3) / 148);
    }
  }
  return 0;
}
