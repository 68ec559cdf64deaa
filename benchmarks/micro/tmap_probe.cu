// Probe: does cuTensorMapEncodeTiled accept a 2-D fp64 map whose rows OVERLAP in memory
// (globalDim[0] = whole buffer, globalStrides[0] = L*8 bytes with L << globalDim[0]) and does a
// {64 x 8} box at an arbitrary 64-byte-aligned x0 then deliver the 8 lines base + x0 + y*L ?
// (What the packed J/K kernel needs to fetch a (k-chunk, l-block) tile with ONE cp.async.bulk.tensor.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, double* out) {
  __shared__ __align__(128) double tile[8 * 64];
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(tile);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(d), "l"(&tm), "r"(x0), "r"(0), "r"(b) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(8 * 64 * 8) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = tile[i];
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  const uint64_t N = 1 << 20;
  std::vector<double> h(N);
  for (uint64_t i = 0; i < N; ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, N * 8);
  cudaMalloc(&out, 512 * 8);
  cudaMemcpy(d, h.data(), N * 8, cudaMemcpyHostToDevice);
  for (uint64_t L : {72ull, 392ull}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {N, 8};
    cuuint64_t strides[1] = {L * 8};
    cuuint32_t box[2] = {64, 8};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("L=%llu encode -> %d\n", (unsigned long long)L, (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int x0 : {0, 8, 1000, (int)(N - 200)}) {
      probe<<<1, 128>>>(tm, x0, out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<double> o(512);
      cudaMemcpy(o.data(), out, 512 * 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 64; ++x) {
          const uint64_t src = (uint64_t)x0 + x + y * L;
          const double want = (uint64_t)x0 + x < N && src < N ? (double)src : 0.0;   // OOB along dim 0 -> zero fill
          if (o[y * 64 + x] != want) ++bad;
        }
      printf("  x0=%d: %s, mismatches %d (o[0]=%.0f o[64]=%.0f o[511]=%.0f)\n", x0, cudaGetErrorString(e), bad, o[0], o[64], o[511]);
    }
  }
  return 0;
}
