// Microbenchmark: per-SM TMA box-load throughput/latency as a function of boxes in flight and tensor-map rank.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I latent-flexible-video-diffusion-modeling_b200/csrc -I include tools/tma_bench.cu -o gpurun_out/tma_bench -lcuda
#include "tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace fdm;

__global__ void __launch_bounds__(128) bench_kernel(const __grid_constant__ CUtensorMap map, int rank, int depth, int nloads, int box_bytes,
                                                    int frames_per_cta, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < nloads + depth; ++i) {
      int s = i % depth;
      if (i >= depth) mbar_wait(&bar[s], ((i / depth) - 1) & 1);
      if (i < nloads) {
        mbar_expect_tx(&bar[s], box_bytes);
        int n = blockIdx.x * frames_per_cta + (i % frames_per_cta);
        int tap = i % 9;
        if (rank == 4) tma_load_4d(smem + s * box_bytes, &map, &bar[s], 0, tap % 3 - 1, (i % 8) * 4 + tap / 3 - 1, n);
        else tma_load_3d(smem + s * box_bytes, &map, &bar[s], 0, ((i % 8) * 128) , n);
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int N = 148 * 4, H = 32, W = 32, C = 64;
  size_t bytes = (size_t)N * H * W * C * 2;
  void* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
  long long* out; cudaMalloc(&out, 148 * 8);
  EncodeTiledFn enc = nullptr; { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (EncodeTiledFn)p; }
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int rank : {4, 3}) for (int rows : {128, 320}) for (int depth : {1, 2, 4, 8}) {
    if (rows * 128 * depth > 200 * 1024) continue;
    CUtensorMap m;
    if (rank == 4) {
      cuuint64_t dims[4] = {C, W, H, N}; cuuint64_t st[3] = {C * 2, W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {64, 32, (cuuint32_t)(rows / 32), 1}; cuuint32_t es[4] = {1, 1, 1, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      if (rows > 256) continue;
      cuuint64_t dims[3] = {C, (cuuint64_t)H * W, N}; cuuint64_t st[2] = {C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)rows, 1}; cuuint32_t es[3] = {1, 1, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    int nloads = 64;
    for (int rep = 0; rep < 2; ++rep) bench_kernel<<<148, 128, rows * 128 * depth + 1024>>>(m, rank, depth, nloads, rows * 128, 4, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148); cudaMemcpy(h.data(), out, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= 148;
    printf("rank %d box %3d rows (%6d B) depth %d: %8.0f clk per box, %6.1f B/clk/SM  (%s)\n", rank, rows, rows * 128, depth, avg / nloads, rows * 128.0 * nloads / avg, cudaGetErrorString(e));
  }
  return 0;
}
