// write_ceiling2.cu — what separates a row-granular writer (5.6-5.85 TB/s) from a flat stream (6.65 TB/s)?
// Variables: row size/alignment, warps per SM, rows per warp (kernel length), writer (STG.128 vs TMA).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MINB>
__global__ void __launch_bounds__(128, MINB) stg_rows(char* dst, long n_rows, int row_bytes, int interleave) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long)gridDim.x * (blockDim.x >> 5);
  const long per = (n_rows + nw - 1) / nw;
  double2 v = make_double2(1.0, 0.0);
  for (long r = 0; r < per; ++r) {
    long row = interleave ? r * nw + gw : gw * per + r;
    if (row >= n_rows) break;
    double2* p = reinterpret_cast<double2*>(dst + row * row_bytes);
    for (int k = lane; k < row_bytes / 16; k += 32) p[k] = v;
  }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) tma_rows(char* dst, long n_rows, int row_bytes, int interleave) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + (size_t)warp * row_bytes;
  for (int i = lane * 8; i < row_bytes; i += 256) *reinterpret_cast<double*>(stage + i) = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long)gridDim.x * (blockDim.x >> 5);
  const long per = (n_rows + nw - 1) / nw;
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage);
  if (lane == 0) {
    for (long r = 0; r < per; ++r) {
      long row = interleave ? r * nw + gw : gw * per + r;
      if (row >= n_rows) break;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + row * row_bytes), "r"(s), "r"(row_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

int main() {
  const size_t bytes = (size_t)262144 * 2224;
  char* d;
  cudaMalloc(&d, bytes + 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(tma_rows<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(tma_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  struct Cfg { const char* name; int kind; int row_bytes; int blocks; int interleave; };
  Cfg cfgs[] = {
      {"STG rows 2224 B, 1024 blk (7/SM), region", 0, 2224, 1024, 0},
      {"STG rows 2048 B, 1024 blk (7/SM), region", 0, 2048, 1024, 0},
      {"STG rows 2224 B, 2368 blk (16/SM), region", 1, 2224, 2368, 0},
      {"STG rows 2224 B, 2368 blk (16/SM), interleaved", 1, 2224, 2368, 1},
      {"STG rows 2048 B, 2368 blk (16/SM), interleaved", 1, 2048, 2368, 1},
      {"STG rows 2224 B, 4096 blk (16/SM), interleaved", 1, 2224, 4096, 1},
      {"STG rows 2224 B, 592 blk (4/SM), region", 0, 2224, 592, 0},
      {"TMA rows 2224 B, 1024 blk (7/SM), region", 2, 2224, 1024, 0},
      {"TMA rows 2048 B, 1024 blk (7/SM), region", 2, 2048, 1024, 0},
      {"TMA rows 2224 B, 2368 blk (16/SM), interleaved", 3, 2224, 2368, 1},
      {"TMA rows 4448 B, 1024 blk (7/SM), region", 2, 4448, 1024, 0},
      {"TMA rows 1232 B, 1024 blk (7/SM), region", 2, 1232, 1024, 0},
  };
  for (auto& c : cfgs) {
    long n_rows = bytes / c.row_bytes;
    const int reps = 30;
    for (int it = -3; it < reps; ++it) {
      if (it == 0) cudaEventRecord(e0);
      if (c.kind == 0) stg_rows<7><<<c.blocks, 128>>>(d, n_rows, c.row_bytes, c.interleave);
      else if (c.kind == 1) stg_rows<16><<<c.blocks, 128>>>(d, n_rows, c.row_bytes, c.interleave);
      else if (c.kind == 2) tma_rows<7><<<c.blocks, 128, 4 * c.row_bytes>>>(d, n_rows, c.row_bytes, c.interleave);
      else tma_rows<16><<<c.blocks, 128, 4 * c.row_bytes>>>(d, n_rows, c.row_bytes, c.interleave);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-52s %8.2f us %8.1f GB/s (%s)\n", c.name, ms * 1e3, (double)n_rows * c.row_bytes / 1e9 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
