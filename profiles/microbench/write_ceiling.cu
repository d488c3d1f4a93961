// write_ceiling.cu — how fast can one B200 write the observation buffer, compute aside?
//   (a) plain coalesced STG.128 stream   (b) 2224-byte rows handed to the TMA engine from shared memory
//   (c) cudaMemset                        (d) device-to-device copy (the MEASURED_PEAKS.json method)
// Build/run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wc write_ceiling.cu && ./wc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void stg_stream(double2* dst, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double2 v = make_double2(1.0, 0.0);
  for (; i < n2; i += stride) dst[i] = v;
}

// each warp sweeps its own contiguous region with 512-byte STG.128 wavefronts (the kernel's ownership pattern)
__global__ void __launch_bounds__(128, 7) stg_regions(char* dst, int n_rows, int row_bytes, int rows_per_warp, int interleave) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
  double2 v = make_double2(1.0, 0.0);
  for (int r = 0; r < rows_per_warp; ++r) {
    int row = interleave ? r * nw + gw : gw * rows_per_warp + r;
    if (row >= n_rows) break;
    double2* p = reinterpret_cast<double2*>(dst + (size_t)row * row_bytes);
    for (int k = lane; k < row_bytes / 16; k += 32) p[k] = v;
  }
}

__global__ void __launch_bounds__(128, 7) tma_rows(char* dst, int n_rows, int row_bytes, int rows_per_warp, int interleave) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + (size_t)warp * row_bytes;
  for (int i = lane * 8; i < row_bytes; i += 256) *reinterpret_cast<double*>(stage + i) = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage);
  if (lane == 0) {
    for (int r = 0; r < rows_per_warp; ++r) {
      int row = interleave ? r * nw + gw : gw * rows_per_warp + r;
      if (row >= n_rows) break;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)row * row_bytes), "r"(s),
                   "r"(row_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  const int n_rows = 262144, row_bytes = 2224;
  const size_t bytes = (size_t)n_rows * row_bytes;
  char *d, *d2;
  cudaMalloc(&d, bytes); cudaMalloc(&d2, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 50;
  for (int variant = 0; variant < 7; ++variant) {
    for (int it = -5; it < reps; ++it) {
      if (it == 0) cudaEventRecord(e0);
      if (variant == 0) stg_stream<<<148 * 16, 256>>>((double2*)d, bytes / 16);
      else if (variant == 1) tma_rows<<<1024, 128, 4 * row_bytes>>>(d, n_rows, row_bytes, 64, 0);
      else if (variant == 2) cudaMemsetAsync(d, 1, bytes);
      else if (variant == 3) cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice);
      else if (variant == 4) tma_rows<<<1024, 128, 4 * row_bytes>>>(d, n_rows, row_bytes, 64, 1);
      else if (variant == 5) stg_regions<<<1024, 128>>>(d, n_rows, row_bytes, 64, 0);
      else stg_regions<<<1024, 128>>>(d, n_rows, row_bytes, 64, 1);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = time_ms(e0, e1) / reps;
    const char* name[] = {"STG.128 stream", "TMA rows, region per warp", "cudaMemset", "D2D copy (read+write bytes)",
                          "TMA rows, interleaved warps", "STG rows, region per warp", "STG rows, interleaved warps"};
    double gb = (variant == 3 ? 2.0 : 1.0) * bytes / 1e9;
    printf("%-32s %8.3f us  %8.1f GB/s  (err %s)\n", name[variant], ms * 1e3, gb / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
