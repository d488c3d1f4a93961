// write_ceiling3.cu — flat STG.128 streams: does the sweep pattern or the occupancy matter?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) flat(double2* dst, size_t n2, int mode, int work) {
  double2 v = make_double2(1.0, 0.0);
  if (mode == 0) {  // global grid-stride sweep
    size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x, stride = (size_t)gridDim.x * THREADS;
    for (; i < n2; i += stride) dst[i] = v;
  } else if (mode == 1) {  // each block sweeps its own contiguous region
    size_t per = (n2 + gridDim.x - 1) / gridDim.x, b = (size_t)blockIdx.x * per, e = b + per < n2 ? b + per : n2;
    for (size_t i = b + threadIdx.x; i < e; i += THREADS) dst[i] = v;
  } else {  // like mode 0 but with `work` dependent integer ops per element (stand-in for the gather arithmetic)
    size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x, stride = (size_t)gridDim.x * THREADS;
    for (; i < n2; i += stride) {
      uint32_t h = (uint32_t)i;
      for (int k = 0; k < work; ++k) h = h * 1664525u + 1013904223u;
      v.y = (h == 0x12345678u) ? 2.0 : 0.0;
      dst[i] = v;
    }
  }
}

int main() {
  const size_t bytes = (size_t)262144 * 2224, n2 = bytes / 16;
  double2* d;
  cudaMalloc(&d, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { const char* name; int kind, blocks, mode, work; };
  Cfg cfgs[] = {{"flat 256thr x 2368 blk, grid-stride", 0, 2368, 0, 0},
                {"flat 128thr x 1036 blk (7/SM), grid-stride", 1, 1036, 0, 0},
                {"flat 128thr x 1036 blk (7/SM), block regions", 1, 1036, 1, 0},
                {"flat 256thr x 2368 blk, block regions", 0, 2368, 1, 0},
                {"flat 256thr x 1184 blk (8/SM), grid-stride", 0, 1184, 0, 0},
                {"flat 256thr x 2368 blk, grid-stride + 16 int ops", 0, 2368, 2, 16},
                {"flat 256thr x 2368 blk, grid-stride + 32 int ops", 0, 2368, 2, 32},
                {"flat 256thr x 2368 blk, grid-stride + 64 int ops", 0, 2368, 2, 64},
                {"flat 256thr x 18944 blk, grid-stride (1 elt/thread x 7.7)", 0, 18944, 0, 0}};
  for (auto& c : cfgs) {
    const int reps = 30;
    for (int it = -3; it < reps; ++it) {
      if (it == 0) cudaEventRecord(e0);
      if (c.kind == 0) flat<256, 8><<<c.blocks, 256>>>(d, n2, c.mode, c.work);
      else flat<128, 7><<<c.blocks, 128>>>(d, n2, c.mode, c.work);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-60s %8.2f us %8.1f GB/s (%s)\n", c.name, ms * 1e3, bytes / 1e9 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
