// l2_atomics.cu -- the rate of red.global.min.u64 (k_project's z-buffer update) against plain 64-bit stores and
// red.global.min.u32, with k_project's access pattern: consecutive threads hit consecutive words of a buffer far larger
// than L2 (256 z-buffers of 640x480 words = 629 MB), every word once.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_atomics l2_atomics.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned long long *z, unsigned int *z32, size_t n, unsigned long long key) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // a small permutation inside 32-word groups, like neighbouring points landing on neighbouring pixels out of order
    const size_t j = (i & ~(size_t)31) | ((i * 5 + 3) & 31);
    if (MODE == 0) asm volatile("red.global.min.u64 [%0], %1;" ::"l"(z + j), "l"(key + i) : "memory");
    if (MODE == 1) z[j] = key + i;
    if (MODE == 2) asm volatile("red.global.min.u32 [%0], %1;" ::"l"(z32 + j), "r"((unsigned int)(key + i)) : "memory");
    if (MODE == 3) asm volatile("red.global.min.u64 [%0], %1;" ::"l"(z + (j & 0xFFFFF)), "l"(key + i) : "memory");  // 8 MB: L2 resident
  }
}

int main() {
  const size_t n = (size_t)256 * 640 * 480;
  unsigned long long *z;
  cudaMalloc(&z, n * 8);
  cudaMemset(z, 0xFF, n * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char *names[4] = {"red.min.u64, 629 MB", "st.u64, 629 MB", "red.min.u32, 315 MB", "red.min.u64, 8 MB window"};
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      const int blocks = 148 * 16;
      if (mode == 0) k<0><<<blocks, 256>>>(z, (unsigned int *)z, n, 1000 - rep);
      if (mode == 1) k<1><<<blocks, 256>>>(z, (unsigned int *)z, n, 1000 - rep);
      if (mode == 2) k<2><<<blocks, 256>>>(z, (unsigned int *)z, n, 1000 - rep);
      if (mode == 3) k<3><<<blocks, 256>>>(z, (unsigned int *)z, n, 1000 - rep);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("%-28s %8.1f us  %6.1f G ops/s\n", names[mode], ms * 1e3, n / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
