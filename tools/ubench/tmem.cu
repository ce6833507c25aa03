// Micro-benchmark: tensor-memory read bandwidth per SM (tcgen05.ld 32x32b.x32, the load the attention softmax uses to
// fetch its fp32 score tile) with 1, 2 and 3 warps per lane quadrant, and tcgen05.st write bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../tclight_b200/csrc -I../../include tmem.cu -o tmem && ./tmem
#include <cstdio>
#include "common.cuh"
using namespace tcl;

template <int MODE>
__global__ void k(unsigned* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = threadIdx.x;
  uint32_t r0[32], r1[32], r2[32], r3[32];
  for (int i = 0; i < 32; ++i) r0[i] = r1[i] = r2[i] = r3[i] = acc + i;
  // initialise the columns that will be read
  for (int c = 0; c < 512; c += 32) tmem_st_32x32b_x32(base + c, r0);
  tmem_st_wait();
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t c0 = (it & 1) * 128 + (warp >> 2) * 256 % 512;
    if (MODE == 0) {
      tmem_ld_32x32b_x32(base + c0 + 0, r0);
      tmem_ld_32x32b_x32(base + c0 + 32, r1);
      tmem_ld_32x32b_x32(base + c0 + 64, r2);
      tmem_ld_32x32b_x32(base + c0 + 96, r3);
      tmem_ld_wait();
      acc ^= r0[0] ^ r1[7] ^ r2[13] ^ r3[31] ^ r0[31] ^ r1[0] ^ r2[1] ^ r3[2];
    } else {
      r0[0] = acc + it;
      tmem_st_32x32b_x32(base + c0 + 0, r0);
      tmem_st_32x32b_x32(base + c0 + 32, r0);
      tmem_st_32x32b_x32(base + c0 + 64, r0);
      tmem_st_32x32b_x32(base + c0 + 96, r0);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}

template <int MODE>
void run(const char* name, int warps) {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, warps * 32>>>(out, cyc, 64);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  long long hc[148];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
  const double bytes = (double)warps * 32 * 128 * 4 * iters;    // per SM
  printf("%-10s %2d warps/SM: %7.1f B/clk/SM   (%.0f clk per 128-column x 32-lane tile per warp)  err=%s\n", name, warps,
         bytes / mx, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
namespace tcl { void set_last_error(const char*, ...) {} void count_launch() {} }
int main() {
  for (int w : {4, 8, 12}) { run<0>("tcgen05.ld", w); run<1>("tcgen05.st", w); }
  return 0;
}
