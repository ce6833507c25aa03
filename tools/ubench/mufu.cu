// Micro-benchmark: MUFU.EX2 throughput per SM for f32, f16x2 and bf16x2 operands, and a degree-3
// polynomial exp2 on the FMA pipe.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 mufu.cu -o mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; h[i] = 0x30003100u + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE >= 4) a[i] = __uint_as_float(__float_as_uint(a[i]) ^ (h[i] & 1u));   // keep the loop live (1 LOP3)
      if (MODE == 4) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); h[i] ^= r; }
      if (MODE == 5) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); h[i] ^= r; }
      if (MODE == 6) {  // MUFU + bf16 pack together (2 ex2 + 1 cvt)
        float e0, e1; unsigned r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a[(i + 1) & 7]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e0), "f"(e1)); h[i] ^= r;
      }
      if (MODE == 7) {  // MUFU + integer round-half-up bf16 pack
        float e0, e1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a[(i + 1) & 7]));
        unsigned u0 = __float_as_uint(e0) + 0x8000u, u1 = __float_as_uint(e1) + 0x8000u;
        h[i] ^= __byte_perm(u0, u1, 0x7632);
      }
      if (MODE == 3) {  // poly: round, frac, horner3, exponent add
        float x = a[i];
        float t = x + 12582912.f;
        float r = t - 12582912.f;
        float f = x - r;
        float p = fmaf(f, 0.0555f, 0.2402f);
        p = fmaf(p, f, 0.6931f);
        p = fmaf(p, f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) * 1e-3f;
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 4096;
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  k<MODE><<<148, warps_per_sm * 32>>>(out, 16, -0.5f);
  cudaEventRecord(s);
  k<MODE><<<148, warps_per_sm * 32>>>(out, iters, -0.5f);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double ops = 148.0 * warps_per_sm * 32 * 8.0 * iters;   // thread-level instructions
  double clk = ms * 1e-3 * 1.965e9;
  printf("%-10s warps/SM %2d: %.2f thread-ops/clk/SM (assuming 1.965 GHz), %.3f ms\n", name, warps_per_sm, ops / 148 / clk, ms);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) { run<0>("ex2.f32", w); run<1>("ex2.f16x2", w); run<2>("ex2.bf16x2", w); run<3>("poly3", w); run<4>("cvt.bf16x2", w); run<5>("cvt.f16x2", w); run<6>("2ex2+cvt", w); run<7>("2ex2+ipack", w); }
  return 0;
}
