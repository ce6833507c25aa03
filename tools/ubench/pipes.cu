// Micro-benchmark: per-SM issue rates of the instructions the attention softmax is built from
// (scalar vs packed fp32 FMA/ADD, MUFU.EX2 in f32 / f16, F2FP packs, 3-input max, LEA) and of two
// candidate per-pair instruction mixes.  Cycle counts come from clock64 inside the kernel, so the
// figures are independent of the SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define F2(lo, hi) (((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo))

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float a[8], b[8];
  unsigned long long p[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
    b[i] = seed * 0.5f - i * 0.02f;
    p[i] = F2(a[i], b[i]);
    h[i] = 0x30003100u + i;
  }
  const float c1 = 0.999f, c2 = 1e-3f;
  const unsigned long long pc1 = F2(c1, c1), pc2 = F2(c2, c2);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2));
      if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc2));
      if (MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 5) { unsigned short hs = (unsigned short)h[i]; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(hs)); h[i] = hs; }
      if (MODE == 6) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(r | 0x3f000000u); }
      if (MODE == 7) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) & 7]));
      if (MODE == 8) asm volatile("{ .reg .b32 t; shl.b32 t, %1, 23; add.s32 %0, %0, t; }" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      if (MODE == 9) {   // per pair: scale (FFMA2), 2 MUFU, pack, 3-input max
        unsigned long long s = p[i];
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(s) : "l"(pc1), "l"(pc2));
        float x, y;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(s));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(x), "f"(y));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(y));
        h[i] ^= r;
      }
      if (MODE == 10 || MODE == 11) {   // per pair: scale (FFMA2), packed Cody-Waite + degree-3 Horner, 2 clamps, 2 exponent adds, pack, max
        unsigned long long s = p[i];
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(s) : "l"(pc1), "l"(pc2));
        float x, y;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(s));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(x), "f"(y));
        x = fmaxf(x, -126.f); y = fmaxf(y, -126.f);
        unsigned long long xy, t, f, pl;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xy) : "f"(x), "f"(y));
        const unsigned long long magic = F2(12582912.f, 12582912.f), nmagic = F2(-12582912.f, -12582912.f);
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xy), "l"(magic));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(t), "l"(nmagic));
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(xy), "l"(f));
        const unsigned long long k3 = F2(0.055838283f, 0.055838283f), k2 = F2(0.24263948f, 0.24263948f),
                                 k1 = F2(0.69313675f, 0.69313675f), k0 = F2(0.99992454f, 0.99992454f);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(k3), "l"(f), "l"(k2));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(pl), "l"(f), "l"(k1));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pl) : "l"(pl), "l"(f), "l"(k0));
        unsigned tl, th, pl_lo, pl_hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(tl), "=r"(th) : "l"(t));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(pl_lo), "=r"(pl_hi) : "l"(pl));
        const float e0 = __uint_as_float(pl_lo + (tl << 23)), e1 = __uint_as_float(pl_hi + (th << 23));
        unsigned r;
        if (MODE == 10) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e0), "f"(e1));
        else r = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);   // truncating bf16 pack on the ALU pipe
        h[i] ^= r;
      }
      if (MODE == 12) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(a[i]));   // immediate operands
      if (MODE == 13) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1));
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __uint_as_float(h[i]) + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm, double per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, 16, -0.5f);
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters, -0.5f);
  long long hc[148];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
  const double units = (double)warps_per_sm * 32 * 8.0 * iters * per_iter;   // per SM
  printf("%-22s warps/SM %2d: %8.2f units/clk/SM   (%.1f clk per warp-instruction-group)\n", name, warps_per_sm, units / mx,
         (double)mx / (8.0 * iters) / (warps_per_sm / 4.0));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA (3 reg)", w, 1);
    run<12>("FFMA (imm)", w, 1);
    run<1>("FFMA2 [pairs]", w, 1);
    run<2>("FADD", w, 1);
    run<3>("FADD2 [pairs]", w, 1);
    run<13>("FMUL2 [pairs]", w, 1);
    run<4>("MUFU.EX2 f32", w, 1);
    run<5>("MUFU.EX2 f16", w, 1);
    run<6>("F2FP bf16x2 [pairs]", w, 1);
    run<7>("FMNMX3", w, 1);
    run<8>("SHL+IADD", w, 1);
    run<9>("mix MUFU [elements]", w, 2);
    run<10>("mix poly [elements]", w, 2);
    run<11>("mix poly+prmt [elem]", w, 2);
  }
  return 0;
}
