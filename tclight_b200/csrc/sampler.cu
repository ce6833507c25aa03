// Per-step elementwise tail of the multi-axis sampler:
//   * AdaIN of the yt-plane noise to the xy-plane noise statistics + variance-preserving blend
//     (generate.py:281-282, utils/general_utils.py:137-156),
//   * overlap rescale of later temporal windows (generate.py:276-278),
//   * DPM-Solver++ (SDE, 2nd-order midpoint, Karras) update (diffusers
//     DPMSolverMultistepScheduler.step as configured at utils/model_utils.py:71-78).
// Latents are [N, 4, h, w] contiguous in the latent dtype (fp16 / bf16 / fp32).  When the latent
// dtype is 16-bit every tensor op of the reference rounds to 16 bit; the kernels reproduce those
// roundings op by op (`R` below) so results track the PyTorch path bit for bit up to reduction
// order.
#include "common.cuh"
#include "tclight.h"

namespace tcl {

template <typename T> struct Lat;
template <> struct Lat<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float R(float v) { return v; }
};
template <> struct Lat<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float R(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct Lat<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float R(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  return sh[0];
}

// one block per (frame, channel) plane
template <typename T>
__global__ void adain_blend_kernel(T* __restrict__ noises_t, T* __restrict__ noises, int plane, float sa, float sb,
                                   float eps) {
  using Lt = Lat<T>;
  __shared__ float sh[32];
  T* ct = noises_t + (long long)blockIdx.x * plane;   // content
  T* st = noises + (long long)blockIdx.x * plane;     // style
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < plane; i += blockDim.x) { s1 += Lt::ld(ct + i); s2 += Lt::ld(st + i); }
  const float cm32 = block_sum(s1, sh) / plane;
  const float sm32 = block_sum(s2, sh) / plane;
  float v1 = 0.f, v2 = 0.f;
  for (int i = threadIdx.x; i < plane; i += blockDim.x) {
    const float a = Lt::ld(ct + i) - cm32, b = Lt::ld(st + i) - sm32;
    v1 += a * a; v2 += b * b;
  }
  const float cvar = block_sum(v1, sh) / (plane - 1);   // Tensor.var default: unbiased
  const float svar = block_sum(v2, sh) / (plane - 1);
  const float cmean = Lt::R(cm32), smean = Lt::R(sm32);
  const float cstd = Lt::R(sqrtf(Lt::R(Lt::R(cvar) + eps)));
  const float sstd = Lt::R(sqrtf(Lt::R(Lt::R(svar) + eps)));
  for (int i = threadIdx.x; i < plane; i += blockDim.x) {
    const float c = Lt::ld(ct + i), s = Lt::ld(st + i);
    const float nrm = Lt::R(Lt::R(c - cmean) / cstd);
    const float ad = Lt::R(Lt::R(nrm * sstd) + smean);           // AdaIN(noises_t, noises)
    const float bl = Lt::R(Lt::R(sa * ad) + Lt::R(sb * s));      // sqrt(a)*noises_t + sqrt(1-a)*noises
    Lt::st(ct + i, ad);
    Lt::st(st + i, bl);
  }
}

template <typename T>
__global__ void scale_kernel(T* __restrict__ x, long long n, float s) {
  using Lt = Lat<T>;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Lt::st(x + i, s * Lt::ld(x + i));
}

struct DpmCoef {
  float sigma_c_hat, alpha_c_hat;  // eps -> x0 of the current step
  float A, B, Cn;                  // x' = A*x + B*D0 (+ 0.5*B*D1) + Cn*z
  float inv_r0;
  int second_order;
};

template <typename T>
__global__ void dpm_step_kernel(const T* __restrict__ eps, const T* __restrict__ x, const T* __restrict__ x0_prev,
                                const float* __restrict__ z, T* __restrict__ x0_out, T* __restrict__ x_out, long long n,
                                DpmCoef c) {
  using Lt = Lat<T>;
  const float halfB = 0.5f * c.B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float e = Lt::ld(eps + i), xs = Lt::ld(x + i);
    const float x0 = Lt::R(Lt::R(xs - Lt::R(c.sigma_c_hat * e)) / c.alpha_c_hat);
    float acc = c.A * xs + Lt::R(c.B * x0);
    if (c.second_order) {
      const float d1 = Lt::R(c.inv_r0 * Lt::R(x0 - Lt::ld(x0_prev + i)));
      acc += Lt::R(halfB * d1);
    }
    acc += c.Cn * z[i];
    Lt::st(x0_out + i, x0);
    Lt::st(x_out + i, acc);
  }
}

// DDIM step in either direction (invert.py:215-244 pred_next_x): x' = mu_out*((x - sig_in*eps)/mu_in) + sig_out*eps,
// every tensor op rounded to the latent dtype like the reference's fp16 expression
template <typename T>
__global__ void ddim_next_kernel(const T* __restrict__ eps, const T* __restrict__ x, T* __restrict__ x_out, long long n,
                                 float mu_in, float sig_in, float mu_out, float sig_out) {
  using Lt = Lat<T>;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float e = Lt::ld(eps + i), xs = Lt::ld(x + i);
    const float x0 = Lt::R(Lt::R(xs - Lt::R(sig_in * e)) / mu_in);
    Lt::st(x_out + i, Lt::R(mu_out * x0) + Lt::R(sig_out * e));
  }
}

static inline int grid_n(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_adain_blend(int latent_dtype, void* noises_t, void* noises, int planes, int plane_elems,
                               float alpha, cudaStream_t stream) {
  TCL_CHECK_ARG(noises_t && noises && planes > 0 && plane_elems > 1, "tcl_adain_blend: args");
  const float sa = (float)sqrt((double)alpha), sb = (float)sqrt(1.0 - (double)alpha);
  const float eps = 1e-5f;
  if (latent_dtype == TCL_LATENT_FP32) adain_blend_kernel<float><<<planes, 256, 0, stream>>>((float*)noises_t, (float*)noises, plane_elems, sa, sb, eps);
  else if (latent_dtype == TCL_LATENT_FP16) adain_blend_kernel<__half><<<planes, 256, 0, stream>>>((__half*)noises_t, (__half*)noises, plane_elems, sa, sb, eps);
  else if (latent_dtype == TCL_LATENT_BF16) adain_blend_kernel<__nv_bfloat16><<<planes, 256, 0, stream>>>((__nv_bfloat16*)noises_t, (__nv_bfloat16*)noises, plane_elems, sa, sb, eps);
  else { set_last_error("tcl_adain_blend: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
  TCL_CHECK_LAUNCH("tcl_adain_blend");
  return TCL_OK;
}

extern "C" int tcl_scale_inplace(int latent_dtype, void* x, long long n, float s, cudaStream_t stream) {
  TCL_CHECK_ARG(x && n >= 0, "tcl_scale_inplace: args");
  if (n == 0) return TCL_OK;
  if (latent_dtype == TCL_LATENT_FP32) scale_kernel<float><<<grid_n(n, 256), 256, 0, stream>>>((float*)x, n, s);
  else if (latent_dtype == TCL_LATENT_FP16) scale_kernel<__half><<<grid_n(n, 256), 256, 0, stream>>>((__half*)x, n, s);
  else if (latent_dtype == TCL_LATENT_BF16) scale_kernel<__nv_bfloat16><<<grid_n(n, 256), 256, 0, stream>>>((__nv_bfloat16*)x, n, s);
  else { set_last_error("tcl_scale_inplace: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
  TCL_CHECK_LAUNCH("tcl_scale_inplace");
  return TCL_OK;
}

extern "C" int tcl_dpm_step(int latent_dtype, const void* eps, const void* x, const void* x0_prev, const float* z,
                            void* x0_out, void* x_out, long long n, float sigma_c_hat, float alpha_c_hat, float A,
                            float B, float Cn, float inv_r0, int second_order, cudaStream_t stream) {
  TCL_CHECK_ARG(eps && x && z && x0_out && x_out && n > 0, "tcl_dpm_step: args");
  TCL_CHECK_ARG(!second_order || x0_prev, "tcl_dpm_step: second order needs the previous x0");
  DpmCoef c{sigma_c_hat, alpha_c_hat, A, B, Cn, inv_r0, second_order};
  const int g = grid_n(n, 256);
  if (latent_dtype == TCL_LATENT_FP32) dpm_step_kernel<float><<<g, 256, 0, stream>>>((const float*)eps, (const float*)x, (const float*)x0_prev, z, (float*)x0_out, (float*)x_out, n, c);
  else if (latent_dtype == TCL_LATENT_FP16) dpm_step_kernel<__half><<<g, 256, 0, stream>>>((const __half*)eps, (const __half*)x, (const __half*)x0_prev, z, (__half*)x0_out, (__half*)x_out, n, c);
  else if (latent_dtype == TCL_LATENT_BF16) dpm_step_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)eps, (const __nv_bfloat16*)x, (const __nv_bfloat16*)x0_prev, z, (__nv_bfloat16*)x0_out, (__nv_bfloat16*)x_out, n, c);
  else { set_last_error("tcl_dpm_step: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
  TCL_CHECK_LAUNCH("tcl_dpm_step");
  return TCL_OK;
}

extern "C" int tcl_ddim_next(int latent_dtype, const void* eps, const void* x, void* x_out, long long n, float mu_in,
                             float sig_in, float mu_out, float sig_out, cudaStream_t stream) {
  TCL_CHECK_ARG(eps && x && x_out && n > 0, "tcl_ddim_next: args");
  TCL_CHECK_ARG(mu_in != 0.f, "tcl_ddim_next: mu_in == 0");
  const int g = grid_n(n, 256);
  if (latent_dtype == TCL_LATENT_FP32) ddim_next_kernel<float><<<g, 256, 0, stream>>>((const float*)eps, (const float*)x, (float*)x_out, n, mu_in, sig_in, mu_out, sig_out);
  else if (latent_dtype == TCL_LATENT_FP16) ddim_next_kernel<__half><<<g, 256, 0, stream>>>((const __half*)eps, (const __half*)x, (__half*)x_out, n, mu_in, sig_in, mu_out, sig_out);
  else if (latent_dtype == TCL_LATENT_BF16) ddim_next_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)eps, (const __nv_bfloat16*)x, (__nv_bfloat16*)x_out, n, mu_in, sig_in, mu_out, sig_out);
  else { set_last_error("tcl_ddim_next: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
  TCL_CHECK_LAUNCH("tcl_ddim_next");
  return TCL_OK;
}
