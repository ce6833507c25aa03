// Support kernels for the VAE encode / decode path (reference call sites
// utils/VidToMe/generate_utils.py:140-172, invert.py:118-149).  The convolutions, GroupNorms and linears of
// diffusers' AutoencoderKL run on tcl_igemm / tcl_groupnorm; this file adds what the UNet path did not need:
//   tcl_softmax_rows   softmax over the key axis of the single-head 512-wide mid-block attention (its scores come
//                      from tcl_igemm as a [T, T_pad] 16-bit matrix: head dim 512 is outside tcl_attention's TMEM budget)
//   tcl_image_to_nhwc  NCHW image / latent (fp32 or 16-bit) -> channel-padded NHWC 16-bit with y = x*scale + shift
//                      (folds `2*imgs - 1`, generate_utils.py:160, and `1/0.18215 * latents`, :143)
//   tcl_nhwc_to_image  NHWC 16-bit -> NCHW with y = x*scale + shift (+ clamp) (folds `(imgs/2 + 0.5).clamp(0, 1)`, :145,
//                      and `posterior.mean * 0.18215`, :162)
// All three are HBM-bound streaming kernels.
#include "common.cuh"
#include "tclight.h"

namespace tcl {

template <bool BF16>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(void* __restrict__ xv, long long rows, int cols, int pitch) {
  using E = Elem<BF16>;
  typename E::T* x = reinterpret_cast<typename E::T*>(xv) + (long long)blockIdx.x * pitch;
  __shared__ float red[8];
  __shared__ float bc;
  const int vec_cols = cols & ~7;
  float m = -INFINITY;
  for (int c = threadIdx.x * 8; c < vec_cols; c += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = E::unpack(w[j]); m = fmaxf(m, fmaxf(f.x, f.y)); }
  }
  for (int c = vec_cols + threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, E::to_f(x[c]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { float t = red[0]; for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]); bc = t; }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int c = threadIdx.x * 8; c < vec_cols; c += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = E::unpack(w[j]); s += __expf(f.x - m) + __expf(f.y - m); }
  }
  for (int c = vec_cols + threadIdx.x; c < cols; c += blockDim.x) s += __expf(E::to_f(x[c]) - m);
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = 1.f / t; }
  __syncthreads();
  const float inv = bc;
  for (int c = threadIdx.x * 8; c < vec_cols; c += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = E::unpack(w[j]); o[j] = E::pack(__expf(f.x - m) * inv, __expf(f.y - m) * inv); }
    *reinterpret_cast<uint4*>(x + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (int c = vec_cols + threadIdx.x; c < pitch; c += blockDim.x)
    x[c] = E::from_f(c < cols ? __expf(E::to_f(x[c]) - m) * inv : 0.f);     // also zeroes the padding columns
}

template <typename T> __device__ __forceinline__ float ld_any(const T* p);
template <> __device__ __forceinline__ float ld_any<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_any<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld_any<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_any(T* p, float v);
template <> __device__ __forceinline__ void st_any<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_any<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void st_any<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// one thread per pixel: reads C planar values (coalesced per plane), writes one padded NHWC pixel (16-byte stores)
template <bool BF16, typename T>
__global__ void image_to_nhwc_kernel(const T* __restrict__ src, int B, int C, long long P, int cpad, float scale, float shift,
                                     void* __restrict__ outv) {
  using E = Elem<BF16>;
  typename E::T* out = reinterpret_cast<typename E::T*>(outv);
  const long long total = (long long)B * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / P, p = i - b * P;
    typename E::T* o = out + i * cpad;
    for (int c0 = 0; c0 < cpad; c0 += 8) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ca = c0 + 2 * j, cb = ca + 1;
        const float a = ca < C ? ld_any(src + (b * C + ca) * P + p) * scale + shift : 0.f;
        const float bb = cb < C ? ld_any(src + (b * C + cb) * P + p) * scale + shift : 0.f;
        w[j] = E::pack(a, bb);
      }
      *reinterpret_cast<uint4*>(o + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

template <bool BF16, typename T>
__global__ void nhwc_to_image_kernel(const void* __restrict__ inv, int B, int C, long long P, int pitch, float scale, float shift,
                                     int do_clamp, float lo, float hi, T* __restrict__ out) {
  using E = Elem<BF16>;
  const typename E::T* in = reinterpret_cast<const typename E::T*>(inv);
  const long long total = (long long)B * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / P, p = i - b * P;
    for (int c = 0; c < C; ++c) {
      float v = E::to_f(in[i * pitch + c]) * scale + shift;
      if (do_clamp) v = fminf(fmaxf(v, lo), hi);
      st_any(out + (b * C + c) * P + p, v);
    }
  }
}

static inline int grid_px(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_softmax_rows(int dtype, void* x, long long rows, int cols, int pitch, cudaStream_t stream) {
  TCL_CHECK_ARG(x && rows > 0 && cols > 0 && pitch >= cols && pitch % 8 == 0, "tcl_softmax_rows: rows=%lld cols=%d pitch=%d", rows, cols, pitch);
  TCL_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "tcl_softmax_rows: misaligned");
  TCL_CHECK_ARG(dtype == TCL_DTYPE_FP16 || dtype == TCL_DTYPE_BF16, "tcl_softmax_rows: dtype");
  TCL_CHECK_ARG(rows < (1ll << 31), "tcl_softmax_rows: too many rows");
  if (dtype == TCL_DTYPE_BF16) softmax_rows_kernel<true><<<(unsigned)rows, 256, 0, stream>>>(x, rows, cols, pitch);
  else softmax_rows_kernel<false><<<(unsigned)rows, 256, 0, stream>>>(x, rows, cols, pitch);
  TCL_CHECK_LAUNCH("tcl_softmax_rows");
  return TCL_OK;
}

extern "C" int tcl_image_to_nhwc(int dtype, int src_dtype, const void* src, int B, int C, int H, int W, int c_pad, float scale,
                                 float shift, void* out, cudaStream_t stream) {
  TCL_CHECK_ARG(src && out && B > 0 && C > 0 && H > 0 && W > 0 && c_pad >= C && c_pad % 8 == 0, "tcl_image_to_nhwc: args");
  TCL_CHECK_ARG(dtype == TCL_DTYPE_FP16 || dtype == TCL_DTYPE_BF16, "tcl_image_to_nhwc: dtype");
  const long long P = (long long)H * W;
  const int g = grid_px((long long)B * P);
  const bool bf = dtype == TCL_DTYPE_BF16;
#define TCL_I2N(T)                                                                                                      \
  do {                                                                                                                  \
    if (bf) image_to_nhwc_kernel<true, T><<<g, 256, 0, stream>>>((const T*)src, B, C, P, c_pad, scale, shift, out);     \
    else image_to_nhwc_kernel<false, T><<<g, 256, 0, stream>>>((const T*)src, B, C, P, c_pad, scale, shift, out);       \
  } while (0)
  if (src_dtype == TCL_LATENT_FP32) TCL_I2N(float);
  else if (src_dtype == TCL_LATENT_FP16) TCL_I2N(__half);
  else if (src_dtype == TCL_LATENT_BF16) TCL_I2N(__nv_bfloat16);
  else { set_last_error("tcl_image_to_nhwc: source dtype %d", src_dtype); return TCL_ERR_ARG; }
#undef TCL_I2N
  TCL_CHECK_LAUNCH("tcl_image_to_nhwc");
  return TCL_OK;
}

extern "C" int tcl_nhwc_to_image(int dtype, const void* in, int B, int C, int H, int W, int pitch, float scale, float shift,
                                 int do_clamp, float lo, float hi, int out_dtype, void* out, cudaStream_t stream) {
  TCL_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0 && pitch >= C, "tcl_nhwc_to_image: args");
  TCL_CHECK_ARG(dtype == TCL_DTYPE_FP16 || dtype == TCL_DTYPE_BF16, "tcl_nhwc_to_image: dtype");
  const long long P = (long long)H * W;
  const int g = grid_px((long long)B * P);
  const bool bf = dtype == TCL_DTYPE_BF16;
#define TCL_N2I(T)                                                                                                                        \
  do {                                                                                                                                    \
    if (bf) nhwc_to_image_kernel<true, T><<<g, 256, 0, stream>>>(in, B, C, P, pitch, scale, shift, do_clamp, lo, hi, (T*)out);            \
    else nhwc_to_image_kernel<false, T><<<g, 256, 0, stream>>>(in, B, C, P, pitch, scale, shift, do_clamp, lo, hi, (T*)out);              \
  } while (0)
  if (out_dtype == TCL_LATENT_FP32) TCL_N2I(float);
  else if (out_dtype == TCL_LATENT_FP16) TCL_N2I(__half);
  else if (out_dtype == TCL_LATENT_BF16) TCL_N2I(__nv_bfloat16);
  else { set_last_error("tcl_nhwc_to_image: output dtype %d", out_dtype); return TCL_ERR_ARG; }
#undef TCL_N2I
  TCL_CHECK_LAUNCH("tcl_nhwc_to_image");
  return TCL_OK;
}
