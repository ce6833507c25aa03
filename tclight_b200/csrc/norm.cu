// HBM-bound helper kernels of the UNet forward: GroupNorm(+SiLU) over
// NHWC with on-the-fly channel concat of the skip connection, LayerNorm, nearest upsampling,
// latent <-> NHWC staging (IC-Light concat of the condition latent, utils/model_utils.py:35-40)
// and the CFG combine of generate.py:349-350.
#include "common.cuh"
#include "tclight.h"

namespace tcl {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// ------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[n][g] = {sum, sumsq} over (pixels, C/G channels).
// Thread -> fixed block of 8 channels; per-channel partials -> smem -> per-group -> global atomics.
// ------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void gn_stats_kernel(const void* __restrict__ x1, int c1, const void* __restrict__ x2, int c2,
                                long long pix_per_img, int groups, float* __restrict__ stats,
                                int pix_per_block) {
  using E = Elem<BF16>;
  extern __shared__ float sh[];  // [2*C]
  const int C = c1 + c2;
  const int vecs = C / 8;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int cb = threadIdx.x % vecs;      // channel block
  const int sub = threadIdx.x / vecs;     // pixel lane
  const int lanes = blockDim.x / vecs;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > pix_per_img) p1 = pix_per_img;
  const int ch = cb * 8;
  const bool first = ch < c1;
  const typename E::T* src = reinterpret_cast<const typename E::T*>(first ? x1 : x2);
  const int cw = first ? c1 : c2;
  const int co = first ? ch : ch - c1;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  if (sub < lanes) {
    auto acc = [&](const uint4& v) {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = E::unpack(w[j]);
        s[2 * j] += f.x; ss[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; ss[2 * j + 1] += f.y * f.y;
      }
    };
    const typename E::T* base = src + (long long)n * pix_per_img * cw + co;
    long long pix = p0 + sub;
    // four independent 16-byte loads in flight per thread (the single-load loop was latency bound at ~1.5 TB/s)
    for (; pix + 3LL * lanes < p1; pix += 4LL * lanes) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(base + pix * cw);
      const uint4 v1 = *reinterpret_cast<const uint4*>(base + (pix + lanes) * cw);
      const uint4 v2 = *reinterpret_cast<const uint4*>(base + (pix + 2LL * lanes) * cw);
      const uint4 v3 = *reinterpret_cast<const uint4*>(base + (pix + 3LL * lanes) * cw);
      acc(v0); acc(v1); acc(v2); acc(v3);
    }
    for (; pix < p1; pix += lanes) acc(*reinterpret_cast<const uint4*>(base + pix * cw));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[ch + j], s[j]);
      atomicAdd(&sh[C + ch + j], ss[j]);
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  if (threadIdx.x < groups) {
    float a = 0.f, b = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { a += sh[c]; b += sh[C + c]; }
    atomicAdd(&stats[((long long)n * groups + threadIdx.x) * 2 + 0], a);
    atomicAdd(&stats[((long long)n * groups + threadIdx.x) * 2 + 1], b);
  }
}

// y = x * a[c] + b[c] with a = rstd*gamma, b = beta - mean*rstd*gamma built once per block in shared memory
// (blockIdx.y = image), so the streaming loop is one 16-byte load, 8 FMAs (+SiLU) and one 16-byte store per vector.
template <bool BF16>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const void* __restrict__ x1, int c1, const void* __restrict__ x2, int c2,
                long long pix_per_img, int n_img, int groups, const float* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                int silu, void* __restrict__ out) {
  using E = Elem<BF16>;
  extern __shared__ float ab[];      // [2][C]
  const int C = c1 + c2;
  const int vecs = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y;
  const float inv_cnt = 1.0f / (float)(pix_per_img * cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sm = stats[((long long)n * groups + g) * 2 + 0];
    const float sq = stats[((long long)n * groups + g) * 2 + 1];
    const float mean = sm * inv_cnt;
    const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
    const float a = rsqrtf(var + eps) * gamma[c];
    ab[c] = a;
    ab[C + c] = beta[c] - mean * a;
  }
  __syncthreads();
  const long long total = pix_per_img * vecs;
  const typename E::T* s1 = reinterpret_cast<const typename E::T*>(x1) + (long long)n * pix_per_img * c1;
  const typename E::T* s2 = reinterpret_cast<const typename E::T*>(x2) + (long long)n * pix_per_img * c2;
  typename E::T* dst = reinterpret_cast<typename E::T*>(out) + (long long)n * pix_per_img * C;
  auto one = [&](long long i) {
    const int cb = (int)(i % vecs);
    const long long pix = i / vecs;
    const int ch = cb * 8;
    const uint4 v = ch < c1 ? *reinterpret_cast<const uint4*>(s1 + pix * c1 + ch)
                            : *reinterpret_cast<const uint4*>(s2 + pix * c2 + (ch - c1));
    const float4 a0 = *reinterpret_cast<const float4*>(ab + ch), a1 = *reinterpret_cast<const float4*>(ab + ch + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(ab + C + ch), b1 = *reinterpret_cast<const float4*>(ab + C + ch + 4);
    const float2 f0 = E::unpack(v.x), f1 = E::unpack(v.y), f2 = E::unpack(v.z), f3 = E::unpack(v.w);
    float y[8] = {fmaf(f0.x, a0.x, b0.x), fmaf(f0.y, a0.y, b0.y), fmaf(f1.x, a0.z, b0.z), fmaf(f1.y, a0.w, b0.w),
                  fmaf(f2.x, a1.x, b1.x), fmaf(f2.y, a1.y, b1.y), fmaf(f3.x, a1.z, b1.z), fmaf(f3.y, a1.w, b1.w)};
    if (silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = silu_f(y[j]);
    }
    uint4 o;
    o.x = E::pack(y[0], y[1]); o.y = E::pack(y[2], y[3]); o.z = E::pack(y[4], y[5]); o.w = E::pack(y[6], y[7]);
    *reinterpret_cast<uint4*>(dst + pix * C + ch) = o;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < total; i += 2 * stride) { one(i); one(i + stride); }
  if (i < total) one(i);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 2048, C % 8 == 0): one warp per row, two-pass in registers.
// ------------------------------------------------------------------------------------------
// LPR lanes share one row (8 / 16 / 32: the smallest that keeps <= VPL vectors of 8 per lane), so a warp normalises
// 32/LPR rows at once with every lane busy; all of a lane's 16-byte loads are issued before the first use.
template <bool BF16, int LPR, int VPL>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, void* __restrict__ out) {
  using E = Elem<BF16>;
  constexpr int RPW = 32 / LPR;                     // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool live = row < rows;
  const int vecs = C / 8;
  const typename E::T* src = reinterpret_cast<const typename E::T*>(x) + (live ? row : 0) * C;
  uint4 raw[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = l + k * LPR;
    raw[k] = (live && v < vecs) ? *reinterpret_cast<const uint4*>(src + v * 8) : make_uint4(0, 0, 0, 0);
  }
  float f[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = E::unpack(w[j]);
      f[k][2 * j] = t.x; f[k][2 * j + 1] = t.y;
      sum += t.x + t.y;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (l + k * LPR < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dlt = f[k][j] - mean; sq += dlt * dlt; }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  if (!live) return;
  typename E::T* dst = reinterpret_cast<typename E::T*>(out) + row * C;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = l + k * LPR;
    if (v < vecs) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (f[k][j] - mean) * rstd * gg[j] + bb[j];
      uint4 o;
      o.x = E::pack(y[0], y[1]); o.y = E::pack(y[2], y[3]); o.z = E::pack(y[4], y[5]); o.w = E::pack(y[6], y[7]);
      *reinterpret_cast<uint4*>(dst + v * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// nearest-neighbour upsampling NHWC (F.interpolate(mode="nearest") index rule: floor(dst*in/out))
// ------------------------------------------------------------------------------------------
__global__ void upsample_nearest_kernel(const uint4* __restrict__ x, int n, int h, int w, int vecs, int oh, int ow,
                                        float sh, float sw, uint4* __restrict__ out) {
  const long long total = (long long)n * oh * ow * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long r = i / vecs;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const int img = (int)(r / oh);
    int sy = (int)floorf(oy * sh); if (sy > h - 1) sy = h - 1;
    int sx = (int)floorf(ox * sw); if (sx > w - 1) sx = w - 1;
    out[i] = x[(((long long)img * h + sy) * w + sx) * vecs + v];
  }
}

// ------------------------------------------------------------------------------------------
// latent staging.  `x` and `cond` are strided views (element strides) of 4-channel latents:
// image i, channel c, row y, col x -> base[i*s_img + c*s_c + y*s_y + x*s_x].
// Output: NHWC [2*F, H, W, 64] 16-bit; channels 0..3 latent, 4..7 condition, 8..63 zero; the
// second CFG half is a copy of the first (generate.py:298, model_utils.py:37-38).
// ------------------------------------------------------------------------------------------
template <bool BF16, typename TIn>
__global__ void stage_latent_kernel(const TIn* __restrict__ x, long long xs_img, long long xs_c, long long xs_y, long long xs_x,
                                    const TIn* __restrict__ cond, long long cs_img, long long cs_c, long long cs_y, long long cs_x,
                                    int F, int H, int W, int halves, void* __restrict__ out) {
  using E = Elem<BF16>;
  const long long total = (long long)F * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % H);
    const int f = (int)(i / ((long long)W * H));
    float v[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = (float)x[f * xs_img + c * xs_c + yy * xs_y + xx * xs_x];
      v[4 + c] = cond ? (float)cond[f * cs_img + c * cs_c + yy * cs_y + xx * cs_x] : 0.f;
    }
    uint4 o;
    o.x = E::pack(v[0], v[1]); o.y = E::pack(v[2], v[3]); o.z = E::pack(v[4], v[5]); o.w = E::pack(v[6], v[7]);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int half = 0; half < halves; ++half) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<typename E::T*>(out) + ((long long)half * total + i) * 64);
      dst[0] = o;
#pragma unroll
      for (int k = 1; k < 8; ++k) dst[k] = z;
    }
  }
}

// CFG combine (generate.py:349-350): noise = uncond + g*(cond - uncond), with the reference's
// per-op rounding when the latent dtype is 16-bit.  eps: NHWC [2F, H, W, pitch]; out: strided
// 4-channel latent view.
template <bool BF16, typename TOut>
__global__ void cfg_store_kernel(const void* __restrict__ eps, int pitch, float gscale, int F, int H, int W,
                                 TOut* __restrict__ out, long long os_img, long long os_c, long long os_y, long long os_x) {
  using E = Elem<BF16>;
  const long long total = (long long)F * H * W;
  const typename E::T* e = reinterpret_cast<const typename E::T*>(eps);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % H);
    const int f = (int)(i / ((long long)W * H));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float u = E::to_f(e[i * pitch + c]);
      const float cnd = E::to_f(e[(total + i) * pitch + c]);
      float r;
      if (sizeof(TOut) == 2) {
        const float d1 = E::to_f(E::from_f(cnd - u));
        const float d2 = E::to_f(E::from_f(gscale * d1));
        r = u + d2;
      } else {
        r = u + gscale * (cnd - u);
      }
      out[f * os_img + c * os_c + yy * os_y + xx * os_x] = (TOut)r;
    }
  }
}

// y = W x + b for a single vector (time embedding MLP and the per-resnet time_emb_proj,
// SURVEY.md B.1).  One warp per output row; optional SiLU on the input and 16-bit rounding of
// the output (the reference runs these nn.Linear layers in the UNet's 16-bit dtype).
template <bool BF16>
__global__ void gemv_kernel(const void* __restrict__ W, const float* __restrict__ x, const float* __restrict__ bias,
                            int N, int K, int silu_in, int round16, float* __restrict__ y) {
  using E = Elem<BF16>;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const typename E::T* w = reinterpret_cast<const typename E::T*>(W) + (long long)row * K;
  float acc = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(w + k);
    const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = E::unpack(ww[j]);
      float x0 = x[k + 2 * j], x1 = x[k + 2 * j + 1];
      if (silu_in) { x0 = silu_f(x0); x1 = silu_f(x1); if (round16) { x0 = E::to_f(E::from_f(x0)); x1 = E::to_f(E::from_f(x1)); } }
      acc += f.x * x0 + f.y * x1;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float r = acc + (bias ? bias[row] : 0.f);
    if (round16) r = E::to_f(E::from_f(r));
    y[row] = r;
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_groupnorm(int dtype, const void* x1, int c1, const void* x2, int c2, int n_img,
                             long long pix_per_img, int groups, const float* gamma, const float* beta,
                             float eps, int silu, float* stats_ws, void* out, cudaStream_t stream) {
  TCL_CHECK_ARG(x1 && out && gamma && beta && stats_ws, "tcl_groupnorm: null pointer");
  TCL_CHECK_ARG(c1 > 0 && c1 % 8 == 0 && c2 >= 0 && c2 % 8 == 0 && (c2 == 0 || x2), "tcl_groupnorm: channels");
  const int C = c1 + c2;
  TCL_CHECK_ARG(groups > 0 && groups <= 32 && C % groups == 0, "tcl_groupnorm: groups");
  TCL_CHECK_ARG(n_img > 0 && pix_per_img > 0, "tcl_groupnorm: empty");
  const int vecs = C / 8;
  TCL_CHECK_ARG(vecs <= 512, "tcl_groupnorm: C too large");
  cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(float) * 2 * groups * n_img, stream);
  if (e != cudaSuccess) { set_last_error("tcl_groupnorm: memset: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
  int k = 512 / vecs; if (k < 1) k = 1;
  const int block = vecs * k;
  // aim for ~4 blocks per SM per image batch
  long long blocks_x = (148LL * 4 + n_img - 1) / n_img;
  long long ppb = (pix_per_img + blocks_x - 1) / blocks_x;
  if (ppb < k) ppb = k;
  blocks_x = (pix_per_img + ppb - 1) / ppb;
  dim3 grid((unsigned)blocks_x, (unsigned)n_img);
  const size_t sh = sizeof(float) * 2 * C;
  const bool bf = dtype == TCL_DTYPE_BF16;
  if (bf) gn_stats_kernel<true><<<grid, block, sh, stream>>>(x1, c1, x2, c2, pix_per_img, groups, stats_ws, (int)ppb);
  else gn_stats_kernel<false><<<grid, block, sh, stream>>>(x1, c1, x2, c2, pix_per_img, groups, stats_ws, (int)ppb);
  TCL_CHECK_LAUNCH("tcl_groupnorm(stats)");
  const long long total = pix_per_img * vecs;                 // vectors per image
  long long ax = (total + 511) / 512;                         // two vectors per thread per trip
  const long long cap = (148LL * 16 + n_img - 1) / n_img;
  if (ax > cap) ax = cap;
  if (ax < 1) ax = 1;
  const dim3 agrid((unsigned)ax, (unsigned)n_img);
  if (bf) gn_apply_kernel<true><<<agrid, 256, sh, stream>>>(x1, c1, x2, c2, pix_per_img, n_img, groups, stats_ws, gamma, beta, eps, silu, out);
  else gn_apply_kernel<false><<<agrid, 256, sh, stream>>>(x1, c1, x2, c2, pix_per_img, n_img, groups, stats_ws, gamma, beta, eps, silu, out);
  TCL_CHECK_LAUNCH("tcl_groupnorm(apply)");
  return TCL_OK;
}

extern "C" int tcl_layernorm(int dtype, const void* x, long long rows, int C, const float* gamma,
                             const float* beta, float eps, void* out, cudaStream_t stream) {
  TCL_CHECK_ARG(x && out && gamma && beta, "tcl_layernorm: null pointer");
  TCL_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 2048, "tcl_layernorm: C=%d", C);
  if (rows <= 0) return TCL_OK;
  TCL_CHECK_ARG(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "tcl_layernorm: gamma/beta must be 16-byte aligned");
  const int wpb = 8;
  const int vecs = C / 8;
  const bool bf = dtype == TCL_DTYPE_BF16;
#define TCL_LN(LPR, VPL)                                                                                               \
  do {                                                                                                                 \
    const long long rpb = (long long)wpb * (32 / LPR);                                                                 \
    const long long blocks = (rows + rpb - 1) / rpb;                                                                   \
    if (bf) layernorm_kernel<true, LPR, VPL><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x, rows, C, gamma, beta, eps, out);  \
    else layernorm_kernel<false, LPR, VPL><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x, rows, C, gamma, beta, eps, out);    \
  } while (0)
  if (vecs <= 8) TCL_LN(8, 1);            // C <= 64
  else if (vecs <= 16) TCL_LN(8, 2);      // C <= 128
  else if (vecs <= 40) TCL_LN(8, 5);      // C <= 320
  else if (vecs <= 80) TCL_LN(16, 5);     // C <= 640
  else if (vecs <= 160) TCL_LN(32, 5);    // C <= 1280
  else TCL_LN(32, 8);                     // C <= 2048
#undef TCL_LN
  TCL_CHECK_LAUNCH("tcl_layernorm");
  return TCL_OK;
}

extern "C" int tcl_upsample_nearest(const void* x, int n, int h, int w, int c, int oh, int ow, void* out,
                                    cudaStream_t stream) {
  TCL_CHECK_ARG(x && out && c % 8 == 0 && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "tcl_upsample_nearest: args");
  const int vecs = c / 8;
  const long long total = (long long)n * oh * ow * vecs;
  // PyTorch: scale = in/out when an explicit size is given; exactly 0.5 for scale_factor=2
  const float sh = (float)h / (float)oh, sw = (float)w / (float)ow;
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), n, h, w, vecs, oh, ow, sh, sw,
                                                                     reinterpret_cast<uint4*>(out));
  TCL_CHECK_LAUNCH("tcl_upsample_nearest");
  return TCL_OK;
}

extern "C" int tcl_stage_latent(int dtype, int latent_dtype, const void* x, const long long* xs, const void* cond,
                                const long long* cs, int F, int H, int W, int duplicate, void* out, cudaStream_t stream) {
  TCL_CHECK_ARG(x && xs && out && F > 0 && H > 0 && W > 0, "tcl_stage_latent: args");
  TCL_CHECK_ARG(cond == nullptr || cs != nullptr, "tcl_stage_latent: cond strides");
  const long long total = (long long)F * H * W;
  const long long z[4] = {0, 0, 0, 0};
  if (!cs) cs = z;
  const bool bf = dtype == TCL_DTYPE_BF16;
  const int g = grid_for(total, 256);
#define TCL_STAGE(BF, T) stage_latent_kernel<BF, T><<<g, 256, 0, stream>>>((const T*)x, xs[0], xs[1], xs[2], xs[3], (const T*)cond, cs[0], cs[1], cs[2], cs[3], F, H, W, duplicate ? 2 : 1, out)
  if (latent_dtype == TCL_LATENT_FP32) { if (bf) TCL_STAGE(true, float); else TCL_STAGE(false, float); }
  else if (latent_dtype == TCL_LATENT_FP16) { if (bf) TCL_STAGE(true, __half); else TCL_STAGE(false, __half); }
  else if (latent_dtype == TCL_LATENT_BF16) { if (bf) TCL_STAGE(true, __nv_bfloat16); else TCL_STAGE(false, __nv_bfloat16); }
  else { set_last_error("tcl_stage_latent: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
#undef TCL_STAGE
  TCL_CHECK_LAUNCH("tcl_stage_latent");
  return TCL_OK;
}

extern "C" int tcl_cfg_store(int dtype, int latent_dtype, const void* eps, int pitch, float guidance_scale, int F, int H,
                             int W, void* out, const long long* os, cudaStream_t stream) {
  TCL_CHECK_ARG(eps && out && os && F > 0 && H > 0 && W > 0 && pitch >= 4, "tcl_cfg_store: args");
  const long long total = (long long)F * H * W;
  const bool bf = dtype == TCL_DTYPE_BF16;
  const int g = grid_for(total, 256);
#define TCL_CFG(BF, T) cfg_store_kernel<BF, T><<<g, 256, 0, stream>>>(eps, pitch, guidance_scale, F, H, W, (T*)out, os[0], os[1], os[2], os[3])
  if (latent_dtype == TCL_LATENT_FP32) { if (bf) TCL_CFG(true, float); else TCL_CFG(false, float); }
  else if (latent_dtype == TCL_LATENT_FP16) { if (bf) TCL_CFG(true, __half); else TCL_CFG(false, __half); }
  else if (latent_dtype == TCL_LATENT_BF16) { if (bf) TCL_CFG(true, __nv_bfloat16); else TCL_CFG(false, __nv_bfloat16); }
  else { set_last_error("tcl_cfg_store: latent dtype %d", latent_dtype); return TCL_ERR_ARG; }
#undef TCL_CFG
  TCL_CHECK_LAUNCH("tcl_cfg_store");
  return TCL_OK;
}

extern "C" int tcl_gemv(int dtype, const void* W, const float* x, const float* bias, int N, int K, int silu_in,
                        int round16, float* y, cudaStream_t stream) {
  TCL_CHECK_ARG(W && x && y && N > 0 && K > 0 && K % 8 == 0, "tcl_gemv: args");
  const int wpb = 8;
  if (dtype == TCL_DTYPE_BF16) gemv_kernel<true><<<(N + wpb - 1) / wpb, wpb * 32, 0, stream>>>(W, x, bias, N, K, silu_in, round16, y);
  else gemv_kernel<false><<<(N + wpb - 1) / wpb, wpb * 32, 0, stream>>>(W, x, bias, N, K, silu_in, round16, y);
  TCL_CHECK_LAUNCH("tcl_gemv");
  return TCL_OK;
}
