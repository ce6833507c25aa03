// Two-stage temporal-consistency optimiser, HBM-bound fp32 kernels.
//   stage 1  exposure_align            generate.py:354-451   per-frame 3x4 affine, Adam
//   stage 2  unique_tensor_optimization generate.py:453-533  Unique-Video-Tensor [U,3] (SH-DC), Adam
// One C call = one optimiser iteration (forward, backward, Adam) with no host synchronisation:
//   produce X (2*Bo frames: batch frames then their predecessors) [stage 2: gather + SH2RGB + clamp;
//   stage 1: affine + clamp]  ->  avg-pool pyramid  ->  relaxed MS-SSIM forward (level sums)  ->
//   loss/coefficients  ->  MS-SSIM backward down the pyramid  ->  fused level-0 kernel (bicubic
//   flow warp fwd+bwd, masked L1, TV, [L1], SSIM grad upsample, gradient sink)  ->  predecessor
//   gradient sink  ->  Adam.
// Reference arithmetic: warp_flow utils/flow_utils.py:5-16 (grid_sample bicubic A=-0.75, zeros,
// align_corners=True), l1_loss / relaxed_ms_ssim / TVLoss utils/loss_utils.py:25, 73-211, 324-339
// (pytorch_msssim 11-tap sigma-1.5 valid separable Gaussian), RGB2SH/SH2RGB utils/sh_utils.py:114-118,
// torch.optim.Adam.
#include "common.cuh"
#include "tclight.h"

namespace tcl {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr int MAXB = TCL_POSTOPT_MAX_BATCH;

__constant__ float c_gauss[11];
static bool g_gauss_ready = false;

struct Batch {
  int n;           // frames in the batch (<= MAXB)
  int idx[MAXB];   // frame index of each batch item
};

struct Pyr {
  int h[5], w[5];          // level sizes (level 0 = full resolution)
  int ph[5], pw[5];        // avg-pool padding used to go from level l to l+1
  long long off[5];        // element offset of level l planes inside a pyramid buffer (levels 1..4)
  long long total;         // elements per (frame, channel) over levels 1..4
};

static void make_pyr(int H, int W, Pyr* p) {
  p->h[0] = H; p->w[0] = W;
  long long off = 0;
  for (int l = 0; l < 4; ++l) {
    p->ph[l] = p->h[l] % 2; p->pw[l] = p->w[l] % 2;
    p->h[l + 1] = (p->h[l] + 2 * p->ph[l] - 2) / 2 + 1;
    p->w[l + 1] = (p->w[l] + 2 * p->pw[l] - 2) / 2 + 1;
  }
  p->off[0] = 0;
  for (int l = 1; l <= 4; ++l) { p->off[l] = off; off += (long long)p->h[l] * p->w[l]; }
  p->total = off;
  p->ph[4] = p->pw[4] = 0;
}

// ------------------------------------------------------------------------------------------
// avg_pool2d(kernel 2, stride 2, padding = size % 2, count_include_pad) over `planes` planes.
// in/out plane strides are given explicitly so the same kernel builds any pyramid level.
// ------------------------------------------------------------------------------------------
__global__ void avgpool2_kernel(const float* __restrict__ in, long long in_stride, int hi, int wi, int ph, int pw,
                                float* __restrict__ out, long long out_stride, int ho, int wo, int planes) {
  const long long total = (long long)planes * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const int oy = (int)((i / wo) % ho);
    const int pl = (int)(i / ((long long)wo * ho));
    const float* src = in + pl * in_stride;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * oy - ph + dy, x = 2 * ox - pw + dx;
        if (y >= 0 && y < hi && x >= 0 && x < wi) s += src[(long long)y * wi + x];
      }
    out[pl * out_stride + (long long)oy * wo + ox] = s * 0.25f;
  }
}

// ------------------------------------------------------------------------------------------
// SSIM forward at one pyramid level: per plane sums of cs_map and ssim_map over the valid region.
// X planes: [Bo*3] batch-local; Y planes: frame-indexed pyramid of the target video.
// ------------------------------------------------------------------------------------------
constexpr int ST = 32;          // output tile
constexpr int SR = ST + 10;     // input tile

struct SsimMaps { float mu1, mu2, e11, e22, e12; };

__device__ __forceinline__ void ssim_point(const SsimMaps& m, float C1, float C2, float& lmap, float& cs) {
  const float mu1_sq = m.mu1 * m.mu1, mu2_sq = m.mu2 * m.mu2, mu12 = m.mu1 * m.mu2;
  const float s11 = m.e11 - mu1_sq, s22 = m.e22 - mu2_sq, s12 = m.e12 - mu12;
  cs = (2.f * s12 + C2) / (s11 + s22 + C2);
  lmap = (2.f * mu12 + C1) / (mu1_sq + mu2_sq + C1);
}

__global__ void __launch_bounds__(256)
ssim_fwd_kernel(const float* __restrict__ X, long long x_plane_stride, const float* __restrict__ Y, long long y_frame_stride,
                long long y_chan_stride, Batch bt, int h, int w, float C1, float C2, float* __restrict__ sums /*[planes][2]*/) {
  __shared__ float sx[SR][SR + 1], sy[SR][SR + 1];
  __shared__ float v[5][ST][SR + 1];
  __shared__ float red[2][8];
  const int plane = blockIdx.z;           // b*3 + c
  const int b = plane / 3, c = plane - 3 * b;
  const float* xp = X + plane * x_plane_stride;
  const float* yp = Y + bt.idx[b] * y_frame_stride + c * y_chan_stride;
  const int oh = h - 10, ow = w - 10;     // valid map size
  const int ty0 = blockIdx.y * ST, tx0 = blockIdx.x * ST;
  for (int i = threadIdx.x; i < SR * SR; i += blockDim.x) {
    const int r = i / SR, cc = i - r * SR;
    const int y = ty0 + r, x = tx0 + cc;
    const bool in = y < h && x < w;
    sx[r][cc] = in ? xp[(long long)y * w + x] : 0.f;
    sy[r][cc] = in ? yp[(long long)y * w + x] : 0.f;
  }
  __syncthreads();
  // vertical (H) pass first, as pytorch_msssim.gaussian_filter does
  for (int i = threadIdx.x; i < ST * SR; i += blockDim.x) {
    const int r = i / SR, cc = i - r * SR;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_gauss[k], xv = sx[r + k][cc], yv = sy[r + k][cc];
      a0 += g * xv; a1 += g * yv; a2 += g * xv * xv; a3 += g * yv * yv; a4 += g * xv * yv;
    }
    v[0][r][cc] = a0; v[1][r][cc] = a1; v[2][r][cc] = a2; v[3][r][cc] = a3; v[4][r][cc] = a4;
  }
  __syncthreads();
  float acc_cs = 0.f, acc_ss = 0.f;
  for (int i = threadIdx.x; i < ST * ST; i += blockDim.x) {
    const int r = i / ST, cc = i - r * ST;
    if (ty0 + r < oh && tx0 + cc < ow) {
      SsimMaps m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = c_gauss[k];
        m.mu1 += g * v[0][r][cc + k]; m.mu2 += g * v[1][r][cc + k]; m.e11 += g * v[2][r][cc + k];
        m.e22 += g * v[3][r][cc + k]; m.e12 += g * v[4][r][cc + k];
      }
      float l, cs;
      ssim_point(m, C1, C2, l, cs);
      acc_cs += cs; acc_ss += l * cs;
    }
  }
  acc_cs = warp_sum(acc_cs); acc_ss = warp_sum(acc_ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc_cs; red[1][threadIdx.x >> 5] = acc_ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; bsum += red[1][i]; }
    atomicAdd(&sums[plane * 2 + 0], a);
    atomicAdd(&sums[plane * 2 + 1], bsum);
  }
}

// ------------------------------------------------------------------------------------------
// loss head: MS-SSIM value per plane and d(loss)/d(level mean) coefficients.
//   sums  [4 levels][planes][2]; coef [4][planes]; scal: accumulators / outputs (see enum)
// ------------------------------------------------------------------------------------------
enum { SC_FLOW_ABS = 0, SC_TV_H = 1, SC_TV_W = 2, SC_L1 = 3, SC_MSSSIM = 4, SC_COUNT = 8 };

__global__ void msssim_head_kernel(const float* __restrict__ sums, int planes, Pyr py, float k_ms /* dLoss/d ms[plane] */,
                                   float* __restrict__ coef, float* __restrict__ scal) {
  const float wts[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
  float total = 0.f;
  for (int p = threadIdx.x; p < planes; p += blockDim.x) {
    float vals[5];
    vals[0] = 1.f;
    for (int l = 1; l <= 4; ++l) {
      const float nvalid = (float)(py.h[l] - 10) * (float)(py.w[l] - 10);
      const float s = sums[((l - 1) * planes + p) * 2 + (l == 4 ? 1 : 0)];
      vals[l] = fmaxf(s / nvalid, 0.f);
    }
    float ms = 1.f;
    for (int l = 1; l <= 4; ++l) ms *= powf(vals[l], wts[l]);
    total += ms;
    for (int l = 1; l <= 4; ++l) {
      const float nvalid = (float)(py.h[l] - 10) * (float)(py.w[l] - 10);
      coef[(l - 1) * planes + p] = vals[l] > 0.f ? k_ms * wts[l] * ms / vals[l] / nvalid : 0.f;
    }
  }
  __shared__ float sh[32];
  total = warp_sum(total);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    scal[SC_MSSSIM] = t / planes;
  }
}

// ------------------------------------------------------------------------------------------
// SSIM backward at one level: dX_l = (SSIM term) + avg-pool backward of dX_{l+1}.
// Output tile BT x BT; recomputes the five filtered maps on the (BT+10)^2 halo.
// ------------------------------------------------------------------------------------------
constexpr int BT = 32;
constexpr int BP = BT + 10;   // partial-map tile
constexpr int BR = BT + 20;   // input tile

__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float* __restrict__ X, long long x_plane_stride, const float* __restrict__ Y, long long y_frame_stride,
                long long y_chan_stride, Batch bt, int h, int w, float C1, float C2, const float* __restrict__ coef /*[planes]*/,
                int use_ssim /* level 4: ssim = l*cs */, const float* __restrict__ d_next, long long dn_plane_stride, int hn, int wn,
                int ph, int pw, float* __restrict__ dX, long long dx_plane_stride) {
  extern __shared__ float smem[];
  float* sx = smem;                                  // [BR][BR]
  float* sy = sx + BR * BR;                          // [BR][BR]
  float* vv = sy + BR * BR;                          // [5][BP][BR]   vertical pass
  float* pm = vv + 5 * BP * BR;                      // [3][BP][BP]   partials dmu1, de11, de12
  float* tt = vv;                                    // [3][BT][BP]   (reuses vv after partials are built)
  const int plane = blockIdx.z;
  const int b = plane / 3, c = plane - 3 * b;
  const float* xp = X + plane * x_plane_stride;
  const float* yp = Y + bt.idx[b] * y_frame_stride + c * y_chan_stride;
  const int oh = h - 10, ow = w - 10;
  const int qy0 = blockIdx.y * BT, qx0 = blockIdx.x * BT;   // output (image-domain) tile origin
  const int iy0 = qy0 - 10, ix0 = qx0 - 10;                 // input tile origin == partial-map tile origin
  const float g_coef = coef[plane];
  for (int i = threadIdx.x; i < BR * BR; i += blockDim.x) {
    const int r = i / BR, cc = i - r * BR;
    const int y = iy0 + r, x = ix0 + cc;
    const bool in = y >= 0 && y < h && x >= 0 && x < w;
    sx[i] = in ? xp[(long long)y * w + x] : 0.f;
    sy[i] = in ? yp[(long long)y * w + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BP * BR; i += blockDim.x) {
    const int r = i / BR, cc = i - r * BR;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_gauss[k], xv = sx[(r + k) * BR + cc], yv = sy[(r + k) * BR + cc];
      a0 += g * xv; a1 += g * yv; a2 += g * xv * xv; a3 += g * yv * yv; a4 += g * xv * yv;
    }
    vv[(0 * BP + r) * BR + cc] = a0; vv[(1 * BP + r) * BR + cc] = a1; vv[(2 * BP + r) * BR + cc] = a2;
    vv[(3 * BP + r) * BR + cc] = a3; vv[(4 * BP + r) * BR + cc] = a4;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BP * BP; i += blockDim.x) {
    const int r = i / BP, cc = i - r * BP;
    const int py_ = iy0 + r, px_ = ix0 + cc;     // map position
    float dmu1 = 0.f, de11 = 0.f, de12 = 0.f;
    if (py_ >= 0 && py_ < oh && px_ >= 0 && px_ < ow) {
      SsimMaps m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = c_gauss[k];
        m.mu1 += g * vv[(0 * BP + r) * BR + cc + k]; m.mu2 += g * vv[(1 * BP + r) * BR + cc + k];
        m.e11 += g * vv[(2 * BP + r) * BR + cc + k]; m.e22 += g * vv[(3 * BP + r) * BR + cc + k];
        m.e12 += g * vv[(4 * BP + r) * BR + cc + k];
      }
      const float mu1_sq = m.mu1 * m.mu1, mu2_sq = m.mu2 * m.mu2, mu12 = m.mu1 * m.mu2;
      const float A1 = 2.f * mu12 + C1, B1 = mu1_sq + mu2_sq + C1;
      const float s11 = m.e11 - mu1_sq, s22 = m.e22 - mu2_sq, s12 = m.e12 - mu12;
      const float A2 = 2.f * s12 + C2, B2 = s11 + s22 + C2;
      const float cs = A2 / B2;
      // d cs / d(e12, e11, mu1)
      const float dcs_e12 = 2.f / B2;
      const float dcs_e11 = -A2 / (B2 * B2);
      const float dcs_mu1 = -2.f * m.mu2 / B2 + 2.f * m.mu1 * A2 / (B2 * B2);
      if (use_ssim) {
        const float l = A1 / B1;
        const float dl_mu1 = (2.f * m.mu2 * B1 - A1 * 2.f * m.mu1) / (B1 * B1);
        dmu1 = g_coef * (cs * dl_mu1 + l * dcs_mu1);
        de11 = g_coef * l * dcs_e11;
        de12 = g_coef * l * dcs_e12;
      } else {
        dmu1 = g_coef * dcs_mu1; de11 = g_coef * dcs_e11; de12 = g_coef * dcs_e12;
      }
    }
    pm[(0 * BP + r) * BP + cc] = dmu1; pm[(1 * BP + r) * BP + cc] = de11; pm[(2 * BP + r) * BP + cc] = de12;
  }
  __syncthreads();
  // transposed filter, vertical: t[m][qy][px] = sum_k g[k] * pm[m][qy + 10 - k][px]   (map row = q - k)
  for (int i = threadIdx.x; i < 3 * BT * BP; i += blockDim.x) {
    const int mI = i / (BT * BP);
    const int rem = i - mI * BT * BP;
    const int r = rem / BP, cc = rem - r * BP;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) a += c_gauss[k] * pm[(mI * BP + r + 10 - k) * BP + cc];
    tt[(mI * BT + r) * BP + cc] = a;
  }
  __syncthreads();
  float* dxp = dX + plane * dx_plane_stride;
  for (int i = threadIdx.x; i < BT * BT; i += blockDim.x) {
    const int r = i / BT, cc = i - r * BT;
    const int y = qy0 + r, x = qx0 + cc;
    if (y < h && x < w) {
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = c_gauss[k];
        r0 += g * tt[(0 * BT + r) * BP + cc + 10 - k];
        r1 += g * tt[(1 * BT + r) * BP + cc + 10 - k];
        r2 += g * tt[(2 * BT + r) * BP + cc + 10 - k];
      }
      const float xv = sx[(r + 10) * BR + cc + 10], yv = sy[(r + 10) * BR + cc + 10];
      float gsum = r0 + 2.f * xv * r1 + yv * r2;
      if (d_next) {
        const int oy = (y + ph) >> 1, ox = (x + pw) >> 1;
        if (oy < hn && ox < wn) gsum += 0.25f * d_next[plane * dn_plane_stride + (long long)oy * wn + ox];
      }
      dxp[(long long)y * w + x] = gsum;
    }
  }
}

// ------------------------------------------------------------------------------------------
// producers of X (batch frames then predecessors)
// ------------------------------------------------------------------------------------------
// stage 2: X = clamp(SH2RGB(fdc[id]), 0, 1); flags bit c = gradient passes (0 <= v <= 1)
__global__ void uvt_gather_kernel(const float* __restrict__ fdc, const int* __restrict__ ids, long long P, Batch bt,
                                  float* __restrict__ X, unsigned char* __restrict__ flags) {
  const long long total = (long long)2 * bt.n * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    int fr = f < bt.n ? bt.idx[f] : bt.idx[f - bt.n] - 1;
    if (fr < 0) fr = 0;
    const int id = ids[(long long)fr * P + p];
    unsigned char fl = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fdc[(long long)id * 3 + c] * SH_C0 + 0.5f;
      if (v >= 0.f && v <= 1.f) fl |= (1u << c);
      X[((long long)f * 3 + c) * P + p] = fminf(fmaxf(v, 0.f), 1.f);
    }
    flags[i] = fl;
  }
}

// stage 1: X_j = clamp(sum_k in_k E[k][j] + E[j][3], 0, 1), E = exposure[frame] (3x4 row-major)
__global__ void exposure_apply_kernel(const float* __restrict__ edited, const float* __restrict__ expo, long long P, Batch bt,
                                      float* __restrict__ X, unsigned char* __restrict__ flags) {
  const long long total = (long long)2 * bt.n * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    int fr = f < bt.n ? bt.idx[f] : bt.idx[f - bt.n] - 1;
    if (fr < 0) fr = 0;
    const float* E = expo + (long long)fr * 12;
    const float* src = edited + (long long)fr * 3 * P + p;
    const float in0 = src[0], in1 = src[P], in2 = src[2 * P];
    unsigned char fl = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = in0 * E[0 * 4 + j] + in1 * E[1 * 4 + j] + in2 * E[2 * 4 + j] + E[j * 4 + 3];
      if (v >= 0.f && v <= 1.f) fl |= (1u << j);
      X[((long long)f * 3 + j) * P + p] = fminf(fmaxf(v, 0.f), 1.f);
    }
    flags[i] = fl;
  }
}

// ------------------------------------------------------------------------------------------
// fused level-0 kernel over the batch frames
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  c[0] = cubic2(t + 1.f); c[1] = cubic1(t); c[2] = cubic1(1.f - t); c[3] = cubic2((1.f - t) + 1.f);
}

// one 16-byte reduction instead of three scalar ones (REDG.E.ADD.F32x4): the predecessor-gradient image and the UVT
// gradient keep 4 floats per pixel / row for this (lane 3 is padding)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

struct L0Params {
  int H, W;
  long long P;
  Batch bt;
  const float* X;                 // [2n,3,P]
  const unsigned char* flags;     // [2n,P]
  const float* flows;             // [N,2,P]
  const float* mask;              // [N,1,P]
  const float* d1;                // level-1 gradient [n*3 planes]
  long long d1_stride; int h1, w1, ph0, pw0;
  float k_flow;                   // lambda_flow / (n_valid*3*P)
  float k_tvh, k_tvw;             // lambda_tv*2/(count_h*n), lambda_tv*2/(count_w*n)
  float k_l1;                     // stage 1: (1-lambda_flow)*(1-lambda_dssim)/(n*3*P); 0 in stage 2
  const float* edited;            // [N,3,P] (stage 1 L1 target and affine input)
  float* G_pre;                   // [n,P,4] atomically accumulated predecessor gradient (lane 3 unused)
  float* scal;
  // sinks
  const int* ids; float* grad_fdc;        // stage 2: [U,4] (lane 3 unused)
  float* grad_expo;                       // stage 1: [N,12]
};

template <int MODE /*0 = UVT, 1 = exposure*/>
__global__ void __launch_bounds__(256)
level0_kernel(L0Params q) {
  __shared__ float red[4][8];
  __shared__ float eg[12];
  const int b = blockIdx.y;
  const int fr = q.bt.idx[b];
  const bool valid = fr > 0;
  const float* Xi = q.X + (long long)b * 3 * q.P;
  const float* Xp = q.X + (long long)(q.bt.n + b) * 3 * q.P;
  float acc_flow = 0.f, acc_tvh = 0.f, acc_tvw = 0.f, acc_l1 = 0.f;
  float ge[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) ge[k] = 0.f;
  if (MODE == 1 && threadIdx.x < 12) eg[threadIdx.x] = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < q.P; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % q.W), y = (int)(p / q.W);
    float xi[3], g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { xi[c] = Xi[c * q.P + p]; g[c] = 0.f; }
    // ---- flow term: warp the predecessor with the backward flow ----
    if (valid) {
      const float fx = q.flows[((long long)fr * 2 + 0) * q.P + p] + (float)x;
      const float fy = q.flows[((long long)fr * 2 + 1) * q.P + p] + (float)y;
      const float gx = (fx / (float)(q.W - 1) - 0.5f) * 2.f;
      const float gy = (fy / (float)(q.H - 1) - 0.5f) * 2.f;
      const float ix = ((gx + 1.f) / 2.f) * (float)(q.W - 1);
      const float iy = ((gy + 1.f) / 2.f) * (float)(q.H - 1);
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      float cx[4], cy[4];
      cubic_coeffs(ix - fx0, cx);
      cubic_coeffs(iy - fy0, cy);
      const int x0 = (int)fx0 - 1, y0 = (int)fy0 - 1;
      float wv[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = y0 + j;
        if (yy < 0 || yy >= q.H) continue;
        float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xx = x0 + i;
          if (xx < 0 || xx >= q.W) continue;
          const long long o = (long long)yy * q.W + xx;
#pragma unroll
          for (int c = 0; c < 3; ++c) row[c] += Xp[c * q.P + o] * cx[i];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) wv[c] += row[c] * cy[j];
      }
      const float m = q.mask[(long long)fr * q.P + p];
      float s[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = wv[c] * m - xi[c] * m;
        acc_flow += fabsf(d);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        s[c] = sg * m * q.k_flow;
        g[c] -= s[c];
      }
      float* Gp = q.G_pre + (long long)b * 4 * q.P;
      if (s[0] != 0.f || s[1] != 0.f || s[2] != 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int yy = y0 + j;
          if (yy < 0 || yy >= q.H) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int xx = x0 + i;
            if (xx < 0 || xx >= q.W) continue;
            const float wgt = cx[i] * cy[j];
            const long long o = (long long)yy * q.W + xx;
            red_add_v4(Gp + o * 4, s[0] * wgt, s[1] * wgt, s[2] * wgt);
          }
        }
      }
    }
    // ---- total variation (stage 2) ----
    if (q.k_tvh != 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pl = Xi + c * q.P;
        const float v = xi[c];
        float gg = 0.f;
        if (y + 1 < q.H) { const float d = pl[p + q.W] - v; acc_tvh += d * d; gg -= q.k_tvh * 2.f * d; }
        if (y > 0) gg += q.k_tvh * 2.f * (v - pl[p - q.W]);
        if (x + 1 < q.W) { const float d = pl[p + 1] - v; acc_tvw += d * d; gg -= q.k_tvw * 2.f * d; }
        if (x > 0) gg += q.k_tvw * 2.f * (v - pl[p - 1]);
        g[c] += gg;
      }
    }
    // ---- L1 to the target (stage 1) ----
    float in[3] = {0.f, 0.f, 0.f};
    if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        in[c] = q.edited[((long long)fr * 3 + c) * q.P + p];
        const float d = xi[c] - in[c];
        acc_l1 += fabsf(d);
        g[c] += q.k_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
    }
    // ---- MS-SSIM gradient from level 1 (avg-pool backward) ----
    {
      const int oy = (y + q.ph0) >> 1, ox = (x + q.pw0) >> 1;
      if (oy < q.h1 && ox < q.w1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] += 0.25f * q.d1[(long long)(b * 3 + c) * q.d1_stride + (long long)oy * q.w1 + ox];
      }
    }
    // ---- sink ----
    const unsigned char fl = q.flags[(long long)b * q.P + p];
    if (MODE == 0) {
      const int id = q.ids[(long long)fr * q.P + p];
      if (fl & 7)
        red_add_v4(q.grad_fdc + (long long)id * 4, (fl & 1) ? g[0] * SH_C0 : 0.f, (fl & 2) ? g[1] * SH_C0 : 0.f,
                   (fl & 4) ? g[2] * SH_C0 : 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float gj = ((fl >> j) & 1) ? g[j] : 0.f;
        ge[0 * 4 + j] += gj * in[0]; ge[1 * 4 + j] += gj * in[1]; ge[2 * 4 + j] += gj * in[2];
        ge[j * 4 + 3] += gj;
      }
    }
  }
  // block reductions
  acc_flow = warp_sum(acc_flow); acc_tvh = warp_sum(acc_tvh); acc_tvw = warp_sum(acc_tvw); acc_l1 = warp_sum(acc_l1);
  if ((threadIdx.x & 31) == 0) {
    const int wq = threadIdx.x >> 5;
    red[0][wq] = acc_flow; red[1][wq] = acc_tvh; red[2][wq] = acc_tvw; red[3][wq] = acc_l1;
  }
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float r = warp_sum(ge[k]);
      if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(&eg[k], r);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f, cc = 0.f, dd = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; bb += red[1][i]; cc += red[2][i]; dd += red[3][i]; }
    atomicAdd(&q.scal[SC_FLOW_ABS], a); atomicAdd(&q.scal[SC_TV_H], bb); atomicAdd(&q.scal[SC_TV_W], cc); atomicAdd(&q.scal[SC_L1], dd);
  }
  if (MODE == 1 && threadIdx.x < 12) atomicAdd(&q.grad_expo[(long long)fr * 12 + threadIdx.x], eg[threadIdx.x]);
}

// predecessor half: G_pre -> clamp mask -> sink; clears G_pre for the next iteration
template <int MODE>
__global__ void __launch_bounds__(256)
pre_sink_kernel(L0Params q) {
  __shared__ float eg[12];
  const int b = blockIdx.y;
  int fr = q.bt.idx[b] - 1;
  if (fr < 0) fr = 0;
  float ge[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) ge[k] = 0.f;
  if (MODE == 1 && threadIdx.x < 12) eg[threadIdx.x] = 0.f;
  float4* Gp = reinterpret_cast<float4*>(q.G_pre + (long long)b * 4 * q.P);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < q.P; p += (long long)gridDim.x * blockDim.x) {
    const float4 gv = Gp[p];
    if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f) continue;
    Gp[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float g[3] = {gv.x, gv.y, gv.z};
    const unsigned char fl = q.flags[(long long)(q.bt.n + b) * q.P + p];
    if (MODE == 0) {
      const int id = q.ids[(long long)fr * q.P + p];
      if (fl & 7)
        red_add_v4(q.grad_fdc + (long long)id * 4, (fl & 1) ? g[0] * SH_C0 : 0.f, (fl & 2) ? g[1] * SH_C0 : 0.f,
                   (fl & 4) ? g[2] * SH_C0 : 0.f);
    } else {
      float in[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) in[c] = q.edited[((long long)fr * 3 + c) * q.P + p];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float gj = ((fl >> j) & 1) ? g[j] : 0.f;
        ge[0 * 4 + j] += gj * in[0]; ge[1 * 4 + j] += gj * in[1]; ge[2 * 4 + j] += gj * in[2];
        ge[j * 4 + 3] += gj;
      }
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float r = warp_sum(ge[k]);
      if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(&eg[k], r);
    }
    __syncthreads();
    if (threadIdx.x < 12) atomicAdd(&q.grad_expo[(long long)fr * 12 + threadIdx.x], eg[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, no weight decay / amsgrad), dense over n elements; clears the gradient.
// Also assembles this iteration's loss scalars (thread 0 of block 0).
// ------------------------------------------------------------------------------------------
struct LossAsm {
  float* scal; float* loss_out;     // loss_out[0..2] = total, flow, photometric
  float inv_flow_cnt, inv_tv_h, inv_tv_w, inv_l1_cnt;
  float lambda_flow, lambda_dssim, lambda_tv_over_n;
  float ms_share;                    // local planes / global planes (1 on a single GPU)
  int stage;                         // 1 or 2
};

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, LossAsm la) {
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi; v[i] = vi; g[i] = 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && la.loss_out) {
    const float flow = la.scal[SC_FLOW_ABS] * la.inv_flow_cnt;
    const float ms = la.scal[SC_MSSSIM];
    float photo, tv = 0.f;
    if (la.stage == 1) {
      photo = la.scal[SC_L1] * la.inv_l1_cnt * (1.f - la.lambda_dssim) + la.ms_share * (1.f - ms) * la.lambda_dssim;
    } else {
      photo = la.ms_share * (1.f - ms) * la.lambda_dssim;
      tv = la.lambda_tv_over_n * 2.f * (la.scal[SC_TV_H] * la.inv_tv_h + la.scal[SC_TV_W] * la.inv_tv_w);
    }
    la.loss_out[0] = (1.f - la.lambda_flow) * photo + la.lambda_flow * flow + tv;
    la.loss_out[1] = flow;
    la.loss_out[2] = photo;
  }
}

// Stage-2 Adam over the UVT rows: p, m, v are [U,3], the gradient is [U,4] (lane 3 padding).  One thread handles four
// rows = three float4 of p/m/v and four float4 of g, so every access is a 16-byte vector.
__global__ void __launch_bounds__(256)
adam_uvt_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long U,
                float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  const long long quads = U / 4;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < quads; t += (long long)gridDim.x * blockDim.x) {
    float4* p4 = reinterpret_cast<float4*>(p) + t * 3;
    float4* m4 = reinterpret_cast<float4*>(m) + t * 3;
    float4* v4 = reinterpret_cast<float4*>(v) + t * 3;
    float4* g4 = reinterpret_cast<float4*>(g) + t * 4;
    float pa[12], ma[12], va[12], ga[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 a = p4[k], b = m4[k], c = v4[k];
      pa[4 * k] = a.x; pa[4 * k + 1] = a.y; pa[4 * k + 2] = a.z; pa[4 * k + 3] = a.w;
      ma[4 * k] = b.x; ma[4 * k + 1] = b.y; ma[4 * k + 2] = b.z; ma[4 * k + 3] = b.w;
      va[4 * k] = c.x; va[4 * k + 1] = c.y; va[4 * k + 2] = c.z; va[4 * k + 3] = c.w;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 a = g4[r];
      ga[3 * r] = a.x; ga[3 * r + 1] = a.y; ga[3 * r + 2] = a.z;
      g4[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) upd(pa[k], ga[k], ma[k], va[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      p4[k] = make_float4(pa[4 * k], pa[4 * k + 1], pa[4 * k + 2], pa[4 * k + 3]);
      m4[k] = make_float4(ma[4 * k], ma[4 * k + 1], ma[4 * k + 2], ma[4 * k + 3]);
      v4[k] = make_float4(va[4 * k], va[4 * k + 1], va[4 * k + 2], va[4 * k + 3]);
    }
  }
  // tail rows (U % 4)
  if (blockIdx.x == 0 && threadIdx.x < 3 * (int)(U - quads * 4)) {
    const long long row = quads * 4 + threadIdx.x / 3;
    const int c = threadIdx.x % 3;
    const long long i = row * 3 + c;
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[row * 4 + c], mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi; g[row * 4 + c] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// UVT initialisation (generate.py:477-479: scatter-mean + RGB2SH) and final render (:529-531)
// ------------------------------------------------------------------------------------------
__global__ void uvt_accum_kernel(const float* __restrict__ edited, const int* __restrict__ ids, long long P, int N,
                                 float* __restrict__ sum, float* __restrict__ cnt) {
  const long long total = (long long)N * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    const int id = ids[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&sum[(long long)id * 3 + c], edited[((long long)f * 3 + c) * P + p]);
    atomicAdd(&cnt[id], 1.f);
  }
}
__global__ void uvt_finish_init_kernel(float* __restrict__ fdc, const float* __restrict__ cnt, long long U) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < U * 3; i += (long long)gridDim.x * blockDim.x) {
    const float c = fmaxf(cnt[i / 3], 1.f);
    fdc[i] = (fdc[i] / c - 0.5f) / SH_C0;
  }
}
__global__ void uvt_render_kernel(const float* __restrict__ fdc, const int* __restrict__ ids, long long P, int N,
                                  float* __restrict__ out) {
  const long long total = (long long)N * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    const int id = ids[i];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[((long long)f * 3 + c) * P + p] = fminf(fmaxf(fdc[(long long)id * 3 + c] * SH_C0 + 0.5f, 0.f), 1.f);
  }
}
// dataset.exposure_align (utils/dataloader.py:39-42): bake the affine into the frames, in place
__global__ void exposure_bake_kernel(float* __restrict__ edited, const float* __restrict__ expo, long long P, int N) {
  const long long total = (long long)N * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    const float* E = expo + (long long)f * 12;
    float* src = edited + (long long)f * 3 * P + p;
    const float in0 = src[0], in1 = src[P], in2 = src[2 * P];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = in0 * E[0 * 4 + j] + in1 * E[1 * 4 + j] + in2 * E[2 * 4 + j] + E[j * 4 + 3];
      src[j * P] = fminf(fmaxf(v, 0.f), 1.f);
    }
  }
}

static inline int gridp(long long total, int block, int cap = 148 * 8) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int ensure_gauss() {
  if (g_gauss_ready) return TCL_OK;
  float g[11];
  float s = 0.f;
  for (int i = 0; i < 11; ++i) { const float c = (float)(i - 5); g[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f)); s += g[i]; }
  for (int i = 0; i < 11; ++i) g[i] /= s;
  cudaError_t e = cudaMemcpyToSymbol(c_gauss, g, sizeof(g));
  if (e != cudaSuccess) { set_last_error("postopt: gaussian upload failed: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
  g_gauss_ready = true;
  return TCL_OK;
}

// workspace carving ---------------------------------------------------------------------
struct Ws {
  float* X; unsigned char* flags; float* G_pre; float* xpyr; float* dpyr; float* sums; float* coef; float* scal;
};
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t ws_bytes(int H, int W, int nb) {
  Pyr py; make_pyr(H, W, &py);
  const size_t P = (size_t)H * W;
  size_t b = 0;
  b += align_up(sizeof(float) * 2 * nb * 3 * P);          // X
  b += align_up((size_t)2 * nb * P);                      // flags
  b += align_up(sizeof(float) * nb * 4 * P);              // G_pre [nb,P,4]
  b += align_up(sizeof(float) * nb * 3 * py.total);       // X pyramid levels 1..4
  b += align_up(sizeof(float) * nb * 3 * py.total);       // dX pyramid
  b += align_up(sizeof(float) * 4 * nb * 3 * 2);          // sums
  b += align_up(sizeof(float) * 4 * nb * 3);              // coef
  b += align_up(sizeof(float) * SC_COUNT);                // scalars
  return b;
}
static void carve(void* base, int H, int W, int nb, Ws* w) {
  Pyr py; make_pyr(H, W, &py);
  const size_t P = (size_t)H * W;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  w->X = (float*)p; p += align_up(sizeof(float) * 2 * nb * 3 * P);
  w->flags = p; p += align_up((size_t)2 * nb * P);
  w->G_pre = (float*)p; p += align_up(sizeof(float) * nb * 4 * P);
  w->xpyr = (float*)p; p += align_up(sizeof(float) * nb * 3 * py.total);
  w->dpyr = (float*)p; p += align_up(sizeof(float) * nb * 3 * py.total);
  w->sums = (float*)p; p += align_up(sizeof(float) * 4 * nb * 3 * 2);
  w->coef = (float*)p; p += align_up(sizeof(float) * 4 * nb * 3);
  w->scal = (float*)p;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_postopt_workspace_bytes(int H, int W, int max_batch) { return ws_bytes(H, W, max_batch); }

extern "C" long long tcl_postopt_pyramid_elems(int H, int W) { Pyr py; make_pyr(H, W, &py); return py.total; }

// Target pyramid: ypyr[frame][channel][levels 1..4] from edited [N,3,H,W].
extern "C" int tcl_postopt_build_pyramid(const float* edited, int N, int H, int W, float* ypyr, cudaStream_t stream) {
  TCL_CHECK_ARG(edited && ypyr && N > 0, "tcl_postopt_build_pyramid: args");
  Pyr py; make_pyr(H, W, &py);
  TCL_CHECK_ARG(py.h[4] >= 11 && py.w[4] >= 11, "tcl_postopt_build_pyramid: image too small for 5-level MS-SSIM (%dx%d)", H, W);
  const int planes = N * 3;
  const long long P = (long long)H * W;
  for (int l = 0; l < 4; ++l) {
    const float* in = l == 0 ? edited : ypyr + py.off[l];
    const long long in_stride = l == 0 ? P : py.total;
    const long long total = (long long)planes * py.h[l + 1] * py.w[l + 1];
    avgpool2_kernel<<<gridp(total, 256), 256, 0, stream>>>(in, in_stride, py.h[l], py.w[l], py.ph[l], py.pw[l],
                                                           ypyr + py.off[l + 1], py.total, py.h[l + 1], py.w[l + 1], planes);
    TCL_CHECK_LAUNCH("tcl_postopt_build_pyramid");
  }
  return TCL_OK;
}

static int run_iteration(int stage, const tcl_postopt_ctx* c, const int* idx_host, int nb,
                         // stage 2
                         const int* ids, long long U, float* fdc, float* grad, float* m, float* v,
                         // stage 1
                         float* expo, float* egrad, float* em, float* ev,
                         float lr, float beta1, float beta2, float eps, int step, float* loss_out, cudaStream_t stream,
                         bool do_adam = true) {
  TCL_CHECK_ARG(c && idx_host && nb > 0 && nb <= MAXB, "postopt: batch size %d (max %d)", nb, MAXB);
  TCL_CHECK_ARG(c->edited && c->past_flows && c->mask_bwd && c->ypyr && c->workspace, "postopt: null context pointer");
  TCL_CHECK_ARG(c->max_batch >= nb && c->max_batch <= MAXB, "postopt: batch %d exceeds ctx.max_batch %d", nb, c->max_batch);
  TCL_CHECK_ARG(c->workspace_bytes >= ws_bytes(c->H, c->W, c->max_batch), "postopt: workspace too small");
  TCL_CHECK_ARG(step >= 1 || !do_adam, "postopt: Adam step must start at 1");
  // data-parallel normalisers: means are over the GLOBAL batch; 0 = this call is the whole batch
  const int nb_g = c->norm_batch > 0 ? c->norm_batch : nb;
  TCL_CHECK_ARG(nb_g >= nb, "postopt: norm_batch %d < local batch %d", nb_g, nb);
  int rc = ensure_gauss();
  if (rc) return rc;
  const int H = c->H, W = c->W;
  const long long P = (long long)H * W;
  Pyr py; make_pyr(H, W, &py);
  TCL_CHECK_ARG(py.h[4] >= 11 && py.w[4] >= 11, "postopt: image too small for 5-level MS-SSIM");
  Ws w; carve(c->workspace, H, W, c->max_batch, &w);   // fixed layout: G_pre must stay zero between iterations
  Batch bt; bt.n = nb;
  int n_valid = 0;
  for (int i = 0; i < nb; ++i) {
    TCL_CHECK_ARG(idx_host[i] >= 0 && idx_host[i] < c->N, "postopt: frame index %d out of range", idx_host[i]);
    bt.idx[i] = idx_host[i];
    n_valid += idx_host[i] > 0;
  }
  for (int i = nb; i < MAXB; ++i) bt.idx[i] = 0;
  const int planes = nb * 3;
  cudaMemsetAsync(w.sums, 0, sizeof(float) * 4 * planes * 2, stream);
  cudaMemsetAsync(w.scal, 0, sizeof(float) * SC_COUNT, stream);
  // 1. produce X
  if (stage == 2) uvt_gather_kernel<<<gridp(2 * nb * P, 256, 148 * 16), 256, 0, stream>>>(fdc, ids, P, bt, w.X, w.flags);
  else exposure_apply_kernel<<<gridp(2 * nb * P, 256, 148 * 16), 256, 0, stream>>>(c->edited, expo, P, bt, w.X, w.flags);
  TCL_CHECK_LAUNCH("postopt(produce)");
  // 2. pyramid of the batch frames
  for (int l = 0; l < 4; ++l) {
    const float* in = l == 0 ? w.X : w.xpyr + py.off[l];
    const long long in_stride = l == 0 ? P : py.total;
    const long long total = (long long)planes * py.h[l + 1] * py.w[l + 1];
    avgpool2_kernel<<<gridp(total, 256), 256, 0, stream>>>(in, in_stride, py.h[l], py.w[l], py.ph[l], py.pw[l],
                                                           w.xpyr + py.off[l + 1], py.total, py.h[l + 1], py.w[l + 1], planes);
    TCL_CHECK_LAUNCH("postopt(pyramid)");
  }
  // 3. SSIM forward per level
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;   // data_range = 1
  for (int l = 1; l <= 4; ++l) {
    dim3 grid((py.w[l] - 10 + ST - 1) / ST, (py.h[l] - 10 + ST - 1) / ST, planes);
    ssim_fwd_kernel<<<grid, 256, 0, stream>>>(w.xpyr + py.off[l], py.total, c->ypyr + py.off[l], 3 * py.total, py.total, bt,
                                              py.h[l], py.w[l], C1, C2, w.sums + (size_t)(l - 1) * planes * 2);
    TCL_CHECK_LAUNCH("postopt(ssim_fwd)");
  }
  // 4. loss head
  const int planes_g = nb_g * 3;
  const float k_ms = -(1.f - c->lambda_flow) * c->lambda_dssim / (float)planes_g;
  msssim_head_kernel<<<1, 64, 0, stream>>>(w.sums, planes, py, k_ms, w.coef, w.scal);
  TCL_CHECK_LAUNCH("postopt(head)");
  // 5. SSIM backward 4 -> 1
  static bool bwd_conf = false;
  const size_t bwd_smem = sizeof(float) * (2 * BR * BR + 5 * BP * BR + 3 * BP * BP);
  if (!bwd_conf) {
    cudaError_t e = cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem);
    if (e != cudaSuccess) { set_last_error("postopt: smem attr: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
    bwd_conf = true;
  }
  for (int l = 4; l >= 1; --l) {
    dim3 grid((py.w[l] + BT - 1) / BT, (py.h[l] + BT - 1) / BT, planes);
    const float* dn = l < 4 ? w.dpyr + py.off[l + 1] : nullptr;
    ssim_bwd_kernel<<<grid, 256, bwd_smem, stream>>>(w.xpyr + py.off[l], py.total, c->ypyr + py.off[l], 3 * py.total, py.total, bt,
                                                     py.h[l], py.w[l], C1, C2, w.coef + (size_t)(l - 1) * planes, l == 4 ? 1 : 0,
                                                     dn, py.total, l < 4 ? py.h[l + 1] : 0, l < 4 ? py.w[l + 1] : 0, py.ph[l], py.pw[l],
                                                     w.dpyr + py.off[l], py.total);
    TCL_CHECK_LAUNCH("postopt(ssim_bwd)");
  }
  // 6. fused level-0 kernel + predecessor sink
  L0Params q;
  memset(&q, 0, sizeof(q));
  q.H = H; q.W = W; q.P = P; q.bt = bt; q.X = w.X; q.flags = w.flags; q.flows = c->past_flows; q.mask = c->mask_bwd;
  q.d1 = w.dpyr + py.off[1]; q.d1_stride = py.total; q.h1 = py.h[1]; q.w1 = py.w[1]; q.ph0 = py.ph[0]; q.pw0 = py.pw[0];
  const int n_valid_g = c->norm_batch > 0 ? c->norm_valid : n_valid;
  const float flow_cnt = (float)n_valid_g * 3.f * (float)P;
  q.k_flow = n_valid_g > 0 ? c->lambda_flow / flow_cnt : 0.f;
  const float count_h = 3.f * (float)(H - 1) * (float)W, count_w = 3.f * (float)H * (float)(W - 1);
  if (stage == 2) { q.k_tvh = c->lambda_tv * 2.f / (count_h * nb_g); q.k_tvw = c->lambda_tv * 2.f / (count_w * nb_g); }
  q.k_l1 = stage == 1 ? (1.f - c->lambda_flow) * (1.f - c->lambda_dssim) / ((float)planes_g * (float)P) : 0.f;
  q.edited = c->edited; q.G_pre = w.G_pre; q.scal = w.scal; q.ids = ids; q.grad_fdc = grad; q.grad_expo = egrad;
  dim3 g0(gridp(P, 256, 148 * 2), nb);
  if (stage == 2) level0_kernel<0><<<g0, 256, 0, stream>>>(q); else level0_kernel<1><<<g0, 256, 0, stream>>>(q);
  TCL_CHECK_LAUNCH("postopt(level0)");
  if (stage == 2) pre_sink_kernel<0><<<g0, 256, 0, stream>>>(q); else pre_sink_kernel<1><<<g0, 256, 0, stream>>>(q);
  TCL_CHECK_LAUNCH("postopt(pre_sink)");
  // 7. Adam (+ loss assembly)
  LossAsm la;
  la.scal = w.scal; la.loss_out = loss_out;
  la.inv_flow_cnt = n_valid_g > 0 ? 1.f / flow_cnt : NAN;   // mean over an empty selection is NaN in the reference
  la.inv_tv_h = 1.f / count_h; la.inv_tv_w = 1.f / count_w; la.inv_l1_cnt = 1.f / ((float)planes_g * (float)P);
  la.lambda_flow = c->lambda_flow; la.lambda_dssim = c->lambda_dssim; la.lambda_tv_over_n = c->lambda_tv / nb_g; la.stage = stage;
  la.ms_share = (float)planes / (float)planes_g;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  if (!do_adam) {
    // gradient only (data parallel: the caller all-reduces grad, then calls tcl_adam_step): n = 0 elements,
    // the single block just assembles this rank's share of the loss
    adam_kernel<<<1, 32, 0, stream>>>(nullptr, nullptr, nullptr, nullptr, 0, 0.f, beta1, beta2, eps, 1.f, 1.f, la);
  } else if (stage == 2) {
    adam_uvt_kernel<<<gridp(U / 4 + 1, 256, 148 * 16), 256, 0, stream>>>(fdc, grad, m, v, U, lr, beta1, beta2, eps, bc1, bc2_sqrt);
    TCL_CHECK_LAUNCH("postopt adam");
    adam_kernel<<<1, 32, 0, stream>>>(nullptr, nullptr, nullptr, nullptr, 0, 0.f, beta1, beta2, eps, 1.f, 1.f, la);   // loss assembly
  }
  else adam_kernel<<<gridp((long long)c->N * 12, 256), 256, 0, stream>>>(expo, egrad, em, ev, (long long)c->N * 12, lr, beta1, beta2, eps, bc1, bc2_sqrt, la);
  TCL_CHECK_LAUNCH("postopt(adam)");
  return TCL_OK;
}

extern "C" int tcl_uvt_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids, long long U,
                                 float* fdc, float* grad, float* m, float* v, float lr, float beta1, float beta2, float eps,
                                 int step, float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(ids && fdc && grad && m && v && U > 0, "tcl_uvt_iteration: null pointer");
  return run_iteration(2, ctx, idx_host, n_batch, ids, U, fdc, grad, m, v, nullptr, nullptr, nullptr, nullptr, lr, beta1, beta2, eps,
                       step, loss_out, stream);
}

extern "C" int tcl_exposure_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, float* exposure, float* grad,
                                      float* m, float* v, float lr, float beta1, float beta2, float eps, int step,
                                      float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(exposure && grad && m && v, "tcl_exposure_iteration: null pointer");
  return run_iteration(1, ctx, idx_host, n_batch, nullptr, 0, nullptr, nullptr, nullptr, nullptr, exposure, grad, m, v, lr, beta1,
                       beta2, eps, step, loss_out, stream);
}

extern "C" int tcl_uvt_init(const float* edited, const int* ids, int N, int H, int W, long long U, float* fdc, float* cnt_ws,
                            cudaStream_t stream) {
  TCL_CHECK_ARG(edited && ids && fdc && cnt_ws && U > 0, "tcl_uvt_init: args");
  const long long P = (long long)H * W;
  cudaMemsetAsync(fdc, 0, sizeof(float) * U * 3, stream);
  cudaMemsetAsync(cnt_ws, 0, sizeof(float) * U, stream);
  uvt_accum_kernel<<<gridp((long long)N * P, 256, 148 * 16), 256, 0, stream>>>(edited, ids, P, N, fdc, cnt_ws);
  TCL_CHECK_LAUNCH("tcl_uvt_init(accum)");
  uvt_finish_init_kernel<<<gridp(U * 3, 256, 148 * 16), 256, 0, stream>>>(fdc, cnt_ws, U);
  TCL_CHECK_LAUNCH("tcl_uvt_init(finish)");
  return TCL_OK;
}

extern "C" int tcl_uvt_render(const float* fdc, const int* ids, int N, int H, int W, float* out, cudaStream_t stream) {
  TCL_CHECK_ARG(fdc && ids && out, "tcl_uvt_render: args");
  const long long P = (long long)H * W;
  uvt_render_kernel<<<gridp((long long)N * P, 256, 148 * 16), 256, 0, stream>>>(fdc, ids, P, N, out);
  TCL_CHECK_LAUNCH("tcl_uvt_render");
  return TCL_OK;
}

extern "C" int tcl_exposure_bake(float* edited, const float* exposure, int N, int H, int W, cudaStream_t stream) {
  TCL_CHECK_ARG(edited && exposure, "tcl_exposure_bake: args");
  const long long P = (long long)H * W;
  exposure_bake_kernel<<<gridp((long long)N * P, 256, 148 * 16), 256, 0, stream>>>(edited, exposure, P, N);
  TCL_CHECK_LAUNCH("tcl_exposure_bake");
  return TCL_OK;
}

// Kernel-level test hooks (used by tests/test_postopt_gpu.py): one SSIM level in isolation.
//   X, Y : [planes, h, w] fp32 (Y is indexed as frame = plane/3, channel = plane%3)
//   sums : [planes][2] (cs sum, ssim sum) accumulated;  dX = d( sum_planes coef[p] * sum_map )/dX
extern "C" int tcl_debug_ssim_level(const float* X, const float* Y, int planes, int h, int w, const float* coef, int use_ssim,
                                    float* sums, float* dX, cudaStream_t stream) {
  TCL_CHECK_ARG(X && Y && planes > 0 && planes % 3 == 0 && planes / 3 <= MAXB && h >= 11 && w >= 11, "tcl_debug_ssim_level: args");
  int rc = ensure_gauss();
  if (rc) return rc;
  Batch bt; bt.n = planes / 3;
  for (int i = 0; i < MAXB; ++i) bt.idx[i] = i < bt.n ? i : 0;
  const float C1 = 1e-4f, C2 = 9e-4f;
  const long long hw = (long long)h * w;
  if (sums) {
    cudaMemsetAsync(sums, 0, sizeof(float) * planes * 2, stream);
    dim3 grid((w - 10 + ST - 1) / ST, (h - 10 + ST - 1) / ST, planes);
    ssim_fwd_kernel<<<grid, 256, 0, stream>>>(X, hw, Y, 3 * hw, hw, bt, h, w, C1, C2, sums);
    TCL_CHECK_LAUNCH("tcl_debug_ssim_level(fwd)");
  }
  if (dX) {
    TCL_CHECK_ARG(coef != nullptr, "tcl_debug_ssim_level: coef");
    const size_t bwd_smem = sizeof(float) * (2 * BR * BR + 5 * BP * BR + 3 * BP * BP);
    cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem);
    dim3 grid((w + BT - 1) / BT, (h + BT - 1) / BT, planes);
    ssim_bwd_kernel<<<grid, 256, bwd_smem, stream>>>(X, hw, Y, 3 * hw, hw, bt, h, w, C1, C2, coef, use_ssim, nullptr, 0, 0, 0, 0, 0, dX, hw);
    TCL_CHECK_LAUNCH("tcl_debug_ssim_level(bwd)");
  }
  return TCL_OK;
}

// ---- data-parallel variants: gradient accumulation only, then a separate Adam step ----
extern "C" int tcl_uvt_gradient(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids, long long U,
                                const float* fdc, float* grad, float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(ids && fdc && grad && U > 0, "tcl_uvt_gradient: null pointer");
  return run_iteration(2, ctx, idx_host, n_batch, ids, U, const_cast<float*>(fdc), grad, nullptr, nullptr, nullptr, nullptr, nullptr,
                       nullptr, 0.f, 0.9f, 0.999f, 1e-15f, 1, loss_out, stream, false);
}

extern "C" int tcl_exposure_gradient(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const float* exposure, float* grad,
                                     float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(exposure && grad, "tcl_exposure_gradient: null pointer");
  return run_iteration(1, ctx, idx_host, n_batch, nullptr, 0, nullptr, nullptr, nullptr, nullptr, const_cast<float*>(exposure), grad,
                       nullptr, nullptr, 0.f, 0.9f, 0.999f, 1e-8f, 1, loss_out, stream, false);
}

extern "C" int tcl_adam_step(float* p, float* grad, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                             int step, cudaStream_t stream) {
  TCL_CHECK_ARG(p && grad && m && v && n > 0 && step >= 1, "tcl_adam_step: args");
  LossAsm la;
  memset(&la, 0, sizeof(la));
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<gridp(n, 256, 148 * 16), 256, 0, stream>>>(p, grad, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, la);
  TCL_CHECK_LAUNCH("tcl_adam_step");
  return TCL_OK;
}

// Data-parallel stage 2: Adam over the UVT rows after the [U,4] gradient has been all-reduced.
extern "C" int tcl_adam_step_uvt(float* fdc, float* grad4, float* m, float* v, long long U, float lr, float beta1, float beta2,
                                 float eps, int step, cudaStream_t stream) {
  TCL_CHECK_ARG(fdc && grad4 && m && v && U > 0 && step >= 1, "tcl_adam_step_uvt: args");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_uvt_kernel<<<gridp(U / 4 + 1, 256, 148 * 16), 256, 0, stream>>>(fdc, grad4, m, v, U, lr, beta1, beta2, eps, bc1, bc2_sqrt);
  TCL_CHECK_LAUNCH("tcl_adam_step_uvt");
  return TCL_OK;
}
