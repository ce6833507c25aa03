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
// Relaxed-SSIM kernels over ALL pyramid levels 1..4 in one launch each (utils/loss_utils.py:73-123, the 11-tap
// sigma-1.5 valid separable Gaussian of pytorch_msssim: vertical pass first, then horizontal).
//   ssim_fwd_all : per (level, plane) sums of the cs / ssim maps, and the per-position partial derivatives
//                  d(map)/d(mu1, E[x^2], E[xy]) (un-scaled: the per-plane loss coefficient is only known once every
//                  level's mean exists, because MS-SSIM is a product over levels), stored for the backward pass.
//   ssim_bwd_all : own_l = coef * (G^T dmu1 + 2 x G^T de11 + y G^T de12): the gradient each level contributes to
//                  its own image; the avg-pool backward chain 0.25^k is summed where level 0 is consumed.
// Both filters run as register strips: a thread produces 8 vertically (4 horizontally) adjacent outputs from 18 (14)
// shared-memory reads instead of 11 reads per output.
// ------------------------------------------------------------------------------------------
constexpr int ST = 32;          // output tile
constexpr int SR = ST + 10;     // input tile
constexpr int SPITCH = SR + 1;  // shared-memory row pitch (odd: the horizontal strips are bank-conflict free)

struct LevelPlan {
  int h[5], w[5];               // image sizes of levels 1..4 (index = level)
  long long off[5];             // element offset of level l inside a pyramid plane
  int tiles_x[5], tiles_y[5];   // tile grid of the launch (map domain for fwd, image domain for bwd)
  int first[6];                 // first block of level l (first[5] = total)
  int planes;
};

// The 11-tap filters run on packed fp32x2 FMAs (Blackwell FFMA2: two outputs per issue slot — these kernels are bound by
// FMA issue, not by memory).  Outputs are handled in adjacent pairs (o, o+1): an input at distance k from output o is at
// distance k-1 from output o+1, so the weight pair is gp[k] = (g[k], g[k-1]) with g[-1] = g[11] = 0.  Each output still
// accumulates its 11 products in ascending tap order, exactly like the scalar loop.
struct GaussPairs { float2 gp[12]; };
__device__ __forceinline__ void load_gauss_pairs(GaussPairs& G) {
#pragma unroll
  for (int k = 0; k < 12; ++k) G.gp[k] = make_float2(k <= 10 ? c_gauss[k] : 0.f, k >= 1 ? c_gauss[k - 1] : 0.f);
}

// vertical pass of one thread: 8 outputs (rows strip*8 .. +7) of NM maps at column cc from 18 input rows.
//   val(i, v[NM]) produces the NM map values of input row i;  dst[m][row][cc] receives the outputs.
template <int NM, typename F>
__device__ __forceinline__ void vstrip8(const GaussPairs& G, F val, float (*dst)[ST][SPITCH], int strip, int cc) {
  float2 acc[4][NM];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[q][m] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 18; ++i) {
    float v[NM];
    val(i, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = i - 2 * q;             // distance to the pair's first output
      if (k >= 0 && k <= 11) {
#pragma unroll
        for (int m = 0; m < NM; ++m) acc[q][m] = __ffma2_rn(G.gp[k], make_float2(v[m], v[m]), acc[q][m]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int m = 0; m < NM; ++m) { dst[m][strip * 8 + 2 * q][cc] = acc[q][m].x; dst[m][strip * 8 + 2 * q + 1][cc] = acc[q][m].y; }
}

// horizontal pass: out[o] = sum_k g[k] src_row[o + k], o < 4
__device__ __forceinline__ void hrow4(const GaussPairs& G, const float* src_row, float (&out)[4]) {
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 14; ++t) {
    const float v = src_row[t];
    const float2 vv = make_float2(v, v);
    if (t <= 11) a0 = __ffma2_rn(G.gp[t], vv, a0);
    if (t >= 2) a1 = __ffma2_rn(G.gp[t - 2], vv, a1);
  }
  out[0] = a0.x; out[1] = a0.y; out[2] = a1.x; out[3] = a1.y;
}

__device__ __forceinline__ bool decode_block(const LevelPlan& lp, int& l, int& plane, int& ty0, int& tx0) {
  const int bid = blockIdx.x;
  l = 1;
#pragma unroll
  for (int i = 2; i <= 4; ++i)
    if (bid >= lp.first[i]) l = i;
  const int rem = bid - lp.first[l];
  const int per = lp.tiles_x[l] * lp.tiles_y[l];
  plane = rem / per;
  const int tile = rem - plane * per;
  ty0 = (tile / lp.tiles_x[l]) * ST;
  tx0 = (tile - (tile / lp.tiles_x[l]) * lp.tiles_x[l]) * ST;
  return true;
}

// Target-side maps, constant over the optimisation (utils/loss_utils.py:104-113: mu2 = G*y, E[y^2] = G*(y*y)): computed once
// per frame by tcl_postopt_build_pyramid and stored next to the pyramid, at the pyramid's offsets.
__global__ void __launch_bounds__(256)
ssim_target_kernel(float* __restrict__ Y, long long plane_stride, long long per, LevelPlan lp) {
  __shared__ float sy[SR][SPITCH];
  __shared__ float v[2][ST][SPITCH];
  GaussPairs G;
  load_gauss_pairs(G);
  int l, plane, ty0, tx0;
  decode_block(lp, l, plane, ty0, tx0);
  const int h = lp.h[l], w = lp.w[l];
  float* yp = Y + plane * plane_stride + lp.off[l];
  const int oh = h - 10, ow = w - 10;
  for (int i = threadIdx.x; i < SR * SR; i += blockDim.x) {
    const int r = i / SR, cc = i - r * SR;
    const int y = ty0 + r, x = tx0 + cc;
    sy[r][cc] = (y < h && x < w) ? yp[(long long)y * w + x] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * SR; t += blockDim.x) {
    const int strip = t / SR, cc = t - strip * SR;
    vstrip8<2>(G, [&](int i, float (&o)[2]) { const float yv = sy[strip * 8 + i][cc]; o[0] = yv; o[1] = yv * yv; }, v, strip, cc);
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 4;
  float f[2][4];
  hrow4(G, &v[0][r][c0], f[0]);
  hrow4(G, &v[1][r][c0], f[1]);
  const int py_ = ty0 + r;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int px_ = tx0 + c0 + o;
    if (py_ < oh && px_ < ow) {
      yp[per + (long long)py_ * w + px_] = f[0][o];
      yp[2 * per + (long long)py_ * w + px_] = f[1][o];
    }
  }
}

__global__ void __launch_bounds__(256)
ssim_fwd_all_kernel(const float* __restrict__ X, long long x_plane_stride, const float* __restrict__ Y, long long y_frame_stride,
                    long long y_chan_stride, long long y_per, Batch bt, LevelPlan lp, float C1, float C2,
                    float* __restrict__ sums /*[4][planes][2]*/, float* __restrict__ pm /*[planes][3][pyr]*/, long long pm_map_stride) {
  __shared__ float sxy[2][SR][SPITCH];
  __shared__ float v[3][ST][SPITCH];
  __shared__ float red[2][8];
  GaussPairs G;
  load_gauss_pairs(G);
  int l, plane, ty0, tx0;
  decode_block(lp, l, plane, ty0, tx0);
  const int h = lp.h[l], w = lp.w[l];
  const int b = plane / 3, c = plane - 3 * b;
  const float* xp = X + plane * x_plane_stride + lp.off[l];
  const float* yp = Y + bt.idx[b] * y_frame_stride + c * y_chan_stride + lp.off[l];
  const int oh = h - 10, ow = w - 10;     // valid map size
  for (int i = threadIdx.x; i < SR * SR; i += blockDim.x) {
    const int r = i / SR, cc = i - r * SR;
    const int y = ty0 + r, x = tx0 + cc;
    const bool in = y < h && x < w;
    sxy[0][r][cc] = in ? xp[(long long)y * w + x] : 0.f;
    sxy[1][r][cc] = in ? yp[(long long)y * w + x] : 0.f;
  }
  __syncthreads();
  // vertical (H) pass first, as pytorch_msssim.gaussian_filter does: maps x, x^2, xy (the target-side maps are precomputed)
  for (int t = threadIdx.x; t < 4 * SR; t += blockDim.x) {
    const int strip = t / SR, cc = t - strip * SR;
    vstrip8<3>(G, [&](int i, float (&o)[3]) {
      const float xv = sxy[0][strip * 8 + i][cc], yv = sxy[1][strip * 8 + i][cc];
      o[0] = xv; o[1] = xv * xv; o[2] = xv * yv;
    }, v, strip, cc);
  }
  __syncthreads();
  float acc_cs = 0.f, acc_ss = 0.f;
  {
    const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 4;
    float f[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m) hrow4(G, &v[m][r][c0], f[m]);
    const int py_ = ty0 + r;
    float* pm0 = pm + (long long)plane * 3 * pm_map_stride + lp.off[l] + (long long)py_ * w;
    const float* ym = yp + y_per + (long long)py_ * w;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px_ = tx0 + c0 + o;
      if (py_ < oh && px_ < ow) {
        const float mu1 = f[0][o], e11 = f[1][o], e12 = f[2][o];
        const float mu2 = ym[px_], e22 = ym[y_per + px_];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float A1 = 2.f * mu12 + C1, B1 = mu1_sq + mu2_sq + C1;
        const float s11 = e11 - mu1_sq, s22 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A2 = 2.f * s12 + C2, B2 = s11 + s22 + C2;
        const float cs = A2 / B2;
        const float lum = A1 / B1;
        acc_cs += cs; acc_ss += lum * cs;
        // d cs / d(e12, e11, mu1)
        const float dcs_e12 = 2.f / B2;
        const float dcs_e11 = -A2 / (B2 * B2);
        const float dcs_mu1 = -2.f * mu2 / B2 + 2.f * mu1 * A2 / (B2 * B2);
        float dmu1, de11, de12;
        if (l == 4) {           // the coarsest level contributes ssim = l * cs, the others cs only
          const float dl_mu1 = (2.f * mu2 * B1 - A1 * 2.f * mu1) / (B1 * B1);
          dmu1 = cs * dl_mu1 + lum * dcs_mu1; de11 = lum * dcs_e11; de12 = lum * dcs_e12;
        } else {
          dmu1 = dcs_mu1; de11 = dcs_e11; de12 = dcs_e12;
        }
        pm0[px_] = dmu1; pm0[pm_map_stride + px_] = de11; pm0[2 * pm_map_stride + px_] = de12;
      }
    }
  }
  acc_cs = warp_sum(acc_cs); acc_ss = warp_sum(acc_ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc_cs; red[1][threadIdx.x >> 5] = acc_ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; bsum += red[1][i]; }
    float* dst = sums + ((size_t)(l - 1) * lp.planes + plane) * 2;
    atomicAdd(&dst[0], a);
    atomicAdd(&dst[1], bsum);
  }
}

__global__ void __launch_bounds__(256)
ssim_bwd_all_kernel(const float* __restrict__ X, long long x_plane_stride, const float* __restrict__ Y, long long y_frame_stride,
                    long long y_chan_stride, Batch bt, LevelPlan lp, const float* __restrict__ coef /*[4][planes]*/,
                    const float* __restrict__ pm, long long pm_map_stride, float* __restrict__ own, long long own_plane_stride) {
  __shared__ float sp[3][SR][SPITCH];
  __shared__ float tt[3][ST][SPITCH];
  GaussPairs G;
  load_gauss_pairs(G);
  int l, plane, qy0, qx0;
  decode_block(lp, l, plane, qy0, qx0);
  const int h = lp.h[l], w = lp.w[l];
  const int b = plane / 3, c = plane - 3 * b;
  const int oh = h - 10, ow = w - 10;
  const float* pm0 = pm + (long long)plane * 3 * pm_map_stride + lp.off[l];
  // map positions [qy0 - 10, qy0 + 31] x [qx0 - 10, qx0 + 31]
  for (int i = threadIdx.x; i < 3 * SR * SR; i += blockDim.x) {
    const int m = i / (SR * SR);
    const int rem = i - m * SR * SR;
    const int r = rem / SR, cc = rem - r * SR;
    const int py_ = qy0 - 10 + r, px_ = qx0 - 10 + cc;
    sp[m][r][cc] = (py_ >= 0 && py_ < oh && px_ >= 0 && px_ < ow) ? pm0[m * pm_map_stride + (long long)py_ * w + px_] : 0.f;
  }
  __syncthreads();
  // transposed filters = the same correlations (the window is symmetric): dX(y) = sum_k g[k] pm(y - k)
  for (int t = threadIdx.x; t < 4 * SR; t += blockDim.x) {
    const int strip = t / SR, cc = t - strip * SR;
    vstrip8<3>(G, [&](int i, float (&o)[3]) { o[0] = sp[0][strip * 8 + i][cc]; o[1] = sp[1][strip * 8 + i][cc]; o[2] = sp[2][strip * 8 + i][cc]; },
               tt, strip, cc);
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 4;
  const int y = qy0 + r;
  if (y < h) {
    float f[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m) hrow4(G, &tt[m][r][c0], f[m]);
    const float g_coef = coef[(size_t)(l - 1) * lp.planes + plane];
    const float* xp = X + plane * x_plane_stride + lp.off[l] + (long long)y * w;
    const float* yp = Y + bt.idx[b] * y_frame_stride + c * y_chan_stride + lp.off[l] + (long long)y * w;
    float* op = own + plane * own_plane_stride + lp.off[l] + (long long)y * w;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int x = qx0 + c0 + o;
      if (x < w) op[x] = g_coef * (f[0][o] + 2.f * xp[x] * f[1][o] + yp[x] * f[2][o]);
    }
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
// producers of X (batch frames then predecessors)
// ------------------------------------------------------------------------------------------
// UVT rows live in per-rank shards: row id -> (owner rank, local row).  With one rank the table has a
// single entry.  Every pointer of the table is addressable from this GPU: the local shard, and the peers' shards mapped
// through CUDA IPC over NVLink, so a gather is a plain (peer) load and a gradient scatter a plain (peer) reduction.
struct Shards {
  float* fdc[TCL_MAX_RANKS];      // [rows_per_rank, 3]
  float* grad[TCL_MAX_RANKS];     // [rows_per_rank, 4]
  int world;
  int rows_per_rank;
  float inv_rows;                 // 1 / rows_per_rank
};

template <bool W1>
__device__ __forceinline__ void shard_of(const Shards& sh, int id, int& owner, int& local) {
  if (W1) { owner = 0; local = id; return; }
  int o = __float2int_rz(__int2float_rz(id) * sh.inv_rows);
  int lo = id - o * sh.rows_per_rank;
  if (lo < 0) { --o; lo += sh.rows_per_rank; }
  else if (lo >= sh.rows_per_rank) { ++o; lo -= sh.rows_per_rank; }
  owner = o; local = lo;
}

// X layout: one float4 per pixel, [2n][P] = {R, G, B, flags} with the three clamped channels and, in lane 3, the bits of
// an int whose bit c says "the clamp passed the gradient of channel c" (0 <= v <= 1).  A bicubic tap, a TV neighbour or a
// pixel's own value is then ONE 16-byte load, and the tap's clamp flags come with it.
__device__ __forceinline__ float4 pack_px(float r, float g, float b, unsigned fl) { return make_float4(r, g, b, __uint_as_float(fl)); }

// stage 2: X = clamp(SH2RGB(fdc[id]), 0, 1).  One thread per 2x2 pixel quad (8-byte id loads, 32-byte X stores); for the
// batch frames the quad mean is the level-1 pyramid value (avg_pool2d of an even-sized image has no padding), written
// in the same pass.
template <bool W1>
__global__ void __launch_bounds__(256)
uvt_gather_quad_kernel(Shards sh, const int* __restrict__ ids, int H, int W, Batch bt, float4* __restrict__ X,
                       float* __restrict__ xpyr1, long long pyr_plane_stride) {
  const int h2 = H >> 1, w2 = W >> 1;
  const long long P = (long long)H * W;
  const long long quads = (long long)h2 * w2;
  const long long total = (long long)2 * bt.n * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / quads);
    const int qd = (int)(i - (long long)f * quads);
    const int qy = qd / w2, qx = qd - qy * w2;
    int fr = f < bt.n ? bt.idx[f] : bt.idx[f - bt.n] - 1;
    if (fr < 0) fr = 0;
    const long long p0 = (long long)(2 * qy) * W + 2 * qx;
    const int2 ia = *reinterpret_cast<const int2*>(ids + (long long)fr * P + p0);
    const int2 ib = *reinterpret_cast<const int2*>(ids + (long long)fr * P + p0 + W);
    const int id4[4] = {ia.x, ia.y, ib.x, ib.y};
    float val[4][3];
    unsigned fl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int owner, local;
      shard_of<W1>(sh, id4[k], owner, local);
      const float* row = sh.fdc[owner] + (long long)local * 3;
      fl[k] = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = row[c] * SH_C0 + 0.5f;
        if (v >= 0.f && v <= 1.f) fl[k] |= (1u << c);
        val[k][c] = fminf(fmaxf(v, 0.f), 1.f);
      }
    }
    float4* dst = X + (long long)f * P + p0;
    dst[0] = pack_px(val[0][0], val[0][1], val[0][2], fl[0]);
    dst[1] = pack_px(val[1][0], val[1][1], val[1][2], fl[1]);
    dst[W] = pack_px(val[2][0], val[2][1], val[2][2], fl[2]);
    dst[W + 1] = pack_px(val[3][0], val[3][1], val[3][2], fl[3]);
    if (f < bt.n) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        xpyr1[((long long)f * 3 + c) * pyr_plane_stride + (long long)qy * w2 + qx] =
            (val[0][c] + val[1][c] + val[2][c] + val[3][c]) * 0.25f;
    }
  }
}

// odd image sizes: one thread per pixel, level 1 is pooled by avgpool2_x4_kernel afterwards
template <bool W1>
__global__ void uvt_gather_kernel(Shards sh, const int* __restrict__ ids, long long P, Batch bt, float4* __restrict__ X) {
  const long long total = (long long)2 * bt.n * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    int fr = f < bt.n ? bt.idx[f] : bt.idx[f - bt.n] - 1;
    if (fr < 0) fr = 0;
    int owner, local;
    shard_of<W1>(sh, ids[(long long)fr * P + p], owner, local);
    const float* row = sh.fdc[owner] + (long long)local * 3;
    unsigned fl = 0;
    float val[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = row[c] * SH_C0 + 0.5f;
      if (v >= 0.f && v <= 1.f) fl |= (1u << c);
      val[c] = fminf(fmaxf(v, 0.f), 1.f);
    }
    X[i] = pack_px(val[0], val[1], val[2], fl);
  }
}

// stage 1: X_j = clamp(sum_k in_k E[k][j] + E[j][3], 0, 1), E = exposure[frame] (3x4 row-major)
__global__ void exposure_apply_kernel(const float* __restrict__ edited, const float* __restrict__ expo, long long P, Batch bt,
                                      float4* __restrict__ X) {
  const long long total = (long long)2 * bt.n * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / P);
    const long long p = i - (long long)f * P;
    int fr = f < bt.n ? bt.idx[f] : bt.idx[f - bt.n] - 1;
    if (fr < 0) fr = 0;
    const float* E = expo + (long long)fr * 12;
    const float* src = edited + (long long)fr * 3 * P + p;
    const float in0 = src[0], in1 = src[P], in2 = src[2 * P];
    unsigned fl = 0;
    float val[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = in0 * E[0 * 4 + j] + in1 * E[1 * 4 + j] + in2 * E[2 * 4 + j] + E[j * 4 + 3];
      if (v >= 0.f && v <= 1.f) fl |= (1u << j);
      val[j] = fminf(fmaxf(v, 0.f), 1.f);
    }
    X[i] = pack_px(val[0], val[1], val[2], fl);
  }
}

// level 0 -> level 1 of the batch frames' pyramid from the interleaved X (avg_pool2d, padding = size % 2)
__global__ void avgpool2_x4_kernel(const float4* __restrict__ X, long long P, int hi, int wi, int ph, int pw,
                                   float* __restrict__ out, long long out_stride, int ho, int wo, int frames) {
  const long long total = (long long)frames * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const int oy = (int)((i / wo) % ho);
    const int f = (int)(i / ((long long)wo * ho));
    const float4* src = X + f * P;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * oy - ph + dy, x = 2 * ox - pw + dx;
        if (y >= 0 && y < hi && x >= 0 && x < wi) { const float4 t = src[(long long)y * wi + x]; s0 += t.x; s1 += t.y; s2 += t.z; }
      }
    float* dst = out + (long long)f * 3 * out_stride + (long long)oy * wo + ox;
    dst[0] = s0 * 0.25f; dst[out_stride] = s1 * 0.25f; dst[2 * out_stride] = s2 * 0.25f;
  }
}

// d(loss)/dX at level 0 coming from MS-SSIM is the avg-pool backward chain of levels 1..4,
//   dX_0 = 1/4 up(own_1 + 1/4 up(own_2 + 1/4 up(own_3 + 1/4 up(own_4))))     (up = nearest, with the pooling padding);
// the bracket is evaluated once per level-1 pixel here (interleaved RGB, one 16-byte load per level-0 pixel later).
__global__ void own_collapse_kernel(const float* __restrict__ own, long long own_stride, Pyr py, int frames, float4* __restrict__ out) {
  const int h1 = py.h[1], w1 = py.w[1];
  const long long total = (long long)frames * h1 * w1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x1 = (int)(i % w1);
    const int y1 = (int)((i / w1) % h1);
    const int f = (int)(i / ((long long)w1 * h1));
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = own[((long long)f * 3 + c) * own_stride + py.off[1] + (long long)y1 * w1 + x1];
    int y = y1, x = x1;
    float wgt = 0.25f;
#pragma unroll
    for (int l = 2; l <= 4; ++l) {
      y = (y + py.ph[l - 1]) >> 1;
      x = (x + py.pw[l - 1]) >> 1;
      if (y >= py.h[l] || x >= py.w[l]) break;
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] += wgt * own[((long long)f * 3 + c) * own_stride + py.off[l] + (long long)y * py.w[l] + x];
      wgt *= 0.25f;
    }
    out[i] = make_float4(0.25f * g[0], 0.25f * g[1], 0.25f * g[2], 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// fused level-0 kernels over the batch frames
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
  int P;
  Batch bt;
  const float4* X;                // [2n][P] {R,G,B,flags}
  const float* flows;             // [N,2,P]
  const float* mask;              // [N,1,P]
  const float4* own1;             // [n][h1*w1] collapsed MS-SSIM gradient at level-1 resolution (own_collapse_kernel)
  int h1, w1, ph0, pw0;
  float k_flow;                   // lambda_flow / (n_valid*3*P)
  float k_tvh, k_tvw;             // lambda_tv*2/(count_h*n), lambda_tv*2/(count_w*n)
  float k_l1;                     // stage 1: (1-lambda_flow)*(1-lambda_dssim)/(n*3*P); 0 in stage 2
  const float* edited;            // [N,3,P] (stage 1 L1 target and affine input)
  float* G_pre;                   // stage 1: [n,P,4] atomically accumulated predecessor gradient (lane 3 unused)
  float* scal;
  // sinks
  const int* ids;                 // stage 2: unq_inv [N*P]
  float* grad_expo;               // stage 1: [N,12]
};

struct Bicubic {
  float cx[4], cy[4];
  int x0, y0;
};
__device__ __forceinline__ void bicubic_setup(const L0Params& q, int fr, int p, int x, int y, Bicubic& bc) {
  const float* fl = q.flows + (long long)fr * 2 * q.P + p;
  const float fx = fl[0] + (float)x;
  const float fy = fl[q.P] + (float)y;
  const float gx = (fx / (float)(q.W - 1) - 0.5f) * 2.f;
  const float gy = (fy / (float)(q.H - 1) - 0.5f) * 2.f;
  const float ix = ((gx + 1.f) / 2.f) * (float)(q.W - 1);
  const float iy = ((gy + 1.f) / 2.f) * (float)(q.H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  cubic_coeffs(ix - fx0, bc.cx);
  cubic_coeffs(iy - fy0, bc.cy);
  bc.x0 = (int)fx0 - 1; bc.y0 = (int)fy0 - 1;
}
// 16 taps of 16 bytes; tf collects the clamp flags of the four column-0 taps, 3 bits per row at bit 3*j (0 for taps
// outside the image: they carry no gradient either)
__device__ __forceinline__ void bicubic_sample(const L0Params& q, const float4* __restrict__ Xp, const Bicubic& bc, float (&wv)[3],
                                               unsigned& tf) {
  wv[0] = wv[1] = wv[2] = 0.f;
  tf = 0u;
  if (bc.x0 >= 0 && bc.x0 + 3 < q.W && bc.y0 >= 0 && bc.y0 + 3 < q.H) {       // interior: no per-tap tests
    const float4* rowp = Xp + (bc.y0 * q.W + bc.x0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t0 = __ldg(rowp), t1 = __ldg(rowp + 1), t2 = __ldg(rowp + 2), t3 = __ldg(rowp + 3);
      rowp += q.W;
      float row[3];
      row[0] = t0.x * bc.cx[0]; row[1] = t0.y * bc.cx[0]; row[2] = t0.z * bc.cx[0];
      row[0] += t1.x * bc.cx[1]; row[1] += t1.y * bc.cx[1]; row[2] += t1.z * bc.cx[1];
      row[0] += t2.x * bc.cx[2]; row[1] += t2.y * bc.cx[2]; row[2] += t2.z * bc.cx[2];
      row[0] += t3.x * bc.cx[3]; row[1] += t3.y * bc.cx[3]; row[2] += t3.z * bc.cx[3];
      tf |= (__float_as_uint(t0.w) & 7u) << (3 * j);
#pragma unroll
      for (int c = 0; c < 3; ++c) wv[c] += row[c] * bc.cy[j];
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = bc.y0 + j;
    if (yy < 0 || yy >= q.H) continue;
    const float4* rowp = Xp + (yy * q.W + bc.x0);
    float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (bc.x0 + i < 0 || bc.x0 + i >= q.W) continue;
      const float4 t = __ldg(rowp + i);
      row[0] += t.x * bc.cx[i]; row[1] += t.y * bc.cx[i]; row[2] += t.z * bc.cx[i];
      if (i == 0) tf |= (__float_as_uint(t.w) & 7u) << (3 * j);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) wv[c] += row[c] * bc.cy[j];
  }
}

// ------------------------------------------------------------------------------------------
// Stage 2, fused level-0 kernel: bicubic flow warp forward + backward, masked L1, TV, MS-SSIM gradient, and the gradient
// sinks — every pixel's dLoss/dX goes straight into the UVT gradient row of its id (local or peer shard).
//
// The bicubic backward is 16 contributions per pixel.  A warp walks 29 consecutive pixels of one image row (lanes 3..31;
// lanes 0..2 re-evaluate the three pixels before them as helpers).  For a smooth flow neighbouring pixels hit neighbouring
// columns: tap i of lane l and tap 0 of lane l+i land on the same predecessor pixel, so lane m sums
// v0(m) + v1(m-1) + v2(m-2) + v3(m-3) with three independent shuffles per channel and row and issues ONE 16-byte
// reduction per row into the predecessor pixel's UVT row — 4 instead of 16 per pixel.  Where the chain is broken (flow
// discontinuity, image border) the producer reduces the tap itself; taps whose consumer lies in the next segment are left to
// that segment's helper lanes.
// ------------------------------------------------------------------------------------------
constexpr int SEG = 29;
template <bool W1>
__global__ void __launch_bounds__(256, 4)
level0_uvt_kernel(L0Params q, Shards sh, float inv_nseg) {
  __shared__ float red[3][8];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);     // warp-uniform for the compiler: the chunk loop is convergent
  const int b = blockIdx.y;
  const int fr = q.bt.idx[b];
  const bool valid = fr > 0;
  const float4* Xi = q.X + (long long)b * q.P;
  const float4* Xp = q.X + (long long)(q.bt.n + b) * q.P;
  const int* ids_cur = q.ids + (long long)fr * q.P;
  const int* ids_pre = q.ids + (long long)(fr > 0 ? fr - 1 : 0) * q.P;
  const float4* own1 = q.own1 + (long long)b * q.h1 * q.w1;
  float* Gp = q.G_pre + (long long)b * 4 * q.P;     // sharded runs only: local image of the predecessor frame's gradient
  float acc_flow = 0.f, acc_tvh = 0.f, acc_tvw = 0.f;

  auto sink = [&](int id, unsigned fl, float g0, float g1, float g2) {
    int owner, local;
    shard_of<W1>(sh, id, owner, local);
    red_add_v4(sh.grad[owner] + (long long)local * 4, (fl & 1) ? g0 * SH_C0 : 0.f, (fl & 2) ? g1 * SH_C0 : 0.f,
               (fl & 4) ? g2 * SH_C0 : 0.f);
  };

  // A block's 8 warps take the same 29-pixel segment of 8 CONSECUTIVE image rows: their bicubic footprints (4 predecessor
  // rows each) overlap by three quarters and the TV neighbours are each other's pixels, so the taps hit L1 instead of L2
  // (the kernel was L2-bandwidth bound: 5 GB of L2 traffic per iteration with row-major chunking).
  const int n_seg = (q.W + SEG - 1) / SEG;
  const int n_items = ((q.H + 7) >> 3) * n_seg;
  for (int bi = blockIdx.x; bi < n_items; bi += gridDim.x) {                        // block-uniform
    int grp = __float2int_rz(((float)bi + 0.5f) * inv_nseg);
    if (grp * n_seg > bi) --grp; else if ((grp + 1) * n_seg <= bi) ++grp;
    const int seg = bi - grp * n_seg;
    const int y = grp * 8 + warp;
    if (y >= q.H) continue;                                                          // warp-uniform
    const int x = seg * SEG - 3 + lane;
    const bool act = x >= 0 && x < q.W;
    const bool real = act && lane >= 3;
    const bool next_exists = (seg + 1) * SEG < q.W;
    const int p = y * q.W + x;
    float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) own = Xi[p];
    float g[3] = {0.f, 0.f, 0.f};
    // ---- flow term: warp the predecessor with the backward flow ----
    if (valid) {
      Bicubic bc;
      bc.x0 = bc.y0 = -(1 << 20);
#pragma unroll
      for (int i = 0; i < 4; ++i) bc.cx[i] = bc.cy[i] = 0.f;
      float s[3] = {0.f, 0.f, 0.f};
      unsigned tf = 0u;
      if (act) {
        bicubic_setup(q, fr, p, x, y, bc);
        float wv[3];
        bicubic_sample(q, Xp, bc, wv, tf);
        const float m = q.mask[(long long)fr * q.P + p];
        const float xi[3] = {own.x, own.y, own.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float d = wv[c] * m - xi[c] * m;
          const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          s[c] = sg * m * q.k_flow;
          if (real) { acc_flow += fabsf(d); g[c] -= s[c]; }
        }
      }
      // chain links: lane l continues lane l-1 iff their footprints are one column apart on the same rows
      const int x0l = __shfl_up_sync(FULL, bc.x0, 1), y0l = __shfl_up_sync(FULL, bc.y0, 1);
      const bool link = lane > 0 && bc.x0 == x0l + 1 && bc.y0 == y0l;
      const unsigned lb = __ballot_sync(FULL, link);
      // lane m accepts the tap-k value of lane m-k iff links m-k+1..m are intact (helpers own no column here)
      float accf[4];
      unsigned ab[4];
#pragma unroll
      for (int k = 1; k <= 3; ++k) {
        const unsigned need = ((1u << k) - 1u) << ((lane - k + 1) & 31);
        const bool a = real && lane >= k && (lb & need) == need;
        accf[k] = a ? 1.f : 0.f;
        ab[k] = __ballot_sync(FULL, a);
      }
      // tap i of this lane is left over (reduced by the producer) iff nobody accepts it; a real lane defers taps whose
      // consumer would sit in the next segment, a helper handles exactly the taps its pixel deferred there
      unsigned left = 0u;
#pragma unroll
      for (int i = 1; i <= 3; ++i) {
        const bool cons = lane + i <= 31 && ((ab[i] >> ((lane + i) & 31)) & 1u);
        if (act && !cons && (lane >= 3 ? (lane + i <= 31 || !next_exists) : (lane + i >= 3))) left |= 1u << i;
      }
      if (s[0] == 0.f && s[1] == 0.f && s[2] == 0.f) left = 0u;
      const bool main_on = real && tf != 0u;
      const int o0 = bc.y0 * q.W + bc.x0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wy = bc.cy[j];
        const unsigned fl = (tf >> (3 * j)) & 7u;
        int id = 0;
        if (W1 && main_on && fl) id = __ldg(ids_pre + o0 + j * q.W);
        float v[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float wgt = bc.cx[i] * wy;
#pragma unroll
          for (int c = 0; c < 3; ++c) v[i][c] = s[c] * wgt;
        }
        float out[3] = {v[0][0], v[0][1], v[0][2]};
#pragma unroll
        for (int k = 1; k <= 3; ++k)
#pragma unroll
          for (int c = 0; c < 3; ++c) out[c] = fmaf(__shfl_up_sync(FULL, v[k][c], k), accf[k], out[c]);
        if (main_on && fl) {
          if (W1) sink(id, fl, out[0], out[1], out[2]);
          else red_add_v4(Gp + (long long)(o0 + j * q.W) * 4, out[0], out[1], out[2]);     // sharded: see pre_sink_uvt_kernel
        }
        if (left) {                                  // rare: broken chain / image border
          const int yy = bc.y0 + j;
          if (yy >= 0 && yy < q.H) {
#pragma unroll
            for (int i = 1; i <= 3; ++i) {
              const int xx = bc.x0 + i;
              if (!((left >> i) & 1u) || xx < 0 || xx >= q.W) continue;
              const unsigned fli = __float_as_uint(__ldg(Xp + yy * q.W + xx).w) & 7u;
              if (!fli) continue;
              if (W1) sink(ids_pre[yy * q.W + xx], fli, v[i][0], v[i][1], v[i][2]);
              else red_add_v4(Gp + (long long)(yy * q.W + xx) * 4, v[i][0], v[i][1], v[i][2]);
            }
          }
        }
      }
    }
    if (real) {
    // ---- total variation ----
    if (q.k_tvh != 0.f) {
      const float xi[3] = {own.x, own.y, own.z};
      float gg[3] = {0.f, 0.f, 0.f};
      if (y + 1 < q.H) {
        const float4 t = __ldg(Xi + p + q.W);
        const float d[3] = {t.x - xi[0], t.y - xi[1], t.z - xi[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) { acc_tvh += d[c] * d[c]; gg[c] -= q.k_tvh * 2.f * d[c]; }
      }
      if (y > 0) {
        const float4 t = __ldg(Xi + p - q.W);
        gg[0] += q.k_tvh * 2.f * (xi[0] - t.x); gg[1] += q.k_tvh * 2.f * (xi[1] - t.y); gg[2] += q.k_tvh * 2.f * (xi[2] - t.z);
      }
      if (x + 1 < q.W) {
        const float4 t = __ldg(Xi + p + 1);
        const float d[3] = {t.x - xi[0], t.y - xi[1], t.z - xi[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) { acc_tvw += d[c] * d[c]; gg[c] -= q.k_tvw * 2.f * d[c]; }
      }
      if (x > 0) {
        const float4 t = __ldg(Xi + p - 1);
        gg[0] += q.k_tvw * 2.f * (xi[0] - t.x); gg[1] += q.k_tvw * 2.f * (xi[1] - t.y); gg[2] += q.k_tvw * 2.f * (xi[2] - t.z);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] += gg[c];
    }
    // ---- MS-SSIM gradient (collapsed chain at level-1 resolution) ----
    {
      const int y1 = (y + q.ph0) >> 1, x1 = (x + q.pw0) >> 1;
      if (y1 < q.h1 && x1 < q.w1) {
        const float4 t = __ldg(own1 + y1 * q.w1 + x1);
        g[0] += t.x; g[1] += t.y; g[2] += t.z;
      }
    }
    // ---- sink ----
    const unsigned fl = __float_as_uint(own.w) & 7u;
    if (fl) sink(ids_cur[p], fl, g[0], g[1], g[2]);
    }
  }
  acc_flow = warp_sum(acc_flow); acc_tvh = warp_sum(acc_tvh); acc_tvw = warp_sum(acc_tvw);
  if (lane == 0) { red[0][warp] = acc_flow; red[1][warp] = acc_tvh; red[2][warp] = acc_tvw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f, cc = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; bb += red[1][i]; cc += red[2][i]; }
    atomicAdd(&q.scal[SC_FLOW_ABS], a); atomicAdd(&q.scal[SC_TV_H], bb); atomicAdd(&q.scal[SC_TV_W], cc);
  }
}

// Sharded stage 2, predecessor half: under data parallelism most UVT rows live in a peer's memory and every reduction is
// an NVLink transaction, so the (up to 4 per pixel) predecessor contributions are first summed in the LOCAL image G_pre
// (local L2 reductions are nearly free) and leave as ONE 16-byte reduction per predecessor pixel: 2 instead of 5 remote
// reductions per pixel.  Clears G_pre for the next iteration.
__global__ void __launch_bounds__(256)
pre_sink_uvt_kernel(L0Params q, Shards sh) {
  const int b = blockIdx.y;
  const int fr = q.bt.idx[b];
  if (fr <= 0) return;
  float4* Gp = reinterpret_cast<float4*>(q.G_pre + (long long)b * 4 * q.P);
  const float4* Xpre = q.X + (long long)(q.bt.n + b) * q.P;
  const int* ids_pre = q.ids + (long long)(fr - 1) * q.P;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < q.P; p += gridDim.x * blockDim.x) {
    const float4 gv = Gp[p];
    if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f) continue;
    Gp[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned fl = __float_as_uint(Xpre[p].w) & 7u;
    if (!fl) continue;
    int owner, local;
    shard_of<false>(sh, ids_pre[p], owner, local);
    red_add_v4(sh.grad[owner] + (long long)local * 4, (fl & 1) ? gv.x * SH_C0 : 0.f, (fl & 2) ? gv.y * SH_C0 : 0.f,
               (fl & 4) ? gv.z * SH_C0 : 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// Stage 1, fused level-0 kernel over the batch frames (the exposure gradient is 12 numbers per frame: the bicubic
// backward is accumulated per predecessor pixel in G_pre and folded by pre_sink_kernel)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
level0_expo_kernel(L0Params q) {
  __shared__ float red[4][8];
  __shared__ float eg[12];
  const int b = blockIdx.y;
  const int fr = q.bt.idx[b];
  const bool valid = fr > 0;
  const float4* Xi = q.X + (long long)b * q.P;
  const float4* Xp = q.X + (long long)(q.bt.n + b) * q.P;
  const float4* own1 = q.own1 + (long long)b * q.h1 * q.w1;
  float acc_flow = 0.f, acc_l1 = 0.f;
  float ge[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) ge[k] = 0.f;
  if (threadIdx.x < 12) eg[threadIdx.x] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < q.P; p += gridDim.x * blockDim.x) {
    const int y = p / q.W, x = p - y * q.W;
    const float4 own = Xi[p];
    const float xi[3] = {own.x, own.y, own.z};
    float g[3] = {0.f, 0.f, 0.f};
    if (valid) {
      Bicubic bc;
      bicubic_setup(q, fr, p, x, y, bc);
      float wv[3];
      unsigned tf;
      bicubic_sample(q, Xp, bc, wv, tf);
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
          const int yy = bc.y0 + j;
          if (yy < 0 || yy >= q.H) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int xx = bc.x0 + i;
            if (xx < 0 || xx >= q.W) continue;
            const float wgt = bc.cx[i] * bc.cy[j];
            red_add_v4(Gp + ((long long)yy * q.W + xx) * 4, s[0] * wgt, s[1] * wgt, s[2] * wgt);
          }
        }
      }
    }
    // ---- L1 to the target ----
    float in[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      in[c] = q.edited[((long long)fr * 3 + c) * q.P + p];
      const float d = xi[c] - in[c];
      acc_l1 += fabsf(d);
      g[c] += q.k_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    {
      const int y1 = (y + q.ph0) >> 1, x1 = (x + q.pw0) >> 1;
      if (y1 < q.h1 && x1 < q.w1) {
        const float4 t = __ldg(own1 + y1 * q.w1 + x1);
        g[0] += t.x; g[1] += t.y; g[2] += t.z;
      }
    }
    const unsigned fl = __float_as_uint(own.w);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float gj = ((fl >> j) & 1) ? g[j] : 0.f;
      ge[0 * 4 + j] += gj * in[0]; ge[1 * 4 + j] += gj * in[1]; ge[2 * 4 + j] += gj * in[2];
      ge[j * 4 + 3] += gj;
    }
  }
  acc_flow = warp_sum(acc_flow); acc_l1 = warp_sum(acc_l1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc_flow; red[3][threadIdx.x >> 5] = acc_l1; }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float r = warp_sum(ge[k]);
    if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(&eg[k], r);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, dd = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; dd += red[3][i]; }
    atomicAdd(&q.scal[SC_FLOW_ABS], a); atomicAdd(&q.scal[SC_L1], dd);
  }
  if (threadIdx.x < 12) atomicAdd(&q.grad_expo[(long long)fr * 12 + threadIdx.x], eg[threadIdx.x]);
}

// stage 1, predecessor half: G_pre -> clamp mask -> exposure gradient of frame idx-1; clears G_pre for the next iteration
__global__ void __launch_bounds__(256)
pre_sink_kernel(L0Params q) {
  __shared__ float eg[12];
  const int b = blockIdx.y;
  int fr = q.bt.idx[b] - 1;
  if (fr < 0) fr = 0;
  float ge[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) ge[k] = 0.f;
  if (threadIdx.x < 12) eg[threadIdx.x] = 0.f;
  float4* Gp = reinterpret_cast<float4*>(q.G_pre + (long long)b * 4 * q.P);
  const float4* Xpre = q.X + (long long)(q.bt.n + b) * q.P;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < q.P; p += gridDim.x * blockDim.x) {
    const float4 gv = Gp[p];
    if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f) continue;
    Gp[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float g[3] = {gv.x, gv.y, gv.z};
    const unsigned fl = __float_as_uint(Xpre[p].w);
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
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float r = warp_sum(ge[k]);
    if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(&eg[k], r);
  }
  __syncthreads();
  if (threadIdx.x < 12) atomicAdd(&q.grad_expo[(long long)fr * 12 + threadIdx.x], eg[threadIdx.x]);
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

// Stage-2 Adam over the UVT rows: p, m, v are [U,3], the gradient is [U,4] (lane 3 padding, never read or written here:
// the reductions only ever add 0 to it).  One thread handles one float4 of the flat p/m/v arrays, so the three big streams
// are perfectly coalesced 16-byte accesses; its four gradient values (rows 4a + {0,0,0,1 | 1,1,2,2 | 2,3,3,3}) are scalar
// loads from a span the warp shares.
__global__ void __launch_bounds__(256)
adam_uvt_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long U,
                float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  const long long quads = U / 4;           // groups of 4 rows = 3 float4 of p/m/v = 16 floats of g
  const long long n4 = quads * 3;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n4; f += (long long)gridDim.x * blockDim.x) {
    const long long a = f / 3;
    const int b = (int)(f - a * 3);
    // gradient offsets inside the 16-float group: element e = 4b + k -> row e / 3, channel e % 3 -> 4 * row + channel
    const int o0 = b == 0 ? 0 : (b == 1 ? 5 : 10);
    const int o1 = b == 0 ? 1 : (b == 1 ? 6 : 12);
    const int o2 = b == 0 ? 2 : (b == 1 ? 8 : 13);
    const int o3 = b == 0 ? 4 : (b == 1 ? 9 : 14);
    float* gg = g + a * 16;
    const float g0 = gg[o0], g1 = gg[o1], g2 = gg[o2], g3 = gg[o3];
    float4 pa = reinterpret_cast<float4*>(p)[f], ma = reinterpret_cast<float4*>(m)[f], va = reinterpret_cast<float4*>(v)[f];
    upd(pa.x, g0, ma.x, va.x); upd(pa.y, g1, ma.y, va.y); upd(pa.z, g2, ma.z, va.z); upd(pa.w, g3, ma.w, va.w);
    reinterpret_cast<float4*>(p)[f] = pa; reinterpret_cast<float4*>(m)[f] = ma; reinterpret_cast<float4*>(v)[f] = va;
    gg[o0] = 0.f; gg[o1] = 0.f; gg[o2] = 0.f; gg[o3] = 0.f;
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
  float4* X; float* G_pre; float* xpyr; float* pm; float* own; float4* own1; float* sums; float* coef; float* scal;
};
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t ws_layout(void* base, int H, int W, int nb, Ws* w) {
  Pyr py; make_pyr(H, W, &py);
  const size_t P = (size_t)H * W;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += align_up(bytes); return r; };
  Ws t;
  t.X = (float4*)take(sizeof(float4) * 2 * nb * P);              // {R,G,B,flags} per pixel, batch frames then predecessors
  t.G_pre = (float*)take(sizeof(float) * nb * 4 * P);             // stage 1 only: [nb,P,4], zero between iterations
  t.xpyr = (float*)take(sizeof(float) * nb * 3 * py.total);       // X pyramid levels 1..4
  t.pm = (float*)take(sizeof(float) * nb * 3 * 3 * py.total);     // d(map)/d(mu1, e11, e12) per level
  t.own = (float*)take(sizeof(float) * nb * 3 * py.total);        // per-level MS-SSIM gradient
  t.own1 = (float4*)take(sizeof(float4) * nb * (size_t)py.h[1] * py.w[1]);   // its pooled-down chain at level-1 resolution
  t.sums = (float*)take(sizeof(float) * (4 * nb * 3 * 2 + SC_COUNT));   // level sums, then the scalars (one memset)
  t.scal = t.sums + 4 * nb * 3 * 2;
  t.coef = (float*)take(sizeof(float) * 4 * nb * 3);
  if (w) *w = t;
  return (size_t)(p - reinterpret_cast<uint8_t*>(base));
}
static size_t ws_bytes(int H, int W, int nb) { return ws_layout(nullptr, H, W, nb, nullptr); }

static void make_plan(const Pyr& py, int planes, bool map_domain, int only_level, LevelPlan* lp) {
  int first = 0;
  for (int l = 1; l <= 4; ++l) {
    lp->h[l] = py.h[l]; lp->w[l] = py.w[l]; lp->off[l] = py.off[l];
    const int dh = map_domain ? py.h[l] - 10 : py.h[l], dw = map_domain ? py.w[l] - 10 : py.w[l];
    lp->tiles_x[l] = (dw + ST - 1) / ST; lp->tiles_y[l] = (dh + ST - 1) / ST;
    lp->first[l] = first;
    if (only_level == 0 || only_level == l) first += lp->tiles_x[l] * lp->tiles_y[l] * planes;
  }
  lp->h[0] = lp->w[0] = 0; lp->off[0] = 0; lp->tiles_x[0] = lp->tiles_y[0] = 0; lp->first[0] = 0;
  lp->first[5] = first;
  lp->planes = planes;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_postopt_workspace_bytes(int H, int W, int max_batch) { return ws_bytes(H, W, max_batch); }

extern "C" long long tcl_postopt_pyramid_elems(int H, int W) { Pyr py; make_pyr(H, W, &py); return py.total; }
extern "C" long long tcl_postopt_target_elems(int H, int W) { Pyr py; make_pyr(H, W, &py); return 3 * py.total; }

// Target side of the photometric loss, built once: ypyr[frame][channel] = { levels 1..4 of the avg-pool pyramid | mu2 = G*y per
// level | E[y^2] = G*(y*y) per level }, each tcl_postopt_pyramid_elems long (tcl_postopt_target_elems in total).
extern "C" int tcl_postopt_build_pyramid(const float* edited, int N, int H, int W, float* ypyr, cudaStream_t stream) {
  TCL_CHECK_ARG(edited && ypyr && N > 0, "tcl_postopt_build_pyramid: args");
  Pyr py; make_pyr(H, W, &py);
  TCL_CHECK_ARG(py.h[4] >= 11 && py.w[4] >= 11, "tcl_postopt_build_pyramid: image too small for 5-level MS-SSIM (%dx%d)", H, W);
  int rc = ensure_gauss();
  if (rc) return rc;
  const long long P = (long long)H * W;
  const long long stride = 3 * py.total;
  for (int l = 0; l < 4; ++l) {
    const float* in = l == 0 ? edited : ypyr + py.off[l];
    const long long in_stride = l == 0 ? P : stride;
    const long long total = (long long)N * 3 * py.h[l + 1] * py.w[l + 1];
    avgpool2_kernel<<<gridp(total, 256), 256, 0, stream>>>(in, in_stride, py.h[l], py.w[l], py.ph[l], py.pw[l],
                                                           ypyr + py.off[l + 1], stride, py.h[l + 1], py.w[l + 1], N * 3);
    TCL_CHECK_LAUNCH("tcl_postopt_build_pyramid");
  }
  // a launch handles at most MAXB frames' worth of planes so that the block count stays well inside the grid limit
  for (int f0 = 0; f0 < N; f0 += MAXB) {
    const int nf = N - f0 < MAXB ? N - f0 : MAXB;
    LevelPlan lp;
    make_plan(py, nf * 3, true, 0, &lp);
    ssim_target_kernel<<<lp.first[5], 256, 0, stream>>>(ypyr + (long long)f0 * 3 * stride, stride, py.total, lp);
    TCL_CHECK_LAUNCH("tcl_postopt_build_pyramid(target maps)");
  }
  return TCL_OK;
}

static int check_shards(const tcl_uvt_shards* s, const char* who) {
  TCL_CHECK_ARG(s && s->world >= 1 && s->world <= TCL_MAX_RANKS && s->rank >= 0 && s->rank < s->world && s->rows_per_rank > 0,
                "%s: bad shard table (world %d, rank %d)", who, s ? s->world : -1, s ? s->rank : -1);
  TCL_CHECK_ARG(s->world == 1 || (s->rows_per_rank >= 256 && s->rows_per_rank % 4 == 0),
                "%s: rows_per_rank must be a multiple of 4 and >= 256 when world > 1", who);
  for (int r = 0; r < s->world; ++r) TCL_CHECK_ARG(s->fdc[r] && s->grad[r], "%s: null shard pointer (rank %d)", who, r);
  return TCL_OK;
}

// One optimiser iteration.  stage 2: `shards` names where the UVT rows and their gradient live (one entry on a single
// GPU); stage 1: expo/egrad.  `adam`: 0 = gradient only, 1 = gradient + Adam over the LOCAL shard / the exposure table.
static int run_iteration(int stage, const tcl_postopt_ctx* c, const int* idx_host, int nb,
                         const int* ids, const tcl_uvt_shards* shards, float* m, float* v,
                         float* expo, float* egrad, float* em, float* ev,
                         float lr, float beta1, float beta2, float eps, int step, float* loss_out, cudaStream_t stream,
                         bool do_adam) {
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
  TCL_CHECK_ARG(P < (1ll << 30), "postopt: frames of %lld pixels are not supported (32-bit pixel offsets)", P);
  Ws w; ws_layout(c->workspace, H, W, c->max_batch, &w);   // fixed layout: G_pre must stay zero between iterations
  Batch bt; bt.n = nb;
  int n_valid = 0;
  for (int i = 0; i < nb; ++i) {
    TCL_CHECK_ARG(idx_host[i] >= 0 && idx_host[i] < c->N, "postopt: frame index %d out of range", idx_host[i]);
    bt.idx[i] = idx_host[i];
    n_valid += idx_host[i] > 0;
  }
  for (int i = nb; i < MAXB; ++i) bt.idx[i] = 0;
  const int planes = nb * 3;
  Shards sh;
  memset(&sh, 0, sizeof(sh));
  bool w1 = true;
  if (stage == 2) {
    sh.world = shards->world; sh.rows_per_rank = (int)shards->rows_per_rank; sh.inv_rows = 1.f / (float)shards->rows_per_rank;
    for (int r = 0; r < shards->world; ++r) { sh.fdc[r] = shards->fdc[r]; sh.grad[r] = shards->grad[r]; }
    w1 = shards->world == 1;
  }
  cudaMemsetAsync(w.sums, 0, sizeof(float) * (4 * c->max_batch * 3 * 2 + SC_COUNT), stream);
  // 1. produce X (and, for even sizes in stage 2, level 1 of the pyramid in the same pass)
  int first_pool = 0;
  if (stage == 2) {
    if (H % 2 == 0 && W % 2 == 0) {
      const long long quads = (long long)2 * nb * (P / 4);
      if (w1) uvt_gather_quad_kernel<true><<<gridp(quads, 256, 148 * 16), 256, 0, stream>>>(sh, ids, H, W, bt, w.X, w.xpyr + py.off[1], py.total);
      else uvt_gather_quad_kernel<false><<<gridp(quads, 256, 148 * 16), 256, 0, stream>>>(sh, ids, H, W, bt, w.X, w.xpyr + py.off[1], py.total);
      first_pool = 1;
    } else {
      if (w1) uvt_gather_kernel<true><<<gridp(2 * nb * P, 256, 148 * 16), 256, 0, stream>>>(sh, ids, P, bt, w.X);
      else uvt_gather_kernel<false><<<gridp(2 * nb * P, 256, 148 * 16), 256, 0, stream>>>(sh, ids, P, bt, w.X);
    }
  } else {
    exposure_apply_kernel<<<gridp(2 * nb * P, 256, 148 * 16), 256, 0, stream>>>(c->edited, expo, P, bt, w.X);
  }
  TCL_CHECK_LAUNCH("postopt(produce)");
  // 2. pyramid of the batch frames
  if (first_pool == 0) {
    avgpool2_x4_kernel<<<gridp((long long)nb * py.h[1] * py.w[1], 256), 256, 0, stream>>>(w.X, P, H, W, py.ph[0], py.pw[0],
                                                                                        w.xpyr + py.off[1], py.total, py.h[1], py.w[1], nb);
    TCL_CHECK_LAUNCH("postopt(pyramid)");
  }
  for (int l = 1; l < 4; ++l) {
    const long long total = (long long)planes * py.h[l + 1] * py.w[l + 1];
    avgpool2_kernel<<<gridp(total, 256), 256, 0, stream>>>(w.xpyr + py.off[l], py.total, py.h[l], py.w[l], py.ph[l], py.pw[l],
                                                           w.xpyr + py.off[l + 1], py.total, py.h[l + 1], py.w[l + 1], planes);
    TCL_CHECK_LAUNCH("postopt(pyramid)");
  }
  // 3. relaxed-SSIM forward, all levels in one launch
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;   // data_range = 1
  LevelPlan lpf, lpb;
  make_plan(py, planes, true, 0, &lpf);
  make_plan(py, planes, false, 0, &lpb);
  ssim_fwd_all_kernel<<<lpf.first[5], 256, 0, stream>>>(w.xpyr, py.total, c->ypyr, 9 * py.total, 3 * py.total, py.total, bt, lpf, C1, C2,
                                                        w.sums, w.pm, py.total);
  TCL_CHECK_LAUNCH("postopt(ssim_fwd)");
  // 4. loss head
  const int planes_g = nb_g * 3;
  const float k_ms = -(1.f - c->lambda_flow) * c->lambda_dssim / (float)planes_g;
  msssim_head_kernel<<<1, 64, 0, stream>>>(w.sums, planes, py, k_ms, w.coef, w.scal);
  TCL_CHECK_LAUNCH("postopt(head)");
  // 5. relaxed-SSIM backward, all levels in one launch
  ssim_bwd_all_kernel<<<lpb.first[5], 256, 0, stream>>>(w.xpyr, py.total, c->ypyr, 9 * py.total, 3 * py.total, bt, lpb, w.coef, w.pm,
                                                        py.total, w.own, py.total);
  TCL_CHECK_LAUNCH("postopt(ssim_bwd)");
  own_collapse_kernel<<<gridp((long long)nb * py.h[1] * py.w[1], 256), 256, 0, stream>>>(w.own, py.total, py, nb, w.own1);
  TCL_CHECK_LAUNCH("postopt(collapse)");
  // 6. fused level-0 kernel (+ stage 1: predecessor sink)
  L0Params q;
  memset(&q, 0, sizeof(q));
  q.H = H; q.W = W; q.P = (int)P; q.bt = bt; q.X = w.X; q.flows = c->past_flows; q.mask = c->mask_bwd;
  q.own1 = w.own1; q.h1 = py.h[1]; q.w1 = py.w[1]; q.ph0 = py.ph[0]; q.pw0 = py.pw[0];
  const int n_valid_g = c->norm_batch > 0 ? c->norm_valid : n_valid;
  const float flow_cnt = (float)n_valid_g * 3.f * (float)P;
  q.k_flow = n_valid_g > 0 ? c->lambda_flow / flow_cnt : 0.f;
  const float count_h = 3.f * (float)(H - 1) * (float)W, count_w = 3.f * (float)H * (float)(W - 1);
  if (stage == 2) { q.k_tvh = c->lambda_tv * 2.f / (count_h * nb_g); q.k_tvw = c->lambda_tv * 2.f / (count_w * nb_g); }
  q.k_l1 = stage == 1 ? (1.f - c->lambda_flow) * (1.f - c->lambda_dssim) / ((float)planes_g * (float)P) : 0.f;
  q.edited = c->edited; q.G_pre = w.G_pre; q.scal = w.scal; q.ids = ids; q.grad_expo = egrad;
  dim3 g0(gridp(P, 256, 148 * 2), nb);
  if (stage == 2) {
    const long long items = (long long)((H + 7) / 8) * ((W + SEG - 1) / SEG);      // one item = 8 rows x one segment = one block pass
    dim3 gu(gridp(items * 256, 256, 148 * 2), nb);
    const float inv_nseg = 1.f / (float)((W + SEG - 1) / SEG);
    if (w1) level0_uvt_kernel<true><<<gu, 256, 0, stream>>>(q, sh, inv_nseg);
    else {
      level0_uvt_kernel<false><<<gu, 256, 0, stream>>>(q, sh, inv_nseg);
      TCL_CHECK_LAUNCH("postopt(level0)");
      pre_sink_uvt_kernel<<<g0, 256, 0, stream>>>(q, sh);
    }
    TCL_CHECK_LAUNCH("postopt(level0)");
  } else {
    level0_expo_kernel<<<g0, 256, 0, stream>>>(q);
    TCL_CHECK_LAUNCH("postopt(level0)");
    pre_sink_kernel<<<g0, 256, 0, stream>>>(q);
    TCL_CHECK_LAUNCH("postopt(pre_sink)");
  }
  // 7. Adam (+ loss assembly)
  LossAsm la;
  la.scal = w.scal; la.loss_out = loss_out;
  la.inv_flow_cnt = n_valid_g > 0 ? 1.f / flow_cnt : NAN;   // mean over an empty selection is NaN in the reference
  la.inv_tv_h = 1.f / count_h; la.inv_tv_w = 1.f / count_w; la.inv_l1_cnt = 1.f / ((float)planes_g * (float)P);
  la.lambda_flow = c->lambda_flow; la.lambda_dssim = c->lambda_dssim; la.lambda_tv_over_n = c->lambda_tv / nb_g; la.stage = stage;
  la.ms_share = (float)planes / (float)planes_g;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  if (!do_adam || stage == 2) {
    // loss assembly only (n = 0 elements): gradient-only calls, and stage 2 whose Adam runs in adam_uvt_kernel
    adam_kernel<<<1, 32, 0, stream>>>(nullptr, nullptr, nullptr, nullptr, 0, 0.f, beta1, beta2, eps, 1.f, 1.f, la);
    TCL_CHECK_LAUNCH("postopt(loss)");
    if (do_adam) {
      const long long rows = shards->rows_per_rank;
      adam_uvt_kernel<<<gridp(rows / 4 * 3 + 1, 256, 148 * 16), 256, 0, stream>>>(shards->fdc[shards->rank], shards->grad[shards->rank], m, v,
                                                                             rows, lr, beta1, beta2, eps, bc1, bc2_sqrt);
      TCL_CHECK_LAUNCH("postopt(adam)");
    }
  } else {
    adam_kernel<<<gridp((long long)c->N * 12, 256), 256, 0, stream>>>(expo, egrad, em, ev, (long long)c->N * 12, lr, beta1, beta2, eps,
                                                                      bc1, bc2_sqrt, la);
    TCL_CHECK_LAUNCH("postopt(adam)");
  }
  return TCL_OK;
}

extern "C" int tcl_uvt_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids, long long U,
                                 float* fdc, float* grad, float* m, float* v, float lr, float beta1, float beta2, float eps,
                                 int step, float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(ids && fdc && grad && m && v && U > 0, "tcl_uvt_iteration: null pointer");
  tcl_uvt_shards s;
  memset(&s, 0, sizeof(s));
  s.world = 1; s.rank = 0; s.rows_per_rank = U; s.fdc[0] = fdc; s.grad[0] = grad;
  return run_iteration(2, ctx, idx_host, n_batch, ids, &s, m, v, nullptr, nullptr, nullptr, nullptr, lr, beta1, beta2, eps,
                       step, loss_out, stream, true);
}

extern "C" int tcl_exposure_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, float* exposure, float* grad,
                                      float* m, float* v, float lr, float beta1, float beta2, float eps, int step,
                                      float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(exposure && grad && m && v, "tcl_exposure_iteration: null pointer");
  return run_iteration(1, ctx, idx_host, n_batch, nullptr, nullptr, nullptr, nullptr, exposure, grad, m, v, lr, beta1,
                       beta2, eps, step, loss_out, stream, true);
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

// ---- data-parallel stage 2: the UVT rows and their gradient are sharded by row range across the ranks
// of one NVSwitch box; every rank runs its slice of the batch, gathers rows with (peer) loads and scatters gradients with
// (peer) reductions inside the same kernels, and after a cross-rank barrier applies Adam to its own shard only ----
extern "C" int tcl_uvt_gradient_sharded(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids,
                                        const tcl_uvt_shards* shards, float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(ids != nullptr, "tcl_uvt_gradient_sharded: null pointer");
  int rc = check_shards(shards, "tcl_uvt_gradient_sharded");
  if (rc) return rc;
  return run_iteration(2, ctx, idx_host, n_batch, ids, shards, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.9f,
                       0.999f, 1e-15f, 1, loss_out, stream, false);
}

extern "C" int tcl_exposure_gradient(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const float* exposure, float* grad,
                                     float* loss_out, cudaStream_t stream) {
  TCL_CHECK_ARG(exposure && grad, "tcl_exposure_gradient: null pointer");
  return run_iteration(1, ctx, idx_host, n_batch, nullptr, nullptr, nullptr, nullptr, const_cast<float*>(exposure), grad,
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

// Adam over a UVT row shard: fdc, m, v are [rows,3], grad4 the [rows,4] gradient; leaves the gradient zeroed.
extern "C" int tcl_adam_step_uvt(float* fdc, float* grad4, float* m, float* v, long long U, float lr, float beta1, float beta2,
                                 float eps, int step, cudaStream_t stream) {
  TCL_CHECK_ARG(fdc && grad4 && m && v && U > 0 && step >= 1, "tcl_adam_step_uvt: args");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_uvt_kernel<<<gridp(U / 4 * 3 + 1, 256, 148 * 16), 256, 0, stream>>>(fdc, grad4, m, v, U, lr, beta1, beta2, eps, bc1, bc2_sqrt);
  TCL_CHECK_LAUNCH("tcl_adam_step_uvt");
  return TCL_OK;
}
