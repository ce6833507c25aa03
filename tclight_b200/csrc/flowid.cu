// Producer side of the stage-2 optimiser: everything is fp32 /
// int32 streaming work bound by HBM bandwidth.
//
//   tcl_warp_bicubic     warp_flow                 utils/flow_utils.py:5-16
//   tcl_soft_mask_bwd    get_soft_mask_bwds        utils/flow_utils.py:40-54
//   tcl_flow_ids         get_flowid                utils/flow_utils.py:56-92
//   tcl_unique_inverse   voxelization(voxel_size=None) = torch.unique(dim=0, return_inverse=True) on one id
//                        column                    utils/general_utils.py:223-233
//   tcl_max_f32          tensor.max() feeding the two thresholds (flow_utils.py:52, 71) without a host sync
//
// Design.  The reference walks the frames sequentially with boolean-mask indexing (a host sync per frame) and
// then sorts N*H*W keys.  Here the per-frame dependency is reduced to its minimum:
//   1. parents (all frames in parallel): every pixel of frame i-1 votes for the pixel of frame i its rounded
//      forward flow lands on (atomicMax of the source index = "last writer in row-major order wins", the
//      reference's CPU semantics; its CUDA winner is unspecified);
//   2. fresh ids: a pixel without a parent gets `H*W + rank` where rank counts parent-less pixels of frames
//      1..i in frame-major / row-major order — one exclusive scan over the whole clip (block sums + carry);
//   3. propagation: ids[i] = ids[i-1][parent] — N-1 tiny dependent gathers (3.7 MB each at 720p).
// The ids produced this way are dense in [0, U), so torch.unique's inverse is the identity; the general
// tcl_unique_inverse (presence bitmap + scan + gather) still implements the reference call for arbitrary ids.
#include "common.cuh"
#include "tclight.h"

namespace tcl {

static inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- bicubic sampling (grid_sample: bicubic, zeros padding, align_corners=True, A = -0.75) -------------
__device__ __forceinline__ float fcubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float fcubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

struct Taps {
  int x0, y0;
  float cx[4], cy[4];
};

// pixel (x, y) displaced by (fx, fy): the reference normalises to [-1, 1] and grid_sample maps back
// (flow_utils.py:12-13); both roundings are reproduced.
__device__ __forceinline__ Taps make_taps(int x, int y, float fx, float fy, int H, int W) {
  const float px = fx + (float)x, py = fy + (float)y;
  const float gx = (px / (float)(W - 1) - 0.5f) * 2.f;
  const float gy = (py / (float)(H - 1) - 0.5f) * 2.f;
  const float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  const float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float tx = ix - fx0, ty = iy - fy0;
  Taps t;
  // clamp far-out coordinates so the int conversion is defined; every tap is out of bounds there anyway
  t.x0 = (int)fminf(fmaxf(fx0, -8.f), (float)W + 8.f) - 1;
  t.y0 = (int)fminf(fmaxf(fy0, -8.f), (float)H + 8.f) - 1;
  t.cx[0] = fcubic2(tx + 1.f); t.cx[1] = fcubic1(tx); t.cx[2] = fcubic1(1.f - tx); t.cx[3] = fcubic2((1.f - tx) + 1.f);
  t.cy[0] = fcubic2(ty + 1.f); t.cy[1] = fcubic1(ty); t.cy[2] = fcubic1(1.f - ty); t.cy[3] = fcubic2((1.f - ty) + 1.f);
  return t;
}

template <int C>
__device__ __forceinline__ void sample_planes(const float* __restrict__ src, long long P, int H, int W, const Taps& t,
                                              float (&out)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = t.y0 + j;
    if (yy < 0 || yy >= H) continue;
    float row[C];
#pragma unroll
    for (int c = 0; c < C; ++c) row[c] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = t.x0 + i;
      if (xx < 0 || xx >= W) continue;
      const long long o = (long long)yy * W + xx;
#pragma unroll
      for (int c = 0; c < C; ++c) row[c] += __ldg(src + c * P + o) * t.cx[i];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] += row[c] * t.cy[j];
  }
}

__global__ void __launch_bounds__(256)
warp_bicubic_kernel(const float* __restrict__ frames, const float* __restrict__ flows, int N, int C, int H, int W,
                    float* __restrict__ out) {
  const long long P = (long long)H * W;
  const long long total = (long long)N * P;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(g / P);
    const long long p = g - (long long)n * P;
    const int x = (int)(p % W), y = (int)(p / W);
    const Taps t = make_taps(x, y, flows[((long long)n * 2 + 0) * P + p], flows[((long long)n * 2 + 1) * P + p], H, W);
    for (int c = 0; c < C; ++c) {
      float v[1];
      sample_planes<1>(frames + ((long long)n * C + c) * P, P, H, W, t, v);
      out[((long long)n * C + c) * P + p] = v[0];
    }
  }
}

// ---- max reduction (threshold source) ---------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // order-preserving for any sign: positives compare as ints, negatives as reversed unsigned
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_f32_kernel(float* p, float v) { *p = v; }

__global__ void __launch_bounds__(256) max_f32_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float m = -INFINITY;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomic_max_float(out, m);
  }
}

// ---- soft backward masks ---------------------------------------------------------------------------
// One thread per pixel of frames 1..N-1: the forward flow of frame i-1 and the image of frame i-1 are
// sampled at the same backward-flow position (shared taps), then the two sigmoids are multiplied.
__global__ void __launch_bounds__(256)
soft_mask_kernel(const float* __restrict__ images, const float* __restrict__ flows, const float* __restrict__ past,
                 int N, int H, int W, float alpha, float beta, double diff_threshold, const float* __restrict__ img_max,
                 float* __restrict__ out) {
  const long long P = (long long)H * W;
  const long long total = (long long)N * P;
  const float thr = (float)((double)(*img_max) * diff_threshold);      // python float * float, then cast by the op
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(g / P);
    if (n == 0) { out[g] = 1.f; continue; }
    const long long p = g - (long long)n * P;
    const int x = (int)(p % W), y = (int)(p / W);
    const float pfx = past[((long long)n * 2 + 0) * P + p], pfy = past[((long long)n * 2 + 1) * P + p];
    const Taps t = make_taps(x, y, pfx, pfy, H, W);
    float f2b[2], im[3];
    sample_planes<2>(flows + (long long)(n - 1) * 2 * P, P, H, W, t, f2b);
    sample_planes<3>(images + (long long)(n - 1) * 3 * P, P, H, W, t, im);
    const float sx = pfx + f2b[0], sy = pfy + f2b[1];
    const float a = sqrtf(sx * sx + sy * sy);
    const float b = sqrtf(pfx * pfx + pfy * pfy);
    const float c = sqrtf(f2b[0] * f2b[0] + f2b[1] * f2b[1]);
    const float e1 = -beta * (a - ((b + c) + 1.f) * alpha);
    const float m1 = 1.f / (1.f + expf(-e1));
    float d = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) d = fmaxf(d, fabsf(im[ch] - images[((long long)n * 3 + ch) * P + p]));
    const float e2 = -beta * (d - thr);
    const float m2 = 1.f / (1.f + expf(-e2));
    out[g] = (1.f * m1) * m2;
  }
}

// ---- flow ids ----------------------------------------------------------------------------------------
__global__ void fill_i32_kernel(int* __restrict__ p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// votes: source pixel (frame i-1, linear index s) -> target pixel of frame i
__global__ void __launch_bounds__(256)
flowid_parent_kernel(const float* __restrict__ frames, const float* __restrict__ flows, const float* __restrict__ mask,
                     int N, int H, int W, double rgb_threshold, const float* __restrict__ frames_max,
                     int* __restrict__ parent /* [N, P], pre-filled with -1 */) {
  const long long P = (long long)H * W;
  const long long total = (long long)(N - 1) * P;
  const float thr = (float)((double)(*frames_max) * rgb_threshold);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(g / P) + 1;                    // target frame
    const long long s = g - (long long)(i - 1) * P;    // source pixel in frame i-1
    const int gx = (int)(s % W), gy = (int)(s / W);
    const float tx = rintf((float)gx + flows[((long long)(i - 1) * 2 + 0) * P + s]);
    const float ty = rintf((float)gy + flows[((long long)(i - 1) * 2 + 1) * P + s]);
    if (!(tx >= 0.f && tx < (float)W && ty >= 0.f && ty < (float)H)) continue;
    if (!(mask[(long long)i * P + s] > 0.5f)) continue;
    const long long t = (long long)(int)ty * W + (int)tx;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      d = fmaxf(d, fabsf(frames[((long long)i * 3 + c) * P + t] - frames[((long long)(i - 1) * 3 + c) * P + s]));
    if (d < thr) atomicMax(&parent[(long long)i * P + t], (int)s);
  }
}

// Exclusive scan of a 0/1 flag derived from an int array, in chunks of SCAN_CHUNK elements per block.
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_BLOCK * SCAN_ITEMS;

// MODE 0: flag = element < 0 or index < first_all (frame 0 of the flow ids);  MODE 1: flag = element != 0
template <int MODE>
__device__ __forceinline__ int scan_flag(int v, long long i, long long first_all) {
  return MODE == 0 ? ((v < 0 || i < first_all) ? 1 : 0) : (v != 0 ? 1 : 0);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int wsum[SCAN_BLOCK / 32];
  __shared__ int tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < SCAN_BLOCK / 32; ++k) { const int t = wsum[k]; wsum[k] = run; run += t; }
    tot = run;
  }
  __syncthreads();
  const int res = inc - v + wsum[w];
  *total = tot;
  __syncthreads();
  return res;
}

template <int MODE>
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_count_kernel(const int* __restrict__ a, long long n, long long first_all, int* __restrict__ sums) {
  const long long base = (long long)blockIdx.x * SCAN_CHUNK;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = base + k * SCAN_BLOCK + threadIdx.x;
    if (i < n) c += scan_flag<MODE>(a[i], i, first_all);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int red[SCAN_BLOCK / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < SCAN_BLOCK / 32; ++k) t += red[k];
    sums[blockIdx.x] = t;
  }
}

// one block: exclusive scan of the chunk sums in place; the grand total goes to total_out (int64)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(int* __restrict__ sums, int nb, long long* __restrict__ total_out) {
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    if (i < nb) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// Frame i of the flow ids, in place over the parent array: parent >= 0 -> id of the parent in frame i-1,
// else H*W-offset fresh id (frame 0: rank == linear index).  Element k of a chunk owns items
// [k*ITEMS, (k+1)*ITEMS) so that the ranks follow row-major order.
__global__ void __launch_bounds__(SCAN_BLOCK)
flowid_assign_kernel(int* __restrict__ ids /* [N,P] parent in, id out */, long long P, int frame,
                     const int* __restrict__ sums) {
  const long long fbase = (long long)frame * P;
  const long long cbase = fbase + (long long)blockIdx.x * SCAN_CHUNK;     // chunks are frame-aligned (see host)
  const long long fend = fbase + P;
  int v[SCAN_ITEMS], f[SCAN_ITEMS], cnt = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = cbase + (long long)threadIdx.x * SCAN_ITEMS + k;
    v[k] = i < fend ? ids[i] : 0;
    f[k] = (i < fend && (frame == 0 || v[k] < 0)) ? 1 : 0;
    cnt += f[k];
  }
  int tot;
  int rank = block_exclusive_scan(cnt, &tot) + sums[(long long)frame * gridDim.x + blockIdx.x];
  const int* prev = ids + fbase - P;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = cbase + (long long)threadIdx.x * SCAN_ITEMS + k;
    if (i >= fend) continue;
    if (f[k]) { ids[i] = rank; ++rank; }
    else ids[i] = prev[v[k]];
  }
}

// per-frame chunked count for the flow-id scan (chunks restart at every frame so that the assign kernel's
// blocks never straddle two frames)
__global__ void __launch_bounds__(SCAN_BLOCK)
flowid_count_kernel(const int* __restrict__ parent, long long P, int chunks_per_frame, int* __restrict__ sums) {
  const int frame = blockIdx.y;
  const long long fbase = (long long)frame * P;
  const long long cbase = fbase + (long long)blockIdx.x * SCAN_CHUNK;
  const long long fend = fbase + P;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = cbase + k * SCAN_BLOCK + threadIdx.x;
    if (i < fend) c += (frame == 0 || parent[i] < 0) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int red[SCAN_BLOCK / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < SCAN_BLOCK / 32; ++k) t += red[k];
    sums[(long long)frame * chunks_per_frame + blockIdx.x] = t;
  }
}

// ---- unique inverse ---------------------------------------------------------------------------------
__global__ void mark_present_kernel(const int* __restrict__ ids, long long n, int* __restrict__ present) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    present[ids[i]] = 1;
}

// present[] (0/1) -> rank[] in place
__global__ void __launch_bounds__(SCAN_BLOCK)
rank_write_kernel(int* __restrict__ present, long long n, const int* __restrict__ sums) {
  const long long cbase = (long long)blockIdx.x * SCAN_CHUNK;
  int f[SCAN_ITEMS], cnt = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = cbase + (long long)threadIdx.x * SCAN_ITEMS + k;
    f[k] = (i < n && present[i] != 0) ? 1 : 0;
    cnt += f[k];
  }
  int tot;
  int rank = block_exclusive_scan(cnt, &tot) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const long long i = cbase + (long long)threadIdx.x * SCAN_ITEMS + k;
    if (i < n) { present[i] = rank; rank += f[k]; }
  }
}

__global__ void gather_rank_kernel(const int* __restrict__ ids, long long n, const int* __restrict__ rank,
                                   long long* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (long long)rank[ids[i]];
}

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_warp_bicubic(const float* frames, const float* flows, int N, int C, int H, int W, float* out,
                                cudaStream_t stream) {
  TCL_CHECK_ARG(frames && flows && out, "tcl_warp_bicubic: null pointer");
  TCL_CHECK_ARG(N > 0 && C > 0 && H > 1 && W > 1, "tcl_warp_bicubic: shape N=%d C=%d H=%d W=%d", N, C, H, W);
  warp_bicubic_kernel<<<grid_for((long long)N * H * W, 256), 256, 0, stream>>>(frames, flows, N, C, H, W, out);
  TCL_CHECK_LAUNCH("tcl_warp_bicubic");
  return TCL_OK;
}

extern "C" int tcl_max_f32(const float* x, long long n, float* out, cudaStream_t stream) {
  TCL_CHECK_ARG(x && out && n > 0, "tcl_max_f32: args");
  TCL_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "tcl_max_f32: x must be 16-byte aligned");
  fill_f32_kernel<<<1, 1, 0, stream>>>(out, -INFINITY);
  TCL_CHECK_LAUNCH("tcl_max_f32");
  max_f32_kernel<<<grid_for(n / 4 + 1, 256, 148 * 8), 256, 0, stream>>>(x, n, out);
  TCL_CHECK_LAUNCH("tcl_max_f32");
  return TCL_OK;
}

extern "C" int tcl_soft_mask_bwd(const float* images, const float* flows, const float* past_flows, int N, int H, int W,
                                 float alpha, float beta, double diff_threshold, const float* images_max, float* out,
                                 cudaStream_t stream) {
  TCL_CHECK_ARG(images && flows && past_flows && images_max && out, "tcl_soft_mask_bwd: null pointer");
  TCL_CHECK_ARG(N > 0 && H > 1 && W > 1, "tcl_soft_mask_bwd: shape N=%d H=%d W=%d", N, H, W);
  soft_mask_kernel<<<grid_for((long long)N * H * W, 256), 256, 0, stream>>>(images, flows, past_flows, N, H, W, alpha, beta,
                                                                            diff_threshold, images_max, out);
  TCL_CHECK_LAUNCH("tcl_soft_mask_bwd");
  return TCL_OK;
}

extern "C" size_t tcl_flow_ids_workspace_bytes(int N, int H, int W) {
  const long long P = (long long)H * W;
  const long long cpf = (P + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return al256(sizeof(int) * (size_t)N * cpf) + 256;
}

extern "C" int tcl_flow_ids(const float* frames, const float* flows, const float* mask_bwds, int N, int H, int W,
                            double rgb_threshold, const float* frames_max, int* ids, long long* num_ids, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream) {
  TCL_CHECK_ARG(frames && flows && mask_bwds && frames_max && ids && workspace, "tcl_flow_ids: null pointer");
  TCL_CHECK_ARG(N > 0 && H > 0 && W > 0, "tcl_flow_ids: shape");
  const long long P = (long long)H * W;
  TCL_CHECK_ARG((long long)N * P < (1ll << 31), "tcl_flow_ids: N*H*W >= 2^31 needs int64 ids (not implemented)");
  TCL_CHECK_ARG(workspace_bytes >= tcl_flow_ids_workspace_bytes(N, H, W), "tcl_flow_ids: workspace too small");
  const int cpf = (int)((P + SCAN_CHUNK - 1) / SCAN_CHUNK);
  int* sums = reinterpret_cast<int*>(workspace);
  fill_i32_kernel<<<grid_for((long long)N * P, 256), 256, 0, stream>>>(ids, (long long)N * P, -1);
  TCL_CHECK_LAUNCH("tcl_flow_ids");
  if (N > 1) {
    flowid_parent_kernel<<<grid_for((long long)(N - 1) * P, 256), 256, 0, stream>>>(frames, flows, mask_bwds, N, H, W,
                                                                                    rgb_threshold, frames_max, ids);
    TCL_CHECK_LAUNCH("tcl_flow_ids");
  }
  flowid_count_kernel<<<dim3(cpf, N), SCAN_BLOCK, 0, stream>>>(ids, P, cpf, sums);
  TCL_CHECK_LAUNCH("tcl_flow_ids");
  scan_sums_kernel<<<1, SCAN_BLOCK, 0, stream>>>(sums, N * cpf, num_ids);
  TCL_CHECK_LAUNCH("tcl_flow_ids");
  for (int f = 0; f < N; ++f) {
    flowid_assign_kernel<<<cpf, SCAN_BLOCK, 0, stream>>>(ids, P, f, sums);
    TCL_CHECK_LAUNCH("tcl_flow_ids");
  }
  return TCL_OK;
}

extern "C" size_t tcl_unique_inverse_workspace_bytes(long long id_range) {
  const long long nb = (id_range + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return al256(sizeof(int) * (size_t)id_range) + al256(sizeof(int) * (size_t)nb) + 256;
}

extern "C" int tcl_unique_inverse(const int* ids, long long n, long long id_range, long long* inverse,
                                  long long* num_unique, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  TCL_CHECK_ARG(ids && inverse && workspace && n > 0 && id_range > 0, "tcl_unique_inverse: args");
  TCL_CHECK_ARG(id_range < (1ll << 31), "tcl_unique_inverse: id range >= 2^31");
  TCL_CHECK_ARG(workspace_bytes >= tcl_unique_inverse_workspace_bytes(id_range), "tcl_unique_inverse: workspace too small");
  int* present = reinterpret_cast<int*>(workspace);
  int* sums = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + al256(sizeof(int) * (size_t)id_range));
  const int nb = (int)((id_range + SCAN_CHUNK - 1) / SCAN_CHUNK);
  fill_i32_kernel<<<grid_for(id_range, 256), 256, 0, stream>>>(present, id_range, 0);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  mark_present_kernel<<<grid_for(n, 256), 256, 0, stream>>>(ids, n, present);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  scan_count_kernel<1><<<nb, SCAN_BLOCK, 0, stream>>>(present, id_range, 0, sums);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  scan_sums_kernel<<<1, SCAN_BLOCK, 0, stream>>>(sums, nb, num_unique);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  rank_write_kernel<<<nb, SCAN_BLOCK, 0, stream>>>(present, id_range, sums);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  gather_rank_kernel<<<grid_for(n, 256), 256, 0, stream>>>(ids, n, present, inverse);
  TCL_CHECK_LAUNCH("tcl_unique_inverse");
  return TCL_OK;
}
