// VidToMe bipartite soft matching on tcgen05:
// reference utils/VidToMe/vidtome/merge.py:84-108 (local) and :389-412 (global) materialise
//   scores = a @ b^T  [B, n_src, n_dst]  (2.5-4 GB per block at 720p), then scores.max(-1).
// Here the score matrix never leaves TMEM: a GEMM tile's epilogue rounds each score to the
// 16-bit activation type (one rounding from the fp32 accumulator, as a 16-bit matmul output
// would be) and folds it into a running (max, argmax) per src row.  Ties keep the lowest
// dst index.  A second tiny kernel folds the per-(batch, split) partials: with align_batch the
// dst axes of the batch samples are concatenated (index = batch*n_dst + j, merge.py:96-97).
//
// Also here: the row normalisation (merge.py:84 `metric / metric.norm`) fused with the src/dst
// split, and the token gather kernels implementing merge (merge.py:119-133) and unmerge
// (:135-155) as pure row gathers.
#include "common.cuh"
#include "tmap.h"
#include "tclight.h"

namespace tcl {

constexpr int MT_BN = 256;
constexpr int MT_STAGES = 4;
constexpr int MT_THREADS = 192;

struct MatchParams {
  int batch, n_src, n_dst, k_blocks;
  int m_tiles, splits, tiles_per_split, n_dst_tiles;
  int total_units;
  float* part_max;  // [batch][splits][n_src]
  int* part_idx;
};

struct MatchTmaps {
  CUtensorMap a, b;
};

template <bool BF16>
__global__ void __launch_bounds__(MT_THREADS, 1)
match_kernel(const __grid_constant__ MatchTmaps tm, const __grid_constant__ MatchParams p) {
  using E = Elem<BF16>;
  constexpr uint32_t A_STAGE = 128 * 64 * 2;
  constexpr uint32_t B_STAGE = MT_BN * 64 * 2;
  constexpr uint32_t STAGE_BYTES = A_STAGE + B_STAGE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MT_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MT_STAGES;
  uint64_t* tfull_bar = empty_bar + MT_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
    for (int s = 0; s < MT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler

  auto decode = [&](int unit, int& b, int& mt, int& t0, int& t1) {
    const int sp = unit % p.splits;
    int r = unit / p.splits;
    mt = r % p.m_tiles;
    b = r / p.m_tiles;
    t0 = sp * p.tiles_per_split;
    t1 = t0 + p.tiles_per_split;
    if (t1 > p.n_dst_tiles) t1 = p.n_dst_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        int b, mt, t0, t1;
        decode(unit, b, mt, t0, t1);
        for (int dt = t0; dt < t1; ++dt) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_3d(a_dst, &tm.a, &full_bar[stage], kb * 64, mt * 128, b);
            tma_load_3d(a_dst + A_STAGE, &tm.b, &full_bar[stage], kb * 64, dt * MT_BN, b);
            if (++stage == MT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // converged issuer warp, one elected lane issues (operands stay in uniform registers; see igemm.cu)
    {
      constexpr uint32_t idesc = umma_idesc_f16(BF16, 128, MT_BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        int b, mt, t0, t1;
        decode(unit, b, mt, t0, t1);
        for (int dt = t0; dt < t1; ++dt) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * MT_BN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
              const uint64_t a_desc = umma_desc_k_sw128(a_addr);
              const uint64_t b_desc = umma_desc_k_sw128(a_addr + A_STAGE);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              umma_commit(&empty_bar[stage]);
              if (kb + 1 == p.k_blocks) umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == MT_STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
      int b, mt, t0, t1;
      decode(unit, b, mt, t0, t1);
      float best = -INFINITY;
      int best_j = 0x7fffffff;
      for (int dt = t0; dt < t1; ++dt) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * MT_BN;
        const int jbase = dt * MT_BN;
#pragma unroll 1
        for (int c0 = 0; c0 < MT_BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          const int lim = p.n_dst - (jbase + c0);  // columns < lim are real dst tokens
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = E::to_f(E::from_f(__uint_as_float(v[i])));
            if (i < lim && s > best) { best = s; best_j = jbase + c0 + i; }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int r = mt * 128 + row;
      if (r < p.n_src) {
        const int sp = unit % p.splits;
        const long long o = ((long long)b * p.splits + sp) * p.n_src + r;
        p.part_max[o] = best;
        p.part_idx[o] = best_j;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// fold partials.  align==1: one result per src row over all batches (index b*n_dst + j);
// align==0: one result per (batch, row).
__global__ void match_fold_kernel(const float* __restrict__ pm, const int* __restrict__ pi, int batch, int splits,
                                  int n_src, int n_dst, int align, float* __restrict__ node_max,
                                  long long* __restrict__ node_idx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_src) return;
  if (align) {
    float best = -INFINITY;
    long long bi = 0;
    for (int b = 0; b < batch; ++b)
      for (int s = 0; s < splits; ++s) {
        const long long o = ((long long)b * splits + s) * n_src + r;
        const float m = pm[o];
        if (m > best) { best = m; bi = (long long)b * n_dst + pi[o]; }
      }
    node_max[r] = best;
    node_idx[r] = bi;
  } else {
    for (int b = 0; b < batch; ++b) {
      float best = -INFINITY;
      long long bi = 0;
      for (int s = 0; s < splits; ++s) {
        const long long o = ((long long)b * splits + s) * n_src + r;
        const float m = pm[o];
        if (m > best) { best = m; bi = pi[o]; }
      }
      node_max[(long long)b * n_src + r] = best;
      node_idx[(long long)b * n_src + r] = bi;
    }
  }
}

// ------------------------------------------------------------------------------------------
// row normalisation + src/dst split.  tokens come from up to two tensors concatenated along
// the token axis ([x0 | x1], merge.py callers patch.py:64-70); dst is the contiguous token range
// [d0, d1), src is every other token in order.  One warp per token row.
// Arithmetic follows torch on 16-bit tensors: norm accumulated in fp32 and rounded to 16 bit,
// then each element = round16(float(x) / float(norm16)).
// ------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void normalize_split_kernel(const void* __restrict__ x0, long long n0, const void* __restrict__ x1, long long n1,
                                       int batch, int C, long long d0, long long d1,
                                       void* __restrict__ a_out, void* __restrict__ b_out) {
  using E = Elem<BF16>;
  const long long N = n0 + n1;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N * batch) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / N);
  const long long t = row - (long long)b * N;
  const typename E::T* src = t < n0 ? reinterpret_cast<const typename E::T*>(x0) + ((long long)b * n0 + t) * C
                                    : reinterpret_cast<const typename E::T*>(x1) + ((long long)b * n1 + (t - n0)) * C;
  const long long n_dst = d1 - d0, n_src = N - n_dst;
  typename E::T* dst;
  if (t >= d0 && t < d1) dst = reinterpret_cast<typename E::T*>(b_out) + ((long long)b * n_dst + (t - d0)) * C;
  else dst = reinterpret_cast<typename E::T*>(a_out) + ((long long)b * n_src + (t < d0 ? t : t - n_dst)) * C;
  const int vecs = C / 8;
  float ss = 0.f;
  for (int v = lane; v < vecs; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + v * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = E::unpack(w[j]); ss += f.x * f.x + f.y * f.y; }
  }
  ss = warp_sum(ss);
  const float nrm = E::to_f(E::from_f(sqrtf(ss)));
  for (int v = lane; v < vecs; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + v * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = E::unpack(w[j]); o[j] = E::pack(f.x / nrm, f.y / nrm); }
    *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// row gather: out[b, i, :] = (map[i] < n0 ? x0[b, map[i]] : x1[b, map[i]-n0]) (+ add[b, i, :])
// `map` is shared by all batch samples (align_batch) or per batch (map_per_batch).
// ------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void gather_rows_kernel(const void* __restrict__ x0, long long n0, const void* __restrict__ x1, long long n1,
                                   const int* __restrict__ map, int map_per_batch, long long n_out, int batch, int C,
                                   const void* __restrict__ add, void* __restrict__ out) {
  using E = Elem<BF16>;
  const int vecs = C / 8;
  const long long total = (long long)batch * n_out * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long r = i / vecs;
    const int b = (int)(r / n_out);
    const long long o = r - (long long)b * n_out;
    const long long m = map[map_per_batch ? r : o];
    const typename E::T* src = m < n0 ? reinterpret_cast<const typename E::T*>(x0) + ((long long)b * n0 + m) * C
                                      : reinterpret_cast<const typename E::T*>(x1) + ((long long)b * n1 + (m - n0)) * C;
    uint4 u = *reinterpret_cast<const uint4*>(src + v * 8);
    if (add) {
      const uint4 a4 = *reinterpret_cast<const uint4*>(reinterpret_cast<const typename E::T*>(add) + r * C + v * 8);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, aw[4] = {a4.x, a4.y, a4.z, a4.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = E::unpack(uw[j]), g = E::unpack(aw[j]);
        ow[j] = E::pack(f.x + g.x, f.y + g.y);
      }
      u = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<typename E::T*>(out) + r * C + v * 8) = u;
  }
}

// ------------------------------------------------------------------------------------------
// matching plan -> gather maps (shared across the batch: align_batch semantics).
//   edge   : argsort(node_max, descending) [n_src]   (merge.py:98)
//   node_idx [n_src] in [0, batch*n_dst)
//   r      : number of merged src tokens
// tokens: N = n_src + n_dst, dst = [d0, d0+n_dst).  src row s <-> token (s < d0 ? s : s + n_dst).
//   merge_map  [n_src - r + n_dst] : merged row -> token           (merge.py:126-133)
//   unmerge_map[N]                 : token -> merged row           (merge.py:141-153)
// ------------------------------------------------------------------------------------------
__global__ void plan_maps_kernel(const long long* __restrict__ edge, const long long* __restrict__ node_idx,
                                 int n_src, int n_dst, int r, long long d0, int* __restrict__ merge_map,
                                 int* __restrict__ unmerge_map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int unm = n_src - r;
  if (i < n_src) {
    const int s = (int)edge[i];                       // src row ranked i-th
    const int tok = s < d0 ? s : s + n_dst;
    if (i < r) {
      unmerge_map[tok] = unm + (int)(node_idx[s] % n_dst);  // merged src reads its dst (replace mode)
    } else {
      merge_map[i - r] = tok;
      unmerge_map[tok] = i - r;
    }
  }
  if (i < n_dst) {
    merge_map[unm + i] = (int)d0 + i;
    unmerge_map[d0 + i] = unm + i;
  }
}

static inline int grid_for2(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_vidtome_match_workspace_bytes(int batch, int n_src) {
  return (size_t)batch * 8 * (size_t)n_src * (sizeof(float) + sizeof(int));
}

extern "C" int tcl_vidtome_normalize_split(int dtype, const void* x0, long long n0, const void* x1, long long n1,
                                           int batch, int C, long long d0, long long d1, void* a_out, void* b_out,
                                           cudaStream_t stream) {
  TCL_CHECK_ARG(x0 && a_out && b_out && n0 > 0 && n1 >= 0 && (n1 == 0 || x1), "tcl_vidtome_normalize_split: pointers");
  TCL_CHECK_ARG(C % 8 == 0 && batch > 0 && d0 >= 0 && d1 > d0 && d1 <= n0 + n1, "tcl_vidtome_normalize_split: ranges");
  const long long rows = (n0 + n1) * batch;
  const int wpb = 8;
  const long long blocks = (rows + wpb - 1) / wpb;
  if (dtype == TCL_DTYPE_BF16) normalize_split_kernel<true><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x0, n0, x1, n1, batch, C, d0, d1, a_out, b_out);
  else normalize_split_kernel<false><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x0, n0, x1, n1, batch, C, d0, d1, a_out, b_out);
  TCL_CHECK_LAUNCH("tcl_vidtome_normalize_split");
  return TCL_OK;
}

extern "C" int tcl_vidtome_match(int dtype, const void* a, const void* b, int batch, int n_src, int n_dst, int C,
                                 int align_batch, float* node_max, long long* node_idx, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  TCL_CHECK_ARG(a && b && node_max && node_idx && workspace, "tcl_vidtome_match: null pointer");
  TCL_CHECK_ARG(batch > 0 && n_src > 0 && n_dst > 0 && C % 64 == 0, "tcl_vidtome_match: shapes (C must be a multiple of 64)");
  if (workspace_bytes < tcl_vidtome_match_workspace_bytes(batch, n_src)) {
    set_last_error("tcl_vidtome_match: workspace too small");
    return TCL_ERR_WORKSPACE;
  }
  const bool bf16 = dtype == TCL_DTYPE_BF16;
  MatchTmaps tm;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)n_src, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)n_src * C * 2};
    const uint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tm.a, a, bf16, 3, dims, str, box, es, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)n_dst, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)n_dst * C * 2};
    const uint32_t box[3] = {64, MT_BN, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tm.b, b, bf16, 3, dims, str, box, es, 128);
    if (rc) return rc;
  }
  MatchParams p;
  p.batch = batch; p.n_src = n_src; p.n_dst = n_dst; p.k_blocks = C / 64;
  p.m_tiles = (n_src + 127) / 128;
  p.n_dst_tiles = (n_dst + MT_BN - 1) / MT_BN;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // choose the dst split that best fills whole waves of `sms` persistent CTAs
  int best_s = 1; double best_eff = -1;
  for (int s = 1; s <= 8 && s <= p.n_dst_tiles; ++s) {
    const long long units = (long long)batch * p.m_tiles * s;
    const long long waves = (units + sms - 1) / sms;
    const int tps = (p.n_dst_tiles + s - 1) / s;
    const double eff = (double)batch * p.m_tiles * p.n_dst_tiles / ((double)waves * sms * tps);
    if (eff > best_eff + 1e-6) { best_eff = eff; best_s = s; }
  }
  p.splits = best_s;
  p.tiles_per_split = (p.n_dst_tiles + best_s - 1) / best_s;
  p.total_units = batch * p.m_tiles * p.splits;
  p.part_max = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_max + (size_t)batch * 8 * n_src);
  constexpr size_t smem = MT_STAGES * (128 * 64 * 2 + MT_BN * 64 * 2) + 1024 + 256;
  static bool conf[2] = {false, false};
  if (!conf[bf16]) {
    cudaError_t e = bf16 ? cudaFuncSetAttribute(match_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(match_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error("tcl_vidtome_match: smem attr: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
    conf[bf16] = true;
  }
  const int grid = p.total_units < sms ? p.total_units : sms;
  if (bf16) match_kernel<true><<<grid, MT_THREADS, smem, stream>>>(tm, p);
  else match_kernel<false><<<grid, MT_THREADS, smem, stream>>>(tm, p);
  TCL_CHECK_LAUNCH("tcl_vidtome_match");
  match_fold_kernel<<<(n_src + 255) / 256, 256, 0, stream>>>(p.part_max, p.part_idx, batch, p.splits, n_src, n_dst,
                                                             align_batch, node_max, node_idx);
  TCL_CHECK_LAUNCH("tcl_vidtome_match(fold)");
  return TCL_OK;
}

extern "C" int tcl_vidtome_plan(const long long* edge, const long long* node_idx, int n_src, int n_dst, int r,
                                long long d0, int* merge_map, int* unmerge_map, cudaStream_t stream) {
  TCL_CHECK_ARG(edge && node_idx && merge_map && unmerge_map, "tcl_vidtome_plan: null pointer");
  TCL_CHECK_ARG(n_src > 0 && n_dst > 0 && r >= 0 && r <= n_src && d0 >= 0 && d0 <= n_src, "tcl_vidtome_plan: args");
  const int n = n_src > n_dst ? n_src : n_dst;
  plan_maps_kernel<<<(n + 255) / 256, 256, 0, stream>>>(edge, node_idx, n_src, n_dst, r, d0, merge_map, unmerge_map);
  TCL_CHECK_LAUNCH("tcl_vidtome_plan");
  return TCL_OK;
}

extern "C" int tcl_gather_rows(int dtype, const void* x0, long long n0, const void* x1, long long n1, const int* map,
                               int map_per_batch, long long n_out, int batch, int C, const void* add, void* out,
                               cudaStream_t stream) {
  TCL_CHECK_ARG(x0 && map && out && n0 > 0 && n1 >= 0 && (n1 == 0 || x1), "tcl_gather_rows: pointers");
  TCL_CHECK_ARG(C % 8 == 0 && batch > 0 && n_out > 0, "tcl_gather_rows: shapes");
  const long long total = (long long)batch * n_out * (C / 8);
  if (dtype == TCL_DTYPE_BF16) gather_rows_kernel<true><<<grid_for2(total, 256), 256, 0, stream>>>(x0, n0, x1, n1, map, map_per_batch, n_out, batch, C, add, out);
  else gather_rows_kernel<false><<<grid_for2(total, 256), 256, 0, stream>>>(x0, n0, x1, n1, map, map_per_batch, n_out, batch, C, add, out);
  TCL_CHECK_LAUNCH("tcl_gather_rows");
  return TCL_OK;
}
