// tcgen05 implicit GEMM for the UNet contractions:
//   D[pixel, n] = sum over K-segments ( A_seg[pixel (+tap offset), c] * W[n, k] )
// One kernel covers nn.Linear, 1x1 conv, 3x3 conv (pad 1, stride 1 or 2) and the fused
// "conv2 + 1x1 shortcut" of a ResnetBlock2D, by describing K as a list of segments, each a
// (NHWC tensor, taps) pair.  Activations are NHWC 16-bit; weights are [N, Ktotal] K-major.
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0      TMA producer: 4-D box {64 ch, tw, th, tn} of A (zero-filled halo = conv padding)
//               + 2-D box {64, BN} of W per pipeline stage, 128-byte swizzle.
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16 x4 per stage,
//               fp32 accumulators in TMEM, double buffered (2 x BN columns).
//   warps 2..9  epilogue (two warps per TMEM lane quarter, interleaved 32-column chunks):
//               tcgen05.ld -> bias / residual / GEGLU / head-split -> swizzled smem -> TMA store
//               (or direct 16-bit stores for the head-split layouts).
#include "common.cuh"
#include "tmap.h"
#include "tclight.h"

namespace tcl {

constexpr int IG_BM = 128;
constexpr int IG_BK = 64;
constexpr int IG_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)

struct IgSrc {
  int taps;     // 1 or 9
  int cchunks;  // channels / 64
  int stride;   // 1 or 2
  int pad;      // 0 (taps==1) or 1
};

struct IgParams {
  int n_img, out_h, out_w;
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n;
  int m_tiles, n_tiles, total_tiles;
  int N;
  int num_src;
  IgSrc src[TCL_IGEMM_MAX_SRC];
  int k_blocks;
  uint32_t a_bytes;  // bytes of one A box
  // epilogue
  int mode;
  const float* bias;          // [N] or null
  const void* residual;       // NHWC 16-bit or null
  long long res_pitch;        // elements per pixel in residual
  void* out;                  // NHWC
  long long out_pitch;        // elements per pixel in out
  float out_scale;
  int vec_ok;                 // out/residual pitches allow 16-byte accesses
  int tma_store;              // epilogue stages the tile in smem and writes it with TMA (coalesced)
  // head-split modes
  void* sec_ptr[3];
  int sec_vt[3];
  int sec_cols;               // C (columns per section)
  int heads, d, d_pad;
  long long tok_per_batch;    // T
  long long tok_pitch;        // Tp (rows allocated per (b,head) in Q layout; pitch of V^T rows)
};

struct IgTmaps {
  CUtensorMap a[TCL_IGEMM_MAX_SRC];
  CUtensorMap b;
  CUtensorMap c;   // output (TMA-store epilogue): box {32 ch, tw, th, tn}, 64-byte swizzle
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// far below the 16-bit rounding of the result): one reciprocal + one exp2 instead of erff's ~40 instructions.
__device__ __forceinline__ float gelu_exact(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = exp2f(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float erf_x = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_x);
}

template <int BN, int STAGES, bool BF16>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ IgTmaps tm, const __grid_constant__ IgParams p) {
  using E = Elem<BF16>;
  constexpr uint32_t A_STAGE = IG_BM * IG_BK * 2;  // 16 KB
  constexpr uint32_t B_STAGE = BN * IG_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_STAGE + B_STAGE;
  constexpr uint32_t ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE < 32 ? 32 : 2 * ACC_STRIDE;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of the dynamic window is required by the 128B swizzle.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr uint32_t STAGING_BYTES = BN * 256;     // BN/32 chunks of 128 rows x 64 B (16-bit output tile)
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&tm.a[s]);
    tma_prefetch_desc(&tm.b);
    if (p.tma_store) tma_prefetch_desc(&tm.c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler

  if (warp == 0) {
    // ===================== TMA producer (converged warp, elected lane issues) =====================
    {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int ti_w = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int ti_h = mt % p.tiles_h;
        const int ti_n = mt / p.tiles_h;
        const int x0 = ti_w * p.tw, y0 = ti_h * p.th, img0 = ti_n * p.tn;
        int kb = 0;
        for (int s = 0; s < p.num_src; ++s) {
          const IgSrc sd = p.src[s];
          for (int tap = 0; tap < sd.taps; ++tap) {
            const int dy = sd.taps == 9 ? tap / 3 : 0;
            const int dx = sd.taps == 9 ? tap % 3 : 0;
            const int cx = x0 * sd.stride + dx - sd.pad;
            const int cy = y0 * sd.stride + dy - sd.pad;
            for (int cc = 0; cc < sd.cchunks; ++cc, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                uint8_t* a_dst = smem + stage * STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_STAGE;
                mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + B_STAGE);
                tma_load_4d(a_dst, &tm.a[s], &full_bar[stage], cc * IG_BK, cx, cy, img0);
                tma_load_2d(b_dst, &tm.b, &full_bar[stage], kb * IG_BK, nt * BN);
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the control flow converged and one elected lane issues: operands then live in uniform
    // registers.  (Issuing from inside `if (lane == 0)` made the compiler wrap every tcgen05.mma in an
    // ELECT / R2UR.BROADCAST waterfall, ~70 clk per instruction — more than a BN<=160 MMA takes to execute.)
    {
      constexpr uint32_t idesc = umma_idesc_f16(BF16, IG_BM, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
            const uint64_t a_desc = umma_desc_k_sw128(a_addr);
            const uint64_t b_desc = umma_desc_k_sw128(a_addr + A_STAGE);
#pragma unroll
            for (int k = 0; k < IG_BK / 16; ++k) {
              // advance 16 elements (32 B) inside the 128 B swizzle row: +2 in 16-byte units
              umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);
            if (kb + 1 == p.k_blocks) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which of the two warps sharing that quarter (takes every other chunk)
    const int row = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int ti_w = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ti_h = mt % p.tiles_h;
      const int ti_n = mt / p.tiles_h;
      const int rw = row % p.tw;
      const int rh = (row / p.tw) % p.th;
      const int rn = row / (p.tw * p.th);
      const int x = ti_w * p.tw + rw, y = ti_h * p.th + rh, img = ti_n * p.tn + rn;
      const bool valid = (rn < p.tn) && (x < p.out_w) && (y < p.out_h) && (img < p.n_img);
      const long long pix = (static_cast<long long>(img) * p.out_h + y) * p.out_w + x;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * ACC_STRIDE;

      if (p.tma_store) {
        // ---- coalesced epilogue: TMEM -> registers -> swizzled smem staging -> TMA store ----
        constexpr int HALF = BN / 2;
        const bool geglu = p.mode == TCL_EPI_GEGLU;
        const int out_cols = geglu ? HALF : BN;
        const int epi_tid = threadIdx.x - 64;
        if (epi_tid == 0) bulk_wait_read0();                     // previous tile's stores have read the staging
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint32_t stg_row = smem_u32(staging) + row * 64;
        const int sw = (row >> 1) & 3;
        const typename E::T* res = (p.residual && valid)
                                       ? reinterpret_cast<const typename E::T*>(p.residual) + pix * p.res_pitch : nullptr;
#pragma unroll 1
        for (int c0 = half * 32; c0 < out_cols; c0 += 64) {
          uint32_t v[32];
          uint32_t pk[16];
          tmem_ld_32x32b_x32(t_row + c0, v);
          if (geglu) {
            uint32_t g[32];
            tmem_ld_32x32b_x32(t_row + HALF + c0, g);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float v0 = __uint_as_float(v[j]), v1 = __uint_as_float(v[j + 1]);
              float g0 = __uint_as_float(g[j]), g1 = __uint_as_float(g[j + 1]);
              if (p.bias) {
                v0 += p.bias[nt * BN + c0 + j];
                v1 += p.bias[nt * BN + c0 + j + 1];
                g0 += p.bias[nt * BN + HALF + c0 + j];
                g1 += p.bias[nt * BN + HALF + c0 + j + 1];
              }
              // the reference rounds the projection and gelu(gate) to 16 bit (nn.Linear / F.gelu outputs)
              v0 = E::to_f(E::from_f(v0)); v1 = E::to_f(E::from_f(v1));
              g0 = E::to_f(E::from_f(gelu_exact(E::to_f(E::from_f(g0)))));
              g1 = E::to_f(E::from_f(gelu_exact(E::to_f(E::from_f(g1)))));
              pk[j / 2] = E::pack(v0 * g0, v1 * g1);
            }
          } else {
            // bias / residual loads are issued before waiting on the TMEM load so their latencies overlap
            const int col0 = nt * BN + c0;
            float4 bb[8];
            uint4 rr4[4];
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = col0 + g8 * 8;
              const bool in = col + 8 <= p.N;
              bb[2 * g8] = (p.bias && in) ? *reinterpret_cast<const float4*>(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
              bb[2 * g8 + 1] = (p.bias && in) ? *reinterpret_cast<const float4*>(p.bias + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              rr4[g8] = (res && in) ? *reinterpret_cast<const uint4*>(res + col) : make_uint4(0, 0, 0, 0);
            }
            tmem_ld_wait();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g8 * 8 + j]);
              const float4 b0 = bb[2 * g8], b1 = bb[2 * g8 + 1];
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              const uint32_t rr[4] = {rr4[g8].x, rr4[g8].y, rr4[g8].z, rr4[g8].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 rf = E::unpack(rr[j]);
                f[2 * j] += rf.x;
                f[2 * j + 1] += rf.y;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) pk[g8 * 4 + j] = E::pack(f[2 * j] * p.out_scale, f[2 * j + 1] * p.out_scale);
            }
          }
          const uint32_t dst = stg_row + (c0 / 32) * 8192;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((u ^ sw) * 16)), "r"(pk[u * 4 + 0]),
                         "r"(pk[u * 4 + 1]), "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                         : "memory");
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);            // accumulator drained: MMA may reuse it
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (epi_tid == 0) {
          const int x0 = ti_w * p.tw, y0 = ti_h * p.th, img0 = ti_n * p.tn;
          const int colb = geglu ? nt * HALF : nt * BN;
          for (int c0 = 0; c0 < out_cols; c0 += 32)
            if (colb + c0 < (geglu ? p.N / 2 : p.N)) tma_store_4d(&tm.c, smem_u32(staging) + (c0 / 32) * 8192, colb + c0, x0, y0, img0);
          bulk_commit();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if (p.mode == TCL_EPI_GEGLU) {
        // columns [0, BN/2) = value, [BN/2, BN) = gate for the same BN/2 output channels
        constexpr int HALF = BN / 2;
        typename E::T* out = reinterpret_cast<typename E::T*>(p.out);
#pragma unroll 1
        for (int c0 = half * 16; c0 < HALF; c0 += 32) {
          uint32_t v[16], g[16];
          tmem_ld_32x32b_x16(t_row + c0, v);
          tmem_ld_32x32b_x16(t_row + HALF + c0, g);
          tmem_ld_wait();
          const int ocol = nt * HALF + c0;  // output channel
          if (valid && ocol < p.N / 2) {
            uint32_t packed[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float v0 = __uint_as_float(v[j]), v1 = __uint_as_float(v[j + 1]);
              float g0 = __uint_as_float(g[j]), g1 = __uint_as_float(g[j + 1]);
              if (p.bias) {
                v0 += p.bias[nt * BN + c0 + j];
                v1 += p.bias[nt * BN + c0 + j + 1];
                g0 += p.bias[nt * BN + HALF + c0 + j];
                g1 += p.bias[nt * BN + HALF + c0 + j + 1];
              }
              // reference rounds the projection to 16 bit before the gate (nn.Linear output)
              v0 = E::to_f(E::from_f(v0)); v1 = E::to_f(E::from_f(v1));
              g0 = E::to_f(E::from_f(g0)); g1 = E::to_f(E::from_f(g1));
              packed[j / 2] = E::pack(v0 * gelu_exact(g0), v1 * gelu_exact(g1));
            }
            uint4* dst = reinterpret_cast<uint4*>(out + pix * p.out_pitch + ocol);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          const int col0 = nt * BN + c0;
          if (!valid) continue;
          if (p.mode == TCL_EPI_NHWC) {
            typename E::T* out = reinterpret_cast<typename E::T*>(p.out) + pix * p.out_pitch;
            const typename E::T* res =
                p.residual ? reinterpret_cast<const typename E::T*>(p.residual) + pix * p.res_pitch
                           : nullptr;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = col0 + g8 * 8;
              if (col >= p.N) break;
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g8 * 8 + j]);
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (col + j < p.N) f[j] += p.bias[col + j];
              }
              if (col + 8 <= p.N && p.vec_ok) {
                if (res) {
                  const uint4 r4 = *reinterpret_cast<const uint4*>(res + col);
                  const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 rf = E::unpack(rr[j]);
                    f[2 * j] += rf.x;
                    f[2 * j + 1] += rf.y;
                  }
                }
                uint4 o;
                o.x = E::pack(f[0] * p.out_scale, f[1] * p.out_scale);
                o.y = E::pack(f[2] * p.out_scale, f[3] * p.out_scale);
                o.z = E::pack(f[4] * p.out_scale, f[5] * p.out_scale);
                o.w = E::pack(f[6] * p.out_scale, f[7] * p.out_scale);
                *reinterpret_cast<uint4*>(out + col) = o;
              } else {
                for (int j = 0; j < 8 && col + j < p.N; ++j) {
                  float t = f[j];
                  if (res) t += E::to_f(res[col + j]);
                  out[col + j] = E::from_f(t * p.out_scale);
                }
              }
            }
          } else {  // TCL_EPI_HEADS: split columns into sections (q / k / v^T), heads and head-dim
            const long long b = pix / p.tok_per_batch;
            const long long t = pix - b * p.tok_per_batch;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = col0 + g8 * 8;
              if (col >= p.N) break;
              const int sec = col / p.sec_cols;
              const int within = col - sec * p.sec_cols;
              const int head = within / p.d;
              const int dd = within - head * p.d;  // multiple of 8 because d % 8 == 0
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[j] = __uint_as_float(v[g8 * 8 + j]);
                if (p.bias) f[j] += p.bias[col + j];
              }
              typename E::T* base = reinterpret_cast<typename E::T*>(p.sec_ptr[sec]);
              const long long bh = b * p.heads + head;
              if (!p.sec_vt[sec]) {
                uint4 o;
                o.x = E::pack(f[0], f[1]);
                o.y = E::pack(f[2], f[3]);
                o.z = E::pack(f[4], f[5]);
                o.w = E::pack(f[6], f[7]);
                *reinterpret_cast<uint4*>(base + (bh * p.tok_pitch + t) * p.d_pad + dd) = o;
              } else {
                typename E::T* dst = base + (bh * p.d_pad + dd) * p.tok_pitch + t;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j * p.tok_pitch] = E::from_f(f[j]);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && threadIdx.x == 64) bulk_wait0();   // staging must outlive the last TMA store
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Pick the M-tile box (tw, th, tn), tw*th*tn <= 128, that wastes the fewest MMA rows.
static void choose_tile(int n_img, int h, int w, int* tw_o, int* th_o, int* tn_o) {
  double best = -1.0;
  int btw = 1, bth = 1, btn = 1;
  for (int tw = 1; tw <= w && tw <= 128; ++tw) {
    const int th_max = 128 / tw;
    for (int th = 1; th <= th_max && th <= h; ++th) {
      int tn = 1;
      if (tw == w && th == h) tn = 128 / (tw * th) < n_img ? 128 / (tw * th) : n_img;
      if (tn < 1) tn = 1;
      // box dims are limited to 256 each: always true here
      const long long tiles = static_cast<long long>((w + tw - 1) / tw) * ((h + th - 1) / th) *
                              ((n_img + tn - 1) / tn);
      const double eff = static_cast<double>(n_img) * h * w / (static_cast<double>(tiles) * 128.0);
      // prefer wider boxes on ties (longer contiguous runs per TMA row)
      if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > btw)) {
        best = eff; btw = tw; bth = th; btn = tn;
      }
    }
  }
  *tw_o = btw; *th_o = bth; *tn_o = btn;
}

template <int BN, int STAGES, bool BF16>
static int launch_igemm(const IgTmaps& tm, const IgParams& p, cudaStream_t stream) {
  constexpr size_t smem = STAGES * (IG_BM * IG_BK * 2 + BN * IG_BK * 2) + BN * 256 + 1024 + 256;
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<BN, STAGES, BF16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_last_error("igemm: cudaFuncSetAttribute(%zu B smem) failed: %s", smem, cudaGetErrorString(e));
      return TCL_ERR_CUDA;
    }
    configured = true;
  }
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  igemm_kernel<BN, STAGES, BF16><<<grid, IG_THREADS, smem, stream>>>(tm, p);
  TCL_CHECK_LAUNCH("tcl_igemm");
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_igemm(const tcl_igemm_desc* d, cudaStream_t stream) {
  TCL_CHECK_ARG(d != nullptr, "tcl_igemm: null descriptor");
  TCL_CHECK_ARG(d->num_src >= 1 && d->num_src <= TCL_IGEMM_MAX_SRC, "tcl_igemm: num_src=%d", d->num_src);
  TCL_CHECK_ARG(d->n_img > 0 && d->out_h > 0 && d->out_w > 0 && d->N > 0, "tcl_igemm: empty problem");
  TCL_CHECK_ARG(d->weight != nullptr, "tcl_igemm: null weight");
  const bool bf16 = d->dtype == TCL_DTYPE_BF16;
  TCL_CHECK_ARG(d->dtype == TCL_DTYPE_BF16 || d->dtype == TCL_DTYPE_FP16, "tcl_igemm: dtype");

  IgParams p;
  memset(&p, 0, sizeof(p));
  IgTmaps tm;
  p.n_img = d->n_img; p.out_h = d->out_h; p.out_w = d->out_w; p.N = d->N;
  choose_tile(d->n_img, d->out_h, d->out_w, &p.tw, &p.th, &p.tn);
  p.tiles_w = (p.out_w + p.tw - 1) / p.tw;
  p.tiles_h = (p.out_h + p.th - 1) / p.th;
  p.tiles_n = (p.n_img + p.tn - 1) / p.tn;
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.a_bytes = (uint32_t)(p.tw * p.th * p.tn) * IG_BK * 2;

  // tile width in N
  int BN;
  if (d->mode == TCL_EPI_GEGLU) {
    TCL_CHECK_ARG(d->N % 256 == 0, "tcl_igemm: GEGLU needs N %% 256 == 0 (N=%d)", d->N);
    BN = 256;
  } else if (d->N <= 64) BN = 64;
  else if (d->N % 256 == 0) BN = 256;
  else if (d->N % 160 == 0) BN = 160;
  else if (d->N <= 128) BN = 128;
  else BN = 256;
  p.n_tiles = (d->N + BN - 1) / BN;
  p.total_tiles = p.m_tiles * p.n_tiles;

  long long ktot = 0;
  p.num_src = d->num_src;
  for (int s = 0; s < d->num_src; ++s) {
    const tcl_igemm_src& S = d->src[s];
    TCL_CHECK_ARG(S.ptr != nullptr, "tcl_igemm: src %d null", s);
    TCL_CHECK_ARG(S.taps == 1 || S.taps == 9, "tcl_igemm: src %d taps=%d", s, S.taps);
    TCL_CHECK_ARG(S.c > 0 && S.c % IG_BK == 0, "tcl_igemm: src %d channels %lld not a multiple of 64", s, (long long)S.c);
    TCL_CHECK_ARG(S.pitch >= S.c && S.pitch % 8 == 0, "tcl_igemm: src %d pitch", s);
    TCL_CHECK_ARG(S.stride == 1 || S.stride == 2, "tcl_igemm: src %d stride", s);
    TCL_CHECK_ARG((reinterpret_cast<uintptr_t>(S.ptr) & 15) == 0, "tcl_igemm: src %d misaligned", s);
    p.src[s].taps = S.taps;
    p.src[s].cchunks = (int)(S.c / IG_BK);
    p.src[s].stride = S.stride;
    p.src[s].pad = (S.taps == 9 && !S.no_lead_pad) ? 1 : 0;
    ktot += (long long)S.taps * S.c;
    const uint64_t dims[4] = {(uint64_t)S.c, (uint64_t)S.w, (uint64_t)S.h, (uint64_t)S.n};
    const uint64_t strides[3] = {(uint64_t)S.pitch * 2, (uint64_t)S.w * S.pitch * 2,
                                 (uint64_t)S.h * S.w * S.pitch * 2};
    const uint32_t st = (uint32_t)S.stride;
    const uint32_t box[4] = {IG_BK, (uint32_t)p.tw * st - (st - 1), (uint32_t)p.th * st - (st - 1), (uint32_t)p.tn};
    const uint32_t estr[4] = {1, st, st, 1};
    int rc = make_tmap(&tm.a[s], S.ptr, bf16, 4, dims, strides, box, estr, 128);
    if (rc) return rc;
  }
  TCL_CHECK_ARG(ktot == d->K, "tcl_igemm: K=%lld but segments sum to %lld", (long long)d->K, ktot);
  p.k_blocks = (int)(ktot / IG_BK);
  {
    const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->N};
    const uint64_t strides[1] = {(uint64_t)ktot * 2};
    const uint32_t box[2] = {IG_BK, (uint32_t)BN};
    const uint32_t estr[2] = {1, 1};
    int rc = make_tmap(&tm.b, d->weight, bf16, 2, dims, strides, box, estr, 128);
    if (rc) return rc;
  }

  p.mode = d->mode;
  p.bias = d->bias;
  p.residual = d->residual;
  p.res_pitch = d->res_pitch;
  p.out = d->out;
  p.out_pitch = d->out_pitch;
  p.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  if (d->mode == TCL_EPI_HEADS) {
    TCL_CHECK_ARG(d->d % 8 == 0 && d->d_pad >= d->d && d->heads > 0, "tcl_igemm: head split d=%d d_pad=%d", d->d, d->d_pad);
    TCL_CHECK_ARG(d->sec_cols == d->heads * d->d && d->N % d->sec_cols == 0 && d->N / d->sec_cols <= 3,
                  "tcl_igemm: sections");
    TCL_CHECK_ARG(d->tok_per_batch > 0 && d->tok_pitch >= d->tok_per_batch, "tcl_igemm: tokens");
    for (int i = 0; i < d->N / d->sec_cols; ++i) {
      TCL_CHECK_ARG(d->sec_ptr[i] != nullptr, "tcl_igemm: section %d null", i);
      p.sec_ptr[i] = d->sec_ptr[i];
      p.sec_vt[i] = d->sec_vt[i];
    }
    p.sec_cols = d->sec_cols; p.heads = d->heads; p.d = d->d; p.d_pad = d->d_pad;
    p.tok_per_batch = d->tok_per_batch; p.tok_pitch = d->tok_pitch;
  } else {
    TCL_CHECK_ARG(d->out != nullptr, "tcl_igemm: null out");
    p.vec_ok = (d->out_pitch % 8 == 0) && (d->residual == nullptr || d->res_pitch % 8 == 0) &&
               ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0);
    if (d->mode == TCL_EPI_GEGLU) TCL_CHECK_ARG(p.vec_ok, "tcl_igemm: GEGLU output must be 16-byte aligned with pitch %% 8 == 0");
    const int n_out = d->mode == TCL_EPI_GEGLU ? d->N / 2 : d->N;
    p.tma_store = p.vec_ok && (n_out % 8 == 0) && (d->bias == nullptr || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0);
    if (p.tma_store) {
      const uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n_img};
      const uint64_t strides[3] = {(uint64_t)d->out_pitch * 2, (uint64_t)d->out_w * d->out_pitch * 2,
                                   (uint64_t)d->out_h * d->out_w * d->out_pitch * 2};
      const uint32_t box[4] = {32, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
      const uint32_t estr[4] = {1, 1, 1, 1};
      int rc = make_tmap(&tm.c, d->out, bf16, 4, dims, strides, box, estr, 64);
      if (rc) return rc;
    }
  }

#define TCL_IG_DISPATCH(BN_, ST_)                                            \
  return bf16 ? launch_igemm<BN_, ST_, true>(tm, p, stream)                  \
              : launch_igemm<BN_, ST_, false>(tm, p, stream)
  switch (BN) {
    case 64: TCL_IG_DISPATCH(64, 8);
    case 128: TCL_IG_DISPATCH(128, 6);
    case 160: TCL_IG_DISPATCH(160, 5);
    default: TCL_IG_DISPATCH(256, 3);
  }
#undef TCL_IG_DISPATCH
}
