// tcgen05 attention for the SD-1.5 transformer blocks (diffusers
// Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention, call site
// utils/model_utils.py:66; restated in utils/VidToMe/pnp_utils.py:40-97).
//   O[b, t, h*d + :] = softmax(Q K^T / sqrt(d)) V      (no mask, no dropout)
// Operands come from tcl_igemm's head-split epilogue:
//   Q  [B*H, tq_pitch, d_pad]   K [Bkv*H, tk_pitch, d_pad]   V^T [Bkv*H, d_pad, tk_pitch]
// with the head dim zero-padded to d_pad in {64,128,192} (head_dim 40/80/160) so every tile is
// a 128-byte-swizzled K-major UMMA operand.
//
// Single-pass online softmax with lazy rescaling: per KV tile S = Q K^T lands in TMEM, the
// softmax thread that owns a row takes the row max, forms P = exp2(S*scale - m_ref) in registers
// (m_ref is only advanced when the running max grows by more than 2^8, so P <= 256 and the
// fp32 accumulator is rescaled rarely), stores P as 16-bit into swizzled smem and the MMA warp
// accumulates O += P V in TMEM.  The softmax denominator rides along as one extra output column:
// row `d` of every V^T tile is overwritten with ones in shared memory, so O[:, d] = sum_j P_ij
// (from the same rounded P as the numerator) and is rescaled together with O.
// Warp roles: NQ softmax warpgroups (one 128-row Q tile each), one TMA warp, NQ MMA-issuer warps (one per Q
// tile, running converged with one elected lane issuing).
#include "common.cuh"
#include "tmap.h"
#include "tclight.h"

namespace tcl {

struct AttnParams {
  int tq, tk;          // valid query / key rows per (batch, head)
  int heads, d;        // true head dim
  int kv_batch_div;    // kv batch = q batch / kv_batch_div (cross-attention text broadcast)
  int n_kv_tiles;
  int qk_steps;        // K = 16 MMA steps of Q K^T that carry live head-dim columns: ceil(d / 16)
  int n_o;             // live rows of a V^T tile = N of the P V MMA: round_up(d + 1, 16) (row d = denominator)
  uint32_t idesc_o;    // instruction descriptor of the P V MMA (M = 128, N = n_o)
  float scale_log2;    // log2(e) / sqrt(d)
  void* out;           // [B, tq, heads*d]
  long long out_pitch; // heads*d
  long long* trace;    // TCL_ATTN_TRACE builds only: clock64 event log of CTA (0,0)
  // 1-D grid: CTA id -> (work item = cta_base + id / kv_parts, KV part = id % kv_parts); item -> (bh, Q tile group)
  int q_groups;        // Q tile groups (NQ tiles each) per (batch, head)
  int cta_base;        // first work item of this launch
  int kv_parts;        // 1 = a CTA walks the whole key range and writes normalised output; > 1 = KV-split tail (below)
  float* part_o;       // kv_parts > 1: un-normalised partial O [item - cta_base][part][NQ*128 rows][64] fp32 (column d = denominator)
  float* part_m;       //               reference maximum (scaled, log2 domain) [item - cta_base][part][NQ*128 rows]
};

#ifdef TCL_ATTN_TRACE
#define TRACE_EV(role, ev, j)                                                                          \
  do {                                                                                                 \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (j) >= 64 && (j) < 72)                        \
      p.trace[(((role) * 8 + ((j) - 64)) * 8 + (ev))] = clock64();                                     \
  } while (0)
#else
#define TRACE_EV(role, ev, j) do {} while (0)
#endif

struct AttnTmaps {
  CUtensorMap q, k, vt;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], 2^f by a minimax
// polynomial (degree 3: rel. error 1.0e-4 for bf16 P; degree 4: 3.5e-6 for fp16 P), exponent added as
// integer.  Valid for -126 <= x <= ~100; callers pass x <= 8.
template <bool BF16>
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;          // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.f);
  float pl;
  if (BF16) {
    pl = fmaf(0.055838283f, f, 0.24263948f);
    pl = fmaf(pl, f, 0.69313675f);
    pl = fmaf(pl, f, 0.99992454f);
  } else {
    pl = fmaf(0.009666368f, f, 0.055921976f);
    pl = fmaf(pl, f, 0.2402235f);
    pl = fmaf(pl, f, 0.693121f);
    pl = fmaf(pl, f, 1.0f);
  }
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}

// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (P, 16-bit pairs packed in 32-bit TMEM columns, one row per
// lane) is read straight from tensor memory, so P never crosses shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Packed-pair variants (Blackwell FFMA2 / FADD2: two fp32 lanes per issue slot).  The FMA pipe does 128 lanes/clk/SM either
// way (tools/ubench/pipes.cu), but the packed forms halve the issue slots of the scale and polynomial arithmetic, which
// is what the softmax warps run out of next to MUFU.EX2 (16 lanes/clk/SM).
template <bool BF16>
__device__ __forceinline__ float2 poly_exp2_pair(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(r, make_float2(-1.f, -1.f), x);          // x - round(x), exact
  float2 pl;
  if (BF16) {
    pl = __ffma2_rn(make_float2(0.055838283f, 0.055838283f), f, make_float2(0.24263948f, 0.24263948f));
    pl = __ffma2_rn(pl, f, make_float2(0.69313675f, 0.69313675f));
    pl = __ffma2_rn(pl, f, make_float2(0.99992454f, 0.99992454f));
  } else {
    pl = __ffma2_rn(make_float2(0.009666368f, 0.009666368f), f, make_float2(0.055921976f, 0.055921976f));
    pl = __ffma2_rn(pl, f, make_float2(0.2402235f, 0.2402235f));
    pl = __ffma2_rn(pl, f, make_float2(0.693121f, 0.693121f));
    pl = __ffma2_rn(pl, f, make_float2(1.0f, 1.0f));
  }
  return make_float2(__int_as_float(__float_as_int(pl.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(pl.y) + (__float_as_int(t.y) << 23)));
}

// P is handed to the P V MMA through TMEM (tcgen05.st + A-from-TMEM MMA), never through shared memory.
// POLY  : 8-bit mask; bit b set = unit b of every 8 consecutive units runs exp2 as an FMA-pipe polynomial instead of
//         MUFU.EX2.  A unit is one score (PACKED = false) or one pair of adjacent scores (PACKED = true).
// PACKED: scale-subtract and polynomial arithmetic in packed fp32x2 instructions.
// STALE : tiles after the first form P against the reference of the previous tile while their own row maximum is reduced
//         concurrently (the max leaves the dependent chain in front of the exponentials).  The score tile is only
//         released to the next Q K^T once the maximum is known, so a row whose maximum jumped by more than the
//         representable headroom re-reads its scores from TMEM and redoes the tile against the new reference.
// Register budget: with two Q tiles the block is three warpgroups (2 x softmax, 1 x {TMA warp, 2 MMA-issuer warps, 1 idle
// warp}); after setup the service warpgroup shrinks to 88 registers per thread (setmaxnreg.dec) and the softmax
// warpgroups grow to 200 (setmaxnreg.inc), which is what lets a whole 128-column score row, its packed P and the
// temporaries of the overlapped max / exp2 stay in registers (at the 168 the launch bound allows they spill).
template <int NQ, int REGS>
constexpr int attn_threads() { return REGS > 0 ? 384 : NQ * 128 + 32 + NQ * 32; }

template <int NQ, int DPAD, int KST, int VST, bool BF16, int POLY, bool PACKED, bool STALE, int MINB, int REGS, bool SPLIT = false>
__global__ void __launch_bounds__(attn_threads<NQ, REGS>(), MINB)
attn_kernel(const __grid_constant__ AttnTmaps tm, const __grid_constant__ AttnParams p) {
  using E = Elem<BF16>;
  constexpr int NC = DPAD / 64;                    // 64-wide chunks of the head dim
  constexpr uint32_t QK_TILE = 128 * DPAD * 2;     // one Q or K tile (NC chunk tiles of 16 KB)
  constexpr uint32_t V_CHUNK = DPAD * 128;         // one 64-kv chunk of V^T: up to DPAD rows x 128 B (n_o rows are loaded)
  constexpr uint32_t V_TILE = 2 * V_CHUNK;
  constexpr uint32_t OFF_Q = 0;
  constexpr uint32_t OFF_K = OFF_Q + NQ * QK_TILE;
  constexpr uint32_t OFF_V = OFF_K + KST * QK_TILE;
  constexpr uint32_t OFF_BAR = OFF_V + VST * V_TILE;
  constexpr uint32_t TMEM_P = NQ * (128 + DPAD);      // 64 columns of packed P per Q tile
  constexpr uint32_t TMEM_NEED = NQ * (128 + DPAD + 64);
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
  constexpr int SOFT_THREADS = NQ * 128;
  // largest jump of a row maximum that P (16-bit) and the fp32 accumulator absorb without the slow path
  constexpr float STALE_LIMIT = BF16 ? 60.f : 6.f;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = q_full + 1;            // KST
  uint64_t* k_empty = k_full + KST;         // KST
  uint64_t* v_full = k_empty + KST;         // VST
  uint64_t* v_empty = v_full + VST;         // VST
  uint64_t* s_full = v_empty + VST;         // NQ
  uint64_t* s_empty = s_full + NQ;          // NQ
  uint64_t* p_full = s_empty + NQ;          // NQ
  uint64_t* p_empty = p_full + NQ;          // NQ
  uint64_t* o_full = p_empty + NQ;          // NQ
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NQ);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Normal launches: 2-D grid (Q tile group, batch*head) — measured 1.5-2.5 % faster than the same items on a 1-D grid; work
  // items from p.cta_base on belong to the KV-split tail launch (SPLIT, 1-D grid: item = cta_base + id / kv_parts).
  const int kv_parts = SPLIT ? p.kv_parts : 1;
  const int part = SPLIT ? blockIdx.x % kv_parts : 0;
  const int item = SPLIT ? p.cta_base + blockIdx.x / kv_parts : blockIdx.y * gridDim.x + blockIdx.x;
  if (!SPLIT && item >= p.cta_base) return;        // (cta_base = all items when nothing is split off)
  const int bh = SPLIT ? item / p.q_groups : blockIdx.y;
  const int b = bh / p.heads, head = bh - b * p.heads;
  const int kv_bh = (b / p.kv_batch_div) * p.heads + head;
  const int q_row0 = (SPLIT ? item - bh * p.q_groups : blockIdx.x) * (NQ * 128);
  const int kv0 = SPLIT ? (int)((long long)part * p.n_kv_tiles / kv_parts) : 0;          // this CTA's KV tiles [kv0, kv0 + n_kv)
  const int n_kv = SPLIT ? (int)((long long)(part + 1) * p.n_kv_tiles / kv_parts) - kv0 : p.n_kv_tiles;

  if (threadIdx.x == SOFT_THREADS) {
    tma_prefetch_desc(&tm.q);
    tma_prefetch_desc(&tm.k);
    tma_prefetch_desc(&tm.vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < KST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], NQ); }
    for (int i = 0; i < VST; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], NQ); }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == SOFT_THREADS / 32 + 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // warp-uniform for the compiler (a plain shared-memory load is treated as divergent)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= SOFT_THREADS / 32) {
  // the service warpgroup gives registers back (every warp of the warpgroup executes the same setmaxnreg)
  if constexpr (REGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == SOFT_THREADS / 32) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t v_bytes = 2u * (uint32_t)p.n_o * 128u;          // only the n_o live rows of V^T travel
      mbar_arrive_expect_tx(q_full, NQ * QK_TILE);
      for (int q = 0; q < NQ; ++q)
        for (int c = 0; c < NC; ++c)
          tma_load_3d(smem + OFF_Q + q * QK_TILE + c * 16384, &tm.q, q_full, c * 64, q_row0 + q * 128, bh);
      int kn = 0, vn = 0;
      for (int j = 0; j < n_kv; ++j) {
        {
          const int st = kn % KST;
          mbar_wait(&k_empty[st], ((kn / KST) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], QK_TILE);
          for (int c = 0; c < NC; ++c)
            tma_load_3d(smem + OFF_K + st * QK_TILE + c * 16384, &tm.k, &k_full[st], c * 64, (kv0 + j) * 128, kv_bh);
          ++kn;
        }
        {
          const int st = vn % VST;
          mbar_wait(&v_empty[st], ((vn / VST) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[st], v_bytes);
          for (int c = 0; c < 2; ++c)
            tma_load_3d(smem + OFF_V + st * V_TILE + c * V_CHUNK, &tm.vt, &v_full[st], (kv0 + j) * 128 + c * 64, 0, kv_bh);
          ++vn;
        }
      }
    }
  } else if (warp > SOFT_THREADS / 32 && warp <= SOFT_THREADS / 32 + NQ) {
    // ===================== MMA issuers: one converged warp per Q tile =====================
    // All 32 lanes run the control flow (so every operand stays in uniform registers and no per-MMA
    // ELECT/R2UR.BROADCAST waterfall is generated — with a single `if (lane == 0)` issuer that waterfall plus
    // ~150-clk barrier probes made this thread, not the softmax, the critical path: 3 800 clk per KV tile);
    // one elected lane issues the tcgen05 instructions.
    const int q = warp - (SOFT_THREADS / 32 + 1);
    constexpr uint32_t idesc_s = umma_idesc_f16(BF16, 128, 128);
    const uint32_t idesc_o = p.idesc_o;
    const int qk_steps = p.qk_steps;
    const uint32_t q_base = smem_u32(smem + OFF_Q) + q * QK_TILE;
    const uint32_t k_base = smem_u32(smem + OFF_K);
    const uint32_t v_base = smem_u32(smem + OFF_V);
    const uint32_t t_s = tmem_base + q * 128;
    const uint32_t t_o = tmem_base + NQ * 128 + q * DPAD;
    const uint32_t t_p = tmem_base + TMEM_P + q * 64;

    auto issue_qk = [&](int jn) {
      const int st = jn % KST;
      mbar_wait(&k_full[st], (jn / KST) & 1);
      TRACE_EV(2, q * 4 + 0, jn);
      mbar_wait(&s_empty[q], (jn & 1) ^ 1);
      TRACE_EV(2, q * 4 + 1, jn);
      tcgen05_fence_after();
      if (elect_one()) {
        // only the K = 16 steps that carry live head-dim columns (d = 40: 3 of the 4 steps of the 64-wide chunk)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint64_t a = umma_desc_k_sw128(q_base + c * 16384);
          const uint64_t bd = umma_desc_k_sw128(k_base + st * QK_TILE + c * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c * 4 + k < qk_steps) umma_f16_ss(t_s, a + 2 * k, bd + 2 * k, idesc_s, (c | k) != 0);
        }
        umma_commit(&s_full[q]);
        umma_commit(&k_empty[st]);       // k_empty counts NQ commits
      }
      __syncwarp();
    };

    mbar_wait(q_full, 0);
    tcgen05_fence_after();
    issue_qk(0);                          // scores run one tile ahead of P V
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j % VST;
      mbar_wait(&v_full[st], (j / VST) & 1);
      // row `d` of V^T := 1  => O[:, d] accumulates the softmax denominator (swizzle only permutes 16-byte units
      // inside a 128-byte row, so filling the whole row is layout-safe).  Every issuer warp writes the same
      // values; 16 lanes store one 16-byte unit each.
      if (lane < 16) {
        const uint32_t one2 = BF16 ? 0x3F803F80u : 0x3C003C00u;
        const uint32_t rowa = v_base + st * V_TILE + (lane >> 3) * V_CHUNK + p.d * 128 + (lane & 7) * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rowa), "r"(one2) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      TRACE_EV(2, q * 4 + 2, j);
      mbar_wait(&p_full[q], j & 1);
      TRACE_EV(2, q * 4 + 3, j);
      tcgen05_fence_after();
      if (elect_one()) {
        // N = n_o (d + 1 rounded up to 16): the zero-padded rows of V^T are neither loaded nor multiplied
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t bd = umma_desc_k_sw128(v_base + st * V_TILE + c * V_CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(t_o, t_p + (c * 4 + k) * 8, bd + 2 * k, idesc_o, (j | c | k) != 0);
        }
        umma_commit(&p_empty[q]);
        umma_commit(&v_empty[st]);       // v_empty counts NQ commits
        if (j + 1 == n_kv) umma_commit(&o_full[q]);
      }
      __syncwarp();
    }
  }
  } else {
    // ===================== softmax warpgroups =====================
    if constexpr (REGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
    const int q = warp >> 2;                 // which Q tile
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + q * 128;
    const uint32_t t_o = t_lane + NQ * 128 + q * DPAD;
    const uint32_t t_p = t_lane + TMEM_P + q * 64;
    float m_ref = -INFINITY;                 // reference max (scaled, log2 domain) used by exp2
    float pend = 1.f;                        // STALE: factor owed to O from the previous tile's maximum
    const bool tracer = threadIdx.x == q * 128;
    const float sc = p.scale_log2;

    // exp2(s * scale - ref) of 32 scores -> 16 packed 16-bit pairs
    auto exp_chunk = [&](const uint32_t (&s)[32], uint32_t* pk, float ref) {
      if constexpr (PACKED) {
        const float2 sc2 = make_float2(sc, sc), nref2 = make_float2(-ref, -ref);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, nref2);
          float2 e;
          if ((POLY >> ((i >> 1) & 7)) & 1) e = poly_exp2_pair<BF16>(a);
          else e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
          pk[i >> 1] = E::pack(e.x, e.y);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = fmaf(__uint_as_float(s[i]), sc, -ref);
          const float a1 = fmaf(__uint_as_float(s[i + 1]), sc, -ref);
          const float e0 = ((POLY >> (i & 7)) & 1) ? poly_exp2<BF16>(a0) : fast_exp2(a0);
          const float e1 = ((POLY >> ((i + 1) & 7)) & 1) ? poly_exp2<BF16>(a1) : fast_exp2(a1);
          pk[i >> 1] = E::pack(e0, e1);
        }
      }
    };
    auto max_chunk = [&](const uint32_t (&s)[32], float& a, float& bq) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        a = fmaxf(a, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
        bq = fmaxf(bq, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
      }
    };
    auto mask_chunk = [&](uint32_t (&s)[32], int c0, int valid_cols) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i >= valid_cols) s[i] = 0xff800000u;   // -inf
    };

    // A barrier probe costs ~250 clk even when the phase is long complete (measured with the clock64 trace), so
    // both per-tile barriers are probed early with the non-blocking form and the result is consumed later: the
    // blocking wait only runs when the early probe failed.
    bool s_ready = false;
    for (int j = 0; j < n_kv; ++j) {
      if (tracer) TRACE_EV(q, 0, j);
      if (!s_ready) mbar_wait(&s_full[q], j & 1);
      if (tracer) TRACE_EV(q, 1, j);
      tcgen05_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld_32x32b_x32(t_s + 0, s0);
      tmem_ld_32x32b_x32(t_s + 32, s1);
      tmem_ld_32x32b_x32(t_s + 64, s2);
      tmem_ld_32x32b_x32(t_s + 96, s3);
      const bool pe_ready = mbar_test_wait(&p_empty[q], (j & 1) ^ 1);   // P V of the previous tile: consumed after the exps
      tmem_ld_wait();
      const int valid_cols = p.tk - (kv0 + j) * 128;
      const bool tail = valid_cols < 128;      // warp-uniform: only the last KV tile
      if (tail) { mask_chunk(s0, 0, valid_cols); mask_chunk(s1, 32, valid_cols); mask_chunk(s2, 64, valid_cols); mask_chunk(s3, 96, valid_cols); }
      uint32_t pk[64];
      float factor = 1.f;
      if (!STALE || j == 0) {
        // exact reference: row maximum first (independent chains: one 128-long dependent chain costs ~500 clk)
        tcgen05_fence_before();
        mbar_arrive(&s_empty[q]);
        if (tracer) TRACE_EV(q, 2, j);
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        max_chunk(s0, mx0, mx1); max_chunk(s1, mx2, mx3); max_chunk(s2, mx0, mx1); max_chunk(s3, mx2, mx3);
        const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        if (j == 0) {
          m_ref = m_new;
        } else if (m_new > m_ref + 8.f) {
          factor = fast_exp2(m_ref - m_new);
          m_ref = m_new;
        }
        exp_chunk(s0, pk + 0, m_ref);
        exp_chunk(s1, pk + 16, m_ref);
        exp_chunk(s2, pk + 32, m_ref);
        exp_chunk(s3, pk + 48, m_ref);
      } else {
        // stale reference: the first quarter of the exponentials runs next to the max reduction of the whole tile
        factor = pend;
        pend = 1.f;
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        max_chunk(s0, mx0, mx1); max_chunk(s1, mx2, mx3); max_chunk(s2, mx0, mx1); max_chunk(s3, mx2, mx3);
        exp_chunk(s0, pk + 0, m_ref);
        const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        float m_next = m_ref;
        if (__any_sync(0xffffffffu, m_new > m_ref + STALE_LIMIT)) {
          // rare: this tile would overflow against the old reference -> move the reference NOW (O is rescaled before
          // this tile's P V) and redo the first quarter from the scores still held in TMEM
          if (m_new > m_ref + 8.f) {
            factor *= fast_exp2(m_ref - m_new);
            m_ref = m_new;
            m_next = m_new;
          }
          tmem_ld_32x32b_x32(t_s + 0, s0);
          tmem_ld_wait();
          if (tail) mask_chunk(s0, 0, valid_cols);
          exp_chunk(s0, pk + 0, m_ref);
        } else if (m_new > m_ref + 8.f) {
          pend = fast_exp2(m_ref - m_new);     // applied to O before the NEXT tile's P V
          m_next = m_new;
        }
        tcgen05_fence_before();
        mbar_arrive(&s_empty[q]);
        if (tracer) TRACE_EV(q, 2, j);
        exp_chunk(s1, pk + 16, m_ref);
        exp_chunk(s2, pk + 32, m_ref);
        exp_chunk(s3, pk + 48, m_ref);
        m_ref = m_next;
      }
      // the previous P V MMA must be done before P or O are touched.  Waiting here (not before the exponentials)
      // gives it the whole softmax of this tile to complete: the ncu source view of the earlier placement showed
      // a third of all stall samples on this barrier.
      if (tracer) TRACE_EV(q, 3, j);
      s_ready = (j + 1 < n_kv) && mbar_test_wait(&s_full[q], (j + 1) & 1);   // next scores: consumed at the loop top
      if (!pe_ready) mbar_wait(&p_empty[q], (j & 1) ^ 1);
      if (tracer) TRACE_EV(q, 4, j);
      if (__any_sync(0xffffffffu, factor != 1.f)) {
        tcgen05_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < p.n_o; c0 += 16) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(t_o + c0, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st_32x32b_x16(t_o + c0, o);
        }
        tmem_st_wait();
      }
      // P -> TMEM: lane = row, 32-bit column c holds keys (2c, 2c+1) — the K-major A operand of the P V MMA
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[32]);
        tmem_st_32x32b_x32(t_p, lo);
        tmem_st_32x32b_x32(t_p + 32, hi);
        tmem_st_wait();
      }
      tcgen05_fence_before();
      mbar_arrive(&p_full[q]);
      if (tracer) TRACE_EV(q, 5, j);
    }
    // ---- epilogue: O[:, :d] / O[:, d] ----
    mbar_wait(&o_full[q], 0);
    tcgen05_fence_after();
    const int t = q_row0 + q * 128 + row;
    if (SPLIT) {
      // KV-split tail: this CTA saw only a slice of the keys.  O is consistent with m_ref here (the lazy rescale is applied
      // before every P V; the stale-reference variants, which owe a factor, are never split), so (O, m_ref) can be merged.
      const long long prow = ((long long)(blockIdx.x / kv_parts) * kv_parts + part) * (NQ * 128) + q * 128 + row;
      float* po = p.part_o + prow * 64;
      p.part_m[prow] = m_ref;
#pragma unroll 1
      for (int c0 = 0; c0 < p.n_o; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_o + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(po + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                                                __uint_as_float(v[i + 3]));
      }
    } else {
    float inv_l;
    {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_o + (p.d & ~15), v);
      tmem_ld_wait();
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.d & 15)) l = __uint_as_float(v[i]);
      inv_l = 1.0f / l;        // (a factor still owed to O after the last tile cancels in O / l)
    }
    typename E::T* out = reinterpret_cast<typename E::T*>(p.out) +
                         (static_cast<long long>(b) * p.tq + t) * p.out_pitch + head * p.d;
#pragma unroll 1
    for (int c0 = 0; c0 < p.d; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_o + c0, v);
      tmem_ld_wait();
      if (t < p.tq) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = E::pack(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
        *reinterpret_cast<uint4*>(out + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        if (c0 + 8 < p.d) *reinterpret_cast<uint4*>(out + c0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == SOFT_THREADS / 32 + 1) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Persistent cross-attention on the text tokens (L = 154 or 77 -> 160 / 80 key columns; d_pad = 64, i.e. the ds-1 blocks
// that carry 90 % of the cross-attention bytes).  HBM-bound: Q in, O out, nothing else.
// One CTA per SM walks a contiguous range of (batch*head, Q tile) work items; the K / V^T tiles of the current
// (text, head) stay resident in shared memory (39 KB) and are re-loaded only when the range crosses into another
// (text, head); Q tiles stream through a 4-deep TMA ring; one TMEM allocation per CTA holds two {S|P, O} buffers, one
// per softmax group, so softmax + epilogue of tile i overlap the MMAs of tile i+1.  Two issuer warps (Q K^T, P V) so that
// neither kind of MMA waits behind the other's operands.
// The softmax runs on 16 warps = FOUR threads per score row (warps w, w+4, w+8, w+12 share TMEM lanes 32*(w%4)...): each
// thread holds a quarter of the row's keys in registers after ONE tcgen05.ld round trip (a one-thread-per-row version
// spent its time in 13 dependent TMEM round trips per tile at 17 % occupancy), the quarter maxima meet in shared memory,
// P = exp2(S*scale - max) is written IN PLACE over S (16-bit pairs; every quarter has read S before any writes), O = P V
// with P as the TMEM A operand, the denominator riding along as the ones row of V^T like in the self-attention kernel.
// ------------------------------------------------------------------------------------------
struct XAttnParams {
  int tq, tk, heads, d, kv_batch_div;
  int q_tiles, total_tiles, tiles_per_cta;
  int n_keys;                    // 160 or 80: N of Q K^T and K of P V
  int qk_steps, n_o;
  uint32_t idesc_s, idesc_o;
  float scale_log2;
  void* out;
  long long out_pitch;
};

constexpr int XQST = 4;                       // Q ring depth
constexpr uint32_t X_OFF_K = 0;               // 160 keys x 128 B
constexpr uint32_t X_OFF_V = 20480;           // 3 chunks of 64 keys: 64 rows x 128 B each (n_o rows loaded)
constexpr uint32_t X_VCHUNK = 8192;
constexpr uint32_t X_OFF_Q = X_OFF_V + 3 * X_VCHUNK;
constexpr uint32_t X_OFF_MAX = X_OFF_Q + XQST * 16384;      // float [2 buffers][4 quarters][128 rows]
constexpr uint32_t X_OFF_BAR = X_OFF_MAX + 2 * 4 * 128 * 4;
constexpr size_t X_SMEM = X_OFF_BAR + 256 + 1024;
constexpr int X_THREADS = 16 * 32 + 4 * 32;     // 4 softmax warpgroups + one service warpgroup (TMA, Q K^T, P V, idle)

// N (32 / 16 / 8) scores -> exp2 -> N/2 packed columns, stored at once so the scores' registers retire early
template <bool BF16, int POLY, int N>
__device__ __forceinline__ void xexp_store(const uint32_t* s, uint32_t dst, float2 sc2, float2 nref2) {
  using E = Elem<BF16>;
  constexpr int NP = N / 2 < 8 ? 8 : N / 2;
  uint32_t pk[NP];
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    const float2 a = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, nref2);
    float2 e;
    if ((POLY >> ((i >> 1) & 7)) & 1) e = poly_exp2_pair<BF16>(a);
    else e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
    pk[i >> 1] = E::pack(e.x, e.y);
  }
  if constexpr (N == 32) tmem_st_32x32b_x16(dst, pk);
  else if constexpr (N == 16) tmem_st_32x32b_x8(dst, pk);
  else {
    // 8 keys = 4 packed columns: padded to an x8 store with zeros (the extra columns belong to padding keys)
#pragma unroll
    for (int k = 4; k < 8; ++k) pk[k] = 0u;
    tmem_st_32x32b_x8(dst, pk);
  }
}

// one quarter row: NK keys starting at S column k_off; `valid` = how many of them are real keys (the rest is padding).
// xm = this buffer's [4 quarters][128 rows] maxima.
template <bool BF16, int POLY, int NK>
__device__ __forceinline__ void xsoft_part(uint32_t t_s, int k_off, int valid, float sc, float* xm, int quarter, int row) {
  uint32_t s[NK];
  tmem_ld_cols<NK>(t_s + k_off, s);
  tmem_ld_wait();
  if (valid < NK) {
#pragma unroll
    for (int k = 0; k < NK; ++k)
      if (k >= valid) s[k] = 0xff800000u;   // -inf
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int k = 0; k < NK; k += 4) {
    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[k]), __uint_as_float(s[k + 1])));
    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[k + 2]), __uint_as_float(s[k + 3])));
  }
  xm[quarter * 128 + row] = fmaxf(mx0, mx1);
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const float ref = fmaxf(fmaxf(xm[row], xm[128 + row]), fmaxf(xm[256 + row], xm[384 + row])) * sc;
  const float2 sc2 = make_float2(sc, sc), nref2 = make_float2(-ref, -ref);
  constexpr int n32 = NK / 32, rem = NK % 32;
#pragma unroll
  for (int c = 0; c < n32; ++c) xexp_store<BF16, POLY, 32>(&s[32 * c], t_s + ((k_off + 32 * c) >> 1), sc2, nref2);
  if constexpr (rem >= 16) xexp_store<BF16, POLY, 16>(&s[32 * n32], t_s + ((k_off + 32 * n32) >> 1), sc2, nref2);
  static_assert(rem % 16 == 0, "quarter widths are multiples of 16");
  tmem_st_wait();
}

template <bool BF16, int POLY, int K0, int K1, int K2, int K3>
__global__ void __launch_bounds__(X_THREADS, 1)
xattn_kernel(const __grid_constant__ AttnTmaps tm, const __grid_constant__ XAttnParams p) {
  using E = Elem<BF16>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X_OFF_BAR);
  uint64_t* kv_full = bars;                 // 1
  uint64_t* kv_empty = kv_full + 1;         // 1
  uint64_t* q_full = kv_empty + 1;          // XQST
  uint64_t* q_empty = q_full + XQST;        // XQST
  uint64_t* s_full = q_empty + XQST;        // 2
  uint64_t* p_full = s_full + 2;            // 2
  uint64_t* o_full = p_full + 2;            // 2
  uint64_t* s_free = o_full + 2;            // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  float* xmax = reinterpret_cast<float*>(smem + X_OFF_MAX);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.total_tiles);
  const int n_tiles = t_end - t_begin;

  auto kv_of = [&](int t) { const int bh = t / p.q_tiles; const int b = bh / p.heads; return (b / p.kv_batch_div) * p.heads + (bh - b * p.heads); };

  if (threadIdx.x == 512) {
    tma_prefetch_desc(&tm.q);
    tma_prefetch_desc(&tm.k);
    tma_prefetch_desc(&tm.vt);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < XQST; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 16); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 16); }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int n_vchunks = (p.n_keys + 63) / 64;

  if (warp >= 16) {
  if (warp == 16) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_tiles > 0) {
      const uint32_t kv_bytes = (uint32_t)p.n_keys * 128u + (uint32_t)n_vchunks * (uint32_t)p.n_o * 128u;
      int seg = 0, prev_kv = -1;
      for (int i = 0; i < n_tiles; ++i) {
        const int t = t_begin + i;
        const int kv = kv_of(t);
        if (kv != prev_kv) {
          if (seg > 0) mbar_wait(kv_empty, (seg - 1) & 1);      // every MMA that read the old K / V has completed
          mbar_arrive_expect_tx(kv_full, kv_bytes);
          tma_load_3d(smem + X_OFF_K, &tm.k, kv_full, 0, 0, kv);
          for (int c = 0; c < n_vchunks; ++c) tma_load_3d(smem + X_OFF_V + c * X_VCHUNK, &tm.vt, kv_full, c * 64, 0, kv);
          prev_kv = kv;
          ++seg;
        }
        const int st = i % XQST;
        mbar_wait(&q_empty[st], ((i / XQST) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[st], 16384);
        const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
        tma_load_3d(smem + X_OFF_Q + st * 16384, &tm.q, &q_full[st], 0, qt * 128, bh);
      }
    }
  } else if (warp == 17) {
    // ===================== Q K^T issuer (converged warp, one elected lane issues) =====================
    const uint32_t k_base = smem_u32(smem + X_OFF_K);
    const uint32_t v_base = smem_u32(smem + X_OFF_V);
    const uint32_t q_base = smem_u32(smem + X_OFF_Q);
    int seg = 0, prev_kv = -1;
    for (int i = 0; i < n_tiles; ++i) {
      const int kv = kv_of(t_begin + i);
      if (kv != prev_kv) {
        mbar_wait(kv_full, seg & 1);
        // row `d` of V^T := 1  => O[:, d] = softmax denominator (same rounded P as the numerator)
        if (lane < 8 * n_vchunks) {
          const uint32_t one2 = BF16 ? 0x3F803F80u : 0x3C003C00u;
          const uint32_t rowa = v_base + (lane >> 3) * X_VCHUNK + p.d * 128 + (lane & 7) * 16;
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rowa), "r"(one2) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        prev_kv = kv;
        ++seg;
      }
      const int st = i % XQST, b = i & 1;
      mbar_wait(&q_full[st], (i / XQST) & 1);
      // the S | P columns of this buffer are free as soon as the P V of tile i-2 has retired (its O is read later, from
      // other columns): waiting for the epilogue instead left the whole Q K^T latency exposed in front of every softmax
      if (i >= 2) mbar_wait(&o_full[b], ((i - 2) >> 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t a = umma_desc_k_sw128(q_base + st * 16384);
        const uint64_t bd = umma_desc_k_sw128(k_base);
        for (int k = 0; k < p.qk_steps; ++k) umma_f16_ss(tmem_base + b * 256, a + 2 * k, bd + 2 * k, p.idesc_s, k != 0);
        umma_commit(&s_full[b]);
        umma_commit(&q_empty[st]);
      }
      __syncwarp();
    }
  } else if (warp == 18) {
    // ===================== P V issuer =====================
    const uint32_t v_base = smem_u32(smem + X_OFF_V);
    const int pv_steps = p.n_keys / 16;
    for (int i = 0; i < n_tiles; ++i) {
      const int b = i & 1;
      mbar_wait(&p_full[b], (i >> 1) & 1);
      mbar_wait(&s_free[b], ((i >> 1) & 1) ^ 1);      // the epilogue of tile i-2 has read this buffer's O columns
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t t_o = tmem_base + b * 256 + 192, t_p = tmem_base + b * 256;
        for (int ks = 0; ks < pv_steps; ++ks) {
          const uint64_t bd = umma_desc_k_sw128(v_base + (ks >> 2) * X_VCHUNK) + 2 * (ks & 3);
          umma_f16_ts(t_o, t_p + ks * 8, bd, p.idesc_o, ks != 0);
        }
        umma_commit(&o_full[b]);
        // last tile of a (text, head) segment: once these MMAs retire the K / V tiles may be replaced
        if (i + 1 < n_tiles && kv_of(t_begin + i + 1) != kv_of(t_begin + i)) umma_commit(kv_empty);
      }
      __syncwarp();
    }
  }
  } else {
    // ===================== softmax: 16 warps, four threads per score row =====================
    // Software-pipelined over the two TMEM buffers: softmax(i) then epilogue(i-1), so the P V of tile i runs under the
    // epilogue of tile i-1 and the Q K^T of tile i+1 under the softmax of tile i.
    const int quarter = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float sc = p.scale_log2;
    const int ng = (p.d + 15) >> 4;           // 16-column groups of O; group g is written by quarter g
    for (int i = 0; i <= n_tiles; ++i) {
      if (i < n_tiles) {
        const int b = i & 1;
        mbar_wait(&s_full[b], (i >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t t_s = t_lane + b * 256;
        float* xm = xmax + b * 512;
        switch (quarter) {
          case 0: xsoft_part<BF16, POLY, K0>(t_s, 0, p.tk, sc, xm, 0, row); break;
          case 1: xsoft_part<BF16, POLY, K1>(t_s, K0, p.tk - K0, sc, xm, 1, row); break;
          case 2: xsoft_part<BF16, POLY, K2>(t_s, K0 + K1, p.tk - K0 - K1, sc, xm, 2, row); break;
          default: xsoft_part<BF16, POLY, K3>(t_s, K0 + K1 + K2, p.tk - K0 - K1 - K2, sc, xm, 3, row); break;
        }
        // one arrival per warp (512 per-thread arrivals on one shared-memory word serialise: ~2 000 clk per tile)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
      }
      if (i > 0) {
        // ---- epilogue of tile i-1: O[:, :d] / O[:, d] ----
        const int j = i - 1, b = j & 1;
        mbar_wait(&o_full[b], (j >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t t_o = t_lane + b * 256 + 192;
        const int t = t_begin + j;
        const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
        const int bi = bh / p.heads, head = bh - bi * p.heads;
        const int tok = qt * 128 + row;
        uint32_t lv[16], v[16];
        tmem_ld_32x32b_x16(t_o + (p.d & ~15), lv);
        if (quarter < ng) tmem_ld_32x32b_x16(t_o + 16 * quarter, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[b]);   // this buffer's O is in registers: the P V of tile j+2 may overwrite it
        if (quarter < ng && tok < p.tq) {
          float l = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k == (p.d & 15)) l = __uint_as_float(lv[k]);
          const float inv_l = 1.0f / l;
          typename E::T* out = reinterpret_cast<typename E::T*>(p.out) + (static_cast<long long>(bi) * p.tq + tok) * p.out_pitch + head * p.d;
          uint32_t o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = E::pack(__uint_as_float(v[2 * k]) * inv_l, __uint_as_float(v[2 * k + 1]) * inv_l);
          *reinterpret_cast<uint4*>(out + 16 * quarter) = make_uint4(o[0], o[1], o[2], o[3]);
          if (16 * quarter + 8 < p.d) *reinterpret_cast<uint4*>(out + 16 * quarter + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 17) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool BF16, int POLY, int K0, int K1, int K2, int K3>
static int launch_xattn(const AttnTmaps& tm, const XAttnParams& p, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(xattn_kernel<BF16, POLY, K0, K1, K2, K3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X_SMEM);
    if (e != cudaSuccess) {
      set_last_error("cross-attention: cudaFuncSetAttribute(%zu B) failed: %s", X_SMEM, cudaGetErrorString(e));
      return TCL_ERR_CUDA;
    }
    configured = true;
  }
  xattn_kernel<BF16, POLY, K0, K1, K2, K3><<<grid, X_THREADS, X_SMEM, stream>>>(tm, p);
  TCL_CHECK_LAUNCH("tcl_attention(cross)");
  return TCL_OK;
}

// Merge of the KV-split tail: out[row] = sum_s 2^(m_s - m) O_s / sum_s 2^(m_s - m) l_s  (l_s = column d of O_s).
// grid (rows / 32, tail items), block (8 column groups, 32 rows): thread (cg, r) handles columns cg*8 .. cg*8+7 of one row.
template <bool BF16>
__global__ void attn_merge_kernel(AttnParams p, int rows_per_item) {
  using E = Elem<BF16>;
  const int it = blockIdx.y;
  const int row = blockIdx.x * 32 + threadIdx.y;
  const int item = p.cta_base + it;
  const int bh = item / p.q_groups;
  const int b = bh / p.heads, head = bh - b * p.heads;
  const int t = (item - bh * p.q_groups) * rows_per_item + row;
  if (t >= p.tq) return;
  const int c0 = threadIdx.x * 8;
  if (c0 >= p.d) return;
  const long long base = (long long)it * p.kv_parts * rows_per_item + row;
  float m = -INFINITY;
  for (int s = 0; s < p.kv_parts; ++s) m = fmaxf(m, p.part_m[base + (long long)s * rows_per_item]);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float l = 0.f;
  for (int s = 0; s < p.kv_parts; ++s) {
    const long long pr = base + (long long)s * rows_per_item;
    const float w = fast_exp2(p.part_m[pr] - m);
    const float* po = p.part_o + pr * 64;
    l += w * po[p.d];
    const float4 a = *reinterpret_cast<const float4*>(po + c0), bq = *reinterpret_cast<const float4*>(po + c0 + 4);
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * bq.x; acc[5] += w * bq.y; acc[6] += w * bq.z; acc[7] += w * bq.w;
  }
  const float inv_l = 1.0f / l;
  typename E::T* out = reinterpret_cast<typename E::T*>(p.out) + (static_cast<long long>(b) * p.tq + t) * p.out_pitch + head * p.d + c0;
  *reinterpret_cast<uint4*>(out) = make_uint4(E::pack(acc[0] * inv_l, acc[1] * inv_l), E::pack(acc[2] * inv_l, acc[3] * inv_l),
                                              E::pack(acc[4] * inv_l, acc[5] * inv_l), E::pack(acc[6] * inv_l, acc[7] * inv_l));
}

template <int NQ, int DPAD, int KST, int VST, bool BF16, int POLY, bool PACKED, bool STALE, int MINB = 1, int REGS = 0>
static int launch_attn(const AttnTmaps& tm, const AttnParams& p, int q_tiles, int bh, cudaStream_t stream,
                       void* split_ws = nullptr, size_t split_ws_bytes = 0) {
  constexpr size_t smem = (size_t)NQ * 128 * DPAD * 2 + (size_t)KST * 128 * DPAD * 2 + (size_t)VST * 2 * DPAD * 128 + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<NQ, DPAD, KST, VST, BF16, POLY, PACKED, STALE, MINB, REGS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_last_error("attention: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
      return TCL_ERR_CUDA;
    }
    configured = true;
  }
  static_assert(REGS == 0 || (NQ == 2 && REGS % 8 == 0 && (4 * 88 + 8 * REGS) * 32 <= 65536), "register split");
  AttnParams q = p;
  q.q_groups = (q_tiles + NQ - 1) / NQ;
  const int items = q.q_groups * bh;
  q.cta_base = 0; q.kv_parts = 1;
  int main_items = items;
  // KV-split tail (no stale reference: see the epilogue).  The last, partial wave of CTAs would keep most SMs idle for a whole
  // tile time (T = 47 520, d = 40: 2 976 CTAs = 20.1 waves of 148); its work items are cut into kv_parts slices of the key
  // range that run side by side, and a small kernel merges the (O, max, denominator) partials.
  int parts = 1, tail = 0;
  if (!STALE && split_ws && NQ == 2) {
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    tail = items % sms;
    // only when the tail leaves at least three quarters of the SMs idle: with 2-3 slices the extra launches and the merge
    // cost more than the shorter last wave saves (measured at T = 11 520: 0.605 -> 0.620 ms with 2 slices)
    if (items > sms && tail > 0 && tail * 4 <= sms) {
      parts = sms / tail;
      if (parts > 16) parts = 16;
      while (parts > 1 && p.n_kv_tiles / parts < 8) --parts;          // a slice keeps at least 8 KV tiles
      const size_t need = (size_t)tail * parts * NQ * 128 * (64 + 1) * sizeof(float);
      if (parts > 1 && need <= split_ws_bytes) main_items = items - tail; else parts = 1;
    }
  }
  q.cta_base = main_items;                     // the normal launch skips the items the tail launch takes
  attn_kernel<NQ, DPAD, KST, VST, BF16, POLY, PACKED, STALE, MINB, REGS><<<dim3(q.q_groups, bh), attn_threads<NQ, REGS>(), smem, stream>>>(tm, q);
  TCL_CHECK_LAUNCH("tcl_attention");
  if (parts > 1) {
    q.kv_parts = parts;
    q.part_o = reinterpret_cast<float*>(split_ws);
    q.part_m = q.part_o + (size_t)tail * parts * NQ * 128 * 64;
    if constexpr (!STALE && NQ == 2) {
      static bool configured_split = false;
      if (!configured_split) {
        cudaError_t e = cudaFuncSetAttribute(attn_kernel<NQ, DPAD, KST, VST, BF16, POLY, PACKED, STALE, MINB, REGS, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
          set_last_error("attention: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
          return TCL_ERR_CUDA;
        }
        configured_split = true;
      }
      attn_kernel<NQ, DPAD, KST, VST, BF16, POLY, PACKED, STALE, MINB, REGS, true><<<tail * parts, attn_threads<NQ, REGS>(), smem, stream>>>(tm, q);
    }
    TCL_CHECK_LAUNCH("tcl_attention(tail)");
    attn_merge_kernel<BF16><<<dim3(NQ * 128 / 32, tail), dim3(8, 32), 0, stream>>>(q, NQ * 128);
    TCL_CHECK_LAUNCH("tcl_attention(merge)");
  }
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

// The product library ships ONE configuration per head-dim class.  The alternatives the sweep in
// profiles/r02_attention_variants.txt compared (scalar vs packed arithmetic, FMA-pipe exp2 share, stale reference, untrimmed MMA
// shapes) are compiled only into libtclight_tuning.so (-DTCL_ATTN_TUNING, `make tuning`; include/tclight_tuning.h).
#ifdef TCL_ATTN_TUNING
static int g_attn_variant = -1;  // -1 = shipped configuration; >= 0 selects a tuning variant (tools/bench_attn_variants.py)
static int g_attn_trim = 1;      // 0 = issue the full padded MMA shapes (tuning reference)
static long long* g_attn_trace = nullptr;
#else
constexpr int g_attn_variant = -1;
constexpr int g_attn_trim = 1;
constexpr long long* g_attn_trace = nullptr;
#endif

extern "C" size_t tcl_attention_workspace_bytes(void) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  return (size_t)sms * 256 * (64 + 1) * sizeof(float);        // <= one wave of KV-split CTAs x 256 rows x (O row + max)
}

extern "C" int tcl_attention(const tcl_attn_desc* a, cudaStream_t stream) {
  TCL_CHECK_ARG(a != nullptr, "tcl_attention: null descriptor");
  TCL_CHECK_ARG(a->q && a->k && a->vt && a->out, "tcl_attention: null pointer");
  TCL_CHECK_ARG(a->dtype == TCL_DTYPE_FP16 || a->dtype == TCL_DTYPE_BF16, "tcl_attention: dtype");
  TCL_CHECK_ARG(a->d_pad == 64 || a->d_pad == 128 || a->d_pad == 192, "tcl_attention: d_pad=%d (64/128/192)", a->d_pad);
  TCL_CHECK_ARG(a->d > 0 && a->d < a->d_pad && a->d % 8 == 0, "tcl_attention: d=%d (needs d < d_pad: column d carries the softmax denominator)", a->d);
  TCL_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->tq > 0 && a->tk > 0, "tcl_attention: empty problem");
  TCL_CHECK_ARG(a->kv_batch_div >= 1 && a->batch % a->kv_batch_div == 0, "tcl_attention: kv_batch_div");
  TCL_CHECK_ARG(a->tq_pitch >= a->tq && a->tk_pitch >= a->tk && a->tk_pitch % 8 == 0, "tcl_attention: pitches");
  const bool bf16 = a->dtype == TCL_DTYPE_BF16;
  const int bh = a->batch * a->heads;
  const int kv_bh = (a->batch / a->kv_batch_div) * a->heads;
  AttnParams p;
  p.tq = a->tq; p.tk = a->tk; p.heads = a->heads; p.d = a->d;
  p.kv_batch_div = a->kv_batch_div;
  p.n_kv_tiles = (a->tk + 127) / 128;
  p.qk_steps = g_attn_trim ? (a->d + 15) / 16 : a->d_pad / 16;
  p.n_o = g_attn_trim ? (a->d + 1 + 15) / 16 * 16 : a->d_pad;
  p.idesc_o = umma_idesc_f16(bf16, 128, (uint32_t)p.n_o);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)a->d);
  p.out = a->out;
  p.out_pitch = (long long)a->heads * a->d;
  p.trace = g_attn_trace;
  p.q_groups = 0; p.cta_base = 0; p.kv_parts = 1; p.part_o = nullptr; p.part_m = nullptr;      // set per launch
  AttnTmaps tm;
  {
    const uint64_t dims[3] = {(uint64_t)a->d_pad, (uint64_t)a->tq, (uint64_t)bh};
    const uint64_t str[2] = {(uint64_t)a->d_pad * 2, (uint64_t)a->tq_pitch * a->d_pad * 2};
    const uint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tm.q, a->q, bf16, 3, dims, str, box, es, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->d_pad, (uint64_t)a->tk, (uint64_t)kv_bh};
    const uint64_t str[2] = {(uint64_t)a->d_pad * 2, (uint64_t)a->tk_pitch * a->d_pad * 2};
    const uint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tm.k, a->k, bf16, 3, dims, str, box, es, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->tk, (uint64_t)a->d_pad, (uint64_t)kv_bh};
    const uint64_t str[2] = {(uint64_t)a->tk_pitch * 2, (uint64_t)a->d_pad * a->tk_pitch * 2};
    const uint32_t box[3] = {64, (uint32_t)p.n_o, 1}, es[3] = {1, 1, 1};      // only the live rows of V^T
    int rc = make_tmap(&tm.vt, a->vt, bf16, 3, dims, str, box, es, 128);
    if (rc) return rc;
  }
  const int q_tiles = (a->tq + 127) / 128;
  const int var = g_attn_variant;
  if (a->d_pad == 64) {
    // short key sequences (cross-attention on the 77 / 154 text tokens): the per-CTA latency chain (TMEM alloc, Q / K / V
    // loads, two serial tiles, epilogue) dominates, so run one Q tile per CTA and two CTAs per SM (256 TMEM columns,
    // 85 KB of shared memory each) to overlap the chains of neighbouring tiles
    const int n_keys16 = (a->tk + 15) / 16 * 16;
    if ((n_keys16 == 160 || n_keys16 == 80) && g_attn_variant != 0) {
      // persistent cross-attention (L = 154 / 77): K / V^T resident per (text, head), Q streamed, one TMEM allocation per CTA
      XAttnParams x;
      x.tq = a->tq; x.tk = a->tk; x.heads = a->heads; x.d = a->d; x.kv_batch_div = a->kv_batch_div;
      x.q_tiles = q_tiles; x.total_tiles = q_tiles * bh;
      int sms = 148;
      { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
      const int grid = x.total_tiles < sms ? x.total_tiles : sms;
      x.tiles_per_cta = (x.total_tiles + grid - 1) / grid;
      x.n_keys = n_keys16;
      x.qk_steps = p.qk_steps; x.n_o = p.n_o;
      x.idesc_s = umma_idesc_f16(bf16, 128, (uint32_t)x.n_keys); x.idesc_o = p.idesc_o;
      x.scale_log2 = p.scale_log2; x.out = p.out; x.out_pitch = p.out_pitch;
      AttnTmaps xm = tm;
      {
        const uint64_t dims[3] = {(uint64_t)a->d_pad, (uint64_t)a->tk, (uint64_t)kv_bh};
        const uint64_t str[2] = {(uint64_t)a->d_pad * 2, (uint64_t)a->tk_pitch * a->d_pad * 2};
        const uint32_t box[3] = {64, (uint32_t)x.n_keys, 1}, es[3] = {1, 1, 1};
        int rc = make_tmap(&xm.k, a->k, bf16, 3, dims, str, box, es, 128);
        if (rc) return rc;
      }
      const int grid_used = (x.total_tiles + x.tiles_per_cta - 1) / x.tiles_per_cta;
      if (x.n_keys == 160)
        return bf16 ? launch_xattn<true, 0x11, 48, 32, 48, 32>(xm, x, grid_used, stream) : launch_xattn<false, 0x11, 48, 32, 48, 32>(xm, x, grid_used, stream);
      return bf16 ? launch_xattn<true, 0x11, 32, 16, 16, 16>(xm, x, grid_used, stream) : launch_xattn<false, 0x11, 32, 16, 16, 16>(xm, x, grid_used, stream);
    }
    if (a->tk <= 256) {
      return bf16 ? launch_attn<1, 64, 2, 2, true, 0x11, true, false, 2>(tm, p, q_tiles, bh, stream)
                  : launch_attn<1, 64, 2, 2, false, 0x00, true, false, 2>(tm, p, q_tiles, bh, stream);
    }
    // long key sequences (merged self-attention): two Q tiles per CTA, packed-pair arithmetic, 2 of every 8 pairs of
    // exponentials on the FMA pipe, 200 registers per softmax thread (measured sweep: profiles/r02_attention_variants.txt)
#ifdef TCL_ATTN_TUNING
    if (!bf16) {
      switch (var) {
        case 0: return launch_attn<2, 64, 4, 3, false, 0x00, false, false>(tm, p, q_tiles, bh, stream);
        case 8: return launch_attn<2, 64, 4, 3, false, 0x11, true, true, 1, 200>(tm, p, q_tiles, bh, stream);
        default: break;
      }
    } else {
      switch (var) {
        case 0: return launch_attn<2, 64, 4, 3, true, 0x03, false, false>(tm, p, q_tiles, bh, stream);
        case 5: return launch_attn<2, 64, 4, 3, true, 0x49, true, false, 1, 200>(tm, p, q_tiles, bh, stream);
        case 8: return launch_attn<2, 64, 4, 3, true, 0x11, true, true, 1, 200>(tm, p, q_tiles, bh, stream);
        default: break;
      }
    }
#endif
    return bf16 ? launch_attn<2, 64, 4, 3, true, 0x11, true, false, 1, 200>(tm, p, q_tiles, bh, stream, a->workspace, a->workspace_bytes)
                : launch_attn<2, 64, 4, 3, false, 0x11, true, false, 1, 200>(tm, p, q_tiles, bh, stream, a->workspace, a->workspace_bytes);
  } else if (a->d_pad == 128) {
    // one Q tile per CTA (TMEM: 128 S + 128 O + 64 P): registers are plentiful, the stale-reference softmax wins (+8 %)
#ifdef TCL_ATTN_TUNING
    if (var == 0)
      return bf16 ? launch_attn<1, 128, 3, 2, true, 0, false, false>(tm, p, q_tiles, bh, stream)
                  : launch_attn<1, 128, 3, 2, false, 0, false, false>(tm, p, q_tiles, bh, stream);
#endif
    return bf16 ? launch_attn<1, 128, 3, 2, true, 0x11, true, true>(tm, p, q_tiles, bh, stream)
                : launch_attn<1, 128, 3, 2, false, 0x00, true, true>(tm, p, q_tiles, bh, stream);
  } else {
    return bf16 ? launch_attn<1, 192, 2, 1, true, 0x00, true, false>(tm, p, q_tiles, bh, stream)
                : launch_attn<1, 192, 2, 1, false, 0x00, true, false>(tm, p, q_tiles, bh, stream);
  }
}

#ifdef TCL_ATTN_TUNING
// Tuning hooks (tools/bench_attn_variants.py): kernel variant (-1 = shipped) and MMA-shape trimming; return the previous value.
extern "C" int tcl_debug_attention_variant(int v) {
  const int old = g_attn_variant;
  g_attn_variant = v;
  return old;
}
extern "C" int tcl_debug_attention_trim(int on) {
  const int old = g_attn_trim;
  g_attn_trim = on;
  return old;
}

// Debug hook (effective only in -DTCL_ATTN_TRACE builds): device buffer of 3*8*8 int64 receiving clock64 stamps of
// CTA (0,0) for KV tiles 64..71: role 0/1 = softmax warpgroup q {wait S, got S, loaded S, exp done, got P-empty, P stored},
// role 2 = MMA thread {q0: wait S-empty, got it, wait P-full, got it; q1: the same}.
extern "C" void tcl_debug_attention_trace(long long* buf) { g_attn_trace = buf; }
#endif
