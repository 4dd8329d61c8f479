// Fused multi-head attention (flash style) on tcgen05 for sm_100a.
//
// Replaces the bmm + softmax + bmm of: diffusers Attention/AttnProcessor inside the SD-1.5 UNet (self-attention over
// 4096/1024/256/64 latent tokens and cross-attention over the 77 GILLMapper tokens; gill/models.py:730 ->
// gill/custom_sd.py:633-638) and OPTAttention (causal, gill/models.py:465).
//
// One CTA per (128-query tile, head, batch). Heads are stored padded to HD_PAD (a multiple of 64 columns, zero filled
// by the producing projection) so that every TMA box is a clean 128-byte swizzled row.
//   warp 0 lane 0 : TMA producer (Q once; K/V tiles double buffered)
//   warp 1 lane 0 : MMA issuer    S_j = Q K_j^T  -> TMEM (double buffered);   O += P_j V_j -> TMEM
//   warp 2        : TMEM allocator
//   warps 4..7    : softmax. TMEM lane == query row, so each thread owns one row: no cross-thread reductions.
//                   Online softmax with lazy rescaling (O is only rescaled when the running max grows by > 2^8),
//                   P written to smem in the K-major SWIZZLE_128B layout the P.V MMA reads.
// V tiles are consumed directly as an MN-major B operand (no transpose pass).
#include "../../include/gillb200.h"
#include "gemm_sm100.cuh"
#include "host_common.h"

#include <cstdlib>
#include <cstring>

namespace gb {

struct alignas(64) AttnParams {
  CUtensorMap tma_q, tma_k, tma_v;  // 3D {cols, L, B}, box {64, 128 | BLOCK_KV, 1}
  void* out;
  long long ldo, o_bstride;
  const int* kv_lens;
  int B, H, Lq, Lk;
  int causal, causal_offset;  // key j visible to query i iff j <= i + causal_offset
  int out_dtype;
  int in_dtype;
  int hd_cols;       // columns per head in q/k/v/out (head pitch, multiple of 16, <= HD_PAD): head h starts at column
                     // h * hd_cols; Q.K^T runs hd_cols / 16 K-steps and P.V produces hd_cols output columns. TMA boxes stay
                     // HD_PAD wide (the tail belongs to the next head / is zero filled and is never multiplied)
  int ones_col;      // >= 0: column (inside the head padding) where V holds 1.0, so O[:, ones_col] IS the softmax
                     // denominator (computed by the tensor core from the same rounded P as the numerator); -1: none
  float scale_log2;  // softmax scale * log2(e)
  int q_tiles_per_cta;  // attn4q (short key sequences): consecutive 128-row query tiles one CTA walks with K / V resident
};

template <int HD_PAD, int BLOCK_KV>
struct AttnCfg {
  // HD_PAD == 64 (UNet 64x64 level, the hot shape; ncu: no pipe above 45%, latency-bound with <= 2 softmax warps per
  // scheduler): 64-wide KV tiles and single S/K/V buffers bring a CTA down to 128 TMEM columns and ~49 KB of shared
  // memory, so FOUR CTAs share an SM -- 16 softmax warps per SM hide the TMEM / MUFU / mbarrier latencies and the
  // other CTAs' MMAs fill each CTA's serial S -> softmax -> P.V chain. Wider heads keep the double-buffered 1-CTA form.
  static constexpr int SBUF = HD_PAD == 64 ? 1 : 2;
  static constexpr int VBUF = HD_PAD == 64 ? 1 : 2;
  static constexpr int KBUF = HD_PAD == 64 ? 1 : 2;
  static constexpr int CTAS_PER_SM = HD_PAD == 64 ? 4 : 1;
  static constexpr int KCH = HD_PAD / 64;           // 64-column chunks per head
  static constexpr int Q_BYTES = 128 * HD_PAD * 2;  // chunk-major: KCH x [128 rows x 128 B]
  static constexpr int KV_BYTES = BLOCK_KV * HD_PAD * 2;
  static constexpr int P_BYTES = 128 * BLOCK_KV * 2;  // (BLOCK_KV/64) x [128 rows x 128 B]
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + KBUF * KV_BYTES;
  static constexpr int OFF_P = OFF_V + VBUF * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_S0 = 0;               // S buffers at columns [i*BLOCK_KV, (i+1)*BLOCK_KV)
  static constexpr int TMEM_O = SBUF * BLOCK_KV;  // O accumulator
  static constexpr int TMEM_COLS = (SBUF * BLOCK_KV + HD_PAD) <= 128 ? 128 : (SBUF * BLOCK_KV + HD_PAD) <= 256 ? 256 : 512;
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 228 * 1024, "smem for the intended occupancy");
  static_assert(CTAS_PER_SM * TMEM_COLS <= 512, "TMEM for the intended occupancy");
  static_assert(HD_PAD % 64 == 0 && HD_PAD <= 192, "head pad");
  static_assert(BLOCK_KV == 64 || BLOCK_KV == 128, "kv tile");
  static_assert(SMEM_BYTES <= SMEM_BUDGET, "smem");
};

struct AttnBars {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full, p_empty;
  uint64_t o_done;
  uint32_t tmem_ptr;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// max of 32 raw scores; unmasked tiles use the 3-input FMNMX3 (17 instructions instead of 31)
template <bool MASKED>
__device__ __forceinline__ float attn_max32(const uint32_t (&v)[32], int c, int lim) {
  if (!MASKED) {
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
      t[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    t[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    return fmax3(fmax3(t[0], t[1], t[2]), fmax3(t[3], t[4], t[5]), fmax3(fmax3(t[6], t[7], t[8]), t[9], t[10]));
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (c + i <= lim) mx = fmaxf(mx, __uint_as_float(v[i]));
  return mx;
}

// p = exp2(s * scale - m) for 32 columns -> 16 packed pairs; returns the fp32 row-sum contribution when SUM. Masked
// columns (> lim) give exactly 0. Compile-time variants keep the interior-tile path free of selects.
template <bool MASKED, bool SUM, bool BF16>
__device__ __forceinline__ float attn_exp32(const uint32_t (&v)[32], int c, int lim, float scale_log2, float m,
                                            uint32_t (&pk)[16]) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), scale_log2, -m));
    float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m));
    if (MASKED) {
      p0 = (c + i <= lim) ? p0 : 0.f;
      p1 = (c + i + 1 <= lim) ? p1 : 0.f;
    }
    if (SUM) sum += p0 + p1;
    pk[i >> 1] = BF16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
  }
  return sum;
}

template <int HD_PAD, int BLOCK_KV>
__global__ void __launch_bounds__(256, AttnCfg<HD_PAD, BLOCK_KV>::CTAS_PER_SM) attn_kernel(const __grid_constant__ AttnParams p) {
  using C = AttnCfg<HD_PAD, BLOCK_KV>;
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + C::OFF_BAR);
  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kv_len = p.kv_lens ? p.kv_lens[b] : p.Lk;
  int n_tiles = (kv_len + BLOCK_KV - 1) / BLOCK_KV;
  if (p.causal) {
    const int last_visible = min(kv_len - 1, q0 + 127 + p.causal_offset);
    n_tiles = min(n_tiles, last_visible / BLOCK_KV + 1);
    if (n_tiles < 1) n_tiles = 1;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_empty[i], 4);
    }
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->p_empty, 1);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const int col0 = head * p.hd_cols;
  const int ksteps = p.hd_cols >> 4;

  if (warp == 0) {
    if (lane_id() == 0) {
      // ---------------- TMA producer
      mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < C::KCH; ++c)
        tma_load_3d(smem + C::OFF_Q + c * (128 * 128), &p.tma_q, &bars->q_full, col0 + c * 64, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % C::KBUF, vs = j % C::VBUF;
        const uint32_t kph = (j / C::KBUF) & 1, vph = (j / C::VBUF) & 1;
        mbar_wait(&bars->k_empty[ks], kph ^ 1);
        mbar_arrive_expect_tx(&bars->k_full[ks], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_3d(smem + C::OFF_K + ks * C::KV_BYTES + c * (BLOCK_KV * 128), &p.tma_k, &bars->k_full[ks],
                      col0 + c * 64, j * BLOCK_KV, b);
        mbar_wait(&bars->v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[vs], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_3d(smem + C::OFF_V + vs * C::KV_BYTES + c * (BLOCK_KV * 128), &p.tma_v, &bars->v_full[vs],
                      col0 + c * 64, j * BLOCK_KV, b);
      }
    }
  } else if (warp == 1) {
    if (lane_id() == 0) {
      // ---------------- MMA issuer
      const bool bf16 = p.in_dtype == DT_BF16;
      const uint32_t idesc_s = make_idesc_f16(128, BLOCK_KV, bf16, false);
      const uint32_t idesc_o = make_idesc_f16(128, p.hd_cols, bf16, true);  // B = V, MN-major
      const uint32_t sq = smem_u32(smem + C::OFF_Q);
      const uint32_t sp = smem_u32(smem + C::OFF_P);
      auto issue_s = [&](int j) {
        const int ks = j % C::KBUF, s = j % C::SBUF;
        const uint32_t kph = (j / C::KBUF) & 1, sph = (j / C::SBUF) & 1;
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->s_empty[s], sph ^ 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + C::OFF_K + ks * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD_PAD / 16; ++k) {
          if (k < ksteps) {
            const int c = k >> 2, kk = k & 3;
            const uint64_t da = make_smem_desc_sw128(sq + c * (128 * 128) + kk * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sk + c * (BLOCK_KV * 128) + kk * 32, 16, 1024);
            umma_f16(tmem_base + C::TMEM_S0 + s * BLOCK_KV, da, db, idesc_s, k != 0 ? 1u : 0u);
          }
        }
        umma_commit(&bars->s_full[s]);
        umma_commit(&bars->k_empty[ks]);
      };
      mbar_wait(&bars->q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        // S_{j+1} overlaps the softmax of tile j when S is double buffered; with a single S buffer it is issued as
        // soon as the softmax has drained S_j (s_empty), i.e. concurrently with P_j.V_j below.
        if (C::SBUF == 2 && j + 1 < n_tiles) issue_s(j + 1);
        const int vs = j % C::VBUF;
        const uint32_t vph = (j / C::VBUF) & 1;
        mbar_wait(&bars->v_full[vs], vph);
        mbar_wait(&bars->p_full, j & 1);
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + C::OFF_V + vs * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_KV / 16; ++k) {
          // A = P: K-major, 64-kv chunks of [128 x 128 B]; B = V: MN-major, rows = kv, LBO = 64-column chunk stride
          const uint64_t da = make_smem_desc_sw128(sp + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sv + k * (16 * 128), BLOCK_KV * 128, 1024);
          umma_f16(tmem_base + C::TMEM_O, da, db, idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&bars->v_empty[vs]);
        umma_commit(&bars->p_empty);
        if (j == n_tiles - 1) umma_commit(&bars->o_done);
        if (C::SBUF == 1 && j + 1 < n_tiles) issue_s(j + 1);
      }
    }
  } else if (warp >= 4) {
    // ---------------- softmax / epilogue: thread <-> query row
    const int ewarp = warp & 3;
    const int r = ewarp * 32 + lane_id();
    const int qrow = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(ewarp * 32) << 16;
    const bool bf16 = p.in_dtype == DT_BF16;
    uint8_t* sp = smem + C::OFF_P;
    float m_used = 0.f, l = 0.f;
    const bool use_ones = p.ones_col >= 0;  // softmax denominator comes out of the P.V product (V's ones column)
    const int causal_lim = p.causal ? qrow + p.causal_offset : 0x7fffffff;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % C::SBUF;
      const uint32_t ph = (j / C::SBUF) & 1;
      mbar_wait(&bars->s_full[s], ph);
      tc_fence_after();
      const uint32_t ts = tmem_base + C::TMEM_S0 + s * BLOCK_KV + lane_off;
      const int kv0 = j * BLOCK_KV;
      const int lim = min(kv_len - 1, causal_lim) - kv0;  // columns > lim are masked
      // interior tiles (every column visible to every row of this warp) skip all masking work
      const bool no_mask = __all_sync(0xffffffffu, lim >= BLOCK_KV - 1);

      auto max32 = [&](const uint32_t (&v)[32], int c) -> float {
        return no_mask ? attn_max32<false>(v, c, lim) : attn_max32<true>(v, c, lim);
      };
      // p = exp2(s*scale - m_used) for 32 columns -> packed 16-bit pairs; returns their fp32 sum (0 when use_ones)
      auto exp32 = [&](const uint32_t (&v)[32], int c, uint32_t (&pk)[16]) -> float {
        if (no_mask) {
          if (use_ones) return bf16 ? attn_exp32<false, false, true>(v, c, lim, p.scale_log2, m_used, pk)
                                    : attn_exp32<false, false, false>(v, c, lim, p.scale_log2, m_used, pk);
          return bf16 ? attn_exp32<false, true, true>(v, c, lim, p.scale_log2, m_used, pk)
                      : attn_exp32<false, true, false>(v, c, lim, p.scale_log2, m_used, pk);
        }
        return bf16 ? attn_exp32<true, true, true>(v, c, lim, p.scale_log2, m_used, pk)
                    : attn_exp32<true, true, false>(v, c, lim, p.scale_log2, m_used, pk);
      };
      // P -> smem (K-major SW128: 16-B unit u of row r lands at u ^ (r & 7))
      auto store32 = [&](const uint32_t (&pk)[16], int c) {
        uint8_t* chunk = sp + (c >> 6) * (128 * 128) + r * 128;
        const int u0 = (c & 63) >> 3;  // first 16-B unit of this 32-column group within the 64-column chunk
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int phys = (u0 + u) ^ (r & 7);
          *reinterpret_cast<uint4*>(chunk + phys * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      };
      // O *= alpha (TMEM round trip); only when some row's running max grew by more than 2^8
      auto rescale_o = [&](float alpha) {
        const uint32_t to = tmem_base + C::TMEM_O + lane_off;
#pragma unroll 1
        for (int c = 0; c < p.hd_cols; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(to + c, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x16(to + c, v);
        }
        tmem_wait_st();
      };

      if (j == 0) {
        // first tile: true row max first (two passes over S)
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < BLOCK_KV; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(ts + c, v);
          tmem_wait_ld();
          mx = fmaxf(mx, max32(v, c));
        }
        m_used = mx == -INFINITY ? 0.f : mx * p.scale_log2;
#pragma unroll 1
        for (int c = 0; c < BLOCK_KV; c += 32) {
          uint32_t v[32], pk[16];
          tmem_ld_32x32b_x32(ts + c, v);
          tmem_wait_ld();
          l += exp32(v, c, pk);
          store32(pk, c);
        }
      } else {
        // later tiles: ONE pass. Exponentials are taken against the running reference max while the tile max is
        // tracked on the side (FMNMX3); only if it exceeds the reference by more than 2^8 is the tile redone after
        // rescaling O. The exp work of the first 32 columns overlaps P.V of the previous tile (p_empty wait below).
        float mx = -INFINITY, lt = 0.f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_KV; c += 32) {
          uint32_t v[32], pk[16];
          tmem_ld_32x32b_x32(ts + c, v);
          tmem_wait_ld();
          mx = fmaxf(mx, max32(v, c));
          lt += exp32(v, c, pk);
          if (c == 0) {  // P buffer free and O quiescent once P.V of tile j-1 has retired
            mbar_wait(&bars->p_empty, (j - 1) & 1);
            tc_fence_after();
          }
          store32(pk, c);
        }
        mx *= p.scale_log2;
        const bool need = mx > m_used + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          float alpha = 1.f;
          if (need) {
            alpha = fast_exp2(m_used - mx);
            m_used = mx;
          }
          l *= alpha;
          rescale_o(alpha);
          lt = 0.f;
#pragma unroll 1
          for (int c = 0; c < BLOCK_KV; c += 32) {
            uint32_t v[32], pk[16];
            tmem_ld_32x32b_x32(ts + c, v);
            tmem_wait_ld();
            lt += exp32(v, c, pk);
            store32(pk, c);
          }
        }
        l += lt;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) {
        mbar_arrive(&bars->s_empty[s]);
        mbar_arrive(&bars->p_full);
      }
    }
    // epilogue: O / l
    mbar_wait(&bars->o_done, 0);
    tc_fence_after();
    const uint32_t to = tmem_base + C::TMEM_O + lane_off;
    if (use_ones) {
      l = __uint_as_float(tmem_ld_32x32b_x1(to + p.ones_col));
      tmem_wait_ld();
    }
    const float inv = 1.f / l;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + static_cast<long long>(b) * p.o_bstride +
                     static_cast<long long>(qrow) * p.ldo + col0;
#pragma unroll 1
    for (int c = 0; c < p.hd_cols; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(to + c, v);
      tmem_wait_ld();
      if (qrow < p.Lq) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
        store16(orow + c, f, 16, p.out_dtype);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ================================================================================================================
// attn2: the same math with NO dedicated producer / MMA warps.
//
// ncu on attn_kernel<64,64> (profiles/r01_ncu_attn_*): 65 % of issue slots busy, a third of all executed instructions
// were the mbarrier try_wait / BRA / YIELD loops of the waiting producer and MMA warps, and every KV tile paid two
// barrier hand-offs (softmax -> MMA warp -> softmax). Here a CTA is just the four softmax warps (128 threads = the 128
// query rows): after a tile's P is in shared memory the CTA meets at one bar.sync and thread 0 issues P.V for this
// tile and Q.K^T for the next (8-24 tcgen05.mma + one commit); at the top of a tile, when S has arrived, the same
// thread issues the TMA loads for the next K and the current V -- their buffers are provably free because the tensor
// pipe retires MMAs in order. Nobody spins except on the one barrier that is the true dependency (S ready).
// 128 threads per CTA also lift the register cap to 128, so a whole 64-column score row stays in registers: one TMEM
// read per tile, exact row max before the exponentials, lazy rescale (> 2^8) as before.
// Tried and dropped (measured on B200, 16x8x4096x4096 hd40): issuing Q.K_{j+1}^T in the middle of tile j's softmax
// through two alternating K/V slots (no S wait on the critical path) ran 939 us vs 853 us for this version -- with four
// CTAs per SM the MMA bubble of one CTA is already filled by the others' exponentials; the extra barrier and the probe
// code only add issue slots. Moving every 4th pair of exponentials to an FMA-pipe polynomial (Cody-Waite split + degree-3
// 2^f, 9 instructions per exp) did not help either: 862 us. A K/V-resident cross-attention variant walking 4 query tiles
// per CTA (set-up and K/V loads paid once) lost too: 71 vs 60 us for 16x8x4096x77 -- fewer CTAs per SM to interleave. XU sits at 61 %, issue slots at 48 %: what remains is the
// serial TMEM-load -> max -> exp -> pack -> st.shared -> barrier -> MMA chain of each CTA, overlapped 4 ways per SM.
template <int HD_PAD>
struct Attn2Cfg {
  static constexpr int BKV = 64;
  static constexpr int KCH = HD_PAD / 64;
  static constexpr int Q_BYTES = 128 * HD_PAD * 2;
  static constexpr int KV_BYTES = BKV * HD_PAD * 2;
  static constexpr int P_BYTES = 128 * BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int TMEM_S = 0;
  static constexpr int TMEM_O = BKV;
  static constexpr int TMEM_COLS = BKV + HD_PAD <= 128 ? 128 : 256;
  static constexpr int CTAS_PER_SM = HD_PAD == 64 ? 4 : HD_PAD == 128 ? 2 : 1;
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 228 * 1024, "smem for the intended occupancy");
  static_assert(CTAS_PER_SM * TMEM_COLS <= 512, "TMEM for the intended occupancy");
};

struct Attn2Bars {
  uint64_t q_full, k_full, v_full, s_full;
  uint32_t tmem_ptr;
};

// PT (P through tensor memory): the softmax threads write the packed 16-bit P tile back into the upper half of the S
// columns with tcgen05.st and P.V reads its A operand from there (tcgen05.mma with a TMEM A operand) -- no st.shared of P
// (ncu r01: 7.2 M shared-memory bank conflicts per launch from those stores), no generic->async proxy fence, no P buffer.
// The MMAs of one thread execute in issue order, so S_{j+1} = Q.K_{j+1}^T (issued after P.V_j) overwrites the S columns,
// P included, only after P.V_j has consumed them.
template <int HD_PAD, bool PT>
__global__ void __launch_bounds__(128, Attn2Cfg<HD_PAD>::CTAS_PER_SM) attn2_kernel(const __grid_constant__ AttnParams p) {
  using C = Attn2Cfg<HD_PAD>;
  constexpr int BKV = C::BKV;
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Attn2Bars* bars = reinterpret_cast<Attn2Bars*>(smem + C::OFF_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.kv_lens ? p.kv_lens[b] : p.Lk;
  int n_tiles = (kv_len + BKV - 1) / BKV;
  if (p.causal) {
    const int last_visible = min(kv_len - 1, q0 + 127 + p.causal_offset);
    n_tiles = min(n_tiles, last_visible / BKV + 1);
  }
  if (n_tiles < 1) n_tiles = 1;

  if (tid == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->k_full, 1);
    mbar_init(&bars->v_full, 1);
    mbar_init(&bars->s_full, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const int col0 = head * p.hd_cols;
  const int ksteps = p.hd_cols >> 4;
  const bool bf16 = p.in_dtype == DT_BF16;
  const uint32_t idesc_s = make_idesc_f16(128, BKV, bf16, false);
  const uint32_t idesc_o = make_idesc_f16(128, p.hd_cols, bf16, true);  // B = V, MN-major
  const uint32_t sq = smem_u32(smem + C::OFF_Q), sk = smem_u32(smem + C::OFF_K), sv = smem_u32(smem + C::OFF_V),
                 spa = smem_u32(smem + C::OFF_P);

  // ---- issue helpers (thread 0 only)
  auto load_kv = [&](int j, bool is_v) {
    uint64_t* bar = is_v ? &bars->v_full : &bars->k_full;
    mbar_arrive_expect_tx(bar, C::KV_BYTES);
#pragma unroll
    for (int c = 0; c < C::KCH; ++c)
      tma_load_3d(smem + (is_v ? C::OFF_V : C::OFF_K) + c * (BKV * 128), is_v ? &p.tma_v : &p.tma_k, bar, col0 + c * 64,
                  j * BKV, b);
  };
  auto mma_s = [&]() {
#pragma unroll
    for (int k = 0; k < HD_PAD / 16; ++k) {
      if (k < ksteps) {
        const int c = k >> 2, kk = k & 3;
        umma_f16(tmem_base + C::TMEM_S, make_smem_desc_sw128(sq + c * (128 * 128) + kk * 32, 16, 1024),
                 make_smem_desc_sw128(sk + c * (BKV * 128) + kk * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
      }
    }
  };
  auto mma_pv = [&](int j) {
#pragma unroll
    for (int k = 0; k < BKV / 16; ++k) {
      const uint64_t db = make_smem_desc_sw128(sv + k * (16 * 128), BKV * 128, 1024);
      if (PT)  // A = P in TMEM: packed pairs at S columns [BKV/2, BKV), 8 columns per 16-key step
        umma_f16_ts(tmem_base + C::TMEM_O, tmem_base + C::TMEM_S + BKV / 2 + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);
      else
        umma_f16(tmem_base + C::TMEM_O, make_smem_desc_sw128(spa + k * 32, 16, 1024), db, idesc_o, (j | k) != 0 ? 1u : 0u);
    }
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
#pragma unroll
    for (int c = 0; c < C::KCH; ++c)
      tma_load_3d(smem + C::OFF_Q + c * (128 * 128), &p.tma_q, &bars->q_full, col0 + c * 64, q0, b);
    load_kv(0, false);
    load_kv(0, true);
    mbar_wait(&bars->q_full, 0);
    mbar_wait(&bars->k_full, 0);
    tc_fence_after();
    mma_s();
    umma_commit(&bars->s_full);
  }
  __syncwarp();

  // ---- softmax: thread <-> query row
  const int r = tid;
  const int qrow = q0 + r;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t ts = tmem_base + C::TMEM_S + lane_off;
  const uint32_t to = tmem_base + C::TMEM_O + lane_off;
  uint8_t* sp = smem + C::OFF_P;
  float m_used = 0.f, l = 0.f;
  const bool use_ones = p.ones_col >= 0;
  const int causal_lim = p.causal ? qrow + p.causal_offset : 0x7fffffff;

  for (int j = 0; j < n_tiles; ++j) {
    mbar_wait(&bars->s_full, j & 1);
    tc_fence_after();
    if (tid == 0) {
      if (j + 1 < n_tiles) load_kv(j + 1, false);  // K buffer is free: S_j has retired
      if (j > 0) load_kv(j, true);                 // V buffer is free: P.V_{j-1} retired before S_j
    }
    __syncwarp();
    const int kv0 = j * BKV;
    const int lim = min(kv_len - 1, causal_lim) - kv0;  // columns > lim are masked
    const bool no_mask = __all_sync(0xffffffffu, lim >= BKV - 1);
    uint32_t s0[32], s1[32];
    tmem_ld_32x32b_x32(ts, s0);
    tmem_ld_32x32b_x32(ts + 32, s1);
    tmem_wait_ld();
    float mx = no_mask ? fmaxf(attn_max32<false>(s0, 0, lim), attn_max32<false>(s1, 32, lim))
                       : fmaxf(attn_max32<true>(s0, 0, lim), attn_max32<true>(s1, 32, lim));
    mx *= p.scale_log2;
    // lazy rescale decision (warp-uniform branch because tcgen05.ld/st are .sync.aligned)
    float alpha = 1.f;
    bool need = false;
    if (j == 0) {
      m_used = mx == -INFINITY ? 0.f : mx;
    } else if (mx > m_used + 8.f) {
      alpha = fast_exp2(m_used - mx);
      m_used = mx;
      need = true;
    }
    if (__any_sync(0xffffffffu, need)) {  // O is quiescent: P.V_{j-1} retired before S_j
      l *= alpha;
#pragma unroll 1
      for (int c = 0; c < p.hd_cols; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(to + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st_32x32b_x16(to + c, v);
      }
      tmem_wait_st();
    }
    // p = exp2(s*scale - m) -> P tile in smem (K-major SW128: 16-B unit u of row r lands at u ^ (r & 7))
    uint8_t* prow = sp + r * 128;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t(&sh)[32] = h == 0 ? s0 : s1;
      uint32_t pk[16];
      float part;
      if (no_mask) {
        if (use_ones) part = bf16 ? attn_exp32<false, false, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                                  : attn_exp32<false, false, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
        else part = bf16 ? attn_exp32<false, true, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                         : attn_exp32<false, true, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
      } else {
        part = bf16 ? attn_exp32<true, true, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                    : attn_exp32<true, true, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
      }
      l += part;
      if (PT) {
        tmem_st_32x32b_x16(ts + BKV / 2 + 16 * h, pk);  // this row's 32 keys -> 16 packed columns
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int phys = (4 * h + u) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + phys * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
    }
    if (PT)
      tmem_wait_st();
    else
      fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&bars->v_full, j & 1);
      tc_fence_after();
      mma_pv(j);
      if (j + 1 < n_tiles) {
        mbar_wait(&bars->k_full, (j + 1) & 1);
        tc_fence_after();
        mma_s();
      }
      umma_commit(&bars->s_full);  // completion #(j+1): S_{j+1} ready, or (last tile) O final
    }
    __syncwarp();
  }

  // ---- epilogue: O / l
  mbar_wait(&bars->s_full, n_tiles & 1);
  tc_fence_after();
  if (use_ones) {
    l = __uint_as_float(tmem_ld_32x32b_x1(to + p.ones_col));
    tmem_wait_ld();
  }
  const float inv = 1.f / l;
  uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + static_cast<long long>(b) * p.o_bstride +
                   static_cast<long long>(qrow) * p.ldo + col0;
#pragma unroll 1
  for (int c = 0; c < p.hd_cols; c += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(to + c, v);
    tmem_wait_ld();
    if (qrow < p.Lq) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
      store16(orow + c, f, 16, p.out_dtype);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ================================================================================================================
// attn3: attn2 (64-wide tile, P through tensor memory) with the issue work moved to a FIFTH warp.
//
// In attn2 thread 0 is both a softmax row and the CTA's TMA / MMA issuer: after every tile it alone waits for V and K,
// builds descriptors and issues 7-8 tcgen05.mma + a commit while warps 1-3 already sleep on the S barrier -- so warp 0
// reaches every CTA barrier last and the whole CTA runs at the pace of its most loaded warp. Here warps 0-3 only do
// softmax (thread = query row) and hand P over with one mbarrier arrive per warp (no CTA-wide bar.sync); warp 4 waits
// for that barrier with V / K already checked and issues P.V_j, Q.K_{j+1}^T and the commit immediately.
struct Attn3Bars {
  uint64_t q_full, k_full, v_full, s_full, p_full;
  uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(160, 4) attn3_kernel(const __grid_constant__ AttnParams p) {
  using C = Attn2Cfg<64>;
  constexpr int BKV = C::BKV;
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Attn3Bars* bars = reinterpret_cast<Attn3Bars*>(smem + C::OFF_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.kv_lens ? p.kv_lens[b] : p.Lk;
  int n_tiles = (kv_len + BKV - 1) / BKV;
  if (p.causal) {
    const int last_visible = min(kv_len - 1, q0 + 127 + p.causal_offset);
    n_tiles = min(n_tiles, last_visible / BKV + 1);
  }
  if (n_tiles < 1) n_tiles = 1;

  if (tid == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->k_full, 1);
    mbar_init(&bars->v_full, 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 4);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const int col0 = head * p.hd_cols;
  const int ksteps = p.hd_cols >> 4;
  const bool bf16 = p.in_dtype == DT_BF16;

  if (warp == 4) {
    // ---------------- issuer warp: TMA loads + every tcgen05.mma of the CTA (one lane)
    if (lane_id() == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, BKV, bf16, false);
      const uint32_t idesc_o = make_idesc_f16(128, p.hd_cols, bf16, true);  // B = V, MN-major
      const uint32_t sq = smem_u32(smem + C::OFF_Q), sk = smem_u32(smem + C::OFF_K), sv = smem_u32(smem + C::OFF_V);
      auto load_kv = [&](int j, bool is_v) {
        uint64_t* bar = is_v ? &bars->v_full : &bars->k_full;
        mbar_arrive_expect_tx(bar, C::KV_BYTES);
        tma_load_3d(smem + (is_v ? C::OFF_V : C::OFF_K), is_v ? &p.tma_v : &p.tma_k, bar, col0, j * BKV, b);
      };
      auto mma_s = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < ksteps)
            umma_f16(tmem_base + C::TMEM_S, make_smem_desc_sw128(sq + k * 32, 16, 1024),
                     make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
      };
      mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
      tma_load_3d(smem + C::OFF_Q, &p.tma_q, &bars->q_full, col0, q0, b);
      load_kv(0, false);
      load_kv(0, true);
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full, 0);
      tc_fence_after();
      mma_s();
      umma_commit(&bars->s_full);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&bars->s_full, j & 1);             // S_j (and P.V_{j-1}) retired: K and V buffers are free
        if (j + 1 < n_tiles) load_kv(j + 1, false);
        if (j > 0) load_kv(j, true);
        mbar_wait(&bars->v_full, j & 1);
        if (j + 1 < n_tiles) mbar_wait(&bars->k_full, (j + 1) & 1);
        mbar_wait(&bars->p_full, j & 1);             // all four softmax warps have stored P_j
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_f16_ts(tmem_base + C::TMEM_O, tmem_base + C::TMEM_S + BKV / 2 + k * 8,
                      make_smem_desc_sw128(sv + k * (16 * 128), BKV * 128, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
        if (j + 1 < n_tiles) mma_s();
        umma_commit(&bars->s_full);  // completion #(j+1): S_{j+1} ready, or (last tile) O final
      }
    }
  } else {
    // ---------------- softmax warps: thread <-> query row
    const int r = tid;
    const int qrow = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ts = tmem_base + C::TMEM_S + lane_off;
    const uint32_t to = tmem_base + C::TMEM_O + lane_off;
    float m_used = 0.f, l = 0.f;
    const bool use_ones = p.ones_col >= 0;
    const int causal_lim = p.causal ? qrow + p.causal_offset : 0x7fffffff;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      const int lim = min(kv_len - 1, causal_lim) - kv0;  // columns > lim are masked
      const bool no_mask = __all_sync(0xffffffffu, lim >= BKV - 1);
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(ts, s0);
      tmem_ld_32x32b_x32(ts + 32, s1);
      tmem_wait_ld();
      float mx = no_mask ? fmaxf(attn_max32<false>(s0, 0, lim), attn_max32<false>(s1, 32, lim))
                         : fmaxf(attn_max32<true>(s0, 0, lim), attn_max32<true>(s1, 32, lim));
      mx *= p.scale_log2;
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = mx == -INFINITY ? 0.f : mx;
      } else if (mx > m_used + 8.f) {
        alpha = fast_exp2(m_used - mx);
        m_used = mx;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {  // O is quiescent: P.V_{j-1} retired before S_j
        l *= alpha;
#pragma unroll 1
        for (int c = 0; c < p.hd_cols; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(to + c, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x16(to + c, v);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t(&sh)[32] = h == 0 ? s0 : s1;
        uint32_t pk[16];
        float part;
        if (no_mask) {
          if (use_ones) part = bf16 ? attn_exp32<false, false, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                                    : attn_exp32<false, false, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
          else part = bf16 ? attn_exp32<false, true, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                           : attn_exp32<false, true, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
        } else {
          part = bf16 ? attn_exp32<true, true, true>(sh, 32 * h, lim, p.scale_log2, m_used, pk)
                      : attn_exp32<true, true, false>(sh, 32 * h, lim, p.scale_log2, m_used, pk);
        }
        l += part;
        tmem_st_32x32b_x16(ts + BKV / 2 + 16 * h, pk);  // this row's 32 keys -> 16 packed columns (upper half of S)
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&bars->p_full);
    }
    // ---- epilogue: O / l
    mbar_wait(&bars->s_full, n_tiles & 1);
    tc_fence_after();
    if (use_ones) {
      l = __uint_as_float(tmem_ld_32x32b_x1(to + p.ones_col));
      tmem_wait_ld();
    }
    const float inv = 1.f / l;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + static_cast<long long>(b) * p.o_bstride +
                     static_cast<long long>(qrow) * p.ldo + col0;
#pragma unroll 1
    for (int c = 0; c < p.hd_cols; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(to + c, v);
      tmem_wait_ld();
      if (qrow < p.Lq) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
        store16(orow + c, f, 16, p.out_dtype);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ================================================================================================================
// attn4: 48-key tiles, S MMA decoupled from the softmax (hd_cols <= 48: the UNet's 64x64 level, head pitch 48).
//
// ncu on attn3 (profiles/r02_ncu_full_attn48.txt): XU pipe 70 %, and the hottest stall of the softmax warps is the wait for
// S_{j+1}: with P aliased onto the upper half of the S columns, Q.K_{j+1}^T can only be issued AFTER P.V_j, i.e. after the
// whole softmax of tile j -- every tile pays one MMA round trip (issue + ~200 MMA clocks + commit + wake-up) with the
// warp idle. TMEM has no room for a separate P at 64 keys (64 S + 32 P + 48 O = 144 > 128 columns at 4 CTAs per SM),
// but at 48 keys it has: S [0,48) | P [48,72) | O [72,120). So here the softmax warps release S as soon as the scores
// sit in registers (s_free), the issuer starts Q.K_{j+1}^T at once -- it runs UNDER the exponentials of tile j -- and
// P.V_j follows when P_j arrives. In steady state a softmax warp never waits for the tensor pipe. K and V tiles are
// double buffered (6 KB each) and fetched two / one tiles ahead.
// BKV keys per tile, KCH 64-column chunks of the head (KCH = 1: hd_cols <= 48 at 4 CTAs per SM and 48 keys;
// KCH = 2: hd_cols <= 96 at 2 CTAs per SM and 64 keys, S [0,64) | P [64,96) | O [96,192) of 256 columns -- the UNet's
// 32x32 level, head dim 80 at pitch 96).
template <int BKV_, int KCH_>
struct Attn4Cfg {
  static constexpr int BKV = BKV_, KCH = KCH_;
  static constexpr int Q_CHUNK = 128 * 128;       // one 64-column chunk of the 128 query rows
  static constexpr int KV_CHUNK = BKV * 128;      // one 64-column chunk of a K or V tile
  static constexpr int Q_BYTES = KCH * Q_CHUNK;
  static constexpr int KV_BYTES = KCH * KV_CHUNK;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + 2 * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int HD_MAX = KCH == 1 ? 48 : 96;
  static constexpr int TMEM_S = 0, TMEM_P = BKV, TMEM_O = BKV + BKV / 2;
  static constexpr int TMEM_COLS = TMEM_O + HD_MAX <= 128 ? 128 : 256;
  static constexpr int CTAS_PER_SM = 512 / TMEM_COLS;
  static_assert(TMEM_O + HD_MAX <= TMEM_COLS, "TMEM layout");
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 228 * 1024, "smem for the intended occupancy");
  static_assert(KV_CHUNK % 1024 == 0, "swizzle atoms");
  static_assert(BKV == 48 || BKV == 64, "softmax register tiles are written for 48 / 64 keys");
};
struct Attn4Bars {
  uint64_t q_full, k_full[2], v_full[2], s_full, s_free, p_full, pv_done;
  uint32_t tmem_ptr;
};

// max of N raw scores (N = 32 or 16), unmasked: balanced tree of 3-input FMNMX3 (depth 4 at N = 32; a running
// mx = max3(mx, a, b) chain would be 15 dependent instructions)
template <int N>
__device__ __forceinline__ float attn_max_n(const uint32_t* v) {
  static_assert(N == 32 || N == 16, "chunk width");
  float t[N / 3 + 1];
#pragma unroll
  for (int i = 0; i < N / 3; ++i)
    t[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  if (N == 32) {
    t[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    return fmax3(fmax3(t[0], t[1], t[2]), fmax3(t[3], t[4], t[5]), fmax3(fmax3(t[6], t[7], t[8]), t[9], t[10]));
  }
  return fmax3(fmax3(t[0], t[1], t[2]), fmax3(t[3], t[4], __uint_as_float(v[15])), t[2]);
}
template <int N>
__device__ __forceinline__ float attn_max_n_masked(const uint32_t* v, int c, int lim) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (c + i <= lim) mx = fmaxf(mx, __uint_as_float(v[i]));
  return mx;
}
// p = exp2(s * scale - m) for N columns -> N / 2 packed pairs (see attn_exp32)
template <int N, bool MASKED, bool SUM, bool BF16>
__device__ __forceinline__ float attn_exp_n(const uint32_t* v, int c, int lim, float scale_log2, float m, uint32_t* pk) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), scale_log2, -m));
    float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m));
    if (MASKED) {
      p0 = (c + i <= lim) ? p0 : 0.f;
      p1 = (c + i + 1 <= lim) ? p1 : 0.f;
    }
    if (SUM) sum += p0 + p1;
    pk[i >> 1] = BF16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
  }
  return sum;
}

// 2^x for two lanes on the FMA pipe (no MUFU): round-to-nearest split x = n + f (magic-number add), degree-3 minimax of
// 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the 4.9e-4 rounding of the 16-bit P), exponent added as an integer.
// 4 packed fp32 pair instructions + 2 min + 2 integer per pair. x is clamped at -30 (P below 2^-24 rounds to 0 in fp16 and
// is far below bf16's resolution next to the row maximum).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
  const f32x2 x = pk2(fmaxf(x0, -30.f), fmaxf(x1, -30.f));
  const f32x2 t = add2(x, pk2(12582912.f, 12582912.f));
  const f32x2 n = add2(t, pk2(-12582912.f, -12582912.f));
  const f32x2 f = fma2(n, pk2(-1.f, -1.f), x);
  f32x2 q = fma2(pk2(0.05517132208f, 0.05517132208f), f, pk2(0.24261054397f, 0.24261054397f));
  q = fma2(q, f, pk2(0.69326096773f, 0.69326096773f));
  q = fma2(q, f, pk2(0.99992811680f, 0.99992811680f));
  float t0, t1, q0, q1;
  upk2(t, t0, t1);
  upk2(q, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// Interior (unmasked) tiles: packed scale-and-subtract, and every POLY-th pair of exponentials on the FMA pipe instead of
// the MUFU (POLY = 0: none). The softmax is bound by the 16 ex2 per clock and SM of the XU pipe with issue slots to spare.
template <int N, bool SUM, bool BF16, int POLY>
__device__ __forceinline__ float attn_exp_n_fast(const uint32_t* v, float scale_log2, float m, uint32_t* pk, int phase) {
  const f32x2 sc = pk2(scale_log2, scale_log2), nm = pk2(-m, -m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    float x0, x1, p0, p1;
    upk2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc, nm), x0, x1);
    if (POLY > 0 && ((i >> 1) + phase) % POLY == POLY - 1) {
      exp2_poly2(x0, x1, p0, p1);
    } else {
      p0 = fast_exp2(x0);
      p1 = fast_exp2(x1);
    }
    if (SUM) sum += p0 + p1;
    pk[i >> 1] = BF16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
  }
  return sum;
}

template <int BKV_, int KCH_, int POLY>
__global__ void __launch_bounds__(160, Attn4Cfg<BKV_, KCH_>::CTAS_PER_SM) attn4_kernel(const __grid_constant__ AttnParams p) {
  using C = Attn4Cfg<BKV_, KCH_>;
  constexpr int BKV = C::BKV, KCH = C::KCH;
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Attn4Bars* bars = reinterpret_cast<Attn4Bars*>(smem + C::OFF_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.kv_lens ? p.kv_lens[b] : p.Lk;
  int n_tiles = (kv_len + BKV - 1) / BKV;
  if (p.causal) {
    const int last_visible = min(kv_len - 1, q0 + 127 + p.causal_offset);
    n_tiles = min(n_tiles, last_visible / BKV + 1);
  }
  if (n_tiles < 1) n_tiles = 1;

  if (tid == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->k_full[0], 1);
    mbar_init(&bars->k_full[1], 1);
    mbar_init(&bars->v_full[0], 1);
    mbar_init(&bars->v_full[1], 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->pv_done, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const int col0 = head * p.hd_cols;
  const int ksteps = p.hd_cols >> 4;
  const bool bf16 = p.in_dtype == DT_BF16;

  if (warp == 4) {
    // ---------------- issuer warp: TMA loads + every tcgen05.mma of the CTA (one lane)
    if (lane_id() == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, BKV, bf16, false);
      const uint32_t idesc_o = make_idesc_f16(128, p.hd_cols, bf16, true);  // B = V, MN-major
      const uint32_t sq = smem_u32(smem + C::OFF_Q), sk = smem_u32(smem + C::OFF_K), sv = smem_u32(smem + C::OFF_V);
      auto load_k = [&](int j) {
        mbar_arrive_expect_tx(&bars->k_full[j & 1], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          tma_load_3d(smem + C::OFF_K + (j & 1) * C::KV_BYTES + c * C::KV_CHUNK, &p.tma_k, &bars->k_full[j & 1],
                      col0 + c * 64, j * BKV, b);
      };
      auto load_v = [&](int j) {
        mbar_arrive_expect_tx(&bars->v_full[j & 1], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          tma_load_3d(smem + C::OFF_V + (j & 1) * C::KV_BYTES + c * C::KV_CHUNK, &p.tma_v, &bars->v_full[j & 1],
                      col0 + c * 64, j * BKV, b);
      };
      auto mma_s = [&](int slot) {
#pragma unroll
        for (int k = 0; k < 4 * KCH; ++k)
          if (k < ksteps)
            umma_f16(tmem_base + C::TMEM_S, make_smem_desc_sw128(sq + (k >> 2) * C::Q_CHUNK + (k & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(sk + slot * C::KV_BYTES + (k >> 2) * C::KV_CHUNK + (k & 3) * 32, 16, 1024),
                     idesc_s, k != 0 ? 1u : 0u);
      };
      mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < KCH; ++c)
        tma_load_3d(smem + C::OFF_Q + c * C::Q_CHUNK, &p.tma_q, &bars->q_full, col0 + c * 64, q0, b);
      load_k(0);
      load_v(0);
      if (n_tiles > 1) load_k(1);
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      mma_s(0);
      umma_commit(&bars->s_full);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          mbar_wait(&bars->s_free, j & 1);  // S_j sits in the softmax warps' registers (so Q.K_j^T has retired)
          tc_fence_after();
          mbar_wait(&bars->k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          mma_s((j + 1) & 1);               // runs under the exponentials of tile j
          umma_commit(&bars->s_full);
          if (j + 2 < n_tiles) load_k(j + 2);  // slot j & 1: its reader Q.K_j^T has retired
          if (j >= 1) mbar_wait(&bars->pv_done, (j - 1) & 1);  // P.V_{j-1} has released V slot (j + 1) & 1
          load_v(j + 1);
        }
        mbar_wait(&bars->p_full, j & 1);  // all four softmax warps have stored P_j
        mbar_wait(&bars->v_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_f16_ts(tmem_base + C::TMEM_O, tmem_base + C::TMEM_P + k * 8,
                      make_smem_desc_sw128(sv + (j & 1) * C::KV_BYTES + k * (16 * 128), C::KV_CHUNK, 1024), idesc_o,
                      (j | k) != 0 ? 1u : 0u);
        umma_commit(&bars->pv_done);
      }
    }
  } else {
    // ---------------- softmax warps: thread <-> query row
    const int qrow = q0 + tid;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ts = tmem_base + C::TMEM_S + lane_off;
    const uint32_t tp = tmem_base + C::TMEM_P + lane_off;
    const uint32_t to = tmem_base + C::TMEM_O + lane_off;
    float m_used = 0.f, l = 0.f;
    const bool use_ones = p.ones_col >= 0;
    const int causal_lim = p.causal ? qrow + p.causal_offset : 0x7fffffff;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      constexpr int N1 = BKV - 32;  // second register tile: 16 (48 keys) or 32 (64 keys)
      uint32_t s0[32], s1[N1];
      tmem_ld_32x32b_x32(ts, s0);
      if constexpr (N1 == 16) tmem_ld_32x32b_x16(ts + 32, s1);
      else tmem_ld_32x32b_x32(ts + 32, s1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&bars->s_free);
      const int kv0 = j * BKV;
      const int lim = min(kv_len - 1, causal_lim) - kv0;  // columns > lim are masked
      const bool no_mask = __all_sync(0xffffffffu, lim >= BKV - 1);
      float mx = no_mask ? fmaxf(attn_max_n<32>(s0), attn_max_n<N1>(s1))
                         : fmaxf(attn_max_n_masked<32>(s0, 0, lim), attn_max_n_masked<N1>(s1, 32, lim));
      mx *= p.scale_log2;
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = mx == -INFINITY ? 0.f : mx;
      } else if (mx > m_used + 8.f) {
        alpha = fast_exp2(m_used - mx);
        m_used = mx;
        need = true;
      }
      uint32_t pk[BKV / 2];
      float part;
      if (no_mask) {
        if (use_ones) {
          part = bf16 ? attn_exp_n_fast<32, false, true, POLY>(s0, p.scale_log2, m_used, pk, 0) +
                            attn_exp_n_fast<N1, false, true, POLY>(s1, p.scale_log2, m_used, pk + 16, 16)
                      : attn_exp_n_fast<32, false, false, POLY>(s0, p.scale_log2, m_used, pk, 0) +
                            attn_exp_n_fast<N1, false, false, POLY>(s1, p.scale_log2, m_used, pk + 16, 16);
        } else {
          part = bf16 ? attn_exp_n_fast<32, true, true, POLY>(s0, p.scale_log2, m_used, pk, 0) +
                            attn_exp_n_fast<N1, true, true, POLY>(s1, p.scale_log2, m_used, pk + 16, 16)
                      : attn_exp_n_fast<32, true, false, POLY>(s0, p.scale_log2, m_used, pk, 0) +
                            attn_exp_n_fast<N1, true, false, POLY>(s1, p.scale_log2, m_used, pk + 16, 16);
        }
      } else {
        part = bf16 ? attn_exp_n<32, true, true, true>(s0, 0, lim, p.scale_log2, m_used, pk) +
                          attn_exp_n<N1, true, true, true>(s1, 32, lim, p.scale_log2, m_used, pk + 16)
                    : attn_exp_n<32, true, true, false>(s0, 0, lim, p.scale_log2, m_used, pk) +
                          attn_exp_n<N1, true, true, false>(s1, 32, lim, p.scale_log2, m_used, pk + 16);
      }
      // P.V_{j-1} must have retired before P_j overwrites P_{j-1} and before O is rescaled (long done in steady state)
      if (j >= 1) {
        mbar_wait(&bars->pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, need)) {
        l *= alpha;
#pragma unroll 1
        for (int c = 0; c < p.hd_cols; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(to + c, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x16(to + c, v);
        }
      }
      l += part;
      {
        uint32_t pa[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pa[i] = pk[i];
        tmem_st_32x32b_x16(tp, pa);
        if constexpr (BKV == 48) {
          tmem_st_32x32b_x8(tp + 16, pk + 16);
        } else {
          uint32_t pb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pb[i] = pk[16 + i];
          tmem_st_32x32b_x16(tp + 16, pb);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&bars->p_full);
    }
    // ---- epilogue: O / l
    mbar_wait(&bars->pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    if (use_ones) {
      l = __uint_as_float(tmem_ld_32x32b_x1(to + p.ones_col));
      tmem_wait_ld();
    }
    const float inv = 1.f / l;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + static_cast<long long>(b) * p.o_bstride +
                     static_cast<long long>(qrow) * p.ldo + col0;
#pragma unroll 1
    for (int c = 0; c < p.hd_cols; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(to + c, v);
      tmem_wait_ld();
      if (qrow < p.Lq) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
        store16(orow + c, f, 16, p.out_dtype);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ================================================================================================================
// attn4q: attn4 for SHORT key sequences (at most two key tiles: the UNet's 77-key cross-attention) with the CTA walking
// several query tiles. One (query tile, head, sample) per CTA made the cross-attention pure fixed cost -- TMEM allocation,
// barrier set-up and the Q / K / V round trips to L2 for ~1 us of work: 4096 CTAs in 7 waves, 46 us at 16x8x4096x77
// against a 15 us memory floor. Here K and V (both tiles) are loaded once per CTA and stay in the two slots, the CTA walks
// q_tiles_per_cta query tiles, the next Q tile is requested as soon as the last Q.K^T of the current one has retired, and
// the grid is sized to ONE wave. Barrier parities run on a single iteration counter over (query tile, key tile).
template <int BKV_, int KCH_>
__global__ void __launch_bounds__(160, Attn4Cfg<BKV_, KCH_>::CTAS_PER_SM) attn4q_kernel(const __grid_constant__ AttnParams p) {
  using C = Attn4Cfg<BKV_, KCH_>;
  constexpr int BKV = C::BKV, KCH = C::KCH;
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Attn4Bars* bars = reinterpret_cast<Attn4Bars*>(smem + C::OFF_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nq_all = (p.Lq + 127) / 128;
  const int qt0 = blockIdx.x * p.q_tiles_per_cta;
  const int nq = min(p.q_tiles_per_cta, nq_all - qt0);
  const int kv_len = p.kv_lens ? p.kv_lens[b] : p.Lk;
  int n_tiles = (kv_len + BKV - 1) / BKV;  // 1 or 2 (host-checked: Lk <= 2 * BKV)
  if (n_tiles < 1) n_tiles = 1;

  if (tid == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->k_full[0], 1);
    mbar_init(&bars->k_full[1], 1);
    mbar_init(&bars->v_full[0], 1);
    mbar_init(&bars->v_full[1], 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->pv_done, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  const int col0 = head * p.hd_cols;
  const int ksteps = p.hd_cols >> 4;
  const bool bf16 = p.in_dtype == DT_BF16;

  if (warp == 4) {
    if (lane_id() == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, BKV, bf16, false);
      const uint32_t idesc_o = make_idesc_f16(128, p.hd_cols, bf16, true);
      const uint32_t sq = smem_u32(smem + C::OFF_Q), sk = smem_u32(smem + C::OFF_K), sv = smem_u32(smem + C::OFF_V);
      auto load_q = [&](int t) {
        mbar_arrive_expect_tx(&bars->q_full, C::Q_BYTES);
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          tma_load_3d(smem + C::OFF_Q + c * C::Q_CHUNK, &p.tma_q, &bars->q_full, col0 + c * 64, (qt0 + t) * 128, b);
      };
      auto mma_s = [&](int slot) {
#pragma unroll
        for (int k = 0; k < 4 * KCH; ++k)
          if (k < ksteps)
            umma_f16(tmem_base + C::TMEM_S, make_smem_desc_sw128(sq + (k >> 2) * C::Q_CHUNK + (k & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(sk + slot * C::KV_BYTES + (k >> 2) * C::KV_CHUNK + (k & 3) * 32, 16, 1024),
                     idesc_s, k != 0 ? 1u : 0u);
      };
      load_q(0);
      for (int j = 0; j < n_tiles; ++j) {  // every key / value tile once: resident for all query tiles
        mbar_arrive_expect_tx(&bars->k_full[j], C::KV_BYTES);
        mbar_arrive_expect_tx(&bars->v_full[j], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          tma_load_3d(smem + C::OFF_K + j * C::KV_BYTES + c * C::KV_CHUNK, &p.tma_k, &bars->k_full[j], col0 + c * 64, j * BKV, b);
          tma_load_3d(smem + C::OFF_V + j * C::KV_BYTES + c * C::KV_CHUNK, &p.tma_v, &bars->v_full[j], col0 + c * 64, j * BKV, b);
        }
      }
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&bars->k_full[j], 0);
        mbar_wait(&bars->v_full[j], 0);
      }
      int it = 0;
      for (int t = 0; t < nq; ++t) {
        mbar_wait(&bars->q_full, t & 1);
        if (it > 0) mbar_wait(&bars->s_free, (it - 1) & 1);  // the previous tile's last scores sit in registers
        tc_fence_after();
        mma_s(0);
        umma_commit(&bars->s_full);
        for (int j = 0; j < n_tiles; ++j, ++it) {
          if (j + 1 < n_tiles) {
            mbar_wait(&bars->s_free, it & 1);
            tc_fence_after();
            mma_s(j + 1);
            umma_commit(&bars->s_full);
          }
          if (j == n_tiles - 1 && t + 1 < nq) {
            // the last Q.K^T of this query tile has retired (s_full of iteration `it`): the Q buffer is free
            mbar_wait(&bars->s_full, it & 1);
            load_q(t + 1);
          }
          mbar_wait(&bars->p_full, it & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            umma_f16_ts(tmem_base + C::TMEM_O, tmem_base + C::TMEM_P + k * 8,
                        make_smem_desc_sw128(sv + j * C::KV_BYTES + k * (16 * 128), C::KV_CHUNK, 1024), idesc_o,
                        (j | k) != 0 ? 1u : 0u);
          umma_commit(&bars->pv_done);
        }
      }
    }
  } else {
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t ts = tmem_base + C::TMEM_S + lane_off;
    const uint32_t tp = tmem_base + C::TMEM_P + lane_off;
    const uint32_t to = tmem_base + C::TMEM_O + lane_off;
    const bool use_ones = p.ones_col >= 0;
    int it = 0;
    for (int t = 0; t < nq; ++t) {
      const int qrow = (qt0 + t) * 128 + tid;
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++it) {
        mbar_wait(&bars->s_full, it & 1);
        tc_fence_after();
        constexpr int N1 = BKV - 32;
        uint32_t s0[32], s1[N1];
        tmem_ld_32x32b_x32(ts, s0);
        if constexpr (N1 == 16) tmem_ld_32x32b_x16(ts + 32, s1);
        else tmem_ld_32x32b_x32(ts + 32, s1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&bars->s_free);
        const int lim = kv_len - 1 - j * BKV;  // columns > lim are masked
        const bool no_mask = lim >= BKV - 1;
        float mx = no_mask ? fmaxf(attn_max_n<32>(s0), attn_max_n<N1>(s1))
                           : fmaxf(attn_max_n_masked<32>(s0, 0, lim), attn_max_n_masked<N1>(s1, 32, lim));
        mx *= p.scale_log2;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = mx == -INFINITY ? 0.f : mx;
        } else if (mx > m_used + 8.f) {
          alpha = fast_exp2(m_used - mx);
          m_used = mx;
          need = true;
        }
        uint32_t pk[BKV / 2];
        float part;
        if (no_mask) {
          if (use_ones) {
            part = bf16 ? attn_exp_n_fast<32, false, true, 0>(s0, p.scale_log2, m_used, pk, 0) +
                              attn_exp_n_fast<N1, false, true, 0>(s1, p.scale_log2, m_used, pk + 16, 16)
                        : attn_exp_n_fast<32, false, false, 0>(s0, p.scale_log2, m_used, pk, 0) +
                              attn_exp_n_fast<N1, false, false, 0>(s1, p.scale_log2, m_used, pk + 16, 16);
          } else {
            part = bf16 ? attn_exp_n_fast<32, true, true, 0>(s0, p.scale_log2, m_used, pk, 0) +
                              attn_exp_n_fast<N1, true, true, 0>(s1, p.scale_log2, m_used, pk + 16, 16)
                        : attn_exp_n_fast<32, true, false, 0>(s0, p.scale_log2, m_used, pk, 0) +
                              attn_exp_n_fast<N1, true, false, 0>(s1, p.scale_log2, m_used, pk + 16, 16);
          }
        } else {
          part = bf16 ? attn_exp_n<32, true, true, true>(s0, 0, lim, p.scale_log2, m_used, pk) +
                            attn_exp_n<N1, true, true, true>(s1, 32, lim, p.scale_log2, m_used, pk + 16)
                      : attn_exp_n<32, true, true, false>(s0, 0, lim, p.scale_log2, m_used, pk) +
                            attn_exp_n<N1, true, true, false>(s1, 32, lim, p.scale_log2, m_used, pk + 16);
        }
        if (it >= 1) {  // the previous P.V has retired: P may be overwritten, O may be rescaled
          mbar_wait(&bars->pv_done, (it - 1) & 1);
          tc_fence_after();
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < p.hd_cols; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(to + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(to + c, v);
          }
        }
        l += part;
        {
          uint32_t pa[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pa[i] = pk[i];
          tmem_st_32x32b_x16(tp, pa);
          if constexpr (BKV == 48) {
            tmem_st_32x32b_x8(tp + 16, pk + 16);
          } else {
            uint32_t pb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pb[i] = pk[16 + i];
            tmem_st_32x32b_x16(tp + 16, pb);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&bars->p_full);
      }
      // ---- this query tile's O / l (the next tile's first P.V waits for all four warps' next p_full arrival)
      mbar_wait(&bars->pv_done, (it - 1) & 1);
      tc_fence_after();
      if (use_ones) {
        l = __uint_as_float(tmem_ld_32x32b_x1(to + p.ones_col));
        tmem_wait_ld();
      }
      const float inv = 1.f / l;
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + static_cast<long long>(b) * p.o_bstride +
                       static_cast<long long>(qrow) * p.ldo + col0;
#pragma unroll 1
      for (int c = 0; c < p.hd_cols; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(to + c, v);
        tmem_wait_ld();
        if (qrow < p.Lq) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
          store16(orow + c, f, 16, p.out_dtype);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int BKV, int KCH>
static int launch_attn4q(AttnParams p, cudaStream_t stream) {
  using C = Attn4Cfg<BKV, KCH>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn4q_kernel<BKV, KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  // one wave: as few query tiles per CTA as keeps the whole grid resident
  const int nq = (p.Lq + 127) / 128;
  const long long slots = static_cast<long long>(C::CTAS_PER_SM) * num_sms();
  int qt = static_cast<int>((static_cast<long long>(nq) * p.H * p.B + slots - 1) / slots);
  if (qt < 1) qt = 1;
  while (static_cast<long long>((nq + qt - 1) / qt) * p.H * p.B > slots) ++qt;
  p.q_tiles_per_cta = qt;
  dim3 grid((nq + qt - 1) / qt, p.H, p.B);
  GB_CUDA(launch_pdl_light(attn4q_kernel<BKV, KCH>, grid, dim3(160), C::SMEM_BYTES, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}

template <int BKV, int KCH, int POLY>
static int launch_attn4_t(const AttnParams& p, cudaStream_t stream) {
  using C = Attn4Cfg<BKV, KCH>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn4_kernel<BKV, KCH, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  dim3 grid((p.Lq + 127) / 128, p.H, p.B);
  GB_CUDA(launch_pdl_light(attn4_kernel<BKV, KCH, POLY>, grid, dim3(160), C::SMEM_BYTES, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}
template <int BKV, int KCH>
static int launch_attn4(const AttnParams& p, cudaStream_t stream) {
  static int poly = -1;
  if (poly < 0) {
    // every n-th pair of exponentials on the FMA pipe; 0: all on the MUFU. Measured (profiles/r02_attn4_poly.log, 16 x 8
    // heads): 48-key tiles, 4096^2, hd 40: 665 (0) / 618 (6) / 611 (4) / 606 (3) / 646 us (2); 64-key tiles, 1024^2, hd 80
    // (2 CTAs per SM, bound by the MMA round trips rather than the XU pipe): 88.9 (0) / 91.0 (3) / 90.8 us (4)
    const char* e = getenv("GILLB200_ATTN_POLY");
    poly = e ? atoi(e) : (KCH == 1 ? 3 : 0);
  }
  if (poly == 2) return launch_attn4_t<BKV, KCH, 2>(p, stream);
  if (poly == 3) return launch_attn4_t<BKV, KCH, 3>(p, stream);
  if (poly == 4) return launch_attn4_t<BKV, KCH, 4>(p, stream);
  return launch_attn4_t<BKV, KCH, 0>(p, stream);
}

static int launch_attn3(const AttnParams& p, cudaStream_t stream) {
  using C = Attn2Cfg<64>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  dim3 grid((p.Lq + 127) / 128, p.H, p.B);
  GB_CUDA(launch_pdl_light(attn3_kernel, grid, dim3(160), C::SMEM_BYTES, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}

template <int HD_PAD, bool PT>
static int launch_attn2(const AttnParams& p, cudaStream_t stream) {
  using C = Attn2Cfg<HD_PAD>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn2_kernel<HD_PAD, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  dim3 grid((p.Lq + 127) / 128, p.H, p.B);
  GB_CUDA(launch_pdl_light(attn2_kernel<HD_PAD, PT>, grid, dim3(128), C::SMEM_BYTES, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}

template <int HD_PAD, int BLOCK_KV>
static int launch_attn(const AttnParams& p, cudaStream_t stream) {
  using C = AttnCfg<HD_PAD, BLOCK_KV>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn_kernel<HD_PAD, BLOCK_KV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::SMEM_BYTES));
  }
  dim3 grid((p.Lq + 127) / 128, p.H, p.B);
  GB_CUDA(launch_pdl_light(attn_kernel<HD_PAD, BLOCK_KV>, grid, dim3(256), C::SMEM_BYTES, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}

}  // namespace gb

using namespace gb;

extern "C" int gillb200_attention(const gillb200_attn_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(a && a->q && a->k && a->v && a->out, "null pointer");
  GB_CHECK_ARG(a->hd_pad == 64 || a->hd_pad == 128 || a->hd_pad == 192, "hd_pad must be 64, 128 or 192 (got %d)",
               a->hd_pad);
  GB_CHECK_ARG(a->dtype == DT_BF16 || a->dtype == DT_F16, "attention operands must be bf16 or fp16");
  GB_CHECK_ARG(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "bad attention shape");
  const int hd_cols = a->head_stride > 0 ? a->head_stride : a->hd_pad;
  GB_CHECK_ARG(hd_cols % 16 == 0 && hd_cols >= 16 && hd_cols <= a->hd_pad,
               "head_stride must be a multiple of 16 in [16, hd_pad] (got %d, hd_pad %d)", hd_cols, a->hd_pad);
  GB_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "row strides %% 8");
  GB_CHECK_ARG(a->q_bstride % 8 == 0 && a->k_bstride % 8 == 0 && a->v_bstride % 8 == 0, "batch strides %% 8");
  const bool bf16 = a->dtype == DT_BF16;
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("GILLB200_ATTN");  // A/B aid. "1": warp-specialised attn_kernel for every head size;
    impl = e ? atoi(e) : 0;                   //          "2": 4-warp attn2 for every head size
  }
  // measured on B200 (profiles/r02_attn_variants.log, 16x8 heads): the 4-warp attn2 form wins for 64- and 128-wide tiles
  // (hd 80 self-attention 120 vs 132 us, its 77-key cross-attention 33.5 vs 43.9 us); the 192-wide tile and the causal
  // bf16 OPT prefill keep the warp-specialised kernel (23.3 vs 26.1 us at 16x16)
  const bool use_attn2 = impl == 2 || (impl != 1 && (a->hd_pad == 64 || (a->hd_pad == 128 && !a->causal)));
  static int attn4 = -1;
  if (attn4 < 0) {
    const char* e = getenv("GILLB200_ATTN4");  // "0": 64-key attn3 instead of the 48-key decoupled-S kernel (A/B aid)
    attn4 = e ? atoi(e) : 3;
  }
  // attn4 bit 0: hd_cols <= 48 (48-key tiles, 4 CTAs per SM); bit 1: hd_cols <= 96 under hd_pad 128 (64-key tiles, 2 CTAs)
  const bool use_attn4 = use_attn2 && impl == 0 && (attn4 & 1) && a->hd_pad == 64 && hd_cols <= 48;
  const bool use_attn4w = use_attn2 && impl == 0 && (attn4 & 2) && a->hd_pad == 128 && hd_cols <= 96 && !a->causal;
  const int bkv = use_attn4 ? 48 : (!use_attn2 && a->hd_pad == 128) ? 128 : 64;  // K/V TMA box rows = the kernel's KV tile
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t cols = (uint64_t)a->H * hd_cols;  // boxes that reach past the last head are zero filled by TMA
  {
    const uint64_t dims[3] = {cols, (uint64_t)a->Lq, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->q_bstride * 2};
    const uint32_t box[3] = {64, 128, 1};
    int r = encode_tmap_16bit(&p.tma_q, a->q, 3, dims, str, box, bf16);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {cols, (uint64_t)a->Lk, (uint64_t)a->B};
    const uint64_t strk[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->k_bstride * 2};
    const uint64_t strv[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->v_bstride * 2};
    const uint32_t box[3] = {64, (uint32_t)bkv, 1};
    int r = encode_tmap_16bit(&p.tma_k, a->k, 3, dims, strk, box, bf16);
    if (r) return r;
    r = encode_tmap_16bit(&p.tma_v, a->v, 3, dims, strv, box, bf16);
    if (r) return r;
  }
  p.out = a->out;
  p.ldo = a->ldo;
  p.o_bstride = a->o_bstride;
  p.kv_lens = a->kv_lens;
  p.B = a->B;
  p.H = a->H;
  p.Lq = a->Lq;
  p.Lk = a->Lk;
  p.causal = a->causal;
  p.causal_offset = a->causal_offset;
  p.out_dtype = a->dtype;
  p.in_dtype = a->dtype;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.hd_cols = hd_cols;
  p.ones_col = (a->ones_col > 0 && a->ones_col < hd_cols) ? a->ones_col : -1;
  if (use_attn2) {
    static int ptmem = -1;
    if (ptmem < 0) {
      const char* e = getenv("GILLB200_ATTN_PTMEM");  // "0": P through shared memory (A/B aid)
      ptmem = e ? atoi(e) : 1;
    }
    static int issuer = -1;
    if (issuer < 0) {
      // dedicated issuer warp (attn3) for the 64-wide tile: measured 726 vs 766 us (16x8x4096^2, hd 40) and 51.6 vs 54.2 us
      // for the 77-key cross-attention; "0" falls back to attn2 (A/B aid)
      const char* e = getenv("GILLB200_ATTN_ISSUER");
      issuer = e ? atoi(e) : 1;
    }
    static int attn4q = -1;
    if (attn4q < 0) {
      const char* e = getenv("GILLB200_ATTN4Q");  // "0": one query tile per CTA also for short key sequences (A/B aid)
      attn4q = e ? atoi(e) : 1;
    }
    if (attn4q && !a->causal && a->Lq > 128) {
      if (use_attn4 && a->Lk <= 96) return launch_attn4q<48, 1>(p, stream);
      if (use_attn4w && a->Lk <= 128) return launch_attn4q<64, 2>(p, stream);
    }
    if (use_attn4) return launch_attn4<48, 1>(p, stream);
    if (use_attn4w) return launch_attn4<64, 2>(p, stream);
    if (issuer && a->hd_pad == 64) return launch_attn3(p, stream);
    if (ptmem) {
      if (a->hd_pad == 64) return launch_attn2<64, true>(p, stream);
      if (a->hd_pad == 128) return launch_attn2<128, true>(p, stream);
      return launch_attn2<192, true>(p, stream);
    }
    if (a->hd_pad == 64) return launch_attn2<64, false>(p, stream);
    if (a->hd_pad == 128) return launch_attn2<128, false>(p, stream);
    return launch_attn2<192, false>(p, stream);
  }
  if (a->hd_pad == 64) return launch_attn<64, 64>(p, stream);
  if (a->hd_pad == 128) return launch_attn<128, 128>(p, stream);
  return launch_attn<192, 64>(p, stream);
}
