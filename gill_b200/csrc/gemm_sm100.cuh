// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        A, B: 16-bit (bf16 or fp16), K-major; fp32 accumulate in TMEM.
//
// One CTA per SM, 256 threads:
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16 per instruction)
//   warp 2        : TMEM allocator
//   warps 4..19   : epilogue      (tcgen05.ld 32x32b -> registers -> bias/act/residual -> vectorised global stores);
//                   warps w, w+4, w+8, w+12 share a TMEM lane quadrant and split the tile's 16-column chunks
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// The A operand has two addressing modes:
//   A_PLAIN   : 2D tensor map {K, M}; optional second source (tma_a2) for K-concatenation / split-precision.
//   A_CONV3X3 : implicit GEMM for a 3x3, stride-1, pad-1 convolution over an NHWC activation. The tensor map is
//               4D {C, W, H, B}; each 128-row M tile is a (bw x bh x bb) pixel box and each K block is one
//               (tap, 64-channel) slice fetched with the box shifted by (dx, dy). TMA zero-fills out-of-bounds
//               coordinates, which is exactly the zero padding. No im2col buffer is ever materialised.
#pragma once
#include "ptx.cuh"

namespace gb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 16;  // 4 per SM sub-partition: the CUDA-core epilogue needs the latency hiding
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;  // warps 0-3: producer / MMA / TMEM alloc / spare
constexpr int SMEM_BUDGET = 227 * 1024;

enum { A_PLAIN = 0, A_CONV3X3 = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3, ACT_GEGLU = 4 };
enum { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };

struct alignas(64) GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_a2;
  CUtensorMap tma_b;
  int M, N;
  int num_k_blocks;  // total 64-wide K blocks
  int kb_split;      // blocks [0, kb_split) read tma_a, the rest read tma_a2 (plain 2D) at block (kb - kb_split)
  int b_kb_wrap;     // B k-block = kb % b_kb_wrap  (split-precision A re-reads the same weight block)
  int a_mode;
  int conv_cblocks;  // Cin / 64
  int conv_W, conv_H;
  // epilogue
  void* out;
  void* out_lo;  // optional bf16 "lo" residue: out_lo = bf16(v - float(bf16(v)))   (split-precision activations)
  const float* bias;
  const float* rowbias;
  const void* residual;
  long long ldo, ldr, ld_rowbias;
  int out_dtype, res_dtype;
  int bias_along_m;
  int rows_per_group;
  int act;
  int in_dtype;  // DT_BF16 or DT_F16
  float alpha;
  int tile_order;  // 0 auto, 1 m-fastest, 2 n-inner (see tile_coord)
  int debug_mode;  // 0 normal. 1: no TMA after the ring is primed (MMA ceiling). 2: no MMA issue (TMA-fill ceiling). Results invalid.
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - BAR_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static constexpr int ACC_STRIDE = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-B alignment");
};

// exact-erf GELU. (An Abramowitz-Stegun rcp+ex2 variant was measured 35% SLOWER inside the epilogue: with two epilogue
// warps per scheduler the XU/MUFU pipe and its latency are the scarce resource, erff's FMA-only polynomial is not.)
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_SILU) return v / (1.f + __expf(-v));
  return v;
}

// Shared main loop pieces -------------------------------------------------------------------------------------

struct TileCoord {
  int m_blk, n_blk;
};
// num_m > 0: m-fastest (CTAs running side by side share a B tile). num_m < 0 encodes "n-inner" order with |num_m| = num_n:
// a CTA walks all N tiles of one M block back to back, so the big streaming A operand is fetched from HBM once
// (used when there are only a few N tiles, e.g. N = 320 with BLOCK_N = 160).
__device__ __forceinline__ TileCoord tile_coord(int tile, int num_m) {
  if (num_m > 0) return {tile % num_m, tile / num_m};
  const int num_n = -num_m;
  return {tile / num_n, tile % num_n};
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_producer(const GemmParams& p, uint8_t* smem_tiles, uint64_t* full, uint64_t* empty,
                                              int num_m, int tile_begin, int tile_end, int tile_step) {
  using C = GemmCfg<BLOCK_N>;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
    const TileCoord tc = tile_coord(tile, num_m);
    const int m0 = tc.m_blk * BLOCK_M;
    const int n0 = tc.n_blk * BLOCK_N;
    int cb0 = 0, cy0 = 0, cx0 = 0;
    if (p.a_mode == A_CONV3X3) {
      const int per_img = p.conv_W * p.conv_H;
      cb0 = m0 / per_img;
      cy0 = (m0 % per_img) / p.conv_W;
      cx0 = m0 % p.conv_W;  // non-zero only when W > 128 (a tile is then a 128-pixel row segment)
    }
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      // The whole warp runs this loop (warp-uniform control flow) and one elected lane issues: with a single
      // divergent lane the compiler has to wrap every UTMALDG / UTCHMMA in an ELECT + BRA.U.ANY loop and shuttle
      // operands through R2UR, which made the issue loop ~430 cycles per k-block (measured, profiles/).
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem_tiles + stage * C::STAGE_BYTES;
      uint8_t* sb = sa + C::A_BYTES;
      if (p.debug_mode == 1 && (phase != 0 || tile != tile_begin)) {  // measurement only: reuse stale smem
        if (elect_one()) mbar_arrive(&full[stage]);
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
        continue;
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
        if (kb >= p.kb_split) {
          tma_load_2d(sa, &p.tma_a2, &full[stage], (kb - p.kb_split) * BLOCK_K, m0);
        } else if (p.a_mode == A_CONV3X3) {
          const int tap = kb / p.conv_cblocks;
          const int cb = kb - tap * p.conv_cblocks;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &p.tma_a, &full[stage], cb * BLOCK_K, cx0 + dx, cy0 + dy, cb0);
        } else {
          tma_load_2d(sa, &p.tma_a, &full[stage], kb * BLOCK_K, m0);
        }
        tma_load_2d(sb, &p.tma_b, &full[stage], (kb % p.b_kb_wrap) * BLOCK_K, n0);
      }
      __syncwarp();
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_mma(const GemmParams& p, uint8_t* smem_tiles, uint64_t* full, uint64_t* empty,
                                         uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                         int tile_begin, int tile_end, int tile_step) {
  using C = GemmCfg<BLOCK_N>;
  const uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N, p.in_dtype == DT_BF16, false);
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem_tiles + stage * C::STAGE_BYTES);
      const uint32_t sb = sa + C::A_BYTES;
      const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
      if (elect_one()) {
        if (p.debug_mode != 2) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in 16-B units
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
      }
      __syncwarp();
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
    __syncwarp();
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
}

// Default epilogue ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void store16(void* dst, const float* v, int n, int dtype) {
  // stores n (<=16, multiple of 8 for the vector path) values starting at dst
  if (dtype == DT_F32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i * 4 < n) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i * 8 < n) {
        uint4 u;
        if (dtype == DT_BF16) {
          u.x = pack_bf16x2(v[8 * i], v[8 * i + 1]);
          u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        } else {
          u.x = pack_f16x2(v[8 * i], v[8 * i + 1]);
          u.y = pack_f16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_f16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_f16x2(v[8 * i + 6], v[8 * i + 7]);
        }
        d4[i] = u;
      }
    }
  }
}

__device__ __forceinline__ float load_elem(const void* base, long long idx, int dtype) {
  if (dtype == DT_F32) return reinterpret_cast<const float*>(base)[idx];
  if (dtype == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}
__device__ __forceinline__ void store_elem(void* base, long long idx, float v, int dtype) {
  if (dtype == DT_F32)
    reinterpret_cast<float*>(base)[idx] = v;
  else if (dtype == DT_BF16)
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// Processes 16 accumulator columns [n, n+16) of one row. `v` holds the raw fp32 accumulators.
__device__ __forceinline__ void epilogue_chunk16(const GemmParams& p, int row, int n, float* v) {
  const bool geglu = p.act == ACT_GEGLU;
  const int n_out_total = geglu ? p.N / 2 : p.N;
  if (p.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= p.alpha;
  }
  if (p.bias) {
    if (p.bias_along_m) {
      const float b = p.bias[row];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += b;
    } else if (n + 16 <= p.N && (reinterpret_cast<uintptr_t>(p.bias + n) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(b4 + i);
        v[4 * i] += t.x;
        v[4 * i + 1] += t.y;
        v[4 * i + 2] += t.z;
        v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < p.N) v[j] += __ldg(p.bias + n + j);
    }
  }
  if (p.rowbias) {
    const float* rb = p.rowbias + static_cast<long long>(row / p.rows_per_group) * p.ld_rowbias;
    if (n + 16 <= p.N && (reinterpret_cast<uintptr_t>(rb + n) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(rb + n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(b4 + i);
        v[4 * i] += t.x;
        v[4 * i + 1] += t.y;
        v[4 * i + 2] += t.z;
        v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < p.N) v[j] += __ldg(rb + n + j);
    }
  }
  int cnt = 16, no = n;
  if (geglu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = v[2 * j], g = v[2 * j + 1];
      v[j] = a * gelu_erf(g);
    }
    cnt = 8;
    no = n / 2;
  } else if (p.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_apply(v[j], p.act);
  }
  const bool full = (no + cnt <= n_out_total);
  if (p.residual) {
    const long long roff = static_cast<long long>(row) * p.ldr + no;
    if (full && (p.ldr % 8 == 0)) {
      if (p.res_dtype == DT_F32) {
        const float4* r4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + roff);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i * 4 < cnt) {
            const float4 r = r4[i];
            v[4 * i] += r.x;
            v[4 * i + 1] += r.y;
            v[4 * i + 2] += r.z;
            v[4 * i + 3] += r.w;
          }
        }
      } else {
        const uint4* r4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + roff);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i * 8 < cnt) {
            const uint4 r = r4[i];
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = p.res_dtype == DT_BF16 ? unpack_bf16x2(w[q]) : unpack_f16x2(w[q]);
              v[8 * i + 2 * q] += f.x;
              v[8 * i + 2 * q + 1] += f.y;
            }
          }
        }
      }
    } else {
      for (int j = 0; j < cnt; ++j)
        if (no + j < n_out_total) v[j] += load_elem(p.residual, roff + j, p.res_dtype);
    }
  }
  const long long ooff = static_cast<long long>(row) * p.ldo + no;
  if (full && (p.ldo % 8 == 0)) {
    if (p.out_dtype == DT_F32)
      store16(reinterpret_cast<float*>(p.out) + ooff, v, cnt, DT_F32);
    else
      store16(reinterpret_cast<uint16_t*>(p.out) + ooff, v, cnt, p.out_dtype);
    if (p.out_lo) {
      float lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
      store16(reinterpret_cast<uint16_t*>(p.out_lo) + ooff, lo, cnt, DT_BF16);
    }
  } else {
    for (int j = 0; j < cnt; ++j) {
      if (no + j < n_out_total) {
        store_elem(p.out, ooff + j, v[j], p.out_dtype);
        if (p.out_lo)
          store_elem(p.out_lo, ooff + j, v[j] - __bfloat162float(__float2bfloat16_rn(v[j])), DT_BF16);
      }
    }
  }
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              uint32_t tmem_base, int num_m, int num_tiles) {
  using C = GemmCfg<BLOCK_N>;
  const int ewarp = (threadIdx.x >> 5) & 3;         // TMEM lane quadrant this warp may access
  const int cgrp = ((threadIdx.x >> 5) - 4) >> 2;   // which share of the tile's columns this warp drains
  constexpr int NGRP = GEMM_EPI_WARPS / 4;
  constexpr int NCH = BLOCK_N / 16;                 // 16-column chunks per tile, dealt out as evenly as possible
  const int ch_begin = cgrp * (NCH / NGRP) + min(cgrp, NCH % NGRP);
  const int ch_end = ch_begin + NCH / NGRP + (cgrp < NCH % NGRP ? 1 : 0);
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const TileCoord tc = tile_coord(tile, num_m);
    const int row = tc.m_blk * BLOCK_M + ewarp * 32 + lane_id();
    const int n0 = tc.n_blk * BLOCK_N;
    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + acc * C::ACC_STRIDE + (static_cast<uint32_t>(ewarp * 32) << 16);
#pragma unroll 1
    for (int ch = ch_begin; ch < ch_end; ++ch) {
      const int c = ch * 16;
      uint32_t r[16];
      tmem_ld_32x32b_x16(taddr + c, r);
      tmem_wait_ld();
      if (row < p.M && n0 + c < p.N) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_chunk16(p, row, n0 + c, v);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
}

// Kernel -------------------------------------------------------------------------------------------------------

struct GemmSmemBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_ptr;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_tiles = smem;
  GemmSmemBars* bars = reinterpret_cast<GemmSmemBars*>(smem + C::STAGES * C::STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    if (p.kb_split < p.num_k_blocks) tma_prefetch_desc(&p.tma_a2);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], GEMM_EPI_WARPS);
    }
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

  const bool n_inner = p.tile_order == 2 || (p.tile_order == 0 && num_n <= 4 && num_m >= 2 * static_cast<int>(gridDim.x));
  const int order = n_inner ? -num_n : num_m;  // see tile_coord()
  if (warp == 0) {
    gemm_producer<BLOCK_N>(p, smem_tiles, bars->full, bars->empty, order, blockIdx.x, num_tiles, gridDim.x);
  } else if (warp == 1) {
    gemm_mma<BLOCK_N>(p, smem_tiles, bars->full, bars->empty, bars->tmem_full, bars->tmem_empty, tmem_base,
                      blockIdx.x, num_tiles, gridDim.x);
  } else if (warp >= 4) {
    gemm_epilogue<BLOCK_N>(p, bars->tmem_full, bars->tmem_empty, tmem_base, order, num_tiles);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace gb
