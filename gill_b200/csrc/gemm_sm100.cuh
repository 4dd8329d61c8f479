// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        A, B: 16-bit (bf16 or fp16), K-major; fp32 accumulate in TMEM.
//
// One CTA per SM, 384 threads:
//   warp 0        : TMA producer  (one elected lane: cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1        : MMA issuer    (one elected lane: tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16 per instruction)
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue. Staged form (default): tcgen05.ld -> registers -> bias / activation / residual -> swizzled
//                   shared-memory panel -> TMA store (+ optional GroupNorm statistics, + stream-K partial / fix-up);
//                   direct form (odd layouts only): per-lane vectorised global stores.
//                   Warps w and w+4 share a TMEM lane quadrant and split the tile's 32-column panels.
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// The A operand has two addressing modes:
//   A_PLAIN   : 2D tensor map {K, M}; optional second source (tma_a2) for K-concatenation / split-precision.
//   A_CONV3X3 : implicit GEMM for a 3x3, pad-1 convolution (stride 1 or 2) over an NHWC activation. The tensor map is
//               4D {C, W, H, B}; each 128-row M tile is a (bw x bh x bb) pixel box and each K block is one
//               (tap, 64-channel) slice fetched with the box shifted by (dx, dy). TMA zero-fills out-of-bounds
//               coordinates, which is exactly the zero padding. No im2col buffer is ever materialised.
#pragma once
#include "ptx.cuh"

namespace gb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
#ifndef GILLB200_GELU_ESTRIN
#define GILLB200_GELU_ESTRIN 0  // measured: 137.9 vs 131-135 us (M65536 N2560 K320 GEGLU) -- no gain, Horner keeps the scalar form bit for bit
#endif
#ifndef GILLB200_EPI_PIPELINE
#define GILLB200_EPI_PIPELINE 0
#endif
#ifdef GILLB200_GEMM_TRACE
constexpr bool GEMM_TRACE = true;   // build with -DGILLB200_GEMM_TRACE: debug_mode 3 records CTA 0's wait times (tools/gpu_gemm_trace.py)
#else
constexpr bool GEMM_TRACE = false;  // (default build: the trace code folds away)
#endif
constexpr bool EPI_PIPELINE = GILLB200_EPI_PIPELINE != 0;  // staged epilogue: prefetch the next accumulator chunk (A/B: rebuild with 0)
constexpr int GEMM_EPI_WARPS = 8;  // 2 per SM sub-partition; 384 threads leave 168 registers per thread (16 warps capped it at 96: spills)
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;  // warps 0-3: producer / MMA / TMEM alloc / spare
constexpr int SMEM_BUDGET = 227 * 1024;

// A_CONV3X3_HALO (CTA-pair kernel only, stride 1, W and H multiples of 16): an M tile is a 16 (y) x 8 (x) block of
// output pixels and its A operand for ALL NINE taps of a 64-channel block is ONE 18 x 10 pixel halo tile in shared
// memory (one 4-D TMA box, OOB zero fill = padding); tap (dy, dx) is the same tile read through a descriptor whose
// start address is shifted by (dy * 10 + dx) 128-byte rows, 8-row groups 1280 bytes apart. The SWIZZLE_128B pattern is
// a function of the absolute shared-memory address on both sides (TMA write, tcgen05.mma read), so shifted starts and
// SBO = 1280 read exactly the rows they name (tools/exp/shifted_desc.cu, profiles/r02_shifted_desc_experiment.log).
// The K loop runs 64-channel block outer, tap inner; A traffic per CTA drops from 9 x 16 KB to 23 KB per channel block.
enum { A_PLAIN = 0, A_CONV3X3 = 1, A_CONV3X3_HALO = 2 };
constexpr int HALO_BYTES = 24576;           // slot size (1024-aligned); the box itself is 18 * 10 * 128 = 23040 bytes
constexpr int HALO_TX = 18 * 10 * 128;
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3, ACT_GEGLU = 4, ACT_QUICK_GELU = 5 };
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
  int conv_W, conv_H;  // OUTPUT width / height (= input size / conv_stride)
  int conv_stride;     // 1, or 2: the tensor map walks the input with element strides {1,2,2,1} (no im2col buffer)
  // A_CONV3X3_HALO tap walk: conv_nt taps per channel block; tap t reads the halo tile shifted by
  // (conv_pa + t / conv_ntx) rows and (conv_pb + t % conv_ntx) pixels. 3x3: nt 9, ntx 3, pa = pb = 0. Phase (a, b) of an
  // upsample-fused conv (gillb200_gemm_args::conv_phase): nt 4, ntx 2, pa = a, pb = b.
  int conv_nt, conv_ntx, conv_pa, conv_pb;
  // stats_out slab remap for phase launches: rows [b * stats_hw, (b + 1) * stats_hw) of this launch are sample b's; its
  // slabs go to ((b * 4 + stats_phase) * stats_hw + r) / 32 of the full-resolution tensor's buffer. 0: slab = row / 32.
  int stats_hw, stats_phase;
  // A_CONV3X3_HALO with GroupNorm(+SiLU) of the input applied in shared memory (gemm2_kernel<BN, true, true>): two warps of
  // each CTA rewrite every halo tile in place -- y = x * scale + shift per (sample, channel), SiLU, zero outside the image --
  // before the MMAs read it; ONE pass per input element per CTA tile instead of a separate elementwise kernel over the
  // whole tensor. halo_c0_blocks: channel blocks [0, halo_c0_blocks) come from tma_a, the rest from tma_a2 (cat source).
  const float* gn_ss;
  int gn_silu, gn_C, halo_c0_blocks;
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
  // smem-staged epilogue (epi_tma != 0): each epilogue warp writes 32-row x 32-column output panels into swizzled
  // shared memory and one lane hands them to TMA (tma_out); the residual panel is TMA-loaded into the same buffer
  // ahead of time (tma_res). Global memory only ever sees full-line bulk transactions.
  CUtensorMap tma_out;
  CUtensorMap tma_res;
  CUtensorMap tma_out_lo;  // second store target when out_lo is set (generic staged epilogue, bf16, no residual)
  int epi_stg;        // compile-time 16-bit variants: write staged panels back with per-lane 16-byte stores instead of TMA
  int epi_tma;
  int epi_variant;    // EV_* specialisation of the staged epilogue
  int epi_warps;      // epilogue warps that take part (multiple of 4, <= GEMM_EPI_WARPS)
  int epi_nbuf;       // staging buffers per warp: 2; with a residual 3 (next panel's residual in flight) or 4 (next two)
  int epi_buf_bytes;  // 2048 (16-bit output) or 4096 (fp32 output)
  int num_stages;     // smem ring depth actually used (<= GemmCfg::STAGES)
  // stream-K (sk_per > 0, 1-CTA kernel + staged epilogue only): the tiles' k-blocks form one linear range of
  // num_tiles * num_k_blocks units and CTA c owns units [c * sk_per, (c + 1) * sk_per). A tile cut by a CTA boundary
  // is finished by the CTA that owns its first k-block; the others park fp32 partial accumulators in sk_ws[cta] and
  // raise sk_flags[cta][epilogue warp].
  int sk_per;
  float* sk_ws;
  int* sk_flags;
  // optional GroupNorm statistics of the OUTPUT (staged epilogue, 16-bit output): per 32-row slab and output column the
  // sum and the sum of squares of the rounded values, stats_out[(row / 32) * N_out + col] = {sum, sumsq}. The consumer
  // (gillb200_groupnorm_from_stats) reduces slabs and channels per (sample, group) -- the separate statistics pass
  // over the activation (one full read) disappears.
  float2* stats_out;
  // LayerNorm folded into the NEXT GEMM. Producer side: rowstats_out[panel * rowstats_ld + row] = {sum, sumsq} of this
  // row's 32 output columns of `panel` (global 32-column index) -- the consumer adds the C/32 partials of a row.
  // Consumer side (EV_LN_* variants): the weights carry the LayerNorm scale (W' = W diag(gamma)), so
  //   LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (b + W beta);   ln_cs = colsum(W'), bias = b + W beta.
  float2* rowstats_out;
  const float2* ln_stats;
  const float* ln_cs;
  int rowstats_ld, ln_ld, ln_np, ln_C;
  float ln_eps;
  // split-K of the CTA-pair kernel (ksplit > 1): every 256-row tile is cut into ksplit K slices of kb_per_split k-blocks;
  // slice s of a tile is an independent work unit that writes its fp32 partial tile to rows [s * M, (s + 1) * M) of the
  // (fp32) output, which the host points at a scratch buffer; splitk_reduce_kernel adds the slices and applies the real
  // epilogue. For the UNet's 8x8 / 16x16 convs, whose 16-32 wide tiles cannot fill 74 SM pairs.
  // B-stationary mode of the 1-CTA kernel (b_resident != 0; plain A, short K, n-inner tile order with the grid a multiple
  // of the N-tile count so that a CTA keeps ONE N tile): the CTA's whole [BLOCK_N x K] weight tile is loaded once into
  // shared memory in front of the (A-only) ring; per tile only the 128 x K activation block travels. The 64x64-level
  // linears (K = 320 / 384) are bound by operand traffic: 180 KB per 128 x 160 x 320 tile without this, 80 KB with it.
  int b_resident, b_res_bytes;
  int ksplit, kb_per_split;
  int M_out;  // rows of the output matrix the epilogue may touch: M, or ksplit * M for a split-K launch
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

// Branch-free GELU for the specialised GEGLU epilogue: erf(x) = x * P(x^2) on |x| <= 3 (degree-9 Chebyshev fit, clamped
// outside; |erf error| <= 2.2e-5 = 1 - erf(3)), 16 FMA-pipe instructions instead of erff's ~25 (7 FFMA + 9 FSEL + 4 FMUL
// + MUFU). |gelu error| <= 7e-5 absolute (1.1e-5 * |g|), below the fp16 rounding of the value it multiplies.
__device__ __forceinline__ float gelu_poly(float g) {
  const float x = fminf(fmaxf(g * 0.70710678118654752f, -3.f), 3.f);
  const float t = x * x;
  float p = -4.469938970e-09f;
  p = fmaf(p, t, 2.302152890e-07f);
  p = fmaf(p, t, -5.322654538e-06f);
  p = fmaf(p, t, 7.394709949e-05f);
  p = fmaf(p, t, -7.009955072e-04f);
  p = fmaf(p, t, 4.897189191e-03f);
  p = fmaf(p, t, -2.645343569e-02f);
  p = fmaf(p, t, 1.125671519e-01f);
  p = fmaf(p, t, -3.760564203e-01f);
  p = fmaf(p, t, 1.128376151e+00f);
  const float hg = 0.5f * g;
  return fmaf(hg, p * x, hg);
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction) for the GEGLU epilogue, which is
// bound by FMA-pipe issue slots (ncu r01: tensor pipe 39 %, issue slots 51 % with two epilogue warps per scheduler).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_poly on two gates at once: same polynomial, same clamp, bit-identical per lane to the scalar form (every step
// is the same rounded fp32 operation); 14 packed + 4 min/max instructions per pair instead of 32.
__device__ __forceinline__ f32x2 gelu_poly2(f32x2 g) {
  const f32x2 xs = mul2(g, pk2(0.70710678118654752f, 0.70710678118654752f));
  float x0, x1;
  upk2(xs, x0, x1);
  x0 = fminf(fmaxf(x0, -3.f), 3.f);
  x1 = fminf(fmaxf(x1, -3.f), 3.f);
  const f32x2 x = pk2(x0, x1);
  const f32x2 t = mul2(x, x);
#define GB_C2(c) pk2(c, c)
#if GILLB200_GELU_ESTRIN
  // Estrin evaluation (12 packed operations, dependency depth 5 instead of Horner's 10): the epilogue warps run at ~0.23 IPC
  // on dependent FFMA2 chains that the compiler does not interleave at 168 registers (SASS of the GEGLU epilogue)
  const f32x2 t2 = mul2(t, t), t4 = mul2(t2, t2), t8 = mul2(t4, t4);
  const f32x2 a0 = fma2(GB_C2(-3.760564203e-01f), t, GB_C2(1.128376151e+00f));
  const f32x2 a1 = fma2(GB_C2(-2.645343569e-02f), t, GB_C2(1.125671519e-01f));
  const f32x2 a2 = fma2(GB_C2(-7.009955072e-04f), t, GB_C2(4.897189191e-03f));
  const f32x2 a3 = fma2(GB_C2(-5.322654538e-06f), t, GB_C2(7.394709949e-05f));
  const f32x2 a4 = fma2(GB_C2(-4.469938970e-09f), t, GB_C2(2.302152890e-07f));
  const f32x2 b0 = fma2(a1, t2, a0), b1 = fma2(a3, t2, a2);
  f32x2 p = fma2(b1, t4, b0);
  p = fma2(a4, t8, p);
#else
  f32x2 p = GB_C2(-4.469938970e-09f);
  p = fma2(p, t, GB_C2(2.302152890e-07f));
  p = fma2(p, t, GB_C2(-5.322654538e-06f));
  p = fma2(p, t, GB_C2(7.394709949e-05f));
  p = fma2(p, t, GB_C2(-7.009955072e-04f));
  p = fma2(p, t, GB_C2(4.897189191e-03f));
  p = fma2(p, t, GB_C2(-2.645343569e-02f));
  p = fma2(p, t, GB_C2(1.125671519e-01f));
  p = fma2(p, t, GB_C2(-3.760564203e-01f));
  p = fma2(p, t, GB_C2(1.128376151e+00f));
#endif
  const f32x2 hg = mul2(g, GB_C2(0.5f));
#undef GB_C2
  return fma2(hg, mul2(p, x), hg);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_SILU) return v / (1.f + __expf(-v));
  if (act == ACT_QUICK_GELU) return v / (1.f + __expf(-1.702f * v));
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

// One unit of work for a CTA: k-blocks [kb0, kb1) of `tile`. Plain persistent scheduling hands out whole tiles
// (tile = first, first + step, ...); stream-K hands out one contiguous range of (tile, k-block) units per CTA.
struct Seg {
  int tile, kb0, kb1;
};
struct SegIter {
  int kb_per_tile;
  int tile, step, end;  // whole-tile mode
  int u, u1;            // stream-K mode (u1 > 0)
  __device__ __forceinline__ bool next(Seg& sg) {
    if (u1 == 0) {
      if (tile >= end) return false;
      sg.tile = tile;
      sg.kb0 = 0;
      sg.kb1 = kb_per_tile;
      tile += step;
      return true;
    }
    if (u >= u1) return false;
    sg.tile = u / kb_per_tile;
    sg.kb0 = u - sg.tile * kb_per_tile;
    const int len = min(kb_per_tile - sg.kb0, u1 - u);
    sg.kb1 = sg.kb0 + len;
    u += len;
    return true;
  }
};
__device__ __forceinline__ SegIter make_seg_iter(const GemmParams& p, int num_tiles, int first, int step) {
  SegIter it;
  it.kb_per_tile = p.num_k_blocks;
  it.tile = first;
  it.step = step;
  it.end = num_tiles;
  it.u = 0;
  it.u1 = 0;
  if (p.sk_per > 0) {
    const int total = num_tiles * p.num_k_blocks;
    it.u = min(total, static_cast<int>(blockIdx.x) * p.sk_per);
    it.u1 = min(total, it.u + p.sk_per);
    if (it.u1 == 0) it.u1 = -1, it.u = 0;  // (cannot happen: total > 0) keep stream-K mode distinguishable
  }
  return it;
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_producer(const GemmParams& p, uint8_t* smem_tiles, uint64_t* full, uint64_t* empty,
                                              int num_m, SegIter it, uint8_t* b_res = nullptr, uint64_t* b_full = nullptr) {
  using C = GemmCfg<BLOCK_N>;
  int stage = 0;
  uint32_t phase = 0;
  const int tile_begin = it.tile;
  const bool bres = p.b_resident != 0;
  const int stage_bytes = bres ? C::A_BYTES : C::STAGE_BYTES;
  if (bres) {  // the CTA's single N tile: all K blocks of the weights, once
    SegIter peek = it;
    Seg s0;
    if (peek.next(s0) && elect_one()) {
      const int n0 = tile_coord(s0.tile, num_m).n_blk * BLOCK_N;
      mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(p.b_res_bytes));
      for (int kb = 0; kb < p.num_k_blocks; ++kb) tma_load_2d(b_res + kb * C::B_BYTES, &p.tma_b, b_full, kb * BLOCK_K, n0);
    }
    __syncwarp();
  }
  Seg sg;
  while (it.next(sg)) {
    const int tile = sg.tile;
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
    for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
      // The whole warp runs this loop (warp-uniform control flow) and one elected lane issues: with a single
      // divergent lane the compiler has to wrap every UTMALDG / UTCHMMA in an ELECT + BRA.U.ANY loop and shuttle
      // operands through R2UR, which made the issue loop ~430 cycles per k-block (measured, profiles/).
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem_tiles + stage * stage_bytes;
      uint8_t* sb = sa + C::A_BYTES;
      if (p.debug_mode == 1 && (phase != 0 || tile != tile_begin)) {  // measurement only: reuse stale smem
        if (elect_one()) mbar_arrive(&full[stage]);
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
        continue;
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(stage_bytes));
        if (kb >= p.kb_split) {
          tma_load_2d(sa, &p.tma_a2, &full[stage], (kb - p.kb_split) * BLOCK_K, m0);
        } else if (p.a_mode == A_CONV3X3) {
          const int tap = kb / p.conv_cblocks;
          const int cb = kb - tap * p.conv_cblocks;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &p.tma_a, &full[stage], cb * BLOCK_K, p.conv_stride * cx0 + dx, p.conv_stride * cy0 + dy, cb0);
        } else {
          tma_load_2d(sa, &p.tma_a, &full[stage], kb * BLOCK_K, m0);
        }
        if (!bres) tma_load_2d(sb, &p.tma_b, &full[stage], (kb % p.b_kb_wrap) * BLOCK_K, n0);
      }
      __syncwarp();
      if (++stage == p.num_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

template <int BLOCK_N>
__device__ __forceinline__ void gemm_mma(const GemmParams& p, uint8_t* smem_tiles, uint64_t* full, uint64_t* empty,
                                         uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                         SegIter it, uint8_t* b_res = nullptr, uint64_t* b_full = nullptr) {
  using C = GemmCfg<BLOCK_N>;
  const bool bres = p.b_resident != 0;
  const uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N, p.in_dtype == DT_BF16, false);
  // Measured (GILLB200_GEMM_DEBUG=1, no TMA): a k-block took ~300 cycles of issue-side preparation PLUS 3 x N/2
  // cycles -- tcgen05.mma issue blocks while the previous MMA occupies the pipe, so everything between the last MMA of
  // one k-block and the first of the next is exposed. Hence: smem descriptors are advanced incrementally (their
  // start-address field counts 16-byte units, a stage is STAGE_BYTES/16 further), and the NEXT stage's full barrier is
  // probed (non-blocking test_wait) before this stage's MMAs are issued, so its latency hides under them.
  const uint32_t s0 = smem_u32(smem_tiles);
  const uint64_t da0 = make_smem_desc_sw128(s0, 16, 1024);
  const uint64_t db0 = make_smem_desc_sw128(s0 + C::A_BYTES, 16, 1024);
  const uint64_t DESC_STEP = static_cast<uint64_t>(bres ? C::A_BYTES : C::STAGE_BYTES) >> 4;
  const uint64_t db_res0 = bres ? make_smem_desc_sw128(smem_u32(b_res), 16, 1024) : 0;
  uint64_t da = da0, db = db0;
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  bool ready = false;  // full[stage] of the current phase already seen complete by the look-ahead probe
  bool b_ready = !bres;
  // debug_mode 3 (measurement aid): CTA 0 adds up the clocks its MMA warp and first epilogue warp spend in each wait and
  // leaves them in the stream-K scratch (tools/gpu_gemm_trace.py)
  const bool trace = GEMM_TRACE && p.debug_mode == 3 && blockIdx.x == 0 && p.sk_ws != nullptr;
  long long tr_empty = 0, tr_full = 0, tr_tiles = 0, tr_t0 = 0;
  const long long tr_begin = trace ? clock64() : 0;
  Seg sg;
  while (it.next(sg)) {
    if (!b_ready) {  // (inside the loop: a CTA without tiles never loads, so it must never wait)
      mbar_wait(b_full, 0);
      b_ready = true;
    }
    if (trace) tr_t0 = clock64();
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
    if (trace) tr_empty += clock64() - tr_t0, ++tr_tiles;
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
    for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
      if (trace) tr_t0 = clock64();
      if (!ready) mbar_wait(&full[stage], phase);
      if (trace) tr_full += clock64() - tr_t0;
      tc_fence_after();
      const bool wrap = stage + 1 == p.num_stages;
      const int nstage = wrap ? 0 : stage + 1;
      const uint32_t nphase = wrap ? phase ^ 1 : phase;
      ready = mbar_test_wait(&full[nstage], nphase);
      if (elect_one()) {
        if (p.debug_mode != 2) {
          const uint64_t dbk = bres ? db_res0 + static_cast<uint64_t>(kb) * (C::B_BYTES >> 4) : db;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in 16-B units
            umma_f16(tmem_d, da + 2 * k, dbk + 2 * k, idesc, ((kb - sg.kb0) | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
      }
      __syncwarp();
      da = wrap ? da0 : da + DESC_STEP;
      db = wrap ? db0 : db + DESC_STEP;
      stage = nstage;
      phase = nphase;
    }
    if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
    __syncwarp();
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
  if (trace && lane_id() == 0) {
    long long* d = reinterpret_cast<long long*>(p.sk_ws);
    d[0] = tr_empty, d[1] = tr_full, d[2] = clock64() - tr_begin, d[3] = tr_tiles;
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

// alpha / bias / per-sample row bias / activation on 16 accumulator columns [n, n+16) of one row (in place).
// Returns the number of output values left in v[0..cnt): 8 for GEGLU (value * gelu(gate) pairs), else 16.
__device__ __forceinline__ int epi_math16(const GemmParams& p, int row, int n, float* v) {
  const bool geglu = p.act == ACT_GEGLU;
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
  if (geglu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = v[2 * j], g = v[2 * j + 1];
      v[j] = a * gelu_erf(g);
    }
    return 8;
  }
  if (p.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_apply(v[j], p.act);
  }
  return 16;
}

// Direct epilogue: 16 accumulator columns [n, n+16) of one row -> math -> (+ residual) -> global stores by the
// owning lane. Handles every layout (odd strides, out_lo, mixed residual dtype); each lane touches its own row, so a
// warp store instruction spans 32 cache lines -- used only where the smem-staged TMA epilogue cannot be.
__device__ __forceinline__ void epilogue_chunk16(const GemmParams& p, int row, int n, float* v) {
  const bool geglu = p.act == ACT_GEGLU;
  const int n_out_total = geglu ? p.N / 2 : p.N;
  const int cnt = epi_math16(p, row, n, v);
  const int no = geglu ? n / 2 : n;
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

// Smem-staged TMA epilogue -----------------------------------------------------------------------------------------
//
// Why: with the direct epilogue every lane owns one output row, so each 16-byte store/load instruction of a warp
// touches 32 different cache lines = 32 L1TEX wavefronts; a 128x256 fp16 tile costs >4000 LSU cycles, more than the
// 2560 tensor cycles of a K=320 tile (measured: the UNet's short-K linears ran at 250-600 TFLOP/s). Here the warp
// writes its 32x32 panel into swizzled shared memory (conflict-free 16-byte accesses) and a single lane issues one TMA
// store; residual panels come in the same way, prefetched one panel ahead, and are overwritten in place.
constexpr int EPI_PANEL_COLS = 32;
constexpr int EPI_MAX_NBUF = 4;

// Compile-time specialisations of the staged epilogue for the shapes that dominate the UNet (fp16 output, N % 32 == 0,
// alpha == 1, column bias present); EV_GENERIC keeps every option a runtime branch (fp32 / bf16 outputs, other
// activations, bias along M, N tails). ncu on the first, all-runtime version: ~400 warp instructions per 32x32 panel,
// 44 % issue-slot utilisation and the epilogue warps latency-bound -- the specialised bodies are ~100.
enum { EV_BIAS = 0, EV_BIAS_RES = 1, EV_BIAS_ROWBIAS = 2, EV_GEGLU = 3, EV_GENERIC = 4, EV_LN_BIAS = 5, EV_LN_GEGLU = 6 };

// 16-byte unit `u` of row `r` inside a 32-row panel buffer written/read by TMA with SWIZZLE_64B (16-bit elements,
// 64-byte rows) or SWIZZLE_128B (fp32, 128-byte rows)
__device__ __forceinline__ uint8_t* epi_unit(uint8_t* buf, int r, int u, bool f32) {
  return f32 ? buf + r * 128 + ((u ^ (r & 7)) << 4) : buf + r * 64 + ((u ^ ((r >> 1) & 3)) << 4);
}

// v[0..cnt) -> (+ residual already sitting in the buffer) -> buffer, at output column `oc` of the panel
__device__ __forceinline__ void epi_to_smem(const GemmParams& p, uint8_t* buf, int r, int oc, const float* v, int cnt,
                                            bool has_res, uint8_t* buf_lo = nullptr) {
  if (p.out_dtype == DT_F32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 4 < cnt) {
        float4* q = reinterpret_cast<float4*>(epi_unit(buf, r, (oc >> 2) + i, true));
        float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (has_res) {
          const float4 t = *q;
          o.x += t.x;
          o.y += t.y;
          o.z += t.z;
          o.w += t.w;
        }
        *q = o;
      }
    }
  } else {
    const bool bf16 = p.out_dtype == DT_BF16;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i * 8 < cnt) {
        uint4* q = reinterpret_cast<uint4*>(epi_unit(buf, r, (oc >> 3) + i, false));
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = v[8 * i + j];
        if (has_res) {
          const uint4 t = *q;
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 g = bf16 ? unpack_bf16x2(w[j]) : unpack_f16x2(w[j]);
            f[2 * j] += g.x;
            f[2 * j + 1] += g.y;
          }
        }
        uint4 o;
        if (bf16) {
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          if (buf_lo) {  // bf16 residue of the rounded values: the split-precision operand of the next GEMM
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
            uint32_t lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 h = unpack_bf16x2(ow[j]);
              lw[j] = pack_bf16x2(f[2 * j] - h.x, f[2 * j + 1] - h.y);
            }
            *reinterpret_cast<uint4*>(epi_unit(buf_lo, r, (oc >> 3) + i, false)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        } else {
          o.x = pack_f16x2(f[0], f[1]);
          o.y = pack_f16x2(f[2], f[3]);
          o.z = pack_f16x2(f[4], f[5]);
          o.w = pack_f16x2(f[6], f[7]);
        }
        *q = o;
      }
    }
  }
}

// v[0..8k) fp32 -> fp16 -> 16-byte units [u0, u0 + k) of row `lane` (SWIZZLE_64B panel), adding the fp16 residual that
// already sits there when RES.
template <int K8, bool RES>
__device__ __forceinline__ void epi_f16_units(uint8_t* buf, int lane, int u0, float* f) {
  uint8_t* rowp = buf + lane * 64;
  const int x = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < K8; ++i) {
    uint4* q = reinterpret_cast<uint4*>(rowp + (((u0 + i) ^ x) << 4));
    if (RES) {
      const uint4 t = *q;
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = unpack_f16x2(w[j]);
        f[8 * i + 2 * j] += g.x;
        f[8 * i + 2 * j + 1] += g.y;
      }
    }
    *q = make_uint4(pack_f16x2(f[8 * i], f[8 * i + 1]), pack_f16x2(f[8 * i + 2], f[8 * i + 3]),
                    pack_f16x2(f[8 * i + 4], f[8 * i + 5]), pack_f16x2(f[8 * i + 6], f[8 * i + 7]));
  }
}

struct GemmSmemBars {
  uint64_t b_full;  // B-stationary mode: the resident weight tile has landed
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[GEMM_EPI_WARPS][EPI_MAX_NBUF];
  uint64_t halo_full[3], halo_empty[3];  // A_CONV3X3_HALO: the halo-tile slots (two; three with the GroupNorm transform)
  uint64_t halo_ready[3];                // ... with the GroupNorm transform: both CTAs' transform warps are done with the slot
  uint32_t tmem_ptr;
};
static_assert(sizeof(GemmSmemBars) <= 1024, "barrier block");

// The staged epilogue of one warp over all of its tiles.
//   tile_fn(tile, &row_base, &n0): first accumulator row of this CTA's 128-row block and first column of the tile.
//   release_fn(acc): arrive on the tile's tmem_empty barrier (local, or the pair leader's for cta_group::2).
// Per panel: [lane 0: make sure the staging buffer is free, prefetch the next residual panel] -> bias loads in flight
// -> tcgen05.ld -> math -> swizzled st.shared -> fence.proxy.async -> [lane 0: TMA store].
template <int BLOCK_N, int ACC_STRIDE, int VAR, int NACC, typename Iter, typename TileFn, typename Release>
__device__ __forceinline__ void epilogue_warp_tma(const GemmParams& p, GemmSmemBars* bars, uint8_t* epi_stage,
                                                  uint32_t tmem_base, Iter it, TileFn tile_fn, Release release_fn) {
  const int warp = threadIdx.x >> 5;
  const int ew = warp - 4, quad = warp & 3;
  const int cgrp = ew >> 2, ngrp = p.epi_warps >> 2;
  const int lane = lane_id();
  constexpr bool GEGLU = VAR == EV_GEGLU || VAR == EV_LN_GEGLU;
  constexpr bool LNF = VAR == EV_LN_BIAS || VAR == EV_LN_GEGLU;
  const bool geglu = VAR == EV_GENERIC ? p.act == ACT_GEGLU : GEGLU;
  const bool has_res = VAR == EV_GENERIC ? p.residual != nullptr : VAR == EV_BIAS_RES;
  const int acc_per_panel = geglu ? 2 * EPI_PANEL_COLS : EPI_PANEL_COLS;
  const int n_out_total = geglu ? p.N / 2 : p.N;
  const uint32_t nbuf = static_cast<uint32_t>(p.epi_nbuf);
  const uint32_t panel_bytes = static_cast<uint32_t>(p.epi_buf_bytes);
  uint8_t* stage = epi_stage + ew * nbuf * panel_bytes;
  uint64_t* res_bar = bars->res_full[ew];
  // staging-buffer rotation: `buf` = buffer of the current panel, `par` bit b = parity of res_bar[b]'s next completion
  uint32_t buf = 0, par = 0;
  int acc = 0;
  uint32_t acc_phase = 0;

  const bool trace = GEMM_TRACE && p.debug_mode == 3 && blockIdx.x == 0 && ew == 0 && p.sk_ws != nullptr;
  long long tr_acc = 0, tr_store = 0, tr_ld = 0, tr_res = 0, tr_tiles = 0, tr_t0 = 0, tr_fence = 0, tr_issue = 0;
  const long long tr_begin = trace ? clock64() : 0;
  Seg sg;
  while (it.next(sg)) {
    const int tile = sg.tile;
    int row_base, n0, ncols = BLOCK_N;  // ncols: accumulator columns this tile really holds (a split tail tile: fewer)
    tile_fn(tile, &row_base, &n0, &ncols);
    const int np = (ncols + acc_per_panel - 1) / acc_per_panel;
    const int pb0 = cgrp * (np / ngrp) + min(cgrp, np % ngrp);
    const int pe0 = pb0 + np / ngrp + (cgrp < np % ngrp ? 1 : 0);
    const int row0 = row_base + quad * 32;
    // stream-K: a tile cut by a CTA boundary. The owner of k-block 0 finishes it; everyone else parks partials.
    const bool sk_partial = sg.kb0 > 0;
    const bool sk_reduce = sg.kb0 == 0 && sg.kb1 < it.kb_per_tile;
    int sk_first = 0, sk_last = -1;  // CTAs whose partials this (finishing) CTA adds
    if (sk_reduce) {
      sk_first = static_cast<int>(blockIdx.x) + 1;
      sk_last = ((tile + 1) * it.kb_per_tile - 1) / p.sk_per;
    }
    const int no0 = geglu ? n0 >> 1 : n0;
    const uint32_t taddr = tmem_base + acc * ACC_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
    // panels entirely beyond N (last N tile) or a slab entirely beyond M produce nothing
    int pe = min(pe0, (n_out_total - no0 + EPI_PANEL_COLS - 1) / EPI_PANEL_COLS);
    if (row0 >= p.M_out) pe = pb0;
    const int row = min(row0 + lane, p.M_out - 1);  // rows >= M are clipped by the TMA store; clamp only for bias reads
    const float* rb_row = nullptr;
    if (VAR == EV_BIAS_ROWBIAS) rb_row = p.rowbias + static_cast<long long>(row / p.rows_per_group) * p.ld_rowbias;
    float ln_rstd = 1.f, ln_rm = 0.f;  // folded LayerNorm of this lane's row: rstd and rstd * mean
    if (LNF) {
      float ls = 0.f, lq = 0.f;
      for (int i = 0; i < p.ln_np; ++i) {
        const float2 t = __ldg(p.ln_stats + static_cast<size_t>(i) * p.ln_ld + row);
        ls += t.x;
        lq += t.y;
      }
      const float mean = ls / p.ln_C;
      ln_rstd = rsqrtf(fmaxf(lq / p.ln_C - mean * mean, 0.f) + p.ln_eps);
      ln_rm = ln_rstd * mean;
    }

    // A_CONV3X3_HALO: rows are numbered block by block (sample, 16 x 8 pixel block, y-major inside the block), so this
    // warp's 32 rows are a 4 (y) x 8 (x) patch: panels travel through 3-D {C, W, B * H} tensor maps with {32, 8, 4} boxes
    const bool halo = p.a_mode == A_CONV3X3_HALO;
    int hx = 0, hy = 0;
    if (halo) {
      const int hw = p.conv_W * p.conv_H;
      const int hb = row0 / hw, hr = row0 - hb * hw;
      const int blk = hr >> 7, bxn = p.conv_W >> 3;
      hx = (blk % bxn) * 8;
      hy = hb * p.conv_H + (blk / bxn) * 16 + ((hr & 127) >> 5) * 4;
    }
    auto fetch_res = [&](int pnl, uint32_t b) {  // lane 0 only
      mbar_arrive_expect_tx(&res_bar[b], panel_bytes);
      if (halo) tma_load_3d(stage + b * panel_bytes, &p.tma_res, &res_bar[b], no0 + pnl * EPI_PANEL_COLS, hx, hy);
      else tma_load_2d(stage + b * panel_bytes, &p.tma_res, &res_bar[b], no0 + pnl * EPI_PANEL_COLS, row0);
    };
    // with 4 staging buffers the residual runs TWO panels ahead (the wait-time trace of the N = 320 out-projections showed
    // ~800-1000 clocks per tile spent waiting for residual panels fetched only one panel ahead)
    const bool res2 = nbuf >= 4;
    if (has_res && pb0 < pe && !sk_partial && lane == 0) {  // first residual panel(s) travel while the MMAs finish
      tma_store_wait_read<1>();
      fetch_res(pb0, buf);
      if (res2 && pb0 + 1 < pe) fetch_res(pb0 + 1, buf + 1 == nbuf ? 0 : buf + 1);
    }
    if (trace) tr_t0 = clock64();
    mbar_wait(&bars->tmem_full[acc], acc_phase);
    if (trace) tr_acc += clock64() - tr_t0, ++tr_tiles;
    tc_fence_after();
    if (pb0 >= pe) release_fn(acc);

    if (sk_partial) {
      // park this warp's share of the fp32 accumulators, then raise this warp's flag (the finisher's warp with the same
      // index consumes it). Scratch layout (private to the two warps, so free to choose): sk_ws[cta][quad][32-column
      // chunk][float4 i = 0..7][lane] -- every warp store / load is 512 contiguous bytes. (Round 1 used row-major
      // [row][col]: one 128-byte line PER LANE and instruction = 32 L1 wavefronts per access; the finisher of a tile cut 4
      // ways read 524 KB that way on the critical path of the launch.)
      float* ws = p.sk_ws + (static_cast<size_t>(blockIdx.x) * BLOCK_M + quad * 32) * BLOCK_N + lane * 4;
      for (int pnl = pb0; pnl < pe; ++pnl) {
        const int halves = geglu ? 2 : 1;
#pragma unroll 1
        for (int h = 0; h < halves; ++h) {
          const int acol = pnl * acc_per_panel + h * 32;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + acol, r);
          tmem_wait_ld();
          if (pnl == pe - 1 && h == halves - 1) release_fn(acc);
          float4* dst = reinterpret_cast<float4*>(ws + acol * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            __stcg(dst + i * 32, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                        __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
        }
      }
      __threadfence();
      __syncwarp();
      // raise the flag only when the finisher's same-index warp will consume (and re-arm) it: a warp whose rows lie
      // beyond M or whose panels lie beyond N has pb0 >= pe on both sides and must leave the flag alone
      if (lane == 0 && pb0 < pe) {
        int* flag = p.sk_flags + static_cast<size_t>(blockIdx.x) * GEMM_EPI_WARPS + ew;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
      }
      if (++acc == NACC) {
        acc = 0;
        acc_phase ^= 1;
      }
      continue;
    }
    if (sk_reduce && pb0 < pe) {
      // wait (once per tile) until the same-index warp of every contributing CTA has parked its partials
      if (lane == 0) {
        for (int c = sk_first; c <= sk_last; ++c) {
          const int* flag = p.sk_flags + static_cast<size_t>(c) * GEMM_EPI_WARPS + ew;
          int v;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          } while (v == 0);
        }
      }
      __syncwarp();
    }
    // adds the parked partials of accumulator columns [acol, acol + 32) of this lane's row, in CTA order
    auto sk_add = [&](uint32_t (&r)[32], int acol) {
      for (int c = sk_first; c <= sk_last; ++c) {
        const float4* src = reinterpret_cast<const float4*>(
            p.sk_ws + (static_cast<size_t>(c) * BLOCK_M + quad * 32) * BLOCK_N + acol * 32 + lane * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldcg(src + i * 32);
          r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + t.x);
          r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + t.y);
          r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + t.z);
          r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + t.w);
        }
      }
    };

    // Compile-time variants: the NEXT 32-column accumulator chunk is already travelling TMEM -> registers while this one is
    // converted and staged (a panel used to be one serial chain TMEM load -> math -> st.shared -> TMA store, ~700 clocks
    // with two epilogue warps per scheduler to hide it: the K = 320 linears spent as long draining a tile as filling it).
    constexpr bool PIPE = VAR != EV_GENERIC && EPI_PIPELINE;
    uint32_t rn[32];
    if (PIPE && pb0 < pe) tmem_ld_32x32b_x32(taddr + pb0 * acc_per_panel, rn);

    for (int pnl = pb0; pnl < pe; ++pnl) {
      uint8_t* sbuf = stage + buf * panel_bytes;
      const uint32_t nxt = buf + 1 == nbuf ? 0 : buf + 1;
      if (trace) tr_t0 = clock64();
      if (lane == 0) {
        tma_store_wait_read<1>();  // the store issued two panels ago has drained its staging buffer
        if (has_res && !res2 && pnl + 1 < pe) fetch_res(pnl + 1, nxt);
        if (has_res && res2 && pnl + 2 < pe) fetch_res(pnl + 2, nxt + 1 == nbuf ? 0 : nxt + 1);
      }
      __syncwarp();
      if (trace) tr_store += clock64() - tr_t0;
      const bool last = pnl == pe - 1;
      const int nacc = n0 + pnl * acc_per_panel;  // first accumulator column (global) of this panel

      uint8_t* sbuf_lo = nullptr;
      if (VAR == EV_GENERIC && p.out_lo != nullptr) {  // hi panel in buffer 0, lo panel in buffer 1, both stored per panel
        sbuf = stage;
        sbuf_lo = stage + panel_bytes;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      if (VAR == EV_GENERIC) {
        if (has_res) mbar_wait(&res_bar[buf], (par >> buf) & 1);
        const int halves = geglu ? 2 : 1;
#pragma unroll 1
        for (int h = 0; h < halves; ++h) {
          const int acol = pnl * acc_per_panel + h * 32;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + acol, r);
          tmem_wait_ld();
          if (last && h == halves - 1) release_fn(acc);
          if (sk_reduce) sk_add(r, acol);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[16 * c + j]);
            const int cnt = epi_math16(p, row, n0 + acol + 16 * c, v);
            epi_to_smem(p, sbuf, lane, geglu ? h * 16 + 8 * c : 16 * c, v, cnt, has_res, sbuf_lo);
          }
        }
      } else if (GEGLU) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + nacc + h * 32);
          float4 bv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
          uint32_t r[32];
          if (PIPE) {
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = rn[i];
            if (h == 0) tmem_ld_32x32b_x32(taddr + pnl * 64 + 32, rn);
            else if (!last) tmem_ld_32x32b_x32(taddr + (pnl + 1) * 64, rn);
          } else {
            tmem_ld_32x32b_x32(taddr + pnl * 64 + h * 32, r);
            tmem_wait_ld();
          }
          if (last && h == 1) release_fn(acc);
          if (sk_reduce) sk_add(r, pnl * 64 + h * 32);
          if (LNF) {  // x W'^T -> rstd * (x W'^T - mean * colsum(W')) before the bias
            const float4* c4 = reinterpret_cast<const float4*>(p.ln_cs + nacc + h * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 cs = __ldg(c4 + i);
              r[4 * i] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i]), -ln_rm * cs.x));
              r[4 * i + 1] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 1]), -ln_rm * cs.y));
              r[4 * i + 2] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 2]), -ln_rm * cs.z));
              r[4 * i + 3] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 3]), -ln_rm * cs.w));
            }
          }
          float f[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // (value, gate) pairs are interleaved along N; two outputs per packed op
            const f32x2 val = add2(pk2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 2])), pk2(bv[i].x, bv[i].z));
            const f32x2 gate = add2(pk2(__uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 3])), pk2(bv[i].y, bv[i].w));
            upk2(mul2(val, gelu_poly2(gate)), f[2 * i], f[2 * i + 1]);
          }
          epi_f16_units<2, false>(sbuf, lane, 2 * h, f);
        }
      } else {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + nacc);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        if (VAR == EV_BIAS_ROWBIAS) {
          const float4* r4 = reinterpret_cast<const float4*>(rb_row + nacc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(r4 + i);
            bv[i].x += t.x;
            bv[i].y += t.y;
            bv[i].z += t.z;
            bv[i].w += t.w;
          }
        }
        uint32_t r[32];
        if (PIPE) {
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = rn[i];
          if (!last) tmem_ld_32x32b_x32(taddr + (pnl + 1) * 32, rn);
        } else {
          if (trace) tr_t0 = clock64();
          tmem_ld_32x32b_x32(taddr + pnl * 32, r);
          tmem_wait_ld();
          if (trace) tr_ld += clock64() - tr_t0;
        }
        if (last) release_fn(acc);
        if (sk_reduce) sk_add(r, pnl * 32);
        if (LNF) {
          const float4* c4 = reinterpret_cast<const float4*>(p.ln_cs + nacc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 cs = __ldg(c4 + i);
            r[4 * i] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i]), -ln_rm * cs.x));
            r[4 * i + 1] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 1]), -ln_rm * cs.y));
            r[4 * i + 2] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 2]), -ln_rm * cs.z));
            r[4 * i + 3] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[4 * i + 3]), -ln_rm * cs.w));
          }
        }
        float f[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[4 * i] = __uint_as_float(r[4 * i]) + bv[i].x;
          f[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv[i].y;
          f[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv[i].z;
          f[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv[i].w;
        }
        if (trace) tr_t0 = clock64();
        if (VAR == EV_BIAS_RES) mbar_wait(&res_bar[buf], (par >> buf) & 1);
        if (trace) tr_res += clock64() - tr_t0;
        epi_f16_units<4, VAR == EV_BIAS_RES>(sbuf, lane, 0, f);
        if ((VAR == EV_BIAS || VAR == EV_BIAS_RES) && p.rowstats_out != nullptr && row0 + lane < p.M_out) {
          // f[] now holds this row's 32 final values (residual included): partial LayerNorm sums for the next GEMM
          float rs = 0.f, rq = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rs += f[i];
            rq = fmaf(f[i], f[i], rq);
          }
          p.rowstats_out[static_cast<size_t>((no0 + pnl * EPI_PANEL_COLS) >> 5) * p.rowstats_ld + row0 + lane] =
              make_float2(rs, rq);
        }
      }

      if (p.stats_out != nullptr) {
        // lane = column of this panel: read the 32 rounded values of the column back from the staged panel
        __syncwarp();
        const int su = lane >> 3, se = (lane & 7) * 2;
        const bool sbf = p.out_dtype == DT_BF16;
        float cs = 0.f, cq = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const uint16_t hv = *reinterpret_cast<const uint16_t*>(sbuf + rr * 64 + ((su ^ ((rr >> 1) & 3)) << 4) + se);
          const float xv = sbf ? __uint_as_float(static_cast<uint32_t>(hv) << 16) : __half2float(__ushort_as_half(hv));
          cs += xv;
          cq = fmaf(xv, xv, cq);
        }
        size_t slab = static_cast<size_t>(row0 >> 5);
        if (p.stats_hw) {
          const int sb = row0 / p.stats_hw;
          slab = (static_cast<size_t>(sb * 4 + p.stats_phase) * p.stats_hw + (row0 - sb * p.stats_hw)) >> 5;
        }
        p.stats_out[slab * n_out_total + (no0 + pnl * EPI_PANEL_COLS + lane)] = make_float2(cs, cq);
      }
      if (trace) tr_t0 = clock64();
      if (VAR != EV_GENERIC && p.epi_stg && !halo) {
        // 16-bit panel (32 rows x 64 B, 64-byte swizzle) written back by ALL lanes: lane L takes 16-byte units L, L + 32,
        // L + 64, L + 96 (unit u = row u / 4, quarter u % 4), so every warp store covers 8 rows x 64 contiguous bytes. No
        // proxy fence, no TMA issue (measured ~190 clocks per panel on the issuing lane, profiles/r02_gemm_wait_trace.log),
        // no drain wait before the buffer is reused.
        __syncwarp();
        uint16_t* o16 = static_cast<uint16_t*>(p.out);
        const int pcol = no0 + pnl * EPI_PANEL_COLS;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = i * 32 + lane, rr = u >> 2, su = u & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(sbuf + rr * 64 + ((su ^ ((rr >> 1) & 3)) << 4));
          if (row0 + rr < p.M_out && pcol + su * 8 < n_out_total)
            *reinterpret_cast<uint4*>(o16 + static_cast<long long>(row0 + rr) * p.ldo + pcol + su * 8) = v;
        }
        __syncwarp();
      } else {
        fence_proxy_async_smem();
        __syncwarp();
        if (trace) tr_fence += clock64() - tr_t0, tr_t0 = clock64();
        if (lane == 0) {
          if (halo) tma_store_3d(&p.tma_out, sbuf, no0 + pnl * EPI_PANEL_COLS, hx, hy);
          else tma_store_2d(&p.tma_out, sbuf, no0 + pnl * EPI_PANEL_COLS, row0);
          if (sbuf_lo) tma_store_2d(&p.tma_out_lo, sbuf_lo, no0 + pnl * EPI_PANEL_COLS, row0);
          tma_store_commit();
        }
      }
      if (trace) tr_issue += clock64() - tr_t0;
      par ^= 1u << buf;
      buf = nxt;
    }
    if (sk_reduce && pb0 < pe) {  // consumed: re-arm the flags for the next launch (this warp is their only reader)
      __syncwarp();
      if (lane == 0)
        for (int c = sk_first; c <= sk_last; ++c) p.sk_flags[static_cast<size_t>(c) * GEMM_EPI_WARPS + ew] = 0;
    }
    if (++acc == NACC) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
  if (lane == 0) tma_store_wait_all<0>();  // smem must outlive the bulk reads
  if (trace && lane == 0) {
    long long* d = reinterpret_cast<long long*>(p.sk_ws) + 8;
    d[0] = tr_acc, d[1] = tr_store, d[2] = tr_ld, d[3] = tr_res, d[4] = clock64() - tr_begin, d[5] = tr_tiles, d[6] = tr_fence, d[7] = tr_issue;
  }
}

// Runtime -> compile-time variant dispatch (once per warp, outside the tile loop).
template <int BLOCK_N, int ACC_STRIDE, int NACC = 2, typename Iter, typename TileFn, typename Release>
__device__ __forceinline__ void epilogue_warp_tma_dispatch(const GemmParams& p, GemmSmemBars* bars, uint8_t* epi_stage,
                                                           uint32_t tmem_base, Iter it, TileFn tile_fn,
                                                           Release release_fn) {
  switch (p.epi_variant) {
    case EV_BIAS:
      epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_BIAS, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    case EV_BIAS_RES:
      epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_BIAS_RES, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    case EV_BIAS_ROWBIAS:
      epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_BIAS_ROWBIAS, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    case EV_LN_BIAS:
      epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_LN_BIAS, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    case EV_LN_GEGLU:
      if constexpr (BLOCK_N % 64 == 0)
        epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_LN_GEGLU, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    case EV_GEGLU:
      if constexpr (BLOCK_N % 64 == 0)
        epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_GEGLU, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
      break;
    default:
      epilogue_warp_tma<BLOCK_N, ACC_STRIDE, EV_GENERIC, NACC>(p, bars, epi_stage, tmem_base, it, tile_fn, release_fn);
  }
}

// Kernel -------------------------------------------------------------------------------------------------------


template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_res = smem;  // B-stationary mode: [resident weight tile][A ring]; else the ring starts at `smem`
  uint8_t* smem_tiles = smem + (p.b_resident ? p.b_res_bytes : 0);
  GemmSmemBars* bars = reinterpret_cast<GemmSmemBars*>(smem_tiles + p.num_stages * (p.b_resident ? C::A_BYTES : C::STAGE_BYTES));
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    if (p.kb_split < p.num_k_blocks) tma_prefetch_desc(&p.tma_a2);
    mbar_init(&bars->b_full, 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], p.epi_tma ? p.epi_warps : GEMM_EPI_WARPS);
    }
    if (p.epi_tma) {
      tma_prefetch_desc(&p.tma_out);
      if (p.residual) tma_prefetch_desc(&p.tma_res);
      for (int w = 0; w < GEMM_EPI_WARPS; ++w)
        for (int i = 0; i < EPI_MAX_NBUF; ++i) mbar_init(&bars->res_full[w][i], 1);
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
  pdl_wait();    // barriers, TMEM and descriptors are set up; operands of the previous kernel are visible from here on
  pdl_launch();

  const bool n_inner = p.tile_order == 2 || (p.tile_order == 0 && num_n <= 4 && num_m >= 2 * static_cast<int>(gridDim.x));
  const int order = n_inner ? -num_n : num_m;  // see tile_coord()
  const SegIter seg_it = make_seg_iter(p, num_tiles, blockIdx.x, gridDim.x);
  if (warp == 0) {
    gemm_producer<BLOCK_N>(p, smem_tiles, bars->full, bars->empty, order, seg_it, b_res, &bars->b_full);
  } else if (warp == 1) {
    gemm_mma<BLOCK_N>(p, smem_tiles, bars->full, bars->empty, bars->tmem_full, bars->tmem_empty, tmem_base, seg_it, b_res,
                      &bars->b_full);
  } else if (warp >= 4) {
    if (!p.epi_tma) {
      gemm_epilogue<BLOCK_N>(p, bars->tmem_full, bars->tmem_empty, tmem_base, order, num_tiles);
    } else if (warp - 4 < p.epi_warps) {
      epilogue_warp_tma_dispatch<BLOCK_N, C::ACC_STRIDE>(
          p, bars, epi_stage, tmem_base, seg_it,
          [&](int tile, int* row_base, int* n0, int*) {
            const TileCoord tc = tile_coord(tile, order);
            *row_base = tc.m_blk * BLOCK_M;
            *n0 = tc.n_blk * BLOCK_N;
          },
          [&](int acc) {
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&bars->tmem_empty[acc]);
          });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace gb
