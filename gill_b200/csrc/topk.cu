// Retrieval branch: cosine-similarity scores fused with top-k selection, plus the candidate merge.
//
// Replaces gill/models.py:676-683  (scores = emb_matrix @ ret_emb.T; scores[seen] -= 1000; topk).
//
// scores[q, n] = sum_d Q[q, d] * bank[n, d] is never materialised (3 M x 1024 fp32 would be 12.3 GB). The tcgen05
// main loop of gemm_sm100.cuh runs with the QUERIES on the M side (TMEM lanes) and bank rows on the N side, so
// every epilogue thread owns one query and streams that query's scores out of TMEM 16 columns at a time, keeping a
// sorted top-16 in registers (strict '>' against the current 16th best => ties keep the lowest bank row).
// Each CTA owns one 128-query tile and one contiguous slice of bank rows; slices are merged by topk_merge_kernel,
// which is also the cross-GPU merge (R shards x K candidates per query, order: value desc, index asc).
#include "../../include/gillb200.h"
#include "gemm_sm100.cuh"
#include "host_common.h"

#include <cstdlib>
#include <cstring>

namespace gb {

constexpr int TOPK_BN = 256;
constexpr int KMAX = 16;

struct TopkEpi {
  int Q;
  long long n_local;
  long long index_base;
  const long long* exclude;  // [Q or 1, n_exclude] global row ids (< 0: unused slot)
  long long exclude_ld;      // elements between the lists of consecutive queries; 0: one list shared by every query
  int n_exclude;
  float* part_val;      // [splits, Qpad, KMAX]
  long long* part_idx;  // [splits, Qpad, KMAX]
  int q_pad;
  int tiles_per_split;
  int num_n;
};

__device__ __forceinline__ void topk_insert(float (&tv)[KMAX], int (&ti)[KMAX], float s, int idx) {
#pragma unroll
  for (int j = KMAX - 1; j > 0; --j) {
    if (s > tv[j - 1]) {
      tv[j] = tv[j - 1];
      ti[j] = ti[j - 1];
    } else if (s > tv[j]) {
      tv[j] = s;
      ti[j] = idx;
    }
  }
  if (s > tv[0]) {
    tv[0] = s;
    ti[0] = idx;
  }
}

constexpr int TOPK_THREADS = 384;

__global__ void __launch_bounds__(TOPK_THREADS, 1)
topk_scores_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ TopkEpi e) {
  using C = GemmCfg<TOPK_BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_tiles = smem;
  GemmSmemBars* bars = reinterpret_cast<GemmSmemBars*>(smem + C::STAGES * C::STAGE_BYTES);
  const int warp = threadIdx.x >> 5;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    mbar_init(&bars->b_full, 1);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 8);
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

  // this CTA: query tile m_blk, bank tiles [n_lo, n_hi)
  const int m_blk = blockIdx.x % num_m;
  const int split = blockIdx.x / num_m;
  const int n_lo = split * e.tiles_per_split;
  const int n_hi = min(n_lo + e.tiles_per_split, e.num_n);
  const int tile_begin = n_lo * num_m + m_blk, tile_end = n_hi * num_m, tile_step = num_m;

  if (warp == 0) {
    gemm_producer<TOPK_BN>(p, smem_tiles, bars->full, bars->empty, num_m,
                           SegIter{p.num_k_blocks, tile_begin, tile_step, tile_end, 0, 0});
  } else if (warp == 1) {
    gemm_mma<TOPK_BN>(p, smem_tiles, bars->full, bars->empty, bars->tmem_full, bars->tmem_empty, tmem_base,
                      SegIter{p.num_k_blocks, tile_begin, tile_step, tile_end, 0, 0});
  } else if (warp >= 4) {
    // 8 epilogue warps: warps 4..7 scan columns [0,128) of each tile, warps 8..11 scan [128,256); both groups cover
    // all 128 TMEM lanes (a warp may only touch lanes 32*(warp%4) .. +31). Each thread keeps its own top-KMAX.
    const int ewarp = warp & 3;
    const int half = (warp - 4) >> 2;
    const int qrow = m_blk * BLOCK_M + ewarp * 32 + lane_id();
    float tv[KMAX];
    int ti[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      tv[j] = -INFINITY;
      ti[j] = 0x7fffffff;
    }
    // this query's seen list (exchange buffers carry all-unused lists for most queries: skip the scan for those)
    const long long* my_ex = e.exclude + static_cast<long long>(min(qrow, e.Q - 1)) * e.exclude_ld;
    bool has_ex = false;
    for (int x = 0; x < e.n_exclude; ++x) has_ex |= my_ex[x] >= 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int n_blk = n_lo; n_blk < n_hi; ++n_blk) {
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + acc * C::ACC_STRIDE + half * (TOPK_BN / 2) + (static_cast<uint32_t>(ewarp * 32) << 16);
      const long long n0 = static_cast<long long>(n_blk) * TOPK_BN + half * (TOPK_BN / 2);
#pragma unroll 1
      for (int c = 0; c < TOPK_BN / 2; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c, r);
        tmem_wait_ld();
        float mx = __uint_as_float(r[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        if (__any_sync(0xffffffffu, mx > tv[KMAX - 1])) {
          // Slow path. Per-lane bitmask of columns above this lane's threshold, OR-reduced over the warp; then a real
          // loop over only the columns that hit, re-reading that single column from TMEM (the TMEM address is a
          // run-time operand, so no dynamic register indexing and nothing for the compiler to flatten).
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) m |= (__uint_as_float(r[j]) > tv[KMAX - 1] ? 1u : 0u) << j;
          uint32_t hits = __reduce_or_sync(0xffffffffu, m);
#pragma unroll 1
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            float s = __uint_as_float(tmem_ld_32x32b_x1(taddr + c + j));
            tmem_wait_ld();
            const long long nloc = n0 + c + j;
            if (s > tv[KMAX - 1] && nloc < e.n_local) {
              if (has_ex) {
                const long long gidx = e.index_base + nloc;
                for (int x = 0; x < e.n_exclude; ++x)
                  if (my_ex[x] == gidx) s -= 1000.0f;  // gill/models.py:679-680
              }
              topk_insert(tv, ti, s, static_cast<int>(nloc));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&bars->tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // candidates of (split, column half) for query qrow; rows >= Q carry zeros/-inf and are ignored by the merge
    const long long o = (static_cast<long long>(split * 2 + half) * e.q_pad + qrow) * KMAX;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      e.part_val[o + j] = tv[j];
      e.part_idx[o + j] = tv[j] == -INFINITY ? 0x7fffffffffffffffLL : e.index_base + ti[j];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// Merge R candidate lists per query: cand_{val,idx}[r, q, kc] -> out[q, k]; order (value desc, index asc).
// One warp per query; K rounds of "best candidate strictly after the previous winner".
__global__ void topk_merge_kernel(const float* __restrict__ cand_val, const long long* __restrict__ cand_idx, int R,
                                  long long q_stride /* elements between queries */,
                                  long long r_stride /* elements between lists (values) */,
                                  long long r_stride_i /* elements between lists (indices) */, int kc, int Q, int K,
                                  float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int lane = threadIdx.x & 31;
  float last_v = INFINITY;
  long long last_i = -1;
  const int total = R * kc;
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (int t = lane; t < total; t += 32) {
      const int r = t / kc, j = t - r * kc;
      const float v = cand_val[r * r_stride + q * q_stride + j];
      const long long i = cand_idx[r * r_stride_i + q * q_stride + j];
      const bool after = (v < last_v) || (v == last_v && i > last_i);
      const bool better = (v > bv) || (v == bv && i < bi);
      if (after && better) {
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_val[static_cast<long long>(q) * K + k] = bv;
      out_idx[static_cast<long long>(q) * K + k] = bi;
    }
    last_v = bv;
    last_i = bi;
  }
}

// Same merge with the candidates held in registers (R * kc <= 32 * MAXC): one global read of the lists instead of K. The
// looped form above re-read all R * kc candidates in each of the K rounds -- 46 us under ncu for 1024 queries x 36 lists x 16
// (profiles/r01_ncu_launch_list_summary.csv), a fixed cost that weighed on small bank shards (375 k rows per GPU at N = 8).
template <int MAXC>
__global__ void __launch_bounds__(128) topk_merge_reg_kernel(const float* __restrict__ cand_val,
                                                             const long long* __restrict__ cand_idx, int R,
                                                             long long q_stride, long long r_stride, long long r_stride_i,
                                                             int kc, int Q, int K, float* __restrict__ out_val,
                                                             long long* __restrict__ out_idx) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int lane = threadIdx.x & 31;
  const int total = R * kc;
  constexpr long long EMPTY = 0x7fffffffffffffffLL;
  float v[MAXC];
  long long id[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int t = lane + 32 * c;
    v[c] = -INFINITY;
    id[c] = EMPTY;
    if (t < total) {
      const int r = t / kc, j = t - r * kc;
      v[c] = cand_val[r * r_stride + q * q_stride + j];
      id[c] = cand_idx[r * r_stride_i + q * q_stride + j];
    }
  }
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    long long bi = EMPTY;
    int bc = -1;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (v[c] > bv || (v[c] == bv && id[c] < bi)) {
        bv = v[c];
        bi = id[c];
        bc = c;
      }
    }
    float gv = bv;
    long long gi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, gv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, gi, o);
      if (ov > gv || (ov == gv && oi < gi)) {
        gv = ov;
        gi = oi;
      }
    }
    if (lane == 0) {
      out_val[static_cast<long long>(q) * K + k] = gv;
      out_idx[static_cast<long long>(q) * K + k] = gi;
    }
    // retire every copy of the winner (the looped form's "strictly after the previous winner" emits a pair once)
    if (bc >= 0 && gv == bv && gi == bi) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (v[c] == gv && id[c] == gi) {
          v[c] = -INFINITY;
          id[c] = EMPTY;
        }
    }
  }
}

static cudaError_t launch_topk_merge(const float* cand_val, const long long* cand_idx, int R, long long q_stride,
                                     long long r_stride, long long r_stride_i, int kc, int Q, int K, float* out_val,
                                     long long* out_idx, cudaStream_t stream) {
  const int wpb = 4;
  const dim3 grid((Q + wpb - 1) / wpb), block(wpb * 32);
  const int total = R * kc;
  if (total <= 32 * 8)
    topk_merge_reg_kernel<8><<<grid, block, 0, stream>>>(cand_val, cand_idx, R, q_stride, r_stride, r_stride_i, kc, Q, K, out_val, out_idx);
  else if (total <= 32 * 24)
    topk_merge_reg_kernel<24><<<grid, block, 0, stream>>>(cand_val, cand_idx, R, q_stride, r_stride, r_stride_i, kc, Q, K, out_val, out_idx);
  else
    topk_merge_kernel<<<grid, block, 0, stream>>>(cand_val, cand_idx, R, q_stride, r_stride, r_stride_i, kc, Q, K, out_val, out_idx);
  return cudaGetLastError();
}


// ================================================================================================================
// Small query batches (Q <= 4; the reference's own call shape is Q = 1, K = 3, gill/models.py:676-683).
//
// One pass over the bank, HBM-bound: N * D * 2 bytes against 2 * N * D * Q flops, so the 128-row tensor-core tile would
// be 97-99 % padding and its per-tile hand-offs cap it at ~1/3 of the HBM rate (profiles/r01_topk_check.log). Here the
// bank is streamed by plain coalesced 16-byte loads: a warp owns 8 consecutive rows per step, a lane owns 8 columns of
// every 256-column chunk (512 contiguous bytes per row and warp-load), the queries sit in shared memory as fp32, and the
// 8 x Q per-lane partial dot products are reduced with a transposing butterfly (8Q - 1 + log2(32 / 8Q) shuffles instead
// of 5 per value). bf16 x bf16 products are exact in fp32; sums are fp32 like the tensor-core path.
// Each warp keeps one sorted top-K list per query spread over its lanes (lane j = j-th best); rows arrive in increasing
// order and insertion is strict '>', so ties keep the lowest row. Warp lists are merged per CTA in shared memory, CTA
// lists by topk_merge_kernel (value desc, index asc).
struct StreamParams {
  const uint16_t* bank;
  long long ldb, n_local;
  const uint16_t* q;
  long long ldq;
  int D, Q, K;
  long long index_base;
  const long long* exclude;
  long long exclude_ld;
  int n_exclude;
  int rows_per_cta;  // multiple of 8
  float* part_val;      // [grid, QT, KMAX]
  long long* part_idx;  // [grid, QT, KMAX]
};

constexpr int STREAM_THREADS = 512;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int QT, int NCH>
__global__ void __launch_bounds__(STREAM_THREADS, 1) topk_stream_kernel(const StreamParams p) {
  constexpr int NV = 8 * QT;         // values reduced per step: 8 rows x QT queries
  constexpr int LPV = 32 / NV;       // lanes that end up holding the same value
  constexpr int DP = NCH * 256;      // padded query length in shared memory
  __shared__ __align__(16) float q_s[QT * DP];
  __shared__ float m_val[STREAM_WARPS][QT][KMAX];
  __shared__ int m_idx[STREAM_WARPS][QT][KMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < QT * DP; i += STREAM_THREADS) {
    const int qq = i / DP, d = i - qq * DP;
    float v = 0.f;
    if (qq < p.Q && d < p.D) v = __uint_as_float(static_cast<uint32_t>(p.q[qq * p.ldq + d]) << 16);
    q_s[i] = v;
  }
  __syncthreads();

  const long long cta_lo = static_cast<long long>(blockIdx.x) * p.rows_per_cta;
  const long long cta_hi = min(p.n_local, cta_lo + p.rows_per_cta);
  const int n_steps = cta_hi > cta_lo ? static_cast<int>((cta_hi - cta_lo + 7) >> 3) : 0;

  // per-warp top-K lists, one per query, spread over lanes 0..K-1
  float lv[QT], thr[QT];
  int li[QT];
#pragma unroll
  for (int qq = 0; qq < QT; ++qq) {
    lv[qq] = -INFINITY;
    li[qq] = 0x7fffffff;
    thr[qq] = -INFINITY;
  }
  const int my_i = lane / LPV;   // value index this lane holds after the butterfly
  const int my_r = my_i / QT, my_q = my_i % QT;
  const bool rep = (lane % LPV) == 0 && my_q < p.Q;

  // loads of (step, chunk): 8 rows x 16 bytes per lane; rows beyond the slice / columns beyond D read as zero
  auto load_unit = [&](uint4 (&b)[8], int step, int c) {
    const long long row0 = cta_lo + (static_cast<long long>(step) << 3);
    const int col = c * 256 + lane * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      b[r] = make_uint4(0u, 0u, 0u, 0u);
      if (row0 + r < cta_hi && col < p.D) b[r] = ldg_stream16(p.bank + (row0 + r) * p.ldb + col);
    }
  };

  // Q <= 2: the next unit's 8 loads are in flight while this unit is reduced (double buffer); Q = 4 has no registers
  // left for that (partial sums + query chunk = 64) and relies on the other 15 warps of the SM for overlap
  constexpr bool DB = QT <= 2;
  uint4 cur[8], nxt[DB ? 8 : 1];
  int step = warp;
  if (DB && step < n_steps) load_unit(cur, step, 0);
  for (; step < n_steps; step += STREAM_WARPS) {
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if constexpr (DB) {  // next unit's loads go out before this unit's math
        if (c + 1 < NCH) {
          load_unit(nxt, step, c + 1);
        } else if (step + STREAM_WARPS < n_steps) {
          load_unit(nxt, step + STREAM_WARPS, 0);
        }
      } else {
        load_unit(cur, step, c);
      }
      float qf[QT][8];
#pragma unroll
      for (int qq = 0; qq < QT; ++qq) {
        const float4 a = *reinterpret_cast<const float4*>(&q_s[qq * DP + c * 256 + lane * 8]);
        const float4 b = *reinterpret_cast<const float4*>(&q_s[qq * DP + c * 256 + lane * 8 + 4]);
        qf[qq][0] = a.x, qf[qq][1] = a.y, qf[qq][2] = a.z, qf[qq][3] = a.w;
        qf[qq][4] = b.x, qf[qq][5] = b.y, qf[qq][6] = b.z, qf[qq][7] = b.w;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t w[4] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w};
        float x[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          x[2 * j] = __uint_as_float(w[j] << 16);
          x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
#pragma unroll
        for (int qq = 0; qq < QT; ++qq) {
          float a = acc[r * QT + qq];
#pragma unroll
          for (int j = 0; j < 8; ++j) a = fmaf(x[j], qf[qq][j], a);
          acc[r * QT + qq] = a;
        }
      }
      if constexpr (DB) {
#pragma unroll
        for (int r = 0; r < 8; ++r) cur[r] = nxt[r];
      }
    }
    // transposing butterfly: afterwards lanes [i*LPV, (i+1)*LPV) hold the full sum of value i = row*QT + query
    {
      int off = 16;
#pragma unroll
      for (int n = NV; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
          const float keep = up ? acc[j + half] : acc[j];
          const float send = up ? acc[j] : acc[j + half];
          acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
      }
#pragma unroll
      for (int o = LPV >> 1; o > 0; o >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
    }
    const float s = acc[0];
    const long long row0 = cta_lo + (static_cast<long long>(step) << 3);
    float my_thr = thr[0];
#pragma unroll
    for (int qq = 1; qq < QT; ++qq)
      if (my_q == qq) my_thr = thr[qq];
    unsigned cand = __ballot_sync(0xffffffffu, rep && row0 + my_r < cta_hi && s > my_thr);
    while (cand) {  // rare after the first few steps; everything below is warp-uniform
      const int b = __ffs(cand) - 1;
      cand &= cand - 1;
      float sb = __shfl_sync(0xffffffffu, s, b);
      const int ib = b / LPV, rb = ib / QT, qb = ib % QT;
      const long long rowb = row0 + rb;
      if (p.n_exclude > 0) {
        const long long gidx = p.index_base + rowb;
        const long long* ex = p.exclude + qb * p.exclude_ld;
        for (int x = 0; x < p.n_exclude; ++x)
          if (ex[x] == gidx) sb -= 1000.0f;  // gill/models.py:679-680
      }
#pragma unroll
      for (int qq = 0; qq < QT; ++qq) {
        if (qb == qq && sb > thr[qq]) {
          // entries >= sb stay ahead (they have lower row ids); the rest shift down by one lane
          const int pos = __popc(__ballot_sync(0xffffffffu, lane < p.K && lv[qq] >= sb));
          const float up_v = __shfl_up_sync(0xffffffffu, lv[qq], 1);
          const int up_i = __shfl_up_sync(0xffffffffu, li[qq], 1);
          if (lane < p.K) {
            if (lane > pos) {
              lv[qq] = up_v;
              li[qq] = up_i;
            } else if (lane == pos) {
              lv[qq] = sb;
              li[qq] = static_cast<int>(rowb);
            }
          }
          thr[qq] = __shfl_sync(0xffffffffu, lv[qq], p.K - 1);
        }
      }
    }
  }

  // ---- CTA merge: warp lists -> shared memory -> warp qq merges query qq (value desc, row asc)
#pragma unroll
  for (int qq = 0; qq < QT; ++qq) {
    if (lane < KMAX) {
      m_val[warp][qq][lane] = lane < p.K ? lv[qq] : -INFINITY;
      m_idx[warp][qq][lane] = lane < p.K ? li[qq] : 0x7fffffff;
    }
  }
  __syncthreads();
  if (warp < QT) {
    const int qq = warp;
    float last_v = INFINITY;
    int last_i = -1;
    for (int k = 0; k < p.K; ++k) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int t = lane; t < STREAM_WARPS * KMAX; t += 32) {
        const float v = m_val[t / KMAX][qq][t % KMAX];
        const int i = m_idx[t / KMAX][qq][t % KMAX];
        const bool after = (v < last_v) || (v == last_v && i > last_i);
        const bool better = (v > bv) || (v == bv && i < bi);
        if (after && better) {
          bv = v;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        const long long o = (static_cast<long long>(blockIdx.x) * QT + qq) * KMAX + k;
        p.part_val[o] = bv;
        p.part_idx[o] = bv == -INFINITY ? 0x7fffffffffffffffLL : p.index_base + bi;
      }
      last_v = bv;
      last_i = bi;
    }
  }
}

template <int QT, int NCH>
static int launch_topk_stream(const StreamParams& p, int grid, cudaStream_t stream) {
  topk_stream_kernel<QT, NCH><<<grid, STREAM_THREADS, 0, stream>>>(p);
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}
template <int QT>
static int launch_topk_stream_q(const StreamParams& p, int grid, cudaStream_t stream) {
  switch ((p.D + 255) / 256) {
    case 1: return launch_topk_stream<QT, 1>(p, grid, stream);
    case 2: return launch_topk_stream<QT, 2>(p, grid, stream);
    case 3: return launch_topk_stream<QT, 3>(p, grid, stream);
    default: return launch_topk_stream<QT, 4>(p, grid, stream);
  }
}

static int topk_splits(int Q, long long n_local, int* tiles_per_split, int* num_n_out) {
  const int num_m = (Q + BLOCK_M - 1) / BLOCK_M;
  const int num_n = static_cast<int>((n_local + TOPK_BN - 1) / TOPK_BN);
  int splits = num_sms() / num_m;
  if (splits < 1) splits = 1;
  if (splits > num_n) splits = num_n;
  const int tps = (num_n + splits - 1) / splits;
  splits = (num_n + tps - 1) / tps;
  *tiles_per_split = tps;
  *num_n_out = num_n;
  return splits;
}

}  // namespace gb

using namespace gb;

extern "C" long long gillb200_topk_workspace_bytes(int Q, long long n_local) {
  int tps, num_n;
  const int splits = topk_splits(Q, n_local, &tps, &num_n);
  const int q_pad = (Q + BLOCK_M - 1) / BLOCK_M * BLOCK_M;
  return 2LL * splits * q_pad * KMAX * (sizeof(float) + sizeof(long long)) + 256;
}

extern "C" int gillb200_topk_scores(const void* bank, long long n_local, int d, long long ld_bank, const void* q, int Q,
                                    long long ldq, int K, long long index_base, const long long* exclude_idx,
                                    int n_exclude, long long exclude_ld, void* workspace, float* out_val,
                                    long long* out_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(bank && q && workspace && out_val && out_idx, "null pointer");
  GB_CHECK_ARG(K >= 1 && K <= KMAX, "K=%d out of range [1,%d]", K, KMAX);
  GB_CHECK_ARG(Q >= 1 && n_local >= 1 && d >= 8, "bad shape Q=%d n_local=%lld d=%d", Q, n_local, d);
  GB_CHECK_ARG(n_local < (1LL << 31), "n_local must fit in int32");
  GB_CHECK_ARG(ld_bank % 8 == 0 && ldq % 8 == 0, "row strides must be multiples of 8 elements");
  GB_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace must be 16-byte aligned");
  GB_CHECK_ARG(exclude_ld == 0 || exclude_ld >= n_exclude, "exclude_ld must be 0 (shared list) or >= n_exclude");

  static int env_stream = -1;
  if (env_stream < 0) {
    const char* ev = getenv("GILLB200_TOPK_STREAM");  // "0": always the tensor-core kernel (A/B aid)
    env_stream = ev ? atoi(ev) : 1;
  }
  if (env_stream && Q <= 4 && d <= 1024 && d % 8 == 0 && reinterpret_cast<uintptr_t>(bank) % 16 == 0) {
    // ---- bank-streaming path (HBM-bound)
    StreamParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.bank = static_cast<const uint16_t*>(bank);
    sp.ldb = ld_bank;
    sp.n_local = n_local;
    sp.q = static_cast<const uint16_t*>(q);
    sp.ldq = ldq;
    sp.D = d;
    sp.Q = Q;
    sp.K = K;
    sp.index_base = index_base;
    sp.exclude = exclude_idx;
    sp.exclude_ld = exclude_ld;
    sp.n_exclude = exclude_idx ? n_exclude : 0;
    const int QT = Q == 1 ? 1 : Q == 2 ? 2 : 4;
    int grid = num_sms();
    long long per = (n_local + grid - 1) / grid;
    per = (per + 7) / 8 * 8;
    grid = static_cast<int>((n_local + per - 1) / per);
    sp.rows_per_cta = static_cast<int>(per);
    sp.part_idx = reinterpret_cast<long long*>(workspace);
    sp.part_val = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 1LL * grid * QT * KMAX * sizeof(long long));
    int r = QT == 1 ? launch_topk_stream_q<1>(sp, grid, stream)
            : QT == 2 ? launch_topk_stream_q<2>(sp, grid, stream)
                      : launch_topk_stream_q<4>(sp, grid, stream);
    if (r) return r;
    GB_CUDA(launch_topk_merge(sp.part_val, sp.part_idx, grid, KMAX, static_cast<long long>(QT) * KMAX,
                              static_cast<long long>(QT) * KMAX, K, Q, K, out_val, out_idx, stream));
    GB_COUNT_LAUNCH(1);
    return 0;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = Q;
  p.N = static_cast<int>(n_local);
  p.num_k_blocks = (d + BLOCK_K - 1) / BLOCK_K;
  p.kb_split = p.num_k_blocks;
  p.b_kb_wrap = 1 << 30;
  p.a_mode = A_PLAIN;
  p.in_dtype = DT_BF16;
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)Q};
    const uint64_t strides[1] = {(uint64_t)ldq * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int r = encode_tmap_16bit(&p.tma_a, q, 2, dims, strides, box, true);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)n_local};
    const uint64_t strides[1] = {(uint64_t)ld_bank * 2};
    const uint32_t box[2] = {BLOCK_K, TOPK_BN};
    int r = encode_tmap_16bit(&p.tma_b, bank, 2, dims, strides, box, true);
    if (r) return r;
  }
  p.num_stages = GemmCfg<TOPK_BN>::STAGES;
  TopkEpi e;
  memset(&e, 0, sizeof(e));
  const int num_m = (Q + BLOCK_M - 1) / BLOCK_M;
  int tps, num_n;
  const int splits = topk_splits(Q, n_local, &tps, &num_n);
  e.Q = Q;
  e.n_local = n_local;
  e.index_base = index_base;
  e.exclude = exclude_idx;
  e.exclude_ld = exclude_ld;
  e.n_exclude = exclude_idx ? n_exclude : 0;
  e.q_pad = num_m * BLOCK_M;
  e.tiles_per_split = tps;
  e.num_n = num_n;
  const long long part_elems = 2LL * splits * e.q_pad * KMAX;
  e.part_idx = reinterpret_cast<long long*>(workspace);
  e.part_val = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + part_elems * sizeof(long long));

  using C = GemmCfg<TOPK_BN>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(topk_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  topk_scores_kernel<<<num_m * splits, TOPK_THREADS, C::SMEM_BYTES, stream>>>(p, e);
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  GB_CUDA(launch_topk_merge(e.part_val, e.part_idx, splits * 2, KMAX, static_cast<long long>(e.q_pad) * KMAX,
                            static_cast<long long>(e.q_pad) * KMAX, KMAX, Q, K, out_val, out_idx, stream));
  GB_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int gillb200_topk_merge(const float* cand_val, const long long* cand_idx, int R, int Q, int Kc, int K,
                                   float* out_val, long long* out_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(cand_val && cand_idx && out_val && out_idx, "null pointer");
  GB_CHECK_ARG(R >= 1 && Q >= 1 && K >= 1 && K <= R * Kc, "bad merge shape R=%d Q=%d Kc=%d K=%d", R, Q, Kc, K);
  GB_CUDA(launch_topk_merge(cand_val, cand_idx, R, Kc, static_cast<long long>(Q) * Kc, static_cast<long long>(Q) * Kc, Kc, Q,
                            K, out_val, out_idx, stream));
  GB_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int gillb200_topk_merge_strided(const float* cand_val, long long r_stride_val, const long long* cand_idx,
                                           long long r_stride_idx, long long q_stride, int R, int Q, int Kc, int K,
                                           float* out_val, long long* out_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(cand_val && cand_idx && out_val && out_idx, "null pointer");
  GB_CHECK_ARG(R >= 1 && Q >= 1 && K >= 1 && K <= R * Kc && q_stride >= Kc, "bad merge shape R=%d Q=%d Kc=%d K=%d", R, Q,
               Kc, K);
  GB_CUDA(launch_topk_merge(cand_val, cand_idx, R, q_stride, r_stride_val, r_stride_idx, Kc, Q, K, out_val, out_idx, stream));
  GB_COUNT_LAUNCH(1);
  return 0;
}
