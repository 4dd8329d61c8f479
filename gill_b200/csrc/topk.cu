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

#include <cstring>

namespace gb {

constexpr int TOPK_BN = 256;
constexpr int KMAX = 16;

struct TopkEpi {
  int Q;
  long long n_local;
  long long index_base;
  const long long* exclude;
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
              if (e.n_exclude > 0) {
                const long long gidx = e.index_base + nloc;
                for (int x = 0; x < e.n_exclude; ++x)
                  if (e.exclude[x] == gidx) s -= 1000.0f;  // gill/models.py:679-680
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
                                  long long r_stride /* elements between lists */, int kc, int Q, int K,
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
      const long long off = r * r_stride + q * q_stride + j;
      const float v = cand_val[off];
      const long long i = cand_idx[off];
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
                                    int n_exclude, void* workspace, float* out_val, long long* out_idx,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(bank && q && workspace && out_val && out_idx, "null pointer");
  GB_CHECK_ARG(K >= 1 && K <= KMAX, "K=%d out of range [1,%d]", K, KMAX);
  GB_CHECK_ARG(Q >= 1 && n_local >= 1 && d >= 8, "bad shape Q=%d n_local=%lld d=%d", Q, n_local, d);
  GB_CHECK_ARG(n_local < (1LL << 31), "n_local must fit in int32");
  GB_CHECK_ARG(ld_bank % 8 == 0 && ldq % 8 == 0, "row strides must be multiples of 8 elements");
  GB_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace must be 16-byte aligned");

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
  e.n_exclude = exclude_idx ? n_exclude : 0;
  e.q_pad = num_m * BLOCK_M;
  e.tiles_per_split = tps;
  e.num_n = num_n;
  const long long part_elems = 2LL * splits * e.q_pad * KMAX;
  e.part_idx = reinterpret_cast<long long*>(workspace);
  e.part_val = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + part_elems * sizeof(long long));

  using C = GemmCfg<TOPK_BN>;
  static bool configured = false;
  if (!configured) {
    GB_CUDA(cudaFuncSetAttribute(topk_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  topk_scores_kernel<<<num_m * splits, TOPK_THREADS, C::SMEM_BYTES, stream>>>(p, e);
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  const int warps_per_block = 4;
  topk_merge_kernel<<<(Q + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      e.part_val, e.part_idx, splits * 2, KMAX, static_cast<long long>(e.q_pad) * KMAX, KMAX, Q, K, out_val, out_idx);
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_topk_merge(const float* cand_val, const long long* cand_idx, int R, int Q, int Kc, int K,
                                   float* out_val, long long* out_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(cand_val && cand_idx && out_val && out_idx, "null pointer");
  GB_CHECK_ARG(R >= 1 && Q >= 1 && K >= 1 && K <= R * Kc, "bad merge shape R=%d Q=%d Kc=%d K=%d", R, Q, Kc, K);
  const int warps_per_block = 4;
  topk_merge_kernel<<<(Q + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      cand_val, cand_idx, R, Kc, static_cast<long long>(Q) * Kc, Kc, Q, K, out_val, out_idx);
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}
