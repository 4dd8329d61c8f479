// CTA-pair (cta_group::2) variant of the tcgen05 GEMM / implicit conv: one 256 x BLOCK_N tile per pair of SMs.
//
// Why: measured on B200 (profiles/r01_ncu_*), TMA can fill one SM's shared memory at ~64 B per SM clock, while a
// 1-CTA 128 x BN tile consumes 64 + 8192/BN B per MMA clock -- it can never be MMA-bound (tensor pipe 45 % at BN=160,
// 75 % at BN=256). A CTA pair shares the B tile: each CTA loads its own 128 A rows and HALF of the B rows, the single
// tcgen05.mma.cta_group::2 issued by the leader reads both halves, so a CTA consumes 32 + 8192/BN B per MMA clock.
//
// Protocol (per pair; barriers live at identical smem offsets in both CTAs):
//   full[s]       leader only. Both producers' TMA loads complete_tx on the LEADER's barrier; the leader's producer
//                 arms it with the byte count of both CTAs.
//   empty[s]      one per CTA. The leader's tcgen05.commit multicasts the arrival to both CTAs, each producer waits on
//                 its own copy before refilling its own smem slot.
//   tmem_full[a]  one per CTA (multicast commit); each CTA's epilogue drains its own 128 accumulator rows.
//   tmem_empty[a] leader only, count = 2 x epilogue warps: the peer's epilogue warps arrive remotely (mapa).
#pragma once
#include "gemm_sm100.cuh"

namespace gb {

#ifndef GILLB200_HALO_AHEAD
#define GILLB200_HALO_AHEAD 0  // measured for plain halo convs: 4.99 vs 4.65 ms per UNet evaluation (a third slot costs weight-ring stages)
#endif
constexpr bool HALO_AHEAD_ALWAYS = GILLB200_HALO_AHEAD != 0;

template <int BLOCK_N>
struct Gemm2Cfg {
  // BLOCK_N <= 256: one tcgen05.mma of N = BLOCK_N per K-step, two TMEM accumulator stages.
  // BLOCK_N == 320 ("wide" tile, for the N = 320 / 640 / 1280 implicit convs): two MMAs of N = 160 per K-step into ONE
  //   320-column accumulator stage. Operand fill per MMA clock drops from 81 B (256 x 160 tiles: every A tile fetched once per
  //   N tile) to 56 B per SM, against an L2 -> SM ceiling of ~43 B per SM clock with all 148 SMs pulling (ncu r01: the
  //   256 x 160 conv moved 1.23 GB L2 -> SM per launch, tensor pipe 44 %). The epilogue no longer overlaps the next tile's
  //   main loop -- ~1.7 k of ~29 k clocks per tile at K = 2880.
  static constexpr int NMMA = BLOCK_N > 256 ? 2 : 1;
  static constexpr int MMA_N = BLOCK_N / NMMA;
  static constexpr int NACC = BLOCK_N > 256 ? 1 : 2;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // this CTA's 128 rows
  static constexpr int B_SUB = (MMA_N / 2) * BLOCK_K * 2;        // this CTA's half of one MMA's B rows
  static constexpr int B_BYTES = NMMA * B_SUB;                   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - BAR_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int ACC_STRIDE = BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : BLOCK_N <= 256 ? 256 : 512;
  static constexpr int TMEM_COLS = NACC * ACC_STRIDE;
  // A_CONV3X3_HALO: two resident halo-tile slots in front of a B-only ring
  // (three slots with the GroupNorm transform: load latency + transform of a slot must hide under the MMAs of the others)
  static constexpr int HALO_RES = 2 * HALO_BYTES;
  static constexpr int HALO_RES_GN = 3 * HALO_BYTES;
  static_assert(MMA_N % 32 == 0 && MMA_N >= 64 && MMA_N <= 256 && TMEM_COLS <= 512, "invalid 2-CTA UMMA N");
  static_assert(B_SUB % 1024 == 0, "B sub-tiles must keep 1024-B alignment");
};

// Tile schedule of one CTA pair. Whole rounds hand pair p the tiles p, p + P, p + 2P, ...; when the tiles left for the
// last, partial round of a wide (two-MMA) kernel number at most P / 2, each of them is split into its two N halves so
// that twice as many pairs stay busy (64x64 UNet convs: 256 tiles on 74 pairs = 3 rounds + 34 tiles -> 68 half tiles).
// `code` = tile * 3 + (0: whole tile, 1 / 2: first / second N half).
struct PairSched {
  int kb_per_tile;  // (name shared with SegIter: the staged epilogue reads it)
  int pair, num_pairs, num_tiles, whole_rounds, rem;
  bool split;
  int i;
  __device__ __forceinline__ bool next(Seg& sg) {
    sg.kb0 = 0;
    sg.kb1 = kb_per_tile;
    if (i < whole_rounds) {
      sg.tile = (pair + i * num_pairs) * 3;
      ++i;
      return true;
    }
    if (i > whole_rounds) return false;
    ++i;
    const int base = whole_rounds * num_pairs;
    if (split) {
      if (pair >= 2 * rem) return false;
      sg.tile = (base + (pair >> 1)) * 3 + 1 + (pair & 1);
      return true;
    }
    if (pair >= rem) return false;
    sg.tile = (base + pair) * 3;
    return true;
  }
};
__device__ __forceinline__ PairSched make_pair_sched(int kb_per_tile, int pair, int num_pairs, int num_tiles, bool wide) {
  PairSched s;
  s.kb_per_tile = kb_per_tile;
  s.pair = pair;
  s.num_pairs = num_pairs;
  s.num_tiles = num_tiles;
  s.whole_rounds = num_tiles / num_pairs;
  s.rem = num_tiles - s.whole_rounds * num_pairs;
  s.split = wide && s.rem > 0 && 2 * s.rem <= num_pairs;
  s.i = 0;
  return s;
}

template <int BLOCK_N, bool HALO = false, bool GN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ GemmParams p) {
  static_assert(!GN || HALO, "the GroupNorm transform lives on the halo tile");
  using C = Gemm2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // HALO: [2 halo slots][num_stages x B half-tiles]; else [num_stages x (A + B half-tile)]
  constexpr int RING_STAGE = HALO ? C::B_BYTES : C::STAGE_BYTES;
  constexpr bool AHEAD = GN || HALO_AHEAD_ALWAYS;  // three halo slots, next tile requested one channel block ahead
  constexpr int HSLOTS = AHEAD ? 3 : 2;
  uint8_t* smem_tiles = smem + (HALO ? HSLOTS * HALO_BYTES : 0);
  GemmSmemBars* bars = reinterpret_cast<GemmSmemBars*>(smem_tiles + p.num_stages * RING_STAGE);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int num_tiles = num_m2 * num_n * ksplit;  // work units: (tile, K slice)
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const PairSched sched0 = make_pair_sched(p.num_k_blocks, pair, num_pairs, num_tiles, C::NMMA == 2 && ksplit == 1);
  // unit -> (tile, first k-block, one-past-last k-block, output row offset of the slice)
  auto unit_of = [&](int unit, int* tile, int* kb0, int* kb1, int* row_off) {
    if (ksplit == 1) {
      *tile = unit, *kb0 = 0, *kb1 = p.num_k_blocks, *row_off = 0;
    } else {
      const int sl = unit % ksplit;
      *tile = unit / ksplit;
      *kb0 = sl * p.kb_per_split;
      *kb1 = min(p.num_k_blocks, *kb0 + p.kb_per_split);
      *row_off = sl * p.M;
    }
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    if (p.kb_split < p.num_k_blocks) tma_prefetch_desc(&p.tma_a2);
    mbar_init(&bars->b_full, 1);
    for (int i = 0; i < (HALO ? 8 : C::STAGES); ++i) {  // (the halo mode's B-only ring is deeper than C::STAGES)
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 2 * (p.epi_tma ? p.epi_warps : GEMM_EPI_WARPS));
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->halo_full[i], 1);
      mbar_init(&bars->halo_empty[i], 1);
      mbar_init(&bars->halo_ready[i], 8);  // four transform warps in each CTA of the pair
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
    tmem_alloc_2cta(&bars->tmem_ptr, C::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  __syncthreads();     // (CTA-scope barrier as well: racecheck does not model barrier.cluster as ordering the allocator's
                       // shared-memory write of the TMEM base address against the reads below)
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if constexpr (HALO) {
      // ---------------- TMA producer, halo mode: per 64-channel block ONE halo tile (all nine taps) + nine weight tiles
      int stage = 0, hb = 0;
      uint32_t phase = 0, hphase = 0;
      PairSched sched = sched0;
      Seg sg;
      const int hw = p.conv_W * p.conv_H, bxn = p.conv_W >> 3;
      auto issue_halo = [&](int code, int cb) {  // halo tile of (tile code, channel block) into the next slot
        const int tile = code / 3;
        const int m0 = (tile % num_m2) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
        const int cb0 = m0 / hw, blk = (m0 - cb0 * hw) >> 7;
        const int cx0 = (blk % bxn) * 8, cy0 = (blk / bxn) * 16;
        mbar_wait(&bars->halo_empty[hb], hphase ^ 1);
        const uint32_t hfull_leader = mapa_u32(smem_u32(&bars->halo_full[hb]), 0);
        if (elect_one()) {
          if (GN) {
            // each CTA's transform warps wait for THEIR tile on the local barrier; the leader's MMA warp waits for
            // halo_ready instead
            mbar_arrive_expect_tx(&bars->halo_full[hb], HALO_TX);
            const bool src0 = cb < p.halo_c0_blocks;
            tma_load_4d(smem + hb * HALO_BYTES, src0 ? &p.tma_a : &p.tma_a2, &bars->halo_full[hb],
                        (src0 ? cb : cb - p.halo_c0_blocks) * BLOCK_K, cx0 - 1, cy0 - 1, cb0);
          } else {
            if (leader) mbar_arrive_expect_tx(&bars->halo_full[hb], 2 * HALO_TX);
            tma2_load_4d(smem + hb * HALO_BYTES, &p.tma_a, hfull_leader, cb * BLOCK_K, cx0 - 1, cy0 - 1, cb0);
          }
        }
        __syncwarp();
        if (++hb == HSLOTS) {
          hb = 0;
          hphase ^= 1;
        }
      };
      // GN: the halo tile of the NEXT (tile, channel block) is requested BEFORE this block's weight tiles -- issued after
      // them it could run at most one B-ring depth (~5 taps) ahead of the MMAs, less than TMA latency + transform time
      PairSched ahead = sched0;
      Seg sa;
      bool more = AHEAD && ahead.next(sa);
      int acb = 0;
      auto issue_ahead = [&]() {
        if (!more) return;
        issue_halo(sa.tile, acb);
        if (++acb == p.conv_cblocks) {
          acb = 0;
          more = ahead.next(sa);
        }
      };
      if (AHEAD) issue_ahead();
      while (sched.next(sg)) {
        const int part = sg.tile % 3;
        const int tile = sg.tile / 3;
        const int n0 = (tile / num_m2) * BLOCK_N + (part == 2 ? C::MMA_N : 0) + static_cast<int>(rank) * (C::MMA_N / 2);
        const int nmma = part == 0 ? C::NMMA : 1;
        for (int cb = 0; cb < p.conv_cblocks; ++cb) {
          if (AHEAD) issue_ahead();
          else issue_halo(sg.tile, cb);
          for (int tap = 0; tap < p.conv_nt; ++tap) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sb = smem_tiles + stage * RING_STAGE;
            const uint32_t full_leader = mapa_u32(smem_u32(&bars->full[stage]), 0);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&bars->full[stage], 2 * nmma * C::B_SUB);
              const int bcol = (tap * p.conv_cblocks + cb) * BLOCK_K;  // weight column of (tap, channel block)
              tma2_load_2d(sb, &p.tma_b, full_leader, bcol, n0);
              if (C::NMMA == 2 && nmma == 2) tma2_load_2d(sb + C::B_SUB, &p.tma_b, full_leader, bcol, n0 + C::MMA_N);
            }
            __syncwarp();
            if (++stage == p.num_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else {
      // ---------------- TMA producer (both CTAs); warp-uniform loop, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      PairSched sched = sched0;
      Seg sg;
      while (sched.next(sg)) {
        const int part = sg.tile % 3;  // part 0: whole tile; 1 / 2: one N half of a wide tile
        int tile, kb0, kb1, row_off;
        unit_of(sg.tile / 3, &tile, &kb0, &kb1, &row_off);
        const int m0 = (tile % num_m2) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
        const int n0 = (tile / num_m2) * BLOCK_N + (part == 2 ? C::MMA_N : 0) + static_cast<int>(rank) * (C::MMA_N / 2);
        const int nmma = part == 0 ? C::NMMA : 1;
        int cb0 = 0, cy0 = 0, cx0 = 0;
        if (p.a_mode == A_CONV3X3) {
          const int per_img = p.conv_W * p.conv_H;
          cb0 = m0 / per_img;
          cy0 = (m0 % per_img) / p.conv_W;
          cx0 = m0 % p.conv_W;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem_tiles + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&bars->full[stage]), 0);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&bars->full[stage], 2 * (C::A_BYTES + nmma * C::B_SUB));
            if (kb >= p.kb_split) {
              tma2_load_2d(sa, &p.tma_a2, full_leader, (kb - p.kb_split) * BLOCK_K, m0);
            } else if (p.a_mode == A_CONV3X3) {
              const int tap = kb / p.conv_cblocks;
              const int cb = kb - tap * p.conv_cblocks;
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              tma2_load_4d(sa, &p.tma_a, full_leader, cb * BLOCK_K, p.conv_stride * cx0 + dx, p.conv_stride * cy0 + dy,
                           cb0);
            } else {
              tma2_load_2d(sa, &p.tma_a, full_leader, kb * BLOCK_K, m0);
            }
            const int bcol = (kb % p.b_kb_wrap) * BLOCK_K;
            tma2_load_2d(sb, &p.tma_b, full_leader, bcol, n0);
            if (C::NMMA == 2 && nmma == 2) tma2_load_2d(sb + C::B_SUB, &p.tma_b, full_leader, bcol, n0 + C::MMA_N);
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && HALO) {
    if (leader) {
      // ---------------- MMA issuer, halo mode: A descriptors walk the nine shifted views of the resident halo tile
      const uint32_t idesc = make_idesc_f16(2 * BLOCK_M, C::MMA_N, p.in_dtype == DT_BF16, false);
      const uint64_t dh0 = make_smem_desc_sw128(smem_u32(smem), 16, 1280);  // 8-pixel groups one halo row (10 px) apart
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem_tiles), 16, 1024);
      constexpr uint64_t DESC_STEP = RING_STAGE >> 4;
      constexpr uint64_t SUB_STEP = C::B_SUB >> 4;
      uint64_t db = db0;
      int stage = 0, hb = 0;
      uint32_t phase = 0, hphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ready = false;
      PairSched sched = sched0;
      Seg sg;
      while (sched.next(sg)) {
        const int nmma = sg.tile % 3 == 0 ? C::NMMA : 1;
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
        for (int cb = 0; cb < p.conv_cblocks; ++cb) {
          mbar_wait(GN ? &bars->halo_ready[hb] : &bars->halo_full[hb], hphase);
          tc_fence_after();
          const uint64_t dh = dh0 + static_cast<uint64_t>(hb) * (HALO_BYTES >> 4);
          for (int tap = 0; tap < p.conv_nt; ++tap) {
            if (!ready) mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const bool wrap = stage + 1 == p.num_stages;
            const int nstage = wrap ? 0 : stage + 1;
            const uint32_t nphase = wrap ? phase ^ 1 : phase;
            ready = mbar_test_wait(&bars->full[nstage], nphase);
            const uint64_t da =
                dh + static_cast<uint64_t>((p.conv_pa + tap / p.conv_ntx) * 10 + p.conv_pb + tap % p.conv_ntx) * (128 >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                umma2_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) != 0 ? 1u : 0u);
                if (C::NMMA == 2 && nmma == 2)
                  umma2_f16(tmem_d + C::MMA_N, da + 2 * k, db + SUB_STEP + 2 * k, idesc, (cb | tap | k) != 0 ? 1u : 0u);
              }
              umma2_commit_mcast(&bars->empty[stage], 0b11);
              if (tap == p.conv_nt - 1) umma2_commit_mcast(&bars->halo_empty[hb], 0b11);  // every tap of this halo tile has been read
            }
            __syncwarp();
            db = wrap ? db0 : db + DESC_STEP;
            stage = nstage;
            phase = nphase;
          }
          if (++hb == HSLOTS) {
            hb = 0;
            hphase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_mcast(&bars->tmem_full[acc], 0b11);
        __syncwarp();
        if (++acc == C::NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA only; one thread drives both SMs' tensor cores)
      const uint32_t idesc = make_idesc_f16(2 * BLOCK_M, C::MMA_N, p.in_dtype == DT_BF16, false);
      // incremental descriptors + look-ahead probe of the next stage's barrier: see gemm_mma()
      const uint32_t s0 = smem_u32(smem_tiles);
      const uint64_t da0 = make_smem_desc_sw128(s0, 16, 1024);
      const uint64_t db0 = make_smem_desc_sw128(s0 + C::A_BYTES, 16, 1024);
      constexpr uint64_t DESC_STEP = C::STAGE_BYTES >> 4;
      constexpr uint64_t SUB_STEP = C::B_SUB >> 4;
      uint64_t da = da0, db = db0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ready = false;
      PairSched sched = sched0;
      Seg sg;
      while (sched.next(sg)) {
        const int nmma = sg.tile % 3 == 0 ? C::NMMA : 1;
        int tile_, kb0, kb1, row_off_;
        unit_of(sg.tile / 3, &tile_, &kb0, &kb1, &row_off_);
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!ready) mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const bool wrap = stage + 1 == p.num_stages;
          const int nstage = wrap ? 0 : stage + 1;
          const uint32_t nphase = wrap ? phase ^ 1 : phase;
          ready = mbar_test_wait(&bars->full[nstage], nphase);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma2_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
              if (C::NMMA == 2 && nmma == 2)
                umma2_f16(tmem_d + C::MMA_N, da + 2 * k, db + SUB_STEP + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            }
            umma2_commit_mcast(&bars->empty[stage], 0b11);  // frees this smem slot in BOTH CTAs
          }
          __syncwarp();
          da = wrap ? da0 : da + DESC_STEP;
          db = wrap ? db0 : db + DESC_STEP;
          stage = nstage;
          phase = nphase;
        }
        if (elect_one()) umma2_commit_mcast(&bars->tmem_full[acc], 0b11);  // accumulators ready in both CTAs
        __syncwarp();
        if (++acc == C::NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (GN && (warp == 2 || warp == 3 || warp == 8 || warp == 9)) {
    // ---------------- GroupNorm(+SiLU) of the input, in place on every halo tile (both CTAs, 64 threads each).
    // Thread t owns 16-byte chunk t % 8 (8 channels: their scale / shift sit in registers for the whole channel block) of
    // pixels t / 8, t / 8 + 8, ...; a warp touches 4 consecutive 128-byte rows per access (the swizzle only permutes chunks
    // inside a row). Pixels outside the image become exactly zero: the convolution pads the NORMALISED tensor.
    // FOUR warps on the four schedulers (2, 3, 8, 9 -> SMSP 2, 3, 0, 1; the epilogue runs on warps 4-7 in this mode): the
    // transform is bound by the XU pipe (2 MUFU + the 16-bit conversions per element) and two warps only reach two of them
    const int tt = (warp < 4 ? warp - 2 : warp - 6) * 32 + static_cast<int>(lane_id());
    const int c8 = tt & 7;
    const bool bf = p.in_dtype == DT_BF16;
    int hb = 0;
    uint32_t hphase = 0;
    PairSched sched = sched0;
    Seg sg;
    const int hw = p.conv_W * p.conv_H, bxn = p.conv_W >> 3;
    while (sched.next(sg)) {
      const int tile = sg.tile / 3;
      const int m0 = (tile % num_m2) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
      const int cb0 = m0 / hw, blk = (m0 - cb0 * hw) >> 7;
      const int cx0 = (blk % bxn) * 8, cy0 = (blk / bxn) * 16;
      for (int cb = 0; cb < p.conv_cblocks; ++cb) {
        const float* ss = p.gn_ss + static_cast<size_t>(cb0) * 2 * p.gn_C + cb * BLOCK_K + c8 * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(ss)), a1 = __ldg(reinterpret_cast<const float4*>(ss + 4));
        const float4 d0 = __ldg(reinterpret_cast<const float4*>(ss + p.gn_C)), d1 = __ldg(reinterpret_cast<const float4*>(ss + p.gn_C + 4));
        const float sa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float sd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        mbar_wait(&bars->halo_full[hb], hphase);
        uint8_t* tile_s = smem + hb * HALO_BYTES;
        // four pixels per step, loads first: four independent ex2 / rcp chains per thread (one warp per scheduler has no other
        // way to hide the MUFU latency -- the single-pixel loop took longer than the nine taps of MMAs it should hide under)
        const f32x2 nl2e = pk2(-1.4426950408889634f, -1.4426950408889634f), one2 = pk2(1.f, 1.f);
#pragma unroll 1
        for (int px0 = (p.debug_mode == 4 ? 180 : tt >> 3); px0 < 180; px0 += 64) {  // (debug_mode 4: barriers only, results invalid)
          uint4 v[4];
          uint4* ptr[4];
          bool inb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int px = px0 + 16 * u;
            const int py = px / 10, pxx = px - py * 10;
            const int gy = cy0 - 1 + py, gx = cx0 - 1 + pxx;
            inb[u] = px < 180 && gy >= 0 && gy < p.conv_H && gx >= 0 && gx < p.conv_W;
            ptr[u] = reinterpret_cast<uint4*>(tile_s + px * 128 + ((c8 ^ (px & 7)) << 4));
            v[u] = px < 180 ? *ptr[u] : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = bf ? unpack_bf16x2(w4[j]) : unpack_f16x2(w4[j]);
              f32x2 y = fma2(pk2(f.x, f.y), pk2(sa[2 * j], sa[2 * j + 1]), pk2(sd[2 * j], sd[2 * j + 1]));
              float y0, y1;
              if (p.gn_silu && p.debug_mode != 5) {  // (debug_mode 5: affine only, no MUFU -- measurement aid)
                float z0, z1, e0, e1, s0, s1, r0, r1;
                upk2(mul2(y, nl2e), z0, z1);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
                upk2(add2(pk2(e0, e1), one2), s0, s1);
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s0));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(s1));
                upk2(mul2(y, pk2(r0, r1)), y0, y1);
              } else {
                upk2(y, y0, y1);
              }
              w4[j] = inb[u] ? (bf ? pack_bf16x2(y0, y1) : pack_f16x2(y0, y1)) : 0u;
            }
            v[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (px0 + 16 * u < 180) *ptr[u] = v[u];
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane_id() == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->halo_ready[hb]), 0));
        if (++hb == HSLOTS) {
          hb = 0;
          hphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (both CTAs, own 128 accumulator rows)
    if (p.epi_tma) {
      if (warp - 4 < p.epi_warps) {
        epilogue_warp_tma_dispatch<BLOCK_N, C::ACC_STRIDE, C::NACC>(
            p, bars, epi_stage, tmem_base, sched0,
            [&](int code, int* row_base, int* n0, int* ncols) {
              const int part = code % 3;
              int tile, kb0, kb1, row_off;
              unit_of(code / 3, &tile, &kb0, &kb1, &row_off);
              *row_base = row_off + (tile % num_m2) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
              *n0 = (tile / num_m2) * BLOCK_N + (part == 2 ? C::MMA_N : 0);
              if (part != 0) *ncols = C::MMA_N;
            },
            [&](int acc) {
              tc_fence_before();
              __syncwarp();
              if (lane_id() == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
            });
      }
    } else {
    const int ewarp = warp & 3;
    const int cgrp = (warp - 4) >> 2;
    constexpr int NGRP = GEMM_EPI_WARPS / 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    PairSched sched = sched0;
    Seg sg;
    while (sched.next(sg)) {
      const int tile = sg.tile / 3, part = sg.tile % 3;
      const int nch = (part == 0 ? BLOCK_N : C::MMA_N) / 16;
      const int ch_begin = cgrp * (nch / NGRP) + min(cgrp, nch % NGRP);
      const int ch_end = ch_begin + nch / NGRP + (cgrp < nch % NGRP ? 1 : 0);
      const int row = (tile % num_m2) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M + ewarp * 32 + lane_id();
      const int n0 = (tile / num_m2) * BLOCK_N + (part == 2 ? C::MMA_N : 0);
      mbar_wait(&bars->tmem_full[acc], acc_phase);
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
      if (lane_id() == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
      if (++acc == C::NACC) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading our smem / signalling our barriers until here
  if (warp == 2) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
}

}  // namespace gb
