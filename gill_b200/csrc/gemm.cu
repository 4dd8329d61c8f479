// Host side of the tcgen05 GEMM: tensor-map construction, tile-shape selection, launch, C ABI.
#include "../../include/gillb200.h"
#include "gemm2_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_common.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace gb {

long long g_launch_count = 0;
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// dtype: DT_BF16 / DT_F16 / DT_F32. swizzle_bytes: 128, 64 or 32 (the box's inner extent must not exceed it).
int encode_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_err(-EIO, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return set_err(-EINVAL, "tensor base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) return set_err(-EINVAL, "tensor stride %llu B not a multiple of 16", (unsigned long long)gstr[i - 1]);
    }
  }
  const CUtensorMapDataType dt = dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return 0;
}

int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool is_bf16) {
  return encode_tmap(out, base, is_bf16 ? DT_BF16 : DT_F16, rank, dims, strides_bytes, box, 128, nullptr);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // opt-in: measured on B200 inside the captured UNet graph it buys nothing (22.19 vs 22.02 ms per evaluation, the
    // graph already launches back to back), so it stays off by default
    const char* e = getenv("GILLB200_PDL");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}

bool pdl_light_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GILLB200_PDL_LIGHT");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}

int num_sms() {
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0 && (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0))
    n[dev] = 148;
  return n[dev];
}

// Shared-memory plan: [num_stages x stage][1 KB barriers][epilogue staging] (+1 KB alignment slack). The smem-staged
// epilogue trades ring stages for staging buffers.
static int plan_smem(GemmParams& p, int stage_bytes, int max_stages, int resident_bytes = 0) {
  const int epi_bytes = p.epi_tma ? p.epi_warps * p.epi_nbuf * p.epi_buf_bytes : 0;
  int stages = (SMEM_BUDGET - 2048 - epi_bytes - resident_bytes) / stage_bytes;
  if (stages > max_stages) stages = max_stages;
  if (stages > 8) stages = 8;
  p.num_stages = stages;
  return resident_bytes + stages * stage_bytes + 2048 + epi_bytes;
}

template <int BN>
static int launch_gemm(GemmParams& p, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
  }
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  int grid = tiles < num_sms() ? tiles : num_sms();
  int smem_bytes;
  if (p.b_resident) {
    // B-stationary: the weight tile [BN x K] sits in front of an A-only ring; n-inner order + a grid that is a multiple of
    // the N-tile count pin one N tile to each CTA
    p.b_res_bytes = p.num_k_blocks * C::B_BYTES;
    smem_bytes = plan_smem(p, C::A_BYTES, 8, p.b_res_bytes);
    if (p.num_stages < 3) {
      p.b_resident = 0;
      p.b_res_bytes = 0;
    } else {
      grid = num_sms() / num_n * num_n;
      p.tile_order = 2;
    }
  }
  if (!p.b_resident) smem_bytes = plan_smem(p, C::STAGE_BYTES, C::STAGES);
  GB_CHECK_ARG(p.num_stages >= 2, "no room for a 2-stage ring next to the epilogue staging (BN=%d)", BN);
  if (p.sk_per > 0) {  // stream-K: every SM gets an equal share of the (tile, k-block) units
    const long long total = 1LL * tiles * p.num_k_blocks;
    grid = num_sms();
    p.sk_per = static_cast<int>((total + grid - 1) / grid);
    grid = static_cast<int>((total + p.sk_per - 1) / p.sk_per);
  }
  GB_CUDA(launch_pdl(gemm_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), smem_bytes, stream, p));
  GB_COUNT_LAUNCH(1);
  return 0;
}

template <int BN, bool HALO = false, bool GN = false>
static int launch_gemm2(GemmParams& p, cudaStream_t stream) {
  using C = Gemm2Cfg<BN>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, HALO, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
  }
  // halo mode: two resident halo-tile slots in front of a ring of B half-tiles (8 stages at most: the barrier block)
  const int smem_bytes = HALO ? plan_smem(p, C::B_BYTES, 8, (GN || HALO_AHEAD_ALWAYS) ? C::HALO_RES_GN : C::HALO_RES) : plan_smem(p, C::STAGE_BYTES, C::STAGES);
  GB_CHECK_ARG(p.num_stages >= 2, "no room for a 2-stage ring next to the epilogue staging (pair, BN=%d)", BN);
  const int num_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles = num_m2 * num_n;
  const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  GB_CUDA(launch_pdl(gemm2_kernel<BN, HALO, GN>, dim3(2 * pairs), dim3(GEMM_THREADS), smem_bytes, stream, p));  // cluster (2,1,1)
  GB_COUNT_LAUNCH(1);
  return 0;
}

// Split-K second pass (see GemmParams::ksplit): out = act-free epilogue of sum_s partial[s] -- bias, per-sample row bias,
// residual, 16-bit rounding and the GroupNorm statistics of the rounded output (same layout as the staged epilogue's
// stats_out: per 32-row slab and column {sum, sumsq}). One thread per column, 32 rows per CTA: every load / store of a
// warp is one contiguous row segment.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int S, int M, int N,
                                                            const float* __restrict__ bias, const float* __restrict__ rowbias,
                                                            long long ld_rowbias, int rows_per_group,
                                                            const uint16_t* __restrict__ residual, long long ldr,
                                                            uint16_t* __restrict__ out, long long ldo, int out_bf16,
                                                            float2* __restrict__ stats_out) {
  pdl_wait();
  pdl_launch();
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int row0 = blockIdx.y * 32;
  if (col >= N) return;
  const float b = bias ? bias[col] : 0.f;
  float cs = 0.f, cq = 0.f;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const int row = row0 + r;
    if (row >= M) break;
    float v = b;
    for (int s = 0; s < S; ++s) v += part[(static_cast<size_t>(s) * M + row) * N + col];
    if (rowbias) v += rowbias[static_cast<long long>(row / rows_per_group) * ld_rowbias + col];
    if (residual) {
      const uint16_t h = residual[static_cast<long long>(row) * ldr + col];
      v += out_bf16 ? __uint_as_float(static_cast<uint32_t>(h) << 16) : __half2float(__ushort_as_half(h));
    }
    uint16_t o;
    float back;
    if (out_bf16) {
      const __nv_bfloat16 t = __float2bfloat16_rn(v);
      o = __bfloat16_as_ushort(t);
      back = __bfloat162float(t);
    } else {
      const __half t = __float2half_rn(v);
      o = __half_as_ushort(t);
      back = __half2float(t);
    }
    out[static_cast<long long>(row) * ldo + col] = o;
    cs += back;
    cq = fmaf(back, back, cq);
  }
  if (stats_out) stats_out[static_cast<size_t>(blockIdx.y) * N + col] = make_float2(cs, cq);
}

// Pick the N tile width from a small cost model fitted to measurements on B200 (tools/gpu_sweep_shapes.py,
// profiles/r01_shape_sweep.log): time ~ waves x (k_blocks x BN / eff(BN) + epilogue(BN)), where eff() is the measured
// MMA-issue / smem-fill efficiency of a 128 x BN tile relative to BN = 256 (MMA-only ceilings: 1868 TF/s at BN=256,
// ~1300 at BN=160/128). Padding waste is implicit in the tile count.
// plain GEMMs: minimum K depth (in 64-wide blocks) from which the wide pair tile is picked automatically; tuning aid
static int wide_min_kb() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GILLB200_WIDE_GEMM_KB");
    // measured (tools/gpu_gemm_bench.py, profiles/r02_gemm_shapes.log): from K = 960 up the wide tile wins or ties on
    // every N % 320 == 0 linear with enough tiles (M4096 N1280 K5120 61.7 -> 45.3 us, K2560 35.8 -> 28.3, K1280 24.3 -> 20.8,
    // M65536 N320 K1280 68.0 -> 65.0); at K <= 768 the un-overlapped epilogue costs more than the fill it saves
    v = e ? atoi(e) : 15;
  }
  return v;
}

static int conv_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GILLB200_CONV_HALO");  // "0": one TMA box per tap (round-1 form; A/B aid)
    v = e ? atoi(e) : 1;
  }
  return v;
}

static int pick_block_n(int M, int N, int k_blocks) {
  const int cands[5] = {256, 160, 128, 64, 32};
  const double eff[5] = {1.0, 0.70, 0.70, 0.50, 0.30};
  const int sms = num_sms();
  const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
  double best = 1e30;
  int best_bn = 256;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (bn > 64 && bn / 2 >= N) continue;  // more than half of the tile would be padding
    const int num_n = (N + bn - 1) / bn;
    const long long tiles = 1LL * num_m * num_n;
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * (k_blocks * (bn / eff[i]) + 8.0 * bn + 400.0);
    if (cost < best * 0.98) {
      best = cost;
      best_bn = bn;
    }
  }
  return best_bn;
}

}  // namespace gb

using namespace gb;

extern "C" int gillb200_version(void) { return GILLB200_VERSION; }
extern "C" const char* gillb200_last_error(void) { return gb::err_buf(); }
extern "C" int gillb200_num_sms(void) { return gb::num_sms(); }
extern "C" long long gillb200_launch_count(void) { return gb::g_launch_count; }

// stream-K scratch: [sms x GEMM_EPI_WARPS] arrival flags (padded to 16 KB), then one 128 x 256 fp32 partial tile per SM
static constexpr long long SK_FLAG_BYTES = 16384;
// (also the split-K scratch of the CTA-pair kernel: ksplit fp32 partial copies of the output, at most SPLITK_BYTES)
static constexpr long long SPLITK_BYTES = 64LL << 20;
extern "C" long long gillb200_gemm_streamk_workspace_bytes(void) {
  const long long sk = 1LL * gb::num_sms() * BLOCK_M * 256 * sizeof(float);
  return SK_FLAG_BYTES + (sk > SPLITK_BYTES ? sk : SPLITK_BYTES);
}

extern "C" int gillb200_gemm(const gillb200_gemm_args* a_in, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(a_in != nullptr, "null args");
  gillb200_gemm_args aa = *a_in;  // (a split-K launch redirects the output / epilogue fields of this copy, see below)
  gillb200_gemm_args* a = &aa;
  GB_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "bad GEMM shape M=%d N=%d K=%d", a->M, a->N, a->K);
  GB_CHECK_ARG(a->in_dtype == DT_BF16 || a->in_dtype == DT_F16, "operand dtype must be bf16 or fp16");
  GB_CHECK_ARG(a->out != nullptr && a->a != nullptr && a->b != nullptr, "null operand");
  GB_CHECK_ARG(a->out_dtype >= 0 && a->out_dtype <= 2, "bad out dtype");
  const bool bf16 = a->in_dtype == DT_BF16;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M;
  p.M_out = a->M;
  p.N = a->N;
  int kb_main;
  if (a->conv3x3) {
    GB_CHECK_ARG(a->conv_C % BLOCK_K == 0, "conv3x3 needs C %% 64 == 0 (C=%d)", a->conv_C);
    const int cs = a->conv_stride == 2 ? 2 : 1;
    GB_CHECK_ARG(a->conv_stride == 0 || a->conv_stride == 1 || a->conv_stride == 2, "conv3x3: stride must be 1 or 2");
    GB_CHECK_ARG(a->conv_H % cs == 0 && a->conv_W % cs == 0, "conv3x3: H, W must be multiples of the stride");
    const int W = a->conv_W / cs, H = a->conv_H / cs;  // OUTPUT size: an M tile is a box of output pixels
    GB_CHECK_ARG(a->M == a->conv_B * H * W, "conv3x3: M != B*Ho*Wo");
    GB_CHECK_ARG(a->conv_phase >= 0 && a->conv_phase <= 4, "conv3x3: conv_phase must be 0..4");
    GB_CHECK_ARG(!a->a_cat || (a->gn_scale_shift && a->a_cat_C > 0 && a->a_cat_C < a->conv_C && a->a_cat_C % BLOCK_K == 0 &&
                               (a->conv_C - a->a_cat_C) % BLOCK_K == 0),
                 "a_cat: needs gn_scale_shift and channel counts that are multiples of 64 (conv_C=%d a_cat_C=%d)", a->conv_C,
                 a->a_cat_C);
    GB_CHECK_ARG(!a->gn_scale_shift || (cs == 1 && a->a2_mode == 0 && a->conv_phase == 0),
                 "gn_scale_shift: stride-1 3x3 convs without a K-concatenated second A source");
    GB_CHECK_ARG(a->K == (a->conv_phase ? 4 : 9) * a->conv_C, "conv3x3: K != %d*C", a->conv_phase ? 4 : 9);
    GB_CHECK_ARG(!a->conv_phase || (cs == 1 && a->a2_mode == 0 && !a->residual && !a->rowbias && !a->out_lo && a->ldo == a->N &&
                                    a->act == ACT_NONE),
                 "conv_phase: stride 1, no second A source / residual / row bias / activation, ldo == N");
    int bw, bh, bb;
    if (W >= 128) {
      GB_CHECK_ARG(W % 128 == 0, "conv3x3: W=%d must be a multiple of 128 when >= 128", W);
      bw = 128, bh = 1, bb = 1;
    } else {
      GB_CHECK_ARG(128 % W == 0, "conv3x3: W=%d must divide 128", W);
      bw = W;
      bh = 128 / W;
      if (bh > H) bh = H;
      GB_CHECK_ARG(H % bh == 0, "conv3x3: H=%d not a multiple of the box height %d", H, bh);
      bb = 128 / (bw * bh);
    }
    GB_CHECK_ARG(bw * cs <= 256 && bh * cs <= 256, "conv3x3: strided box exceeds 256");
    const uint64_t dims[4] = {(uint64_t)a->conv_C, (uint64_t)a->conv_W, (uint64_t)a->conv_H, (uint64_t)a->conv_B};
    const uint64_t strides[3] = {(uint64_t)a->conv_C * 2, (uint64_t)a->conv_W * a->conv_C * 2,
                                 (uint64_t)a->conv_H * a->conv_W * a->conv_C * 2};
    // with element strides the box is given in input elements: N loaded elements need a box of N * stride
    const uint32_t box[4] = {64, (uint32_t)(bw * cs), (uint32_t)(bh * cs), (uint32_t)bb};
    const uint32_t es[4] = {1, (uint32_t)cs, (uint32_t)cs, 1};
    int r = encode_tmap(&p.tma_a, a->a, bf16 ? DT_BF16 : DT_F16, 4, dims, strides, box, 128, cs == 2 ? es : nullptr);
    if (r) return r;
    p.conv_stride = cs;
    p.a_mode = A_CONV3X3;
    p.conv_cblocks = a->conv_C / BLOCK_K;
    p.conv_W = W;
    p.conv_H = H;
    p.conv_nt = a->conv_phase ? 4 : 9;
    p.conv_ntx = a->conv_phase ? 2 : 3;
    p.conv_pa = a->conv_phase ? (a->conv_phase - 1) / 2 : 0;
    p.conv_pb = a->conv_phase ? (a->conv_phase - 1) % 2 : 0;
    kb_main = p.conv_nt * p.conv_cblocks;
  } else {
    GB_CHECK_ARG(a->lda % 8 == 0, "lda=%lld must be a multiple of 8 elements", a->lda);
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t strides[1] = {(uint64_t)a->lda * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int r = encode_tmap_16bit(&p.tma_a, a->a, 2, dims, strides, box, bf16);
    if (r) return r;
    p.a_mode = A_PLAIN;
    kb_main = (a->K + BLOCK_K - 1) / BLOCK_K;
  }
  p.kb_split = kb_main;
  p.num_k_blocks = kb_main;
  p.b_kb_wrap = 1 << 30;
  int kb_total_cols = a->K;
  if (a->a2_mode != 0) {
    GB_CHECK_ARG(a->a2 != nullptr && a->k2 > 0, "a2_mode set without a2/k2");
    GB_CHECK_ARG(a->lda2 % 8 == 0, "lda2 must be a multiple of 8 elements");
    GB_CHECK_ARG(a->K % BLOCK_K == 0, "a second A source needs K %% 64 == 0");
    const uint64_t dims[2] = {(uint64_t)a->k2, (uint64_t)a->M};
    const uint64_t strides[1] = {(uint64_t)a->lda2 * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int r = encode_tmap_16bit(&p.tma_a2, a->a2, 2, dims, strides, box, bf16);
    if (r) return r;
    p.num_k_blocks = kb_main + (a->k2 + BLOCK_K - 1) / BLOCK_K;
    if (a->a2_mode == 2) {
      GB_CHECK_ARG(a->k2 == a->K && !a->conv3x3, "split-precision A needs k2 == K and a plain A");
      p.b_kb_wrap = kb_main;
    } else {
      kb_total_cols = a->K + a->k2;
    }
  }
  {
    GB_CHECK_ARG(a->ldb % 8 == 0, "ldb=%lld must be a multiple of 8 elements", a->ldb);
    GB_CHECK_ARG(a->ldb >= kb_total_cols, "ldb=%lld smaller than K=%d", a->ldb, kb_total_cols);
  }

  int bn = a->block_n ? a->block_n : pick_block_n(a->M, a->N, p.num_k_blocks);
  const bool gn_fused = a->conv3x3 && a->gn_scale_shift != nullptr;
  if (gn_fused) {
    // GroupNorm-fused conv: halo-tile pair kernel only; the widest tile gives the two transform warps the most MMA time
    // per halo tile to hide under (9 taps x 640 clocks at 256 x 320 against ~3000 clocks of transform)
    GB_CHECK_ARG(conv_halo_enabled() && a->conv_W % 16 == 0 && a->conv_H % 16 == 0 && a->M % 256 == 0 && a->out_dtype != DT_F32,
                 "gn_scale_shift needs the halo-tile kernel: H, W multiples of 16, M %% 256 == 0, 16-bit output");
    if (!a->block_n) bn = a->N % 320 == 0 ? 320 : a->N % 256 == 0 ? 256 : 128;
  }
  if (a->conv3x3 && a->conv_phase) {
    // upsample-fused phase launch: only the halo-tile pair kernel implements the collapsed tap walk and the strided output
    GB_CHECK_ARG(conv_halo_enabled() && a->conv_W % 16 == 0 && a->conv_H % 16 == 0 && a->M % 256 == 0 && a->out_dtype != DT_F32,
                 "conv_phase needs the halo-tile kernel: H, W multiples of 16, M %% 256 == 0, 16-bit output");
    if (!a->block_n) bn = a->N % 320 == 0 ? (a->conv_C <= 640 ? 160 : 320) : a->N % 256 == 0 ? 256 : 128;
  }
  if (a->act == ACT_GEGLU) {
    GB_CHECK_ARG(a->N % 2 == 0, "GEGLU needs even N");
    // the smem-staged epilogue emits 32-column output panels = 64 accumulator columns under GEGLU
    if (!a->block_n && bn % 64 != 0) bn = bn > 64 ? 128 : 64;
  }
  // CTA-pair (cta_group::2) kernel: 256 x bn tiles, each CTA loads half of the B tile. Used when there are enough
  // 256-row tiles to keep every SM pair busy; small problems keep the 1-CTA kernel (more, smaller tiles).
  // Stream-K candidate (decided first, it needs the 1-CTA kernel): implicit-conv shapes whose whole-tile schedule
  // would leave most of the last wave empty (UNet 16x16 / 8x8 levels: 160 / 40 tiles on 148 SMs). Measured
  // (tools/gpu_sweep_shapes.py): 16x16 C1280 150 -> 111 us, 8x8 C2560 149 -> 102 us; plain short-K linears lose
  // (the fix-up costs more than the idle SMs), so auto mode is limited to convolutions.
  bool want_sk = false;
  static int env_sk = -1;
  if (env_sk < 0) {
    const char* e = getenv("GILLB200_STREAMK");  // "0": never stream-K in auto mode (A/B aid)
    env_sk = e ? atoi(e) : 1;
  }
  // (stream-K partial/flag traffic assumes the previous GEMM of the stream has fully drained: never with PDL overlap)
  if (a->sk_workspace && a->stream_k != 1 && a->stream_k != 3 && (a->stream_k == 2 || env_sk) && a->tile_order != 2 && a->cta_pair != 2 &&
      !pdl_enabled()) {
    if (a->stream_k == 2) {
      want_sk = true;
    } else if (a->conv3x3 && a->out_dtype != DT_F32 && p.num_k_blocks >= 16) {
      if (!a->block_n && a->N % 256 == 0) bn = 256;  // tile quantisation no longer matters: widest tile
      const long long tiles = 1LL * ((a->M + BLOCK_M - 1) / BLOCK_M) * ((a->N + bn - 1) / bn);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      want_sk = static_cast<double>(tiles) / static_cast<double>(waves * num_sms()) < 0.8;
    }
  }
  if ((a->conv3x3 && a->conv_phase) || gn_fused) want_sk = false;
  bool pair = false;
  if (a->cta_pair == 2 || (a->conv3x3 && a->conv_phase) || gn_fused) {
    pair = true;
  } else if (!want_sk && a->cta_pair == 0 && (bn == 128 || bn == 160 || bn == 256)) {
    // measured (tools/gpu_sweep_shapes.py, with the staged epilogue): the pair kernel wins 3..8 % from 10 k-blocks up
    // (K >= 640: M16384 N5120 121 -> 115 us, M4096 N10240 97 -> 91 us) and loses on the K = 320 linears of the 64x64
    // level, whose time is all epilogue (152 vs 170 us)
    const long long tiles2 = 1LL * ((a->M + 255) / 256) * ((a->N + bn - 1) / bn);
    // (3x3 convs that qualify for the halo-tile A path take the pair kernel at any depth: VAE 512x512 C128->128, 18 k-blocks)
    const bool halo_ok = a->conv3x3 && conv_halo_enabled() && p.conv_stride == 1 && a->conv_W % 16 == 0 && a->conv_H % 16 == 0 &&
                         a->M % 256 == 0 && a->a2_mode == 0 && bn >= 128;
    pair = tiles2 >= (num_sms() / 2) * 3 / 4 && (p.num_k_blocks >= (bn == 256 ? 10 : 32) || halo_ok);
  }
  // Wide pair tile (256 x 320, two N = 160 MMAs per K-step, see Gemm2Cfg) for N % 320 == 0. Measured on B200
  // (tools/gpu_conv_bench.py, profiles/r02_conv_wide.log, B = 16): 64x64 C320->320 142 -> 110 us, 32x32 C640->640 134 -> 107,
  // 16x16 C1280->1280 122 (stream-K) -> 101, 16x16 C2560->1280 258 -> 179; all UNet convs of an evaluation 5.97 -> 4.06 ms.
  // It needs enough 256 x 320 tiles for the 74 SM pairs: below ~48 (8x8 level, 16x16 C640) stream-K stays.
  {
    static int env_wide = -1;
    if (env_wide < 0) {
      const char* e = getenv("GILLB200_WIDE");  // "0": never pick the wide tile in auto mode (A/B aid)
      env_wide = e ? atoi(e) : 1;
    }
    const long long tiles_w = 1LL * ((a->M + 255) / 256) * (a->N / 320);
    const bool can_wide = a->N % 320 == 0 && a->act != ACT_GEGLU && a->cta_pair != 1 && a->stream_k != 2;
    const bool auto_wide = a->block_n == 0 && env_wide && p.num_k_blocks >= (a->conv3x3 ? 20 : wide_min_kb()) &&
                           tiles_w >= (num_sms() / 2) * 2 / 3;
    if (can_wide && (a->block_n == 320 || auto_wide)) {
      bn = 320;
      pair = true;
      want_sk = false;
      // With the halo-tile A path (below) the narrow 256 x 160 pair tile no longer pays for re-fetching A per N tile, keeps
      // two accumulator stages (epilogue under the next tile's MMAs) and quantises better (64x64 C320->320: 512 tiles on 74
      // pairs instead of 256). Measured (profiles/r02_conv_halo.log): C320->320 104.9 -> 93.7 us, 64x64 C640->320 177.1 ->
      // 173.0, 32x32 C640->640 92.3 -> 89.5, C320->640 56.5 -> 52.4; from C = 960 up the wide tile wins or ties.
      if (a->block_n == 0 && a->conv3x3 && a->conv_C <= 640 && conv_halo_enabled() && p.conv_stride == 1 && a->conv_W % 16 == 0 &&
          a->conv_H % 16 == 0 && a->M % 256 == 0 && !gn_fused)
        bn = 160;
    }
    // Too few wide tiles for the 74 SM pairs (8x8 level: 16, 16x16 C640: 32): cut K into slices so that tiles x slices
    // fills them, fp32 partials through the scratch buffer, second pass = splitk_reduce_kernel. OPT-IN (stream_k = 3 or
    // GILLB200_SPLITK=1): measured on B200 it LOSES to the 1-CTA stream-K it was meant to replace -- 8x8 C1280->1280, B=16:
    // 104.8 us main launch + 42 us reduce vs 63.6 us (ncu launch list, profiles/r02_splitk_launches.csv). With 64 pairs
    // pulling 4-D halo boxes of the same 2.6 MB activation the per-k-block time rose 4.5x over the 16-pair unsplit run
    // (the aggregate L2 -> SM rate stayed at ~1.1 KB per clock), so more SMs bought nothing.
    static int env_splitk = -1;
    if (env_splitk < 0) {
      const char* e = getenv("GILLB200_SPLITK");
      env_splitk = e ? atoi(e) : 0;
    }
    if (can_wide && a->block_n == 0 && env_wide && (env_splitk || a->stream_k == 3) && a->conv3x3 && a->sk_workspace && a->M % 256 == 0 &&
        tiles_w < (num_sms() / 2) * 2 / 3 && tiles_w >= 8 && a->out_dtype != DT_F32 && a->act == ACT_NONE && a->alpha == 1.f &&
        !a->out_lo && !a->rowstats_out && !a->ln_stats && (!a->residual || a->res_dtype == a->out_dtype) && !a->bias_along_m) {
      int S = static_cast<int>((num_sms() / 2) / tiles_w);
      while (S > 1 && p.num_k_blocks / S < 20) --S;
      while (S > 1 && 1LL * S * a->M * a->N * 4 > SPLITK_BYTES) --S;
      if (S > 1) {
        bn = 320;
        pair = true;
        want_sk = false;
        p.ksplit = S;
        p.kb_per_split = (p.num_k_blocks + S - 1) / S;
        p.M_out = S * a->M;
        // main launch: plain fp32 partial tiles into the scratch buffer; the real epilogue runs in splitk_reduce_kernel
        aa.out = reinterpret_cast<char*>(a_in->sk_workspace) + SK_FLAG_BYTES;
        aa.ldo = a->N;
        aa.out_dtype = DT_F32;
        aa.bias = nullptr;
        aa.rowbias = nullptr;
        aa.residual = nullptr;
        aa.stats_out = nullptr;
      }
    }
    GB_CHECK_ARG(bn != 320 || pair, "block_n 320 exists only as the CTA-pair wide tile (N %% 320 == 0, no GEGLU, no stream-K)");
  }
  if (pair) GB_CHECK_ARG(bn == 64 || bn == 128 || bn == 160 || bn == 256 || bn == 320, "cta_pair needs block_n in {64,128,160,256,320}");
  {
    const uint64_t dims[2] = {(uint64_t)kb_total_cols, (uint64_t)a->N};
    const uint64_t strides[1] = {(uint64_t)a->ldb * 2};
    // a CTA of a pair loads half of each MMA's B rows per box (the wide tile issues two MMAs of N = 160: 80-row boxes)
    const uint32_t box[2] = {BLOCK_K, (uint32_t)(pair ? (bn == 320 ? bn / 4 : bn / 2) : bn)};
    int r = encode_tmap_16bit(&p.tma_b, a->b, 2, dims, strides, box, bf16);
    if (r) return r;
  }
  p.out = a->out;
  p.out_lo = a->out_lo;
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.residual = a->residual;
  p.ldo = a->ldo;
  p.ldr = a->ldr;
  p.ld_rowbias = a->ld_rowbias;
  p.out_dtype = a->out_dtype;
  p.res_dtype = a->res_dtype;
  p.bias_along_m = a->bias_along_m;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.act = a->act;
  p.in_dtype = a->in_dtype;
  p.alpha = a->alpha;
  p.tile_order = a->tile_order;
  {
    const char* dbg = getenv("GILLB200_GEMM_DEBUG");  // measurement aid only (see GemmParams::debug_mode)
    p.debug_mode = dbg ? atoi(dbg) : 0;
  }
  if (a->out_lo) GB_CHECK_ARG(a->out_dtype == DT_BF16, "out_lo requires a bf16 primary output");

  // ---- epilogue selection: smem-staged TMA stores wherever the layout allows, else the direct per-lane epilogue
  {
    static int env_epi = -1, env_warps = 0, env_warps_res = 0;
    if (env_epi < 0) {
      const char* e = getenv("GILLB200_EPI");  // "0" forces the direct epilogue (A/B measurements)
      env_epi = e ? atoi(e) : 1;
      const char* w = getenv("GILLB200_EPI_WARPS");  // tuning aid: epilogue warps of the staged epilogue
      env_warps = w ? atoi(w) : 0;
      if (env_warps != 4 && env_warps != 8 && env_warps != 12 && env_warps != 16) env_warps = 0;
      const char* wr = getenv("GILLB200_EPI_WARPS_RES");  // same, for launches with a residual (3 buffers per warp)
      env_warps_res = wr ? atoi(wr) : env_warps;
      if (env_warps_res != 4 && env_warps_res != 8 && env_warps_res != 12 && env_warps_res != 16) env_warps_res = 0;
    }
    const int esz = a->out_dtype == DT_F32 ? 4 : 2;
    const int n_out = a->act == ACT_GEGLU ? a->N / 2 : a->N;
    bool ok = env_epi != 0 && reinterpret_cast<uintptr_t>(a->out) % 16 == 0 && (a->ldo * esz) % 16 == 0;
    // bf16 hi + lo outputs (split-precision activations of the GILLMapper FFN): the staged form exists (two panels per
    // store) but measured SLOWER than the direct epilogue (M19712 N2048 K512 relu: 266 vs 234 us per launch -- both staging
    // buffers are busy every panel, so each panel waits for the previous pair of stores); opt-in via GILLB200_EPI_LO=1
    if (a->out_lo) {
      static int env_lo = -1;
      if (env_lo < 0) {
        const char* e = getenv("GILLB200_EPI_LO");
        env_lo = e ? atoi(e) : 0;
      }
      ok = ok && env_lo && !a->residual && a->act != ACT_GEGLU && reinterpret_cast<uintptr_t>(a->out_lo) % 16 == 0;
    }
    if (a->residual)
      ok = ok && a->res_dtype == a->out_dtype && reinterpret_cast<uintptr_t>(a->residual) % 16 == 0 &&
           (a->ldr * esz) % 16 == 0;
    if (a->act == ACT_GEGLU && bn % 64 != 0) ok = false;  // explicitly requested odd tile width
    if (ok) {
      const uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)p.M_out};
      const uint32_t box[2] = {EPI_PANEL_COLS, 32};
      const uint64_t so[1] = {(uint64_t)a->ldo * esz};
      int r = encode_tmap(&p.tma_out, a->out, a->out_dtype, 2, dims, so, box, esz == 4 ? 128 : 64, nullptr);
      if (r) return r;
      if (a->residual) {
        const uint64_t sr[1] = {(uint64_t)a->ldr * esz};
        r = encode_tmap(&p.tma_res, a->residual, a->res_dtype, 2, dims, sr, box, esz == 4 ? 128 : 64, nullptr);
        if (r) return r;
      }
      if (a->out_lo) {
        r = encode_tmap(&p.tma_out_lo, a->out_lo, DT_BF16, 2, dims, so, box, 64, nullptr);
        if (r) return r;
      }
      p.epi_tma = 1;
      p.epi_variant = EV_GENERIC;
      const bool plain = a->out_dtype == DT_F16 && a->alpha == 1.f && a->bias && !a->bias_along_m &&
                         reinterpret_cast<uintptr_t>(a->bias) % 16 == 0;
      if (plain && a->act == ACT_GEGLU && a->N % 64 == 0 && !a->rowbias && !a->residual) {
        p.epi_variant = EV_GEGLU;
      } else if (plain && a->act == ACT_NONE && a->N % 32 == 0) {
        if (a->rowbias && !a->residual && a->ld_rowbias % 4 == 0 && reinterpret_cast<uintptr_t>(a->rowbias) % 16 == 0)
          p.epi_variant = EV_BIAS_ROWBIAS;
        else if (a->residual && !a->rowbias)
          p.epi_variant = EV_BIAS_RES;
        else if (!a->residual && !a->rowbias)
          p.epi_variant = EV_BIAS;
      }
      if (a->ln_stats) {  // folded LayerNorm: only the two specialised consumers exist
        const bool lnok = plain && a->ln_cs && a->ln_C > 0 && a->ln_C % 32 == 0 && !a->rowbias && !a->residual &&
                          reinterpret_cast<uintptr_t>(a->ln_cs) % 16 == 0 && a->K == a->ln_C && a->a2_mode == 0;
        if (lnok && a->act == ACT_GEGLU && a->N % 64 == 0)
          p.epi_variant = EV_LN_GEGLU;
        else if (lnok && a->act == ACT_NONE && a->N % 32 == 0)
          p.epi_variant = EV_LN_BIAS;
        else
          return set_err(-EINVAL, "ln_stats: needs fp16 output, aligned bias and ln_cs, K == ln_C, act none/GEGLU, no residual");
        p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
        p.ln_cs = a->ln_cs;
        p.ln_C = a->ln_C;
        p.ln_np = a->ln_C / 32;
        p.ln_ld = a->M;
        p.ln_eps = a->ln_eps;
      }
      if (a->rowstats_out) {
        if (p.epi_variant != EV_BIAS && p.epi_variant != EV_BIAS_RES)
          return set_err(-EINVAL, "rowstats_out needs the bias (+ residual) staged epilogue: fp16 output, N %% 32 == 0");
        p.rowstats_out = reinterpret_cast<float2*>(a->rowstats_out);
        p.rowstats_ld = a->M;
      }
      {
        static int env_var = -2;
        if (env_var == -2) {
          const char* e = getenv("GILLB200_EPI_GENERIC");  // "1": always the all-runtime staged epilogue (A/B)
          env_var = e ? atoi(e) : 0;
        }
        if (env_var && !a->ln_stats && !a->rowstats_out) p.epi_variant = EV_GENERIC;
      }
      {
        static int env_stg = -1;
        if (env_stg < 0) {
          const char* e = getenv("GILLB200_EPI_STG");  // staged panels written back by per-lane 16-byte stores (see epilogue_warp_tma)
          env_stg = e ? atoi(e) : 0;
        }
        p.epi_stg = env_stg && esz == 2 && p.epi_variant != EV_GENERIC && (env_stg > 1 || p.num_k_blocks <= 8);
      }
      {
        static int env_nb = -1;
        if (env_nb < 0) {
          const char* e = getenv("GILLB200_EPI_RES_NBUF");  // staging buffers per epilogue warp with a residual: 3 or 4
          // measured (profiles/r02_gemm_res_nbuf.log): residual two panels ahead (4) = one ahead (3) on every + residual linear
          // and conv (35.0 vs 35.4, 24.5 vs 24.8, 47.1 vs 49.0 us; convs 4.66 vs 4.73 ms per evaluation): 3 keeps the smem
          env_nb = e ? atoi(e) : 3;
          if (env_nb != 3 && env_nb != 4) env_nb = 3;
        }
        p.epi_nbuf = a->residual ? (esz == 2 ? env_nb : 3) : 2;  // (fp32 panels are 4 KB: a fourth buffer would cost ring stages)
      }
      p.epi_buf_bytes = 32 * EPI_PANEL_COLS * esz;
      // as many epilogue warps as leave a >= 3-deep operand ring (fp32 panels are twice as large)
      // measured (tools/gpu_sweep_shapes.py): 8 warps + a deeper operand ring beat 16 warps on every plain shape
      p.epi_warps = GEMM_EPI_WARPS;
      if (a->residual ? env_warps_res : env_warps) p.epi_warps = a->residual ? env_warps_res : env_warps;
      if (p.epi_warps > GEMM_EPI_WARPS) p.epi_warps = GEMM_EPI_WARPS;
      const int stage_bytes = (BLOCK_M + (pair ? bn / 2 : bn)) * BLOCK_K * 2;
      while (p.epi_warps > 4 &&
             (SMEM_BUDGET - 2048 - p.epi_warps * p.epi_nbuf * p.epi_buf_bytes) / stage_bytes < 3)
        p.epi_warps -= 4;
    }
  }

  // ---- stream-K (see want_sk above): only the 1-CTA kernel with the staged epilogue
  GB_CHECK_ARG(p.epi_tma || (!a->ln_stats && !a->rowstats_out), "ln_stats / rowstats_out need the staged epilogue");
  if (a->stats_out) {
    GB_CHECK_ARG(p.epi_tma && a->act != ACT_GEGLU && a->out_dtype != DT_F32 && a->M % 32 == 0 && a->N % 32 == 0,
                 "stats_out needs the staged epilogue: 16-bit output, 16-byte aligned rows, M %% 32 == 0, N %% 32 == 0, no GEGLU");
    p.stats_out = reinterpret_cast<float2*>(a->stats_out);
  }
  if (want_sk && !pair && p.epi_tma) {
    const long long tiles = 1LL * ((a->M + BLOCK_M - 1) / BLOCK_M) * ((a->N + bn - 1) / bn);
    const int sms = num_sms();
    const long long total = tiles * p.num_k_blocks;
    const bool fits = bn <= 256 && total < (1LL << 30) && (total + sms - 1) / sms >= 4 &&
                      p.num_k_blocks <= 8 * ((total + sms - 1) / sms);  // <= ~9 partials per tile
    if (fits) {
      p.sk_per = 1;  // finalised (units per CTA) by launch_gemm
      p.sk_flags = reinterpret_cast<int*>(a->sk_workspace);
      p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(a->sk_workspace) + SK_FLAG_BYTES);
      p.tile_order = 1;  // m-fastest: neighbouring CTAs share weight tiles
    }
  }

  // ---- B-stationary 1-CTA mode (see GemmParams::b_resident): short-K plain linears with many M tiles per CTA and <= 4 N tiles
  {
    static int env_bres = -1;
    if (env_bres < 0) {
      // OPT-IN ("1"): measured on B200 (profiles/r02_gemm_bres.log) it does not pay -- M65536 N320 K384 +res 49.9 us with the
      // resident weight tile vs 45.8 without, K320 34.4 vs 33.3: these linears are bound by the activation stream and the
      // epilogue, the weight tiles were L2 hits all along
      const char* e = getenv("GILLB200_BRES");
      env_bres = e ? atoi(e) : 0;
    }
    const int num_m = (a->M + BLOCK_M - 1) / BLOCK_M, num_n = (a->N + bn - 1) / bn;
    if (env_bres && !pair && p.sk_per == 0 && !a->conv3x3 && a->a2_mode == 0 && p.epi_tma && p.num_k_blocks <= 8 && num_n <= 4 &&
        a->tile_order != 1 && bn >= 128 && 1LL * num_m * num_n >= 3LL * num_sms())
      p.b_resident = 1;
  }
  if (p.ksplit > 1) {
    GB_CHECK_ARG(p.epi_tma && pair && bn == 320, "split-K needs the staged epilogue of the wide pair kernel");
    int r = launch_gemm2<320>(p, stream);
    if (r) return r;
    const gillb200_gemm_args* o = a_in;
    dim3 grid((o->N + 255) / 256, (o->M + 31) / 32);
    GB_CUDA(launch_pdl(splitk_reduce_kernel, grid, dim3(256), 0, stream, reinterpret_cast<const float*>(aa.out), p.ksplit, o->M,
                       o->N, o->bias, o->rowbias, static_cast<long long>(o->ld_rowbias), o->rows_per_group > 0 ? o->rows_per_group : 1,
                       reinterpret_cast<const uint16_t*>(o->residual), static_cast<long long>(o->ldr),
                       reinterpret_cast<uint16_t*>(o->out), static_cast<long long>(o->ldo), o->out_dtype == DT_BF16 ? 1 : 0,
                       reinterpret_cast<float2*>(o->stats_out)));
    GB_COUNT_LAUNCH(1);
    return 0;
  }
  // ---- halo-tile implicit conv (A_CONV3X3_HALO, see gemm_sm100.cuh): stride-1 3x3 convs on the CTA-pair kernel with the
  // staged epilogue, W and H multiples of 16 (an M tile = a 16 x 8 pixel block). One halo tile per 64-channel block feeds
  // all nine taps, so the A traffic of a CTA drops 9 x 16 KB -> 23 KB per channel block.
  {
    const int env_halo = conv_halo_enabled();
    if (env_halo && pair && a->conv3x3 && p.conv_stride == 1 && p.epi_tma && p.ksplit <= 1 && a->a2_mode == 0 &&
        a->conv_W % 16 == 0 && a->conv_H % 16 == 0 && !a->out_lo && (bn == 128 || bn == 160 || bn == 256 || bn == 320) &&
        a->ldo == (a->act == ACT_GEGLU ? a->N / 2 : a->N) && (!a->residual || a->ldr == a->ldo) && a->M % 256 == 0) {
      const int esz = a->out_dtype == DT_F32 ? 4 : 2;
      const int n_out = a->act == ACT_GEGLU ? a->N / 2 : a->N;
      {
        const int c1 = a->a_cat ? a->a_cat_C : 0, c0 = a->conv_C - c1;  // channels of the first / second (cat) source
        const uint32_t box[4] = {64, 10, 18, 1};
        const uint64_t dims[4] = {(uint64_t)c0, (uint64_t)a->conv_W, (uint64_t)a->conv_H, (uint64_t)a->conv_B};
        const uint64_t strides[3] = {(uint64_t)c0 * 2, (uint64_t)a->conv_W * c0 * 2, (uint64_t)a->conv_H * a->conv_W * c0 * 2};
        int r = encode_tmap(&p.tma_a, a->a, bf16 ? DT_BF16 : DT_F16, 4, dims, strides, box, 128, nullptr);
        if (r) return r;
        p.halo_c0_blocks = c0 / BLOCK_K;
        if (c1) {
          const uint64_t dims1[4] = {(uint64_t)c1, (uint64_t)a->conv_W, (uint64_t)a->conv_H, (uint64_t)a->conv_B};
          const uint64_t str1[3] = {(uint64_t)c1 * 2, (uint64_t)a->conv_W * c1 * 2, (uint64_t)a->conv_H * a->conv_W * c1 * 2};
          r = encode_tmap(&p.tma_a2, a->a_cat, bf16 ? DT_BF16 : DT_F16, 4, dims1, str1, box, 128, nullptr);
          if (r) return r;
        }
        p.gn_ss = a->gn_scale_shift;
        p.gn_silu = a->gn_silu;
        p.gn_C = a->conv_C;
      }
      {
        const uint64_t dims[3] = {(uint64_t)n_out, (uint64_t)a->conv_W, (uint64_t)a->conv_B * a->conv_H};
        const uint32_t box[3] = {EPI_PANEL_COLS, 8, 4};
        // phase launch: this launch owns pixels (2i + pa, 2j + pb) of the full-resolution [B, 2H, 2W, N] tensor
        const uint64_t ps = a->conv_phase ? 2 : 1;
        const uint64_t so[2] = {ps * a->ldo * esz, ps * ps * a->conv_W * a->ldo * esz};
        char* obase = reinterpret_cast<char*>(a->out) +
                      (a->conv_phase ? (static_cast<uint64_t>(p.conv_pa) * 2 * a->conv_W + p.conv_pb) * a->ldo * esz : 0);
        int r = encode_tmap(&p.tma_out, obase, a->out_dtype, 3, dims, so, box, esz == 4 ? 128 : 64, nullptr);
        if (r) return r;
        if (a->conv_phase) {
          p.stats_hw = a->conv_H * a->conv_W;
          p.stats_phase = a->conv_phase - 1;
        }
        if (a->residual) {
          const uint64_t sr[2] = {(uint64_t)a->ldr * esz, (uint64_t)a->conv_W * a->ldr * esz};
          r = encode_tmap(&p.tma_res, a->residual, a->res_dtype, 3, dims, sr, box, esz == 4 ? 128 : 64, nullptr);
          if (r) return r;
        }
      }
      p.a_mode = A_CONV3X3_HALO;
      if (gn_fused) {
        p.epi_warps = 4;  // warps 4-7; warps 8 and 9 join warps 2 and 3 as transform warps (one per scheduler)
        switch (bn) {
          case 128: return launch_gemm2<128, true, true>(p, stream);
          case 160: return launch_gemm2<160, true, true>(p, stream);
          case 320: return launch_gemm2<320, true, true>(p, stream);
          default: return launch_gemm2<256, true, true>(p, stream);
        }
      }
      switch (bn) {
        case 128: return launch_gemm2<128, true>(p, stream);
        case 160: return launch_gemm2<160, true>(p, stream);
        case 320: return launch_gemm2<320, true>(p, stream);
        default: return launch_gemm2<256, true>(p, stream);
      }
    }
  }
  if (p.debug_mode == 3 && a->sk_workspace && p.sk_per == 0)  // wait-time trace of CTA 0 (measurement aid, see gemm_mma)
    p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(a->sk_workspace) + SK_FLAG_BYTES);
  GB_CHECK_ARG(!(a->conv3x3 && a->conv_phase), "conv_phase: shape does not qualify for the halo-tile pair kernel (bn=%d)", bn);
  GB_CHECK_ARG(!gn_fused, "gn_scale_shift: shape does not qualify for the halo-tile pair kernel (bn=%d)", bn);
  if (pair) {
    switch (bn) {
      case 64: return launch_gemm2<64>(p, stream);
      case 128: return launch_gemm2<128>(p, stream);
      case 160: return launch_gemm2<160>(p, stream);
      case 320: return launch_gemm2<320>(p, stream);
      default: return launch_gemm2<256>(p, stream);
    }
  }
  switch (bn) {
    case 32: return launch_gemm<32>(p, stream);
    case 64: return launch_gemm<64>(p, stream);
    case 128: return launch_gemm<128>(p, stream);
    case 160: return launch_gemm<160>(p, stream);
    case 256: return launch_gemm<256>(p, stream);
    default: return set_err(-EINVAL, "unsupported block_n %d", bn);
  }
}
