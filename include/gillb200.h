/* libgillb200.so -- C ABI of the B200-native (sm_100a) GILL image-emission hot path.
 *
 * The reference (kohjingyu/gill @ 4600b71) is pure Python and has no FFI of its own; its seams are Python
 * attribute calls into torch / transformers / diffusers. Each entry point below names the reference call site
 * whose arithmetic it replaces. The Python host layer (gill_b200/*.py) binds these with ctypes and mirrors the
 * reference's own class/function surface on top.
 *
 * Conventions
 *   - every function returns 0 on success, a negative errno-style code on failure (-EINVAL bad shape/alignment,
 *     -EIO CUDA error); gillb200_last_error() returns a thread-local description of the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller, 16-byte aligned, row-major / innermost-contiguous.
 *   - `stream` is a cudaStream_t passed as void*. No function synchronises the stream or allocates device memory,
 *     so every call can be captured into a CUDA graph.
 *   - dtype codes: 0 = bf16, 1 = fp16, 2 = fp32.
 */
#ifndef GILLB200_H_
#define GILLB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GILLB200_VERSION 100

enum { GILLB200_BF16 = 0, GILLB200_F16 = 1, GILLB200_F32 = 2 };
enum { GILLB200_ACT_NONE = 0, GILLB200_ACT_RELU = 1, GILLB200_ACT_GELU = 2, GILLB200_ACT_SILU = 3, GILLB200_ACT_GEGLU = 4,
       GILLB200_ACT_QUICK_GELU = 5 /* x * sigmoid(1.702 x): CLIP's activation */ };

int gillb200_version(void);
const char* gillb200_last_error(void);
int gillb200_num_sms(void);
/* number of kernel launches this library has issued so far in this process */
long long gillb200_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05.mma + TMEM accumulators, TMA-fed).
 *
 *   out[M,N] = act( alpha * (A[M,K] . B[N,K]^T) + bias + rowbias ) + residual
 *
 * Replaces every nn.Linear / nn.Conv2d the hot path reaches through third-party modules:
 *   gill/layers.py:42-44 (TextFcLayer fc / tfm linears / model), gill/models.py:465 (OPTForCausalLM linears),
 *   gill/models.py:730 -> gill/custom_sd.py:633-638 (UNet convs / linears), custom_sd.py:388 (VAE decoder).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gillb200_gemm_args {
  /* A operand: [M, K] with row stride lda (elements). With conv3x3 != 0, `a` is an NHWC activation
   * [conv_B, conv_H, conv_W, conv_C] and the GEMM is the 3x3/pad-1 convolution (stride conv_stride): M = B*H*W, K = 9*C,
   * B operand laid out [N, 9*C] with k = (ky*3+kx)*C + c. */
  const void* a;
  long long lda;
  /* optional second A source [M, k2], row stride lda2:
   *   a2_mode 1: K-concatenation -- out = [A | A2] . B^T, B has K + k2 columns
   *   a2_mode 2: split precision  -- out = (A + A2) . B^T, A2 is the bf16 residue of a higher-precision A; k2 == K */
  const void* a2;
  long long lda2;
  int k2;
  int a2_mode;
  const void* b; /* [N, Kb] row stride ldb */
  long long ldb;
  int M, N, K;
  int in_dtype; /* operand dtype of a, a2, b: bf16 or fp16 */
  int conv3x3;
  int conv_B, conv_H, conv_W, conv_C;
  /* epilogue */
  void* out;
  long long ldo;
  int out_dtype;
  void* out_lo;      /* optional: bf16 residue of the (bf16) output, same layout as out */
  const float* bias; /* [N], or [M] when bias_along_m */
  int bias_along_m;
  const float* rowbias; /* [M / rows_per_group, ld_rowbias] added per row group (per-sample time embedding) */
  long long ld_rowbias;
  int rows_per_group;
  const void* residual; /* [M, N_out] row stride ldr */
  long long ldr;
  int res_dtype;
  int act;   /* GEGLU: B rows interleaved (value, gate) pairs; N_out = N/2 */
  float alpha;
  int block_n; /* 0 = auto; else one of 32, 64, 128, 160, 256, or 320 (CTA-pair wide tile: N % 320 == 0) */
  int tile_order; /* 0 = auto; 1 = M-fastest tile order; 2 = N-inner (all N tiles of an M block back to back) */
  int cta_pair; /* 0 = auto; 1 = force the 1-CTA kernel; 2 = force the CTA-pair (tcgen05 cta_group::2, 256-row tile) kernel */
  /* optional stream-K scratch: gillb200_gemm_streamk_workspace_bytes() bytes of device memory, ZERO before its first
   * use (the arrival flags re-arm themselves) and not shared between concurrently running GEMMs. When given, shapes
   * whose tile count fills the SMs badly (the UNet's 8x8 / 16x16 levels) split their K range evenly over all SMs. */
  void* sk_workspace;
  int stream_k; /* 0 = auto (when sk_workspace is given); 1 = never; 2 = always (if the kernel variant supports it);
                 * 3 = split-K over the CTA-pair wide tile + reduce pass (small-M N % 320 == 0 convs; opt-in, see gemm.cu) */
  /* optional: GroupNorm statistics of the output, float [M/32, N, 2] = per 32-row slab and column {sum, sum of squares} of
   * the rounded 16-bit output values (consumed by gillb200_groupnorm_from_stats). Needs a 16-bit output, M % 32 == 0,
   * N % 32 == 0, 16-byte aligned rows and no GEGLU. A slab is 32 rows of ONE sample: 32 consecutive rows, or (3x3
   * convolutions that run on the halo-tile kernel) a 4 x 8 pixel patch; slabs of a sample occupy that sample's M/32-range. */
  void* stats_out;
  /* LayerNorm folded into the NEXT GEMM (UNet transformer blocks).
   * rowstats_out (producer, optional): float [N/32, M, 2] = per 32-column panel and row {sum, sumsq} of the output; needs
   *   the plain bias (+ residual) staged epilogue (fp16 output, N % 32 == 0).
   * ln_stats / ln_cs (consumer, optional, both or none): ln_stats = the producer's rowstats_out of THIS GEMM's A operand
   *   (ln_C = its column count), ln_cs[N] = column sums of the weights, which must already carry the LayerNorm scale
   *   (W' = W diag(gamma)); bias must be b + W beta. The epilogue then computes
   *   rstd * (A W'^T - mean * ln_cs) + bias  ==  LayerNorm(A) W^T + b.  act: none or GEGLU; no residual / rowbias. */
  void* rowstats_out;
  const void* ln_stats;
  const float* ln_cs;
  int ln_C;
  float ln_eps;
  int conv_stride; /* conv3x3 only: 0/1 = stride 1; 2 = stride 2 (conv_H/conv_W stay the INPUT size, M = B*(H/2)*(W/2));
                    * the A tensor map then walks the input with TMA element strides, no im2col buffer */
  int conv_phase;  /* conv3x3 only, 0 = plain. 1..4 = phase (a, b) = ((conv_phase - 1) / 2, (conv_phase - 1) % 2) of
                    * "nearest-neighbour 2x upsample, then 3x3 / pad 1 convolution" (UNet / VAE up blocks, diffusers
                    * Upsample2D behind gill/custom_sd.py:633 and :387) computed on the LOW-RESOLUTION input: output pixel
                    * (2i + a, 2j + b) only ever sees the 2 x 2 low-res pixels (i + a - 1 .. i + a, j + b - 1 .. j + b), so
                    * the nine taps collapse to four with pre-summed weights -- 4/9 of the flops and no upsampled tensor.
                    * a = NHWC [conv_B, conv_H, conv_W, conv_C] (low resolution), b = [N, 4 * conv_C] with
                    * k = (u * 2 + v) * C + c for low-res offset (a - 1 + u, b - 1 + v), M = B*H*W, K = 4*C,
                    * out = BASE of the full-resolution NHWC tensor [conv_B, 2*conv_H, 2*conv_W, N] (ldo = N): the launch writes
                    * its quarter of the pixels. stats_out (optional) = the full-resolution tensor's [4*M/32, N, 2] buffer.
                    * Needs the halo-tile CTA-pair kernel: conv_H, conv_W multiples of 16, M % 256 == 0, >= 55 pair tiles. */
  /* conv3x3 only, optional: GroupNorm(+SiLU) of the INPUT applied on the fly (the resnet pattern norm -> SiLU -> conv of
   * the UNet / VAE, gill/custom_sd.py:633): gn_scale_shift = float [conv_B, 2, conv_C], y = x * scale + shift per sample and
   * channel (from gillb200_groupnorm_scale_shift), gn_silu != 0 applies SiLU; zero padding is applied AFTER the normalisation,
   * as nn.Conv2d does. The raw activation is what `a` points at -- the normalised tensor is never written. a_cat / a_cat_C:
   * optional second NHWC source [conv_B, conv_H, conv_W, a_cat_C] holding the LAST a_cat_C of the conv_C input channels
   * (torch.cat([h, skip], 1) of the up blocks without the concat). Halo-tile CTA-pair kernel only (see conv_phase). */
  const float* gn_scale_shift;
  int gn_silu;
  const void* a_cat;
  int a_cat_C;
} gillb200_gemm_args;

long long gillb200_gemm_streamk_workspace_bytes(void);

int gillb200_gemm(const gillb200_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Retrieval: fused cosine-similarity GEMM + top-k  (gill/models.py:676-683).
 *
 *   scores[q, n] = sum_d q[q, d] * bank[n, d]  (bf16 operands, fp32 accumulate; never materialised)
 *   scores[q, n] -= 1000 for every global row index listed in exclude_idx   (models.py:679-680)
 *   out = top-K per query, value descending, ties -> lowest global index; out_idx = index_base + local row.
 *
 * bank: [n_local, d] bf16 (already normalised and scaled as in models.py:895-900), q: [Q, d] bf16, K <= 16.
 * exclude_idx: [Q or 1, n_exclude] int64 global row ids on the device (entries < 0 are unused slots); exclude_ld = elements
 *   between the lists of consecutive queries, 0 = one list shared by all queries (the reference keeps one `seen_image_idx`
 *   list per conversation, models.py:679; batched callers pass one list per query).
 * `workspace` must hold gillb200_topk_workspace_bytes(Q, n_local) bytes of device memory.
 * Q <= 4 (the reference's call shape is Q = 1, K = 3) takes a bank-streaming kernel bound by HBM; larger batches the
 * tcgen05 kernel bound by the tensor cores.
 * gillb200_topk_merge merges R candidate lists [R, Q, Kc] (e.g. one per GPU after the NCCL all-gather);
 * gillb200_topk_merge_strided does the same over lists that sit inside a packed exchange buffer: list r's values start at
 * cand_val + r * r_stride_val, its indices at cand_idx + r * r_stride_idx, query q at + q * q_stride (all in elements).
 * ------------------------------------------------------------------------------------------------------------- */
long long gillb200_topk_workspace_bytes(int Q, long long n_local);
int gillb200_topk_scores(const void* bank, long long n_local, int d, long long ld_bank, const void* q, int Q,
                         long long ldq, int K, long long index_base, const long long* exclude_idx, int n_exclude,
                         long long exclude_ld, void* workspace, float* out_val, long long* out_idx, void* stream);
int gillb200_topk_merge(const float* cand_val, const long long* cand_idx, int R, int Q, int Kc, int K, float* out_val,
                        long long* out_idx, void* stream);
int gillb200_topk_merge_strided(const float* cand_val, long long r_stride_val, const long long* cand_idx,
                                long long r_stride_idx, long long q_stride, int R, int Q, int Kc, int K, float* out_val,
                                long long* out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused multi-head attention  out = softmax(scale * Q K^T [+ causal / length mask]) V   (flash style, tcgen05).
 *
 * Replaces diffusers' Attention (UNet self/cross attention, gill/custom_sd.py:633-638) and OPTAttention
 * (gill/models.py:465). q: [B, Lq, *], k/v: [B, Lk, *], out: [B, Lq, *]; head h occupies columns
 * [h*hd_pad, (h+1)*hd_pad) of each row (or a tighter pitch, see head_stride), hd_pad in {64, 128, 192}; columns beyond the true head dim must be zero in
 * q, k and v (except v's optional ones column, see ones_col) (the projection weights are zero-padded at load time). Strides are in elements.
 * causal: key j is visible to query i iff j <= i + causal_offset. kv_lens: optional per-batch key count (device).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gillb200_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  long long ldq, ldk, ldv, ldo;
  long long q_bstride, k_bstride, v_bstride, o_bstride;
  const int* kv_lens;
  int B, H, Lq, Lk, hd_pad;
  int causal, causal_offset;
  int dtype;
  float scale;
  int ones_col; /* 0 = unused. > 0: v[..., h*stride + ones_col] == 1.0 for every key (a spare padding column), which
                 * lets the kernel take the softmax denominator from the P.V product (fp16 only) */
  int head_stride; /* 0 = hd_pad. Otherwise the column pitch of the heads in q/k/v/out (multiple of 16, <= hd_pad): head h
                    * occupies columns [h*head_stride, (h+1)*head_stride); hd_pad only selects the kernel's tile width.
                    * SD-1.5's head dims 40 / 80 / 160 are stored at pitch 48 / 96 / 176 instead of 64 / 128 / 192. */
} gillb200_attn_args;

int gillb200_attention(const gillb200_attn_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Normalisation / softmax (fp32 statistics, 16-byte vectorised).
 *   layernorm : F.layer_norm over the last dim; optional bf16 residue output (split-precision GEMM operand).
 *   groupnorm : nn.GroupNorm (+ optional SiLU) over an NHWC tensor that may be the channel concatenation of two
 *               tensors (UNet up blocks, cat([h, skip])); writes the normalised concatenation. `workspace` holds
 *               gillb200_groupnorm_workspace_bytes(B, G) bytes and must be ZERO before its first use (the kernel's
 *               arrival counters reset themselves afterwards).
 * Replace the norms inside nn.Transformer (gill/layers.py:20-22), OPT (gill/models.py:465) and the UNet / VAE
 * (gill/custom_sd.py:633-638, :388).
 * ------------------------------------------------------------------------------------------------------------- */
int gillb200_layernorm(const void* x, long long ldx, int in_dtype, const float* w, const float* b, float eps, int rows,
                       int C, void* out, long long ldo, int out_dtype, void* out_lo, void* stream);
long long gillb200_groupnorm_workspace_bytes(int B, int G);
int gillb200_groupnorm(const void* x0, int C0, const void* x1, int C1, int dtype, int B, int HW, int G, const float* w,
                       const float* b, float eps, int silu, void* out, int out_dtype, void* workspace, void* stream);
/* GroupNorm(+SiLU) whose statistics come from the producing GEMMs' stats_out buffers (stats1 only with a second source):
 * one small reduction per (sample, group) + the elementwise pass -- no statistics read of the activation. */
/* Only the (sample, group) reduction of the producers' statistics: scale_shift [B, 2, C0 + C1] for a consumer that applies
 * the normalisation itself (gillb200_gemm_args::gn_scale_shift). */
int gillb200_groupnorm_scale_shift(int C0, const void* stats0, int C1, const void* stats1, int B, int HW, int G,
                                   const float* w, const float* b, float eps, float* scale_shift, void* stream);
int gillb200_groupnorm_from_stats(const void* x0, int C0, const void* stats0, const void* x1, int C1, const void* stats1,
                                  int dtype, int B, int HW, int G, const float* w, const float* b, float eps, int silu,
                                  void* out, int out_dtype, void* workspace, void* stream);
int gillb200_softmax_rows(const void* x, long long ldx, int in_dtype, float scale, long long rows, int n, void* out,
                          long long ldo, int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Small kernels around the tensor-core path.
 *   gather_add_rows : out[i] = (x ? x[i] : 0) + table[idx[i] + idx_offset]  (embedding / learned-position lookup,
 *                     gill/models.py:620, :529, :709 and OPTLearnedPositionalEmbedding behind models.py:465)
 *   upsample2x      : nearest 2x, NHWC (UNet / VAE up blocks)
 *   im2col3x3       : explicit patches for the stride-2 downsamplers and the 4-channel conv_in
 *   plms_step       : classifier-free guidance + PNDM/PLMS update fused (gill/custom_sd.py:641-646)
 *   image_to_u8     : (x/2+0.5).clamp(0,1) -> uint8 NHWC (gill/custom_sd.py:389-391 + numpy_to_pil)
 *   l2norm_rows     : x / ||x|| (gill/models.py:674)
 *   cast_add        : out = cast(x + y[i % y_period]) with optional bf16 residue (gill/layers.py:32 `x + input_embs`)
 *   attn_small_f32  : fp32 attention for the GILLMapper's short sequences (nn.MultiheadAttention, layers.py:43)
 *   tap_sum3x3      : second half of a 3x3 convolution with very few output channels (UNet conv_out 320 -> 4,
 *                     gill/custom_sd.py:633 `self.unet(...)`'s last layer; VAE conv_out 128 -> 3, custom_sd.py:387):
 *                     y[pixel, tap * Cout + o] holds the per-tap products W_tap x (one plain GEMM that reads the
 *                     activation ONCE instead of nine shifted times), out[b,h,w,o] = bias[o] + sum over the nine taps of
 *                     y at the tap's neighbour pixel (zero outside the image)
 * ------------------------------------------------------------------------------------------------------------- */
int gillb200_gather_add_rows(const void* x, const void* table, const long long* idx, long long idx_offset,
                             long long rows, int D, int dtype, void* out, void* stream);
int gillb200_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream);
int gillb200_im2col3x3(const void* x, int B, int H, int W, int C, int stride, void* out, long long ld_out, void* stream);
int gillb200_plms_step(const void* eps_pair, int eps_dtype, float guidance, float* ets, int head, int mode,
                       float c_sample, float c_eps, float* latents, float* cur_sample, void* lat16_pair, int lat16_dtype,
                       long long n, void* stream);
int gillb200_image_to_u8(const void* x, int dtype, long long pixels, int ldx, int channels, void* out, void* stream);
int gillb200_tap_sum3x3(const float* y, long long ldy, int B, int H, int W, int Cout, const float* bias, void* out,
                        int out_dtype, long long ldo, void* stream);
/* CLIP pre-processing of generated images for the re-rank step (gill/models.py:733-737 `img.resize((224,224))` + the HF
 * feature extractor, gill/utils.py:117-119): uint8 NHWC [B,H,W,3] -> PIL-exact bicubic resize to S x S (8-bit two-pass
 * fixed-point ImagingResample) -> /255 -> (x - mean)/std -> NCHW [B,3,S,S] in out_dtype. mean3 / std3 are HOST pointers.
 * resized_u8 (optional, device) receives the intermediate uint8 [B,S,S,3] image. */
int gillb200_clip_preprocess_u8(const void* img, int B, int H, int W, int S, const float* mean3, const float* std3,
                                void* out, int out_dtype, void* resized_u8, void* stream);
/* Same resample, general geometry: resize to RH x RW, then take the S x S window at (top, left) of the resized image. With
 * RH / RW = (shortest edge -> S, aspect kept) and the centred window this is the HF CLIP feature extractor applied to image
 * PROMPTS (resize shortest edge + centre crop: gill/utils.py:117-119, gill/models.py:608, scripts/extract_img_embs.py:37). */
int gillb200_clip_preprocess_u8_crop(const void* img, int B, int H, int W, int RH, int RW, int top, int left, int S,
                                     const float* mean3, const float* std3, void* out, int out_dtype, void* resized_u8,
                                     void* stream);
int gillb200_l2norm_rows(const float* x, long long ldx, int rows, int n, void* out, long long ldo, int out_dtype,
                         void* stream);
int gillb200_cast_add(const void* x, int x_dtype, const void* y, int y_dtype, long long y_period, void* out,
                      int out_dtype, void* out_lo, long long n, void* stream);
/* out[p,:] = W x[p,:] + b for <= 8 channels (VAE post_quant_conv folded with the 1/0.18215 scale, custom_sd.py:386) */
int gillb200_channel_mix(const float* x, int cin, const float* w, const float* b, int cout, long long n, void* out,
                         int out_dtype, void* stream);
int gillb200_attn_small_f32(const float* q, long long ldq, long long q_bs, const float* k, long long ldk, long long k_bs,
                            const float* v, long long ldv, long long v_bs, int B, int H, int hd, int Lq, int Lk,
                            float scale, void* out, long long ldo, long long o_bs, int out_dtype, void* out_lo,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GILLB200_H_ */
