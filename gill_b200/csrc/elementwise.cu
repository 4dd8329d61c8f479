// Small HBM-bound kernels around the tensor-core path: gathers, resampling, im2col for strided convs, the fused
// CFG + PLMS scheduler step, the VAE uint8 epilogue, L2 normalisation and the fp32 small-sequence attention used by
// the GILLMapper. All are 16-byte vectorised where the layout allows it.
#include "../../include/gillb200.h"
#include "gemm_sm100.cuh"
#include "host_common.h"

namespace gb {

__device__ __forceinline__ uint4 ldg16(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// out[i, :] = (x ? x[i, :] : 0) + table[idx[i] + idx_offset, :]        (16-bit rows, D % 8 == 0)
// OPT: inputs_embeds + embed_positions(pos + 2)  (transformers OPTLearnedPositionalEmbedding; gill/models.py:465),
// and the token-embedding gather gill/models.py:620 / :529 / :709 when x == nullptr.
__global__ void gather_add_rows_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ table,
                                       const long long* __restrict__ idx, long long idx_offset, long long rows, int D,
                                       int is_bf16, uint16_t* __restrict__ out) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const int nvec = D >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / nvec;
    const int v = static_cast<int>(i % nvec);
    const uint4 t = ldg16(table + (idx[r] + idx_offset) * D + v * 8);
    uint4 o = t;
    if (x) {
      const uint4 a = ldg16(x + r * D + v * 8);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, tw[4] = {t.x, t.y, t.z, t.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = is_bf16 ? unpack_bf16x2(aw[j]) : unpack_f16x2(aw[j]);
        const float2 ft = is_bf16 ? unpack_bf16x2(tw[j]) : unpack_f16x2(tw[j]);
        ow[j] = is_bf16 ? pack_bf16x2(fa.x + ft.x, fa.y + ft.y) : pack_f16x2(fa.x + ft.x, fa.y + ft.y);
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    *reinterpret_cast<uint4*>(out + r * D + v * 8) = o;
  }
}

// nearest-neighbour 2x upsample, NHWC 16-bit (F.interpolate(scale_factor=2, mode="nearest") in the UNet/VAE up blocks)
__global__ void upsample2x_kernel(const uint16_t* __restrict__ x, int B, int H, int W, int C,
                                  uint16_t* __restrict__ out) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(B) * 2 * H * 2 * W * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long t = i / nvec;
    const int ox = static_cast<int>(t % (2 * W));
    t /= 2 * W;
    const int oy = static_cast<int>(t % (2 * H));
    const int b = static_cast<int>(t / (2 * H));
    const uint4 val = ldg16(x + ((static_cast<long long>(b) * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8);
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * 2 * H + oy) * 2 * W + ox) * C + v * 8) = val;
  }
}

// im2col for 3x3 / pad 1 convolutions with stride s (UNet downsamplers, stride 2) or tiny channel counts (conv_in,
// C = 4). out[m, (ky*3+kx)*C + c] for m = (b, oy, ox); columns [9*C, ld_out) are zero-filled.
__global__ void im2col3x3_kernel(const uint16_t* __restrict__ x, int B, int H, int W, int C, int stride, int Ho,
                                 int Wo, uint16_t* __restrict__ out, long long ld_out) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const long long total = static_cast<long long>(B) * Ho * Wo * ld_out;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % ld_out);
    long long m = i / ld_out;
    uint16_t val = 0;
    if (col < 9 * C) {
      const int tap = col / C, c = col - tap * C;
      const int ox = static_cast<int>(m % Wo);
      m /= Wo;
      const int oy = static_cast<int>(m % Ho);
      const int b = static_cast<int>(m / Ho);
      const int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = x[((static_cast<long long>(b) * H + iy) * W + ix) * C + c];
    }
    out[i] = val;
  }
}

// Fused classifier-free guidance + PLMS update (gill/custom_sd.py:641-646; diffusers PNDMScheduler.step_plms).
//   eps = eps_u + g * (eps_t - eps_u);  e' = linear multistep over the last <= 4 eps;  x' = c_sample * x - c_eps * e'
// eps_pair: [2, n] 16-bit or fp32 (uncond first). ets: ring of 4 fp32 buffers [4, n]; `head` is the slot that receives
// this step's eps (ignored for mode 1, where eps is NOT pushed: the repeated-timestep half step averages with ets[-1]).
// mode: 0 first step (e' = eps; cur_sample = x saved), 1 second call (e' = (eps + e_-1)/2, x = cur_sample),
//       2/3/4: Adams-Bashforth of that order.
__global__ void plms_step_kernel(const void* __restrict__ eps_pair, int eps_dtype, float guidance, float* __restrict__ ets,
                                 int head, int mode, float c_sample, float c_eps, float* __restrict__ latents,
                                 float* __restrict__ cur_sample, void* __restrict__ lat16_pair, int lat16_dtype,
                                 long long n) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float eu = load_elem(eps_pair, i, eps_dtype), et = load_elem(eps_pair, n + i, eps_dtype);
    const float eps = eu + guidance * (et - eu);
    float x = latents[i];
    float e;
    float* e0 = ets + static_cast<long long>(head) * n;
    const float* e1 = ets + static_cast<long long>((head + 3) & 3) * n;
    const float* e2 = ets + static_cast<long long>((head + 2) & 3) * n;
    const float* e3 = ets + static_cast<long long>((head + 1) & 3) * n;
    if (mode == 1) {
      e = 0.5f * (eps + e1[i]);  // e1 is the most recently pushed eps (head is NOT advanced by this call)
      x = cur_sample[i];
    } else {
      e0[i] = eps;
      if (mode == 0) {
        e = eps;
        cur_sample[i] = x;
      } else if (mode == 2) {
        e = (3.f * eps - e1[i]) * 0.5f;
      } else if (mode == 3) {
        e = (23.f * eps - 16.f * e1[i] + 5.f * e2[i]) * (1.f / 12.f);
      } else {
        e = (55.f * eps - 59.f * e1[i] + 37.f * e2[i] - 9.f * e3[i]) * (1.f / 24.f);
      }
    }
    const float xn = c_sample * x - c_eps * e;
    latents[i] = xn;
    if (lat16_pair) {  // next UNet input: the latent duplicated for the (uncond, text) pair
      store_elem(lat16_pair, i, xn, lat16_dtype);
      store_elem(lat16_pair, n + i, xn, lat16_dtype);
    }
  }
}

// VAE epilogue (gill/custom_sd.py:389-391 + numpy_to_pil): uint8 NHWC = round(clamp(x/2 + 0.5, 0, 1) * 255)
__global__ void image_to_u8_kernel(const void* __restrict__ x, int dtype, long long pixels, int ldx, int channels,
                                   uint8_t* __restrict__ out) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const long long total = pixels * channels;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / channels;
    const int c = static_cast<int>(i % channels);
    float v = load_elem(x, pix * ldx + c, dtype) * 0.5f + 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = static_cast<uint8_t>(rintf(v * 255.f));
  }
}

// y = x / ||x||_2 per row (gill/models.py:674); one warp per row; fp32 in, 16-bit or fp32 out.
__global__ void l2norm_rows_kernel(const float* __restrict__ x, long long ldx, int rows, int n, void* __restrict__ out,
                                   long long ldo, int out_dtype) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float v = x[row * ldx + i];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / sqrtf(s);
  for (int i = lane; i < n; i += 32) store_elem(out, row * ldo + i, x[row * ldx + i] * inv, out_dtype);
}

// elementwise convert / add: out = cast(x (+ y)); optional bf16 residue for split-precision consumers
__global__ void cast_add_kernel(const void* __restrict__ x, int x_dtype, const void* __restrict__ y, int y_dtype,
                                void* __restrict__ out, int out_dtype, void* __restrict__ out_lo, long long n,
                                long long y_period) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = load_elem(x, i, x_dtype);
    if (y) v += load_elem(y, y_period > 0 ? i % y_period : i, y_dtype);
    store_elem(out, i, v, out_dtype);
    if (out_lo) store_elem(out_lo, i, v - __bfloat162float(__float2bfloat16_rn(v)), DT_BF16);
  }
}

// 16-bit -> bf16 (+ bf16 residue), 8 elements per thread: the GILLMapper's fp16 attention output split back into the
// hi + lo operand pair of the next split-precision GEMM (the scalar kernel above took 20 us per 10 M elements).
__global__ void __launch_bounds__(256) cast_split8_kernel(const uint4* __restrict__ x, int x_is_bf16, uint4* __restrict__ out,
                                                          uint4* __restrict__ out_lo, long long nvec) {
  pdl_wait();
  pdl_launch();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = __ldg(x + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = x_is_bf16 ? unpack_bf16x2(w[j]) : unpack_f16x2(w[j]);
      hi[j] = pack_bf16x2(f.x, f.y);
      const float2 h = unpack_bf16x2(hi[j]);
      lo[j] = pack_bf16x2(f.x - h.x, f.y - h.y);
    }
    out[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_lo) out_lo[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// fp32 attention for short sequences (GILLMapper: 4 heads x 128, Lq <= 77, Lk in {8, 77}; gill/layers.py:43).
// One CTA per (batch, head). q/k/v fp32 with row stride ld*, head h at column h*HD. Output fp32 or bf16 hi+lo.
template <int HD>
__global__ void __launch_bounds__(256) attn_small_f32_kernel(const float* __restrict__ q, long long ldq, long long q_bs,
                                                             const float* __restrict__ k, long long ldk, long long k_bs,
                                                             const float* __restrict__ v, long long ldv, long long v_bs,
                                                             int Lq, int Lk, float scale, void* __restrict__ out,
                                                             long long ldo, long long o_bs, int out_dtype,
                                                             void* __restrict__ out_lo) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  extern __shared__ float sm[];
  const int LkP = Lk + 1;
  float* sk = sm;                      // [Lk][HD+1]
  float* sv = sk + Lk * (HD + 1);      // [Lk][HD]
  float* sp = sv + Lk * HD;            // [8 warps][LkP]
  const int b = blockIdx.y, h = blockIdx.x;
  const float* kb = k + b * k_bs + h * HD;
  const float* vb = v + b * v_bs + h * HD;
  for (int i = threadIdx.x; i < Lk * HD; i += blockDim.x) {
    const int r = i / HD, c = i - r * HD;
    sk[r * (HD + 1) + c] = kb[r * ldk + c];
    sv[r * HD + c] = vb[r * ldv + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pw = sp + warp * LkP;
  for (int i = warp; i < Lq; i += 8) {
    const float* qr = q + b * q_bs + static_cast<long long>(i) * ldq + h * HD;
    float qreg[HD / 32];
#pragma unroll
    for (int t = 0; t < HD / 32; ++t) qreg[t] = qr[lane + 32 * t];
    // scores: lane handles keys lane, lane+32, lane+64, lane+96 (Lk <= 128); q elements are broadcast by shuffle
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
    int jj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) jj[u] = min(lane + 32 * u, Lk - 1) * (HD + 1);
#pragma unroll
    for (int t = 0; t < HD / 32; ++t) {
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        const float qv = __shfl_sync(0xffffffffu, qreg[t], l);
        const int c = t * 32 + l;
#pragma unroll
        for (int u = 0; u < 4; ++u) sc[u] += qv * sk[jj[u] + c];
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      if (j < Lk) {
        const float sv_ = sc[u] * scale;
        pw[j] = sv_;
        mx = fmaxf(mx, sv_);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    __syncwarp();
    for (int j = lane; j < Lk; j += 32) {
      const float pj = expf(pw[j] - mx);
      pw[j] = pj;
      sum += pj;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.f / sum;
    float acc[HD / 32];
#pragma unroll
    for (int t = 0; t < HD / 32; ++t) acc[t] = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float pj = pw[j];
#pragma unroll
      for (int t = 0; t < HD / 32; ++t) acc[t] += pj * sv[j * HD + t * 32 + lane];
    }
    const long long o = b * o_bs + static_cast<long long>(i) * ldo + h * HD;
#pragma unroll
    for (int t = 0; t < HD / 32; ++t) {
      const float val = acc[t] * inv;
      store_elem(out, o + t * 32 + lane, val, out_dtype);
      if (out_lo) store_elem(out_lo, o + t * 32 + lane, val - __bfloat162float(__float2bfloat16_rn(val)), DT_BF16);
    }
    __syncwarp();
  }
}

// out[p, :] = W x[p, :] + b for tiny channel counts (<= 8): the VAE's `latents / 0.18215 -> post_quant_conv` 1x1
// (gill/custom_sd.py:386-388), folded into one 4x4 map. fp32 in, 16-bit out.
__global__ void channel_mix_kernel(const float* __restrict__ x, int cin, const float* __restrict__ w,
                                   const float* __restrict__ b, int cout, long long n, void* __restrict__ out,
                                   int out_dtype) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n * cout;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / cout;
    const int co = static_cast<int>(i % cout);
    float acc = b[co];
    for (int ci = 0; ci < cin; ++ci) acc = fmaf(w[co * cin + ci], x[pix * cin + ci], acc);
    store_elem(out, i, acc, out_dtype);
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------------------------------------
// CLIP pre-processing of generated images on the device: uint8 NHWC [B,H,W,3] -> bicubic resize to S x S -> /255 ->
// (x - mean) / std -> NCHW [B,3,S,S]. Replaces `img.resize((224, 224))` + the HF feature extractor of the re-rank step
// (gill/models.py:733-737, gill/utils.py:117-119), i.e. PIL's ImagingResample for 8-bit images, restated exactly:
// separable two-pass (horizontal, then vertical) convolution, support = 2 * scale for down-scaling, cubic a = -0.5,
// per-output-pixel coefficient windows normalised in double, converted to 22-bit fixed point, each pass rounded and
// clipped to uint8. One thread per output pixel recomputes the <= KS horizontal results it needs.
constexpr int RESIZE_KS = 23;  // window size bound: ceil(2 * scale) * 2 + 1 with scale <= 5

__device__ __forceinline__ double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// fixed-point coefficient window of output index `o` (libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc)
__device__ __forceinline__ void pil_coeffs(int in_size, int out_size, int o, int* xmin_out, int* n_out, int* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const double center = (o + 0.5) * scale;
  const double ss = 1.0 / filterscale;
  int xmin = static_cast<int>(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double w[RESIZE_KS];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    w[x] = pil_bicubic((x + xmin - center + 0.5) * ss);
    ww += w[x];
  }
  for (int x = 0; x < xmax; ++x) {
    const double k = ww != 0.0 ? w[x] / ww : w[x];
    kk[x] = k < 0 ? static_cast<int>(-0.5 + k * (1 << 22)) : static_cast<int>(0.5 + k * (1 << 22));
  }
  *xmin_out = xmin;
  *n_out = xmax;
}

__device__ __forceinline__ int pil_clip8(int v) {
  v >>= 22;
  return v < 0 ? 0 : v > 255 ? 255 : v;
}

// The image is resized to RH x RW and the S x S window at (top, left) of the resized image is produced: RH = RW = S,
// top = left = 0 is `img.resize((S, S))`; RH / RW = shortest edge S, centred window = the HF CLIP feature extractor
// (resize shortest edge + centre crop, gill/utils.py:117-119 -> gill/models.py:608).
__global__ void clip_preprocess_u8_kernel(const uint8_t* __restrict__ img, int B, int H, int W, int RH, int RW, int top,
                                          int left, int S, float m0, float m1,
                                          float m2, float s0, float s1, float s2, void* __restrict__ out, int out_dtype,
                                          uint8_t* __restrict__ resized /* optional [B,S,S,3] */) {
  pdl_wait();
  pdl_launch();
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= 1LL * B * S * S) return;
  const int ox = static_cast<int>(t % S), oy = static_cast<int>((t / S) % S), b = static_cast<int>(t / (1LL * S * S));
  int kx[RESIZE_KS], ky[RESIZE_KS], x0, nx, y0, ny;
  pil_coeffs(W, RW, ox + left, &x0, &nx, kx);
  pil_coeffs(H, RH, oy + top, &y0, &ny, ky);
  const uint8_t* base = img + static_cast<long long>(b) * H * W * 3;
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = base + (static_cast<long long>(y0 + j) * W + x0) * 3;
    int h[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int i = 0; i < nx; ++i) {
      h[0] += row[3 * i] * kx[i];
      h[1] += row[3 * i + 1] * kx[i];
      h[2] += row[3 * i + 2] * kx[i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += pil_clip8(h[c]) * ky[j];  // horizontal pass result is an 8-bit image
  }
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v = pil_clip8(acc[c]);
    if (resized) resized[(static_cast<long long>(b) * S * S + static_cast<long long>(oy) * S + ox) * 3 + c] = static_cast<uint8_t>(v);
    // HF image processor: rescale in float64 -> float32, then (x - mean) / std in float32
    const float x = static_cast<float>(static_cast<double>(v) * 0.00392156862745098);
    store_elem(out, ((static_cast<long long>(b) * 3 + c) * S + oy) * S + ox, (x - mean[c]) / sd[c], out_dtype);
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gillb200_gather_add_rows(const void* x, const void* table, const long long* idx, long long idx_offset,
                                        long long rows, int D, int dtype, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(table && idx && out && D % 8 == 0 && rows > 0, "gather_add_rows: bad args");
  GB_CHECK_ARG(dtype == DT_BF16 || dtype == DT_F16, "gather_add_rows: 16-bit only");
  GB_CUDA(launch_pdl_light(gather_add_rows_kernel, dim3(grid_for(rows * (D / 8), 256)), dim3(256), 0, stream, 
      reinterpret_cast<const uint16_t*>(x), reinterpret_cast<const uint16_t*>(table), idx, idx_offset, rows, D,
      dtype == DT_BF16, reinterpret_cast<uint16_t*>(out)));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

// out[b,h,w,o] = bias[o] + sum_{tap} y[(b, h + dy - 1, w + dx - 1), tap * COUT + o]   (tap = dy * 3 + dx; zero padding)
template <int COUT>
__global__ void __launch_bounds__(256) tap_sum3x3_kernel(const float* __restrict__ y, long long ldy, int B, int H, int W,
                                                         const float* __restrict__ bias, void* __restrict__ out,
                                                         int out_dtype, long long ldo) {
  pdl_wait();
  pdl_launch();
  const long long pix = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (pix >= static_cast<long long>(B) * H * W) return;
  const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = bias ? bias[o] : 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int hh = h + dy - 1;
    if (hh < 0 || hh >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ww = w + dx - 1;
      if (ww < 0 || ww >= W) continue;
      const float* src = y + (pix + (dy - 1) * W + (dx - 1)) * ldy + (dy * 3 + dx) * COUT;
      if (COUT == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src));
        acc[0] += t.x;
        acc[1 % COUT] += t.y;
        acc[2 % COUT] += t.z;
        acc[3 % COUT] += t.w;
      } else {
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] += __ldg(src + o);
      }
    }
  }
  if (out_dtype == DT_F32) {
    float* o = static_cast<float*>(out) + pix * ldo;
#pragma unroll
    for (int i = 0; i < COUT; ++i) o[i] = acc[i];
  } else {
    uint16_t* o = static_cast<uint16_t*>(out) + pix * ldo;
#pragma unroll
    for (int i = 0; i < COUT; ++i)
      o[i] = out_dtype == DT_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(acc[i])) : __half_as_ushort(__float2half_rn(acc[i]));
  }
}

extern "C" int gillb200_tap_sum3x3(const float* y, long long ldy, int B, int H, int W, int Cout, const float* bias, void* out,
                                   int out_dtype, long long ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(y && out && B > 0 && H > 0 && W > 0, "tap_sum3x3: bad args");
  GB_CHECK_ARG(Cout == 3 || Cout == 4 || Cout == 8, "tap_sum3x3: Cout must be 3, 4 or 8 (got %d)", Cout);
  GB_CHECK_ARG(ldy >= 9 * Cout && ldo >= Cout, "tap_sum3x3: ldy=%lld ldo=%lld too small for Cout=%d", ldy, ldo, Cout);
  GB_CHECK_ARG(Cout != 4 || (ldy % 4 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0), "tap_sum3x3: Cout 4 needs 16-byte rows");
  const long long n = static_cast<long long>(B) * H * W;
  const dim3 grid(static_cast<unsigned>((n + 255) / 256));
  if (Cout == 3) GB_CUDA(launch_pdl_light(tap_sum3x3_kernel<3>, grid, dim3(256), 0, stream, y, ldy, B, H, W, bias, out, out_dtype, ldo));
  else if (Cout == 4) GB_CUDA(launch_pdl_light(tap_sum3x3_kernel<4>, grid, dim3(256), 0, stream, y, ldy, B, H, W, bias, out, out_dtype, ldo));
  else GB_CUDA(launch_pdl_light(tap_sum3x3_kernel<8>, grid, dim3(256), 0, stream, y, ldy, B, H, W, bias, out, out_dtype, ldo));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && C % 8 == 0, "upsample2x: bad args");
  GB_CUDA(launch_pdl_light(upsample2x_kernel, dim3(grid_for(4LL * B * H * W * (C / 8), 256)), dim3(256), 0, stream, 
      reinterpret_cast<const uint16_t*>(x), B, H, W, C, reinterpret_cast<uint16_t*>(out)));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_im2col3x3(const void* x, int B, int H, int W, int C, int stride, void* out, long long ld_out,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && (stride == 1 || stride == 2) && ld_out >= 9 * C, "im2col3x3: bad args");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  GB_CUDA(launch_pdl_light(im2col3x3_kernel, dim3(grid_for(1LL * B * Ho * Wo * ld_out, 256)), dim3(256), 0, stream, 
      reinterpret_cast<const uint16_t*>(x), B, H, W, C, stride, Ho, Wo, reinterpret_cast<uint16_t*>(out), ld_out));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_plms_step(const void* eps_pair, int eps_dtype, float guidance, float* ets, int head, int mode,
                                  float c_sample, float c_eps, float* latents, float* cur_sample, void* lat16_pair,
                                  int lat16_dtype, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(eps_pair && ets && latents && cur_sample && n > 0, "plms_step: null pointer");
  GB_CHECK_ARG(mode >= 0 && mode <= 4 && head >= 0 && head < 4, "plms_step: bad mode/head");
  GB_CUDA(launch_pdl_light(plms_step_kernel, dim3(grid_for(n, 256)), dim3(256), 0, stream, eps_pair, eps_dtype, guidance, ets, head, mode, c_sample,
                                                         c_eps, latents, cur_sample, lat16_pair, lat16_dtype, n));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_image_to_u8(const void* x, int dtype, long long pixels, int ldx, int channels, void* out,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && pixels > 0 && channels > 0 && ldx >= channels, "image_to_u8: bad args");
  GB_CUDA(launch_pdl_light(image_to_u8_kernel, dim3(grid_for(pixels * channels, 256)), dim3(256), 0, stream, x, dtype, pixels, ldx, channels,
                                                                           reinterpret_cast<uint8_t*>(out)));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_l2norm_rows(const float* x, long long ldx, int rows, int n, void* out, long long ldo,
                                    int out_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && rows > 0 && n > 0, "l2norm_rows: bad args");
  GB_CUDA(launch_pdl_light(l2norm_rows_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream, x, ldx, rows, n, out, ldo, out_dtype));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_cast_add(const void* x, int x_dtype, const void* y, int y_dtype, long long y_period, void* out,
                                 int out_dtype, void* out_lo, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && n > 0, "cast_add: bad args");
  if (!y && x_dtype != DT_F32 && out_dtype == DT_BF16 && n % 8 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0 && reinterpret_cast<uintptr_t>(out_lo) % 16 == 0) {
    GB_CUDA(launch_pdl_light(cast_split8_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, stream, reinterpret_cast<const uint4*>(x),
                       x_dtype == DT_BF16 ? 1 : 0, reinterpret_cast<uint4*>(out), reinterpret_cast<uint4*>(out_lo), n / 8));
    GB_COUNT_LAUNCH(1);
    return 0;
  }
  GB_CUDA(launch_pdl_light(cast_add_kernel, dim3(grid_for(n, 256)), dim3(256), 0, stream, x, x_dtype, y, y_dtype, out, out_dtype, out_lo, n, y_period));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_attn_small_f32(const float* q, long long ldq, long long q_bs, const float* k, long long ldk,
                                       long long k_bs, const float* v, long long ldv, long long v_bs, int B, int H,
                                       int hd, int Lq, int Lk, float scale, void* out, long long ldo, long long o_bs,
                                       int out_dtype, void* out_lo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(q && k && v && out, "attn_small_f32: null pointer");
  GB_CHECK_ARG(hd == 128, "attn_small_f32: head dim must be 128 (got %d)", hd);
  GB_CHECK_ARG(Lk >= 1 && Lk <= 128 && Lq >= 1, "attn_small_f32: Lk must be in [1,128]");
  const size_t smem = (static_cast<size_t>(Lk) * (128 + 1) + static_cast<size_t>(Lk) * 128 + 8 * (Lk + 1)) * sizeof(float);
  static PerDeviceOnce configured;
  if (configured.need()) {
    GB_CUDA(cudaFuncSetAttribute(attn_small_f32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  GB_CUDA(launch_pdl_light(attn_small_f32_kernel<128>, dim3(dim3(H, B)), dim3(256), smem, stream, q, ldq, q_bs, k, ldk, k_bs, v, ldv, v_bs, Lq, Lk, scale,
                                                                  out, ldo, o_bs, out_dtype, out_lo));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_channel_mix(const float* x, int cin, const float* w, const float* b, int cout, long long n,
                                    void* out, int out_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && w && b && out && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8 && n > 0, "channel_mix: bad args");
  GB_CUDA(launch_pdl_light(channel_mix_kernel, dim3(grid_for(n * cout, 256)), dim3(256), 0, stream, x, cin, w, b, cout, n, out, out_dtype));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gillb200_clip_preprocess_u8_crop(const void* img, int B, int H, int W, int RH, int RW, int top, int left,
                                                int S, const float* mean3, const float* std3, void* out, int out_dtype,
                                                void* resized_u8, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(img && out && mean3 && std3 && B > 0 && H > 0 && W > 0 && S > 0, "clip_preprocess_u8: bad args");
  GB_CHECK_ARG(RH >= S && RW >= S && top >= 0 && left >= 0 && top + S <= RH && left + S <= RW,
               "clip_preprocess_u8: the %d x %d window at (%d, %d) leaves the %d x %d resized image", S, S, top, left, RH, RW);
  GB_CHECK_ARG(H <= 5 * RH && W <= 5 * RW, "clip_preprocess_u8: down-scaling factor above 5 (H=%d W=%d -> %d x %d)", H, W, RH,
               RW);
  GB_CHECK_ARG(out_dtype >= 0 && out_dtype <= 2, "clip_preprocess_u8: bad out dtype");
  GB_CUDA(launch_pdl_light(clip_preprocess_u8_kernel, dim3(grid_for(1LL * B * S * S, 128)), dim3(128), 0, stream,
                     reinterpret_cast<const uint8_t*>(img), B, H, W, RH, RW, top, left, S, mean3[0], mean3[1], mean3[2],
                     std3[0], std3[1], std3[2], out, out_dtype, reinterpret_cast<uint8_t*>(resized_u8)));
  GB_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int gillb200_clip_preprocess_u8(const void* img, int B, int H, int W, int S, const float* mean3,
                                           const float* std3, void* out, int out_dtype, void* resized_u8,
                                           void* stream_) {
  return gillb200_clip_preprocess_u8_crop(img, B, H, W, S, S, 0, 0, S, mean3, std3, out, out_dtype, resized_u8, stream_);
}
