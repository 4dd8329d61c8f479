// LayerNorm / GroupNorm(+SiLU) / row softmax: HBM-bound kernels, 16-byte vectorised, fp32 statistics.
//
// Replaces F.layer_norm / nn.GroupNorm / softmax inside the third-party modules on the hot path:
//   nn.Transformer norms of the GILLMapper (gill/layers.py:20-22), OPT decoder layer norms (gill/models.py:465),
//   UNet / VAE GroupNorm+SiLU and transformer-block LayerNorms (gill/custom_sd.py:633-638, :388).
#include "../../include/gillb200.h"
#include "gemm_sm100.cuh"
#include "host_common.h"

#include <cstdlib>

namespace gb {

// ------------------------------------------------------------------------------------------------ helpers
struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const void* base, long long elem_off, int dtype) {
  Vec8 r;
  if (dtype == DT_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + elem_off);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = dtype == DT_BF16 ? unpack_bf16x2(w[i]) : unpack_f16x2(w[i]);
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    }
  }
  return r;
}

__device__ __forceinline__ void store8(void* base, long long elem_off, const Vec8& x, int dtype) {
  if (dtype == DT_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
    p[1] = make_float4(x.v[4], x.v[5], x.v[6], x.v[7]);
  } else {
    uint4 u;
    if (dtype == DT_BF16) {
      u.x = pack_bf16x2(x.v[0], x.v[1]); u.y = pack_bf16x2(x.v[2], x.v[3]);
      u.z = pack_bf16x2(x.v[4], x.v[5]); u.w = pack_bf16x2(x.v[6], x.v[7]);
    } else {
      u.x = pack_f16x2(x.v[0], x.v[1]); u.y = pack_f16x2(x.v[2], x.v[3]);
      u.z = pack_f16x2(x.v[4], x.v[5]); u.w = pack_f16x2(x.v[6], x.v[7]);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + elem_off) = u;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One row per `WARPS` warps; each lane keeps up to MAXV 8-wide vectors of the row in registers (two-pass variance).
template <int WARPS, int MAXV>
__global__ void __launch_bounds__(128) layernorm_kernel(const void* __restrict__ x, long long ldx, int in_dtype,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float eps, int rows, int C, void* __restrict__ out,
                                                        long long ldo, int out_dtype, void* __restrict__ out_lo) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  __shared__ float red[2][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per_block = 4 / WARPS;
  const int row = blockIdx.x * rows_per_block + (WARPS == 1 ? warp : 0);
  const int t = WARPS == 1 ? lane : threadIdx.x;  // thread index within the row team
  const int team = WARPS * 32;
  const int nvec = C >> 3;
  const bool active = row < rows;
  Vec8 buf[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = t + i * team;
    if (active && v < nvec) {
      buf[i] = load8(x, static_cast<long long>(row) * ldx + v * 8, in_dtype);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += buf[i].v[j];
    }
  }
  s = warp_sum(s);
  if (WARPS > 1) {
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  }
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = t + i * team;
    if (active && v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = buf[i].v[j] - mean;
        q += d * d;
      }
    }
  }
  q = warp_sum(q);
  if (WARPS > 1) {
    if (lane == 0) red[1][warp] = q;
    __syncthreads();
    q = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = t + i * team;
    if (active && v < nvec) {
      Vec8 o;
      const float4 w0 = *reinterpret_cast<const float4*>(w + v * 8), w1 = *reinterpret_cast<const float4*>(w + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b + v * 8), b1 = *reinterpret_cast<const float4*>(b + v * 8 + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (buf[i].v[j] - mean) * rstd * ww[j] + bb[j];
      const long long off = static_cast<long long>(row) * ldo + v * 8;
      store8(out, off, o, out_dtype);
      if (out_lo) {
        Vec8 lo;
#pragma unroll
        for (int j = 0; j < 8; ++j) lo.v[j] = o.v[j] - __bfloat162float(__float2bfloat16_rn(o.v[j]));
        store8(out_lo, off, lo, DT_BF16);
      }
    }
  }
}

// Narrow rows (C <= 1280): LPR lanes per row, 32/LPR rows per warp, every lane keeps its <= MAXV 16-byte vectors in
// registers, so one warp has 32*MAXV independent loads in flight (the one-warp-per-row form above leaves most lanes
// idle at C = 320 and reached ~2 TB/s). Sub-warp xor-shuffle reductions, two-pass variance.
template <int LPR, int MAXV>
__global__ void __launch_bounds__(128) layernorm_sub_kernel(const void* __restrict__ x, long long ldx, int in_dtype,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, int rows, int C, void* __restrict__ out,
                                                            long long ldo, int out_dtype, void* __restrict__ out_lo) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = (blockIdx.x * 4 + warp) * RPW + lane / LPR;
  const int t = lane % LPR;
  const int nvec = C >> 3;
  const bool active = row < rows;
  Vec8 buf[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = t + i * LPR;
    if (active && v < nvec) buf[i] = load8(x, static_cast<long long>(row) * ldx + v * 8, in_dtype);
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (active && t + i * LPR < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += buf[i].v[j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (active && t + i * LPR < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = buf[i].v[j] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = t + i * LPR;
    if (active && v < nvec) {
      Vec8 o;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + v * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + v * 8 + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (buf[i].v[j] - mean) * rstd * ww[j] + bb[j];
      const long long off = static_cast<long long>(row) * ldo + v * 8;
      store8(out, off, o, out_dtype);
      if (out_lo) {
        Vec8 lo;
#pragma unroll
        for (int j = 0; j < 8; ++j) lo.v[j] = o.v[j] - __bfloat162float(__float2bfloat16_rn(o.v[j]));
        store8(out_lo, off, lo, DT_BF16);
      }
    }
  }
}

// Persistent form of layernorm_sub_kernel for 16-bit input and output (the UNet's 48 LayerNorms per evaluation): a CTA walks
// row blocks with a grid stride and keeps the NEXT block's raw 16-byte vectors in flight while it reduces and writes the
// current one (the one-shot form ran 4096 short-lived CTAs of 16 rows at 64x64 and reached ~3 TB/s on L2-resident rows).
template <int LPR, int MAXV>
__global__ void __launch_bounds__(128) layernorm_sub16_kernel(const uint16_t* __restrict__ x, long long ldx, int in_dtype,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              float eps, int rows, int C, uint16_t* __restrict__ out,
                                                              long long ldo, int out_dtype) {
  pdl_wait();
  pdl_launch();
  constexpr int RPW = 32 / LPR;          // rows per warp
  constexpr int RPB = 4 * RPW;           // rows per CTA step
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane % LPR, rsub = warp * RPW + lane / LPR;
  const int nvec = C >> 3;
  const int nblk = (rows + RPB - 1) / RPB;
  const bool bf = in_dtype == DT_BF16, obf = out_dtype == DT_BF16;
  uint4 nxt[MAXV];
  auto fetch = [&](int blk) {
    const int row = blk * RPB + rsub;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = t + i * LPR;
      nxt[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row < rows && v < nvec) nxt[i] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx + v * 8));
    }
  };
  int blk = blockIdx.x;
  if (blk < nblk) fetch(blk);
  for (; blk < nblk; blk += gridDim.x) {
    uint4 cur[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) cur[i] = nxt[i];
    if (blk + static_cast<int>(gridDim.x) < nblk) fetch(blk + gridDim.x);
    const int row = blk * RPB + rsub;
    float f[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const uint32_t wd[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 p2 = bf ? unpack_bf16x2(wd[j]) : unpack_f16x2(wd[j]);
        f[i][2 * j] = p2.x;
        f[i][2 * j + 1] = p2.y;
      }
      if (t + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[i][j];
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (t + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[i][j] - mean;
          q += d * d;
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = t + i * LPR;
      if (row < rows && v < nvec) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + v * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + v * 8 + 4));
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float y0 = (f[i][2 * j] - mean) * rstd * ww[2 * j] + bb[2 * j];
          const float y1 = (f[i][2 * j + 1] - mean) * rstd * ww[2 * j + 1] + bb[2 * j + 1];
          r[j] = obf ? pack_bf16x2(y0, y1) : pack_f16x2(y0, y1);
        }
        *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * ldo + v * 8) = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// NHWC input, optionally the channel-concatenation of two tensors (UNet up blocks: cat([h, skip])).
// Pass 1: grid (chunks, B): per-channel partial sums over a pixel chunk -> per-group (sum, sumsq) partials.
// Pass 2: elementwise normalise (+SiLU), reducing the `chunks` partials in a fixed order (deterministic).
struct GnSrc {
  const void* x0;
  const void* x1;
  int C0, C1;  // channels of each source (C1 == 0: single source)
};

__device__ __forceinline__ Vec8 gn_load(const GnSrc& s, long long pix, int v, int dtype) {
  const int v0 = s.C0 >> 3;
  if (v < v0) return load8(s.x0, pix * s.C0 + v * 8, dtype);
  return load8(s.x1, pix * s.C1 + (v - v0) * 8, dtype);
}

__global__ void __launch_bounds__(256) gn_stats_kernel(GnSrc src, int dtype, int HW, int G, int chunks,
                                                       float* __restrict__ partial /* [B, chunks, G, 2] */,
                                                       int* __restrict__ counters /* [B], zero between launches */,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       float eps, float* __restrict__ scale_shift /* [B, 2, C] */) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  // Thread t owns channel vector v and pixel subgroup sg: 256 consecutive threads read 256 consecutive 16-byte
  // vectors (NHWC is pixel-major, so the next pixel's channels follow). Per-thread register sums -> smem
  // [sg][sum|sumsq][C] -> per-group totals summed in a fixed order: no atomics, bit-reproducible.
  extern __shared__ float sm[];
  const int C = src.C0 + src.C1;
  const int nvec = C >> 3;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int pix_per_chunk = (HW + chunks - 1) / chunks;
  const int p0 = chunk * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
  const bool wide = nvec >= 256;
  const int sgs = wide ? 1 : 256 / nvec;
  const int vper = wide ? (nvec + 255) / 256 : 1;
  for (int vi = 0; vi < vper; ++vi) {
    int v, sg;
    bool active;
    if (wide) {
      v = vi * 256 + threadIdx.x;
      sg = 0;
      active = v < nvec;
    } else {
      v = threadIdx.x % nvec;
      sg = threadIdx.x / nvec;
      active = sg < sgs;
    }
    if (active) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
      int pix = p0 + sg;
      for (; pix + 3 * sgs < p1; pix += 4 * sgs) {  // 4 independent 16-byte loads in flight per thread
        Vec8 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = gn_load(src, static_cast<long long>(b) * HW + pix + u * sgs, v, dtype);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += x[u].v[j];
            q[j] += x[u].v[j] * x[u].v[j];
          }
        }
      }
      for (; pix < p1; pix += sgs) {
        const Vec8 x = gn_load(src, static_cast<long long>(b) * HW + pix, v, dtype);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += x.v[j];
          q[j] += x.v[j] * x.v[j];
        }
      }
      float* ps = sm + (sg * 2) * C + v * 8;
      float* pq = sm + (sg * 2 + 1) * C + v * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ps[j] = s[j];
        pq[j] = q[j];
      }
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int sg = 0; sg < sgs; ++sg) {
      const float* ps = sm + (sg * 2) * C;
      const float* pq = sm + (sg * 2 + 1) * C;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += ps[c];
        q += pq[c];
      }
    }
    float* o = partial + ((static_cast<long long>(b) * chunks + chunk) * G + g) * 2;
    o[0] = s;
    o[1] = q;
  }
  // The last CTA of each sample to finish turns the partials into the per-channel affine y = x * scale + shift
  // (summing chunks in a fixed order: deterministic), so no separate finalize launch is needed.
  __shared__ int is_last;
  __shared__ float gmean[64], grstd[64];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&counters[b], 1) == chunks - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      const volatile float* pp = partial + ((static_cast<long long>(b) * chunks + ch) * G + g) * 2;
      s += pp[0];
      q += pp[1];
    }
    const float n = static_cast<float>(cpg) * HW;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    gmean[g] = mean;
    grstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  float* sc = scale_shift + static_cast<long long>(b) * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = grstd[g] * w[c];
    sc[c] = a;
    sc[C + c] = bias[c] - gmean[g] * a;
  }
  if (threadIdx.x == 0) counters[b] = 0;
}

// (kept for reference / debugging) Per-(sample, channel) affine of the normalisation: y = x * scale + shift. One tiny launch per GroupNorm so that the
// elementwise pass below needs no shared memory and no per-CTA recomputation.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, int G, int C, int HW,
                                   const float* __restrict__ w, const float* __restrict__ bias, float eps,
                                   float* __restrict__ scale_shift /* [B, 2, C] */) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  __shared__ float gmean[64], grstd[64];
  const int b = blockIdx.x;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {  // fixed order: deterministic
      const float* pp = partial + ((static_cast<long long>(b) * chunks + ch) * G + g) * 2;
      s += pp[0];
      q += pp[1];
    }
    const float n = static_cast<float>(cpg) * HW;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    gmean[g] = mean;
    grstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  float* sc = scale_shift + static_cast<long long>(b) * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = grstd[g] * w[c];
    sc[c] = a;
    sc[C + c] = bias[c] - gmean[g] * a;
  }
}

// GroupNorm statistics from the producers' epilogues: stats[(sample * slabs + s) * Csrc + c] = {sum, sumsq} of channel c
// over the 32 pixels of slab s (gemm stats_out). One CTA per (group, sample) adds slabs x channels-of-the-group in a
// fixed order (deterministic) and writes the group's part of the per-channel affine y = x * scale + shift.
__global__ void __launch_bounds__(128) gn_finalize_stats_kernel(const float2* __restrict__ st0, int C0,
                                                                const float2* __restrict__ st1, int C1, int slabs,
                                                                int G, int HW, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float eps,
                                                                float* __restrict__ scale_shift /* [B, 2, C] */) {
  pdl_wait();
  pdl_launch();
  __shared__ float rs[128], rq[128];
  const int g = blockIdx.x, b = blockIdx.y;
  const int C = C0 + C1, cpg = C / G;
  const int n = slabs * cpg;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const int sl = i / cpg, c = g * cpg + (i - sl * cpg);
    const float2 v = c < C0 ? st0[(static_cast<size_t>(b) * slabs + sl) * C0 + c]
                            : st1[(static_cast<size_t>(b) * slabs + sl) * C1 + (c - C0)];
    s += v.x;
    q += v.y;
  }
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      rs[threadIdx.x] += rs[threadIdx.x + o];
      rq[threadIdx.x] += rq[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float cnt = static_cast<float>(cpg) * HW;
  const float mean = rs[0] / cnt;
  const float var = fmaxf(rq[0] / cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float* sc = scale_shift + static_cast<long long>(b) * 2 * C;
  for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += 128) {
    const float a = rstd * w[c];
    sc[c] = a;
    sc[C + c] = bias[c] - mean * a;
  }
}

// Elementwise normalise (+SiLU). Thread t owns channel vector v = t % nvec (scale/shift live in 16 registers) and
// walks pixels with stride (256 / nvec): 256 consecutive threads touch 256 consecutive 16-byte vectors.
__global__ void __launch_bounds__(256) gn_apply_kernel(GnSrc src, int dtype, int HW, const float* __restrict__ scale_shift,
                                                       int silu, void* __restrict__ out, int out_dtype,
                                                       int pix_per_block) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  const int C = src.C0 + src.C1;
  const int nvec = C >> 3;
  const int b = blockIdx.y;
  const bool wide = nvec >= 256;
  const int sgs = wide ? 1 : 256 / nvec;
  const int vper = wide ? (nvec + 255) / 256 : 1;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const float* sc = scale_shift + static_cast<long long>(b) * 2 * C;
  for (int vi = 0; vi < vper; ++vi) {
    int v, sg;
    bool active;
    if (wide) {
      v = vi * 256 + threadIdx.x;
      sg = 0;
      active = v < nvec;
    } else {
      v = threadIdx.x % nvec;
      sg = threadIdx.x / nvec;
      active = sg < sgs;
    }
    if (!active) continue;
    float a[8], d[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(sc + v * 8), a1 = *reinterpret_cast<const float4*>(sc + v * 8 + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(sc + C + v * 8), d1 = *reinterpret_cast<const float4*>(sc + C + v * 8 + 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
    }
    int pix = p0 + sg;
    // 4 independent 16-byte loads in flight per thread
    for (; pix + 3 * sgs < p1; pix += 4 * sgs) {
      Vec8 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = gn_load(src, static_cast<long long>(b) * HW + pix + u * sgs, v, dtype);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = fmaf(x[u].v[j], a[j], d[j]);
          if (silu) y = __fdividef(y, 1.f + __expf(-y));
          x[u].v[j] = y;
        }
        store8(out, (static_cast<long long>(b) * HW + pix + u * sgs) * C + v * 8, x[u], out_dtype);
      }
    }
    for (; pix < p1; pix += sgs) {
      Vec8 x = gn_load(src, static_cast<long long>(b) * HW + pix, v, dtype);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(x.v[j], a[j], d[j]);
        if (silu) y = __fdividef(y, 1.f + __expf(-y));
        x.v[j] = y;
      }
      store8(out, (static_cast<long long>(b) * HW + pix) * C + v * 8, x, out_dtype);
    }
  }
}

// y = x * a + d for a pair of channels, then SiLU, in packed fp32 (FFMA2 / FMUL2 / FADD2: the elementwise pass is bound by
// issue slots -- ncu r02: 62 % issue-active, XU 52 %, DRAM 22 % -- so halving the fp32 instruction count is what pays).
// Same rounded fp32 operations per lane as the scalar form: bit-identical results.
__device__ __forceinline__ uint32_t gn_affine_silu2(float2 f, float a0, float a1, float d0, float d1, bool silu, bool obf) {
  f32x2 y = fma2(pk2(f.x, f.y), pk2(a0, a1), pk2(d0, d1));
  float y0, y1;
  if (silu) {
    float z0, z1;
    upk2(mul2(y, pk2(-1.4426950408889634f, -1.4426950408889634f)), z0, z1);
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
    float s0, s1;
    upk2(add2(pk2(e0, e1), pk2(1.f, 1.f)), s0, s1);
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(s1));
    upk2(mul2(y, pk2(r0, r1)), y0, y1);
  } else {
    upk2(y, y0, y1);
  }
  return obf ? pack_bf16x2(y0, y1) : pack_f16x2(y0, y1);
}

// Flat elementwise normalise (+SiLU) for 16-bit activations: a CTA's slice of one sample is a contiguous run of 16-byte
// vectors, thread t takes vectors t, t + 256, ... (every warp load/store is 512 contiguous bytes, no idle lanes for any
// channel count) with UNR independent loads in flight; the per-channel scale / shift of the sample live in shared
// memory. Replaces the channel-owning form above (75 registers, 3 CTAs per SM, 15/16 of the lanes active at C = 320:
// ncu r01 27 % of the warps resident, 2.1 TB/s on tensors that sit in L2).
template <int UNR>
__global__ void __launch_bounds__(256, 4) gn_apply2_kernel(GnSrc src, int dtype, int HW, const float* __restrict__ scale_shift,
                                                        int silu, void* __restrict__ out, int out_dtype,
                                                        long long vec_per_block) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float gn_sm[];  // [2 * C]: scale, shift
  const int C = src.C0 + src.C1;
  const int nvec = C >> 3, v0 = src.C0 >> 3;
  const int b = blockIdx.y;
  const float* sc = scale_shift + static_cast<long long>(b) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += 256) gn_sm[i] = sc[i];
  __syncthreads();
  const long long total = static_cast<long long>(HW) * nvec;
  const long long e0 = blockIdx.x * vec_per_block, e1 = min(total, e0 + vec_per_block);
  const int qstep = 256 / nvec, rstep = 256 % nvec;
  long long e = e0 + threadIdx.x;
  long long pix = e / nvec;
  int v = static_cast<int>(e - pix * nvec);
  const bool bf = dtype == DT_BF16, obf = out_dtype == DT_BF16;
  const uint16_t* x0 = static_cast<const uint16_t*>(src.x0) + static_cast<long long>(b) * HW * src.C0;
  const uint16_t* x1 = static_cast<const uint16_t*>(src.x1) + static_cast<long long>(b) * HW * src.C1;
  uint16_t* o = static_cast<uint16_t*>(out) + static_cast<long long>(b) * HW * C;
  for (; e < e1; e += 256 * UNR) {
    uint4 raw[UNR];
    int vv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      vv[u] = v;
      raw[u] = make_uint4(0u, 0u, 0u, 0u);
      if (e + u * 256 < e1) {
        const uint16_t* ptr = v < v0 ? x0 + pix * src.C0 + v * 8 : x1 + pix * src.C1 + (v - v0) * 8;
        raw[u] = __ldg(reinterpret_cast<const uint4*>(ptr));
      }
      pix += qstep;
      v += rstep;
      if (v >= nvec) {
        v -= nvec;
        ++pix;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (e + u * 256 < e1) {
        const float4 a0 = *reinterpret_cast<const float4*>(gn_sm + vv[u] * 8), a1 = *reinterpret_cast<const float4*>(gn_sm + vv[u] * 8 + 4);
        const float4 d0 = *reinterpret_cast<const float4*>(gn_sm + C + vv[u] * 8), d1 = *reinterpret_cast<const float4*>(gn_sm + C + vv[u] * 8 + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = bf ? unpack_bf16x2(w[j]) : unpack_f16x2(w[j]);
          r[j] = gn_affine_silu2(f, a[2 * j], a[2 * j + 1], d[2 * j], d[2 * j + 1], silu != 0, obf);
        }
        *reinterpret_cast<uint4*>(o + (e + u * 256) * 8) = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
  }
}

// GroupNorm from the producers' statistics in ONE launch, for the UNet's 32x32 / 16x16 / 8x8 levels (HW <= 1024, i.e. at
// most 32 statistics slabs per sample). A CTA owns (sample, block of CB channels = whole groups and whole 16-byte vectors,
// pixel range): its prologue reduces the slabs x CB producer sums of ITS groups (one warp per group, fixed order) into
// scale / shift in shared memory, then it streams its pixels. The two-launch form (gn_finalize_stats_kernel +
// gn_apply2_kernel) cost 15-23 us per GroupNorm at these levels for 2.6-21 MB tensors -- two dependent launches of latency
// for a few microseconds of traffic; 49 of the 61 GroupNorms of a UNet evaluation are at these levels.
template <int UNR>
__global__ void __launch_bounds__(256, 4) gn_fused_kernel(GnSrc src, const float2* __restrict__ st0,
                                                          const float2* __restrict__ st1, int dtype, int HW, int G, int CB,
                                                          int pix_per_cta, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float eps, int silu,
                                                          void* __restrict__ out, int out_dtype) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float gn_sm[];  // [2 * CB]: scale, shift
  const int C = src.C0 + src.C1, cpg = C / G, slabs = HW >> 5;
  const int nblk = C / CB;
  const int cb = blockIdx.x % nblk, pc = blockIdx.x / nblk, b = blockIdx.y;
  const int c0 = cb * CB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float cnt = static_cast<float>(cpg) * HW;
  for (int gl = warp; gl < CB / cpg; gl += 8) {
    const int cg = c0 + gl * cpg;  // first channel of the group
    float sm = 0.f, sq = 0.f;
    const int n = slabs * cpg;
    for (int i = lane; i < n; i += 32) {
      const int sl = i / cpg, c = cg + (i - sl * cpg);
      const float2 v = c < src.C0 ? __ldg(st0 + (static_cast<size_t>(b) * slabs + sl) * src.C0 + c)
                                  : __ldg(st1 + (static_cast<size_t>(b) * slabs + sl) * src.C1 + (c - src.C0));
      sm += v.x;
      sq += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = sm / cnt;
    const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + eps);
    for (int j = lane; j < cpg; j += 32) {
      const float a = rstd * w[cg + j];
      gn_sm[gl * cpg + j] = a;
      gn_sm[CB + gl * cpg + j] = bias[cg + j] - mean * a;
    }
  }
  __syncthreads();
  const int nv = CB >> 3;
  const int p0 = pc * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const int total = (p1 - p0) * nv;
  const bool bf = dtype == DT_BF16, obf = out_dtype == DT_BF16;
  const uint16_t* x0 = static_cast<const uint16_t*>(src.x0) + static_cast<long long>(b) * HW * src.C0;
  const uint16_t* x1 = static_cast<const uint16_t*>(src.x1) + static_cast<long long>(b) * HW * src.C1;
  uint16_t* o = static_cast<uint16_t*>(out) + static_cast<long long>(b) * HW * C;
  for (int e = threadIdx.x; e < total; e += 256 * UNR) {
    uint4 raw[UNR];
    int vv[UNR], px[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int ee = e + u * 256;
      raw[u] = make_uint4(0u, 0u, 0u, 0u);
      px[u] = p0 + ee / nv;
      vv[u] = ee - (ee / nv) * nv;
      if (ee < total) {
        const int c = c0 + vv[u] * 8;
        const uint16_t* ptr = c < src.C0 ? x0 + static_cast<long long>(px[u]) * src.C0 + c
                                         : x1 + static_cast<long long>(px[u]) * src.C1 + (c - src.C0);
        raw[u] = __ldg(reinterpret_cast<const uint4*>(ptr));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (e + u * 256 < total) {
        const float4 a0 = *reinterpret_cast<const float4*>(gn_sm + vv[u] * 8), a1 = *reinterpret_cast<const float4*>(gn_sm + vv[u] * 8 + 4);
        const float4 d0 = *reinterpret_cast<const float4*>(gn_sm + CB + vv[u] * 8), d1 = *reinterpret_cast<const float4*>(gn_sm + CB + vv[u] * 8 + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint32_t wd[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = bf ? unpack_bf16x2(wd[j]) : unpack_f16x2(wd[j]);
          r[j] = gn_affine_silu2(f, a[2 * j], a[2 * j + 1], d[2 * j], d[2 * j + 1], silu != 0, obf);
        }
        *reinterpret_cast<uint4*>(o + static_cast<long long>(px[u]) * C + c0 + vv[u] * 8) = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ row softmax
// out[r, :] = softmax(scale * x[r, :]); one CTA per row; x may be fp32 or 16-bit, out 16-bit. n % 8 == 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const void* __restrict__ x, long long ldx, int in_dtype,
                                                           float scale, int n, void* __restrict__ out, long long ldo,
                                                           int out_dtype) {
  pdl_wait();  // (programmatic dependent launch: inputs of the previous kernel visible from here)
  pdl_launch();
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const int nvec = n >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    const Vec8 a = load8(x, row * ldx + v * 8, in_dtype);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, a.v[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    const Vec8 a = load8(x, row * ldx + v * 8, in_dtype);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf((a.v[j] - mx) * scale);
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  const float inv = 1.f / s;
  for (int v = threadIdx.x; v < nvec; v += 256) {
    Vec8 a = load8(x, row * ldx + v * 8, in_dtype);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] = __expf((a.v[j] - mx) * scale) * inv;
    store8(out, row * ldo + v * 8, a, out_dtype);
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gillb200_layernorm(const void* x, long long ldx, int in_dtype, const float* w, const float* b, float eps,
                                  int rows, int C, void* out, long long ldo, int out_dtype, void* out_lo,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && w && b && out, "null pointer");
  GB_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "layernorm needs C, ldx, ldo multiples of 8");
  GB_CHECK_ARG(C <= 5120, "layernorm supports C <= 5120 (got %d)", C);
  GB_CHECK_ARG(!out_lo || out_dtype == DT_BF16, "out_lo requires bf16 output");
  {
    static int env_p = -1;
    if (env_p < 0) {
      const char* e = getenv("GILLB200_LN_PERSISTENT");  // "0": one-shot CTAs (A/B aid)
      env_p = e ? atoi(e) : 1;
    }
    if (env_p && in_dtype != DT_F32 && out_dtype != DT_F32 && !out_lo && C <= 1280 && rows >= 2048) {
      const uint16_t* xi = static_cast<const uint16_t*>(x);
      uint16_t* oo = static_cast<uint16_t*>(out);
      const int lpr = C <= 320 ? 8 : C <= 640 ? 16 : 32;
      const int nblk = (rows + 4 * (32 / lpr) - 1) / (4 * (32 / lpr));
      const int grid = nblk < 5 * num_sms() ? nblk : 5 * num_sms();  // 94 registers: 5 resident CTAs per SM
      if (lpr == 8)
        GB_CUDA(launch_pdl_light(layernorm_sub16_kernel<8, 5>, dim3(grid), dim3(128), 0, stream, xi, ldx, in_dtype, w, b, eps, rows, C, oo, ldo, out_dtype));
      else if (lpr == 16)
        GB_CUDA(launch_pdl_light(layernorm_sub16_kernel<16, 5>, dim3(grid), dim3(128), 0, stream, xi, ldx, in_dtype, w, b, eps, rows, C, oo, ldo, out_dtype));
      else
        GB_CUDA(launch_pdl_light(layernorm_sub16_kernel<32, 5>, dim3(grid), dim3(128), 0, stream, xi, ldx, in_dtype, w, b, eps, rows, C, oo, ldo, out_dtype));
      GB_COUNT_LAUNCH(1);
      GB_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (C <= 320) {
    GB_CUDA(launch_pdl_light(layernorm_sub_kernel<8, 5>, dim3((rows + 15) / 16), dim3(128), 0, stream, x, ldx, in_dtype, w, b, eps, rows, C, out, ldo,
                                                                     out_dtype, out_lo));
  } else if (C <= 640) {
    GB_CUDA(launch_pdl_light(layernorm_sub_kernel<16, 5>, dim3((rows + 7) / 8), dim3(128), 0, stream, x, ldx, in_dtype, w, b, eps, rows, C, out, ldo,
                                                                    out_dtype, out_lo));
  } else if (C <= 1280) {
    GB_CUDA(launch_pdl_light(layernorm_sub_kernel<32, 5>, dim3((rows + 3) / 4), dim3(128), 0, stream, x, ldx, in_dtype, w, b, eps, rows, C, out, ldo,
                                                                    out_dtype, out_lo));
  } else {
    GB_CUDA(launch_pdl_light(layernorm_kernel<4, 5>, dim3(rows), dim3(128), 0, stream, x, ldx, in_dtype, w, b, eps, rows, C, out, ldo, out_dtype,
                                                      out_lo));
  }
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}

// elementwise pass of both GroupNorm entry points: flat 16-bit form, or the channel-owning form (fp32 tensors, A/B aid)
static int launch_gn_apply(const GnSrc& src, int dtype, int B, int HW, const float* scale_shift, int silu, void* out,
                           int out_dtype, cudaStream_t stream) {
  const int C = src.C0 + src.C1;
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("GILLB200_GN_APPLY");  // "1": round-1 channel-owning kernel
    impl = e ? atoi(e) : 2;
  }
  if (impl == 2 && dtype != DT_F32 && out_dtype != DT_F32 && C / 8 <= 256 * 8) {
    constexpr int UNR = 8;
    const long long total = static_cast<long long>(HW) * (C / 8);
    // ONE wave: at most 4 resident CTAs per SM over the whole batch, equal slices. (Round 1 aimed at 8 CTAs per SM and
    // rounded the slice up to whole unrolled iterations: 64x64 C320 at B = 16 became 640 CTAs on 592 slots -- a second
    // wave of 48 CTAs doubled the kernel's time.) The unrolled loop predicates its tail, so a slice is any multiple of
    // 256 vectors; never less than one full unrolled iteration per CTA.
    static int env_wave = -1;
    if (env_wave < 0) {
      const char* e = getenv("GILLB200_GN_ONEWAVE");  // "0": round-1 sizing (A/B aid)
      env_wave = e ? atoi(e) : 1;
    }
    long long per;
    if (env_wave) {
      int bps = 4 * num_sms() / B;  // CTAs per sample
      if (bps < 1) bps = 1;
      per = (total + bps - 1) / bps;
      per = (per + 255) / 256 * 256;
      if (per < 256LL * UNR) per = 256LL * UNR;
    } else {
      per = (total * B + 8LL * num_sms() - 1) / (8LL * num_sms());
      const long long unit = 256LL * UNR;
      per = (per + unit - 1) / unit * unit;
    }
    const int blocks = static_cast<int>((total + per - 1) / per);
    GB_CUDA(launch_pdl_light(gn_apply2_kernel<UNR>, dim3(blocks, B), dim3(256), static_cast<size_t>(2 * C) * sizeof(float), stream,
                       src, dtype, HW, scale_shift, silu, out, out_dtype, per));
    GB_COUNT_LAUNCH(1);
    return 0;
  }
  const int nvec = C / 8;
  int pix_per_block = (65536 + C - 1) / C;
  const int sgs = nvec >= 256 ? 1 : 256 / nvec;
  const int want_blocks = (8 * num_sms() + B - 1) / B;
  if (pix_per_block * want_blocks > HW) pix_per_block = (HW + want_blocks - 1) / want_blocks;
  pix_per_block = ((pix_per_block + 4 * sgs - 1) / (4 * sgs)) * (4 * sgs);
  const int blocks = (HW + pix_per_block - 1) / pix_per_block;
  GB_CUDA(launch_pdl_light(gn_apply_kernel, dim3(blocks, B), dim3(256), 0, stream, src, dtype, HW, scale_shift, silu, out, out_dtype,
                     pix_per_block));
  GB_COUNT_LAUNCH(1);
  return 0;
}

extern "C" long long gillb200_groupnorm_workspace_bytes(int B, int G) {
  // [B] counters (zero-initialised by the caller, self-resetting) + [B, 128 chunks, G, 2] partial sums
  // + [B, 2, C<=4096] scale/shift
  return (1024 + 128LL * B * G * 2 + 2LL * B * 4096) * sizeof(float);
}

extern "C" int gillb200_groupnorm(const void* x0, int C0, const void* x1, int C1, int dtype, int B, int HW, int G,
                                  const float* w, const float* b, float eps, int silu, void* out, int out_dtype,
                                  void* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x0 && w && b && out && workspace, "null pointer");
  const int C = C0 + C1;
  GB_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C % G == 0 && G <= 64 && C <= 4096, "groupnorm: C0=%d C1=%d G=%d", C0, C1, G);
  GB_CHECK_ARG(C1 == 0 || x1 != nullptr, "groupnorm: second source missing");
  GB_CHECK_ARG(dtype == DT_BF16 || dtype == DT_F16 || dtype == DT_F32, "bad dtype");
  GnSrc src{x0, x1, C0, C1};
  GB_CHECK_ARG(B <= 1024, "groupnorm: batch %d > 1024", B);
  // ~4 CTAs per SM in flight, each thread streaming >= 4 independent 16-byte loads
  int chunks = (4 * num_sms() + B - 1) / B;
  if (chunks > 128) chunks = 128;
  if (chunks > HW) chunks = HW;
  if (chunks < 1) chunks = 1;
  chunks = (HW + (HW + chunks - 1) / chunks - 1) / ((HW + chunks - 1) / chunks);  // drop empty chunks
  int* counters = reinterpret_cast<int*>(workspace);
  float* partial = reinterpret_cast<float*>(workspace) + 1024;
  float* scale_shift = partial + 128LL * B * G * 2;
  const int nvec = C / 8;
  const size_t smem_stats = static_cast<size_t>(nvec >= 256 ? 1 : 256 / nvec) * 2 * C * sizeof(float);
  GB_CUDA(launch_pdl_light(gn_stats_kernel, dim3(dim3(chunks, B)), dim3(256), smem_stats, stream, src, dtype, HW, G, chunks, partial, counters, w, b, eps,
                                                               scale_shift));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return launch_gn_apply(src, dtype, B, HW, scale_shift, silu, out, out_dtype, stream);
}

extern "C" int gillb200_groupnorm_from_stats(const void* x0, int C0, const void* stats0, const void* x1, int C1,
                                             const void* stats1, int dtype, int B, int HW, int G, const float* w,
                                             const float* b, float eps, int silu, void* out, int out_dtype,
                                             void* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x0 && stats0 && w && b && out && workspace, "null pointer");
  const int C = C0 + C1;
  GB_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C % G == 0 && G <= 64 && C <= 4096, "groupnorm: C0=%d C1=%d G=%d", C0, C1, G);
  GB_CHECK_ARG(C1 == 0 || (x1 != nullptr && stats1 != nullptr), "groupnorm: second source / its statistics missing");
  GB_CHECK_ARG(HW % 32 == 0, "groupnorm_from_stats: HW=%d must be a multiple of the 32-row statistics slabs", HW);
  GB_CHECK_ARG(dtype == DT_BF16 || dtype == DT_F16, "groupnorm_from_stats: 16-bit activations only");
  GB_CHECK_ARG(B <= 1024, "groupnorm: batch %d > 1024", B);
  GnSrc src{x0, x1, C0, C1};
  {
    static int env_fused = -1;
    if (env_fused < 0) {
      const char* e = getenv("GILLB200_GN_FUSED");  // "0": always the two-launch form (A/B aid)
      env_fused = e ? atoi(e) : 1;
    }
    // channel block of a CTA: whole groups and whole 16-byte vectors, at least 64 channels when C allows it
    const int cpg = C / G;
    int CB = cpg;
    while (CB % 8) CB += cpg;
    while (CB < 64 && C % (2 * CB) == 0) CB *= 2;
    // measured (tools/gpu_small_level.py through CUDA graphs, B = 16, profiles/r02_gn_fused.log): 8x8 C1280 8.6 -> 5.2 us,
    // 8x8 C2560 9.9 -> 6.7, 16x16 C640 9.3 -> 6.6; from 16x16 C1280 up the flat two-launch form wins (32x32 C640 15.9 vs
    // 19.5 us: 160-byte channel-block segments stream worse than whole rows), so the fused form is kept for small tensors
    if (env_fused && static_cast<long long>(HW) * C <= 200000 && HW <= 1024 && out_dtype != DT_F32 && C % CB == 0 &&
        CB / cpg <= 64) {
      constexpr int UNR = 4;
      const int nblk = C / CB;
      // ~4 CTAs per SM over the whole batch; every CTA at least one full unrolled iteration
      int pchunks = (4 * num_sms() + B * nblk - 1) / (B * nblk);
      const int min_pix = (256 * UNR + CB / 8 - 1) / (CB / 8);
      if (pchunks > (HW + min_pix - 1) / min_pix) pchunks = (HW + min_pix - 1) / min_pix;
      if (pchunks < 1) pchunks = 1;
      const int ppc = (HW + pchunks - 1) / pchunks;
      pchunks = (HW + ppc - 1) / ppc;
      GB_CUDA(launch_pdl_light(gn_fused_kernel<UNR>, dim3(nblk * pchunks, B), dim3(256), static_cast<size_t>(2 * CB) * sizeof(float),
                         stream, src, reinterpret_cast<const float2*>(stats0), reinterpret_cast<const float2*>(stats1), dtype, HW,
                         G, CB, ppc, w, b, eps, silu, out, out_dtype));
      GB_COUNT_LAUNCH(1);
      return 0;
    }
  }
  float* partial = reinterpret_cast<float*>(workspace) + 1024;
  float* scale_shift = partial + 128LL * B * G * 2;  // same workspace layout as gillb200_groupnorm
  GB_CUDA(launch_pdl_light(gn_finalize_stats_kernel, dim3(G, B), dim3(128), 0, stream, reinterpret_cast<const float2*>(stats0), C0,
                     reinterpret_cast<const float2*>(stats1), C1, HW / 32, G, HW, w, b, eps, scale_shift));
  GB_COUNT_LAUNCH(1);
  return launch_gn_apply(src, dtype, B, HW, scale_shift, silu, out, out_dtype, stream);
}

extern "C" int gillb200_groupnorm_scale_shift(int C0, const void* stats0, int C1, const void* stats1, int B, int HW, int G,
                                              const float* w, const float* b, float eps, float* scale_shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = C0 + C1;
  GB_CHECK_ARG(stats0 && w && b && scale_shift, "null pointer");
  GB_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C % G == 0 && G <= 64 && C <= 4096, "groupnorm: C0=%d C1=%d G=%d", C0, C1, G);
  GB_CHECK_ARG(C1 == 0 || stats1 != nullptr, "groupnorm: second source's statistics missing");
  GB_CHECK_ARG(HW % 32 == 0 && B <= 1024, "groupnorm_scale_shift: HW=%d must be a multiple of 32, B <= 1024", HW);
  GB_CUDA(launch_pdl_light(gn_finalize_stats_kernel, dim3(G, B), dim3(128), 0, stream, reinterpret_cast<const float2*>(stats0), C0,
                           reinterpret_cast<const float2*>(stats1), C1, HW / 32, G, HW, w, b, eps, scale_shift));
  GB_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int gillb200_softmax_rows(const void* x, long long ldx, int in_dtype, float scale, long long rows, int n,
                                     void* out, long long ldo, int out_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GB_CHECK_ARG(x && out && n % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "softmax_rows: n, ldx, ldo multiples of 8");
  GB_CHECK_ARG(rows > 0 && rows < (1LL << 31), "softmax_rows: bad row count");
  GB_CUDA(launch_pdl_light(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0, stream, x, ldx, in_dtype, scale, n, out, ldo, out_dtype));
  GB_COUNT_LAUNCH(1);
  GB_CUDA(cudaGetLastError());
  return 0;
}
