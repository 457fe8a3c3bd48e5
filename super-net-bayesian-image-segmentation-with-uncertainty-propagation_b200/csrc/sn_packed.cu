// FAST mode bandwidth-bound kernels around the tensor-core convolution.  All of them work on the packed
// moment layout [n][h][w][3][c] bf16 (planes mean_hi, mean_lo, variance per pixel) with 16-byte accesses:
// pack / unpack / fill, weight preparation, the first convolution (fp32 image in, Cin <= 8), the arg-max
// pooling and the final 1x1 convolution fused with the softmax-Jacobian variance.
#include "sn_common.cuh"
#include "sn_sm100.cuh"

#include <cuda.h>

#include <stdlib.h>
#include <mutex>

namespace sn {

__device__ __forceinline__ float blo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bhi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 8 fp32 means -> hi and lo words (4 x bf16x2 each)
__device__ __forceinline__ void split8(const float (&m)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(m[2 * j]), h1 = __float2bfloat16_rn(m[2 * j + 1]);
    h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[j] = pk2(m[2 * j] - __bfloat162float(h0), m[2 * j + 1] - __bfloat162float(h1));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pk2(v[0], v[1]), pk2(v[2], v[3]), pk2(v[4], v[5]), pk2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = blo(u.x); f[1] = bhi(u.x); f[2] = blo(u.y); f[3] = bhi(u.y);
  f[4] = blo(u.z); f[5] = bhi(u.z); f[6] = blo(u.w); f[7] = bhi(u.w);
}

// ---------------------------------------------------------------------------------------------------------
// pack / unpack / fill: one thread per (pixel, 8-channel group)
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_kernel(size_t pixels, int c, const float* __restrict__ mu, const float* __restrict__ var,
                            __nv_bfloat16* __restrict__ out) {
  const int g = c / 8;
  const size_t total = pixels * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t pix = i / g;
    const int c8 = (int)(i - pix * g) * 8;
    float m[8], v[8];
    const float4 a = *reinterpret_cast<const float4*>(mu + pix * c + c8);
    const float4 b = *reinterpret_cast<const float4*>(mu + pix * c + c8 + 4);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    if (var) {
      const float4 e = *reinterpret_cast<const float4*>(var + pix * c + c8);
      const float4 f = *reinterpret_cast<const float4*>(var + pix * c + c8 + 4);
      v[0] = e.x; v[1] = e.y; v[2] = e.z; v[3] = e.w; v[4] = f.x; v[5] = f.y; v[6] = f.z; v[7] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    uint4 hi, lo;
    split8(m, hi, lo);
    __nv_bfloat16* o = out + pix * 3 * c + c8;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + c) = lo;
    *reinterpret_cast<uint4*>(o + 2 * c) = pack8(v);
  }
}

__global__ void unpack_kernel(size_t pixels, int c, const __nv_bfloat16* __restrict__ in, float* __restrict__ mu,
                              float* __restrict__ var) {
  const int g = c / 8;
  const size_t total = pixels * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t pix = i / g;
    const int c8 = (int)(i - pix * g) * 8;
    const __nv_bfloat16* s = in + pix * 3 * c + c8;
    float h[8], l[8], v[8];
    unpack8(*reinterpret_cast<const uint4*>(s), h);
    unpack8(*reinterpret_cast<const uint4*>(s + c), l);
    unpack8(*reinterpret_cast<const uint4*>(s + 2 * c), v);
    float* pm = mu + pix * c + c8;
    *reinterpret_cast<float4*>(pm) = make_float4(h[0] + l[0], h[1] + l[1], h[2] + l[2], h[3] + l[3]);
    *reinterpret_cast<float4*>(pm + 4) = make_float4(h[4] + l[4], h[5] + l[5], h[6] + l[6], h[7] + l[7]);
    if (var) {
      float* pv = var + pix * c + c8;
      *reinterpret_cast<float4*>(pv) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(pv + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

__global__ void packed_fill_kernel(size_t pixels, int c, float var_fill, __nv_bfloat16* __restrict__ out) {
  const int g = c / 8;
  const size_t total = pixels * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint32_t vv = pk2(var_fill, var_fill);
  const uint4 vfill = make_uint4(vv, vv, vv, vv), zero = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t pix = i / g;
    const int c8 = (int)(i - pix * g) * 8;
    __nv_bfloat16* o = out + pix * 3 * c + c8;
    *reinterpret_cast<uint4*>(o) = zero;
    *reinterpret_cast<uint4*>(o + c) = zero;
    *reinterpret_cast<uint4*>(o + 2 * c) = vfill;
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight preparation: HWIO fp32 -> [3][taps][cout][cin] bf16 (W_hi, W_lo, W^2), s = softplus(w_sigma)
// ---------------------------------------------------------------------------------------------------------
__global__ void prepare_weights_kernel(const float* __restrict__ w, const float* __restrict__ ws, int k, int cin,
                                       int cout, int upconv, __nv_bfloat16* __restrict__ out,
                                       float* __restrict__ s_out) {
  const int taps = k * k;
  const size_t plane = (size_t)taps * cout * cin;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < plane; i += stride) {
    const int ci = (int)(i % cin);
    size_t t = i / cin;
    const int n = (int)(t % cout);
    const int tap = (int)(t / cout);
    int kh = tap / k, kw = tap - kh * k;
    if (upconv) { kh = 1 - kh; kw = 1 - kw; }     // parity (a,b) <- W[1-a, 1-b]
    const float v = w[(((size_t)kh * k + kw) * cin + ci) * cout + n];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    out[i] = hi;
    out[plane + i] = lo;
    out[2 * plane + i] = __float2bfloat16_rn(v * v);   // squared in fp32, rounded once
  }
  for (size_t i = t0; i < (size_t)cout; i += stride) s_out[i] = softplus_f(ws[i]);
}

// ---------------------------------------------------------------------------------------------------------
// first convolution (Brats.py:65-76): fp32 image, Cin <= 8, k <= 3; thread = (pixel, 8 output channels)
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) first_conv_packed_kernel(int B, int H, int W, int cin, int cout, int k,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ ws, sn_packed_view dst,
                                                                int relu) {
  extern __shared__ float sm[];            // weights [K][cout], then s[cout]
  const int K = k * k * cin;
  float* sw = sm;
  float* ss = sm + K * cout;
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) ss[i] = softplus_f(ws[i]);
  __syncthreads();
  const int Ho = H - k + 1, Wo = W - k + 1;
  const int g = cout / 8;
  const size_t total = (size_t)B * Ho * Wo * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int n8 = (int)(i % g) * 8;
    size_t t = i / g;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    float r = 0.f;
    int kk = 0;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const float* px = x + (((size_t)b * H + yo + kh) * W + xo + kw) * cin;
        for (int ci = 0; ci < cin; ++ci, ++kk) {
          const float xv = __ldg(px + ci);
          r = fmaf(xv, xv, r);
          const float4 w0 = *reinterpret_cast<const float4*>(sw + kk * cout + n8);
          const float4 w1 = *reinterpret_cast<const float4*>(sw + kk * cout + n8 + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    float var[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      var[j] = ss[n8 + j] * r;
      if (relu) {
        var[j] = acc[j] > 0.f ? var[j] : 0.f;
        acc[j] = fmaxf(acc[j], 0.f);
      }
    }
    uint4 hi, lo;
    split8(acc, hi, lo);
    __nv_bfloat16* o = out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0 + n8;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + dst.c) = lo;
    *reinterpret_cast<uint4*>(o + 2 * dst.c) = pack8(var);
  }
}

// Specialisation for the shapes the two networks use (k = 3, 32 output channels, Cin = 4 or 1): one thread owns
// PPT = 2 output pixels and all 32 channels.  The 9*CIN input values of each pixel sit in registers; the weights
// are read from shared memory with warp-uniform (broadcast) 16-byte loads, one LDS.128 per 4*PPT FMAs, which is
// what makes the kernel FMA-bound rather than LDS-bound.  When the destination is a whole buffer the block's
// 256 pixels x 192 B are one contiguous 48 KB run: they are staged in shared memory and written with fully
// coalesced 16-byte stores.
constexpr int FC_PPT = 2, FC_THREADS = 128, FC_PIX = FC_PPT * FC_THREADS;

template <int CIN>
__global__ void __launch_bounds__(FC_THREADS) first_conv_k3c32_kernel(int B, int H, int W,
                                                                      const float* __restrict__ x,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ ws,
                                                                      sn_packed_view dst, int relu, int contiguous) {
  constexpr int K = 9 * CIN, COUT = 32;
  constexpr int ROWB = 3 * COUT * 2 + 16;          // one pixel = 192 B, padded to 208 B: conflict-free 16-B accesses
  extern __shared__ __align__(16) uint8_t fc_smem[];
  float* sw = reinterpret_cast<float*>(fc_smem);               // [K][COUT]
  float* ss = sw + K * COUT;                                   // [COUT]
  uint8_t* stage = fc_smem + (K * COUT + COUT) * sizeof(float);  // [FC_PIX][ROWB]
  for (int i = threadIdx.x; i < K * COUT; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < COUT) ss[threadIdx.x] = softplus_f(ws[threadIdx.x]);
  __syncthreads();
  const int Ho = H - 2, Wo = W - 2;
  const size_t total = (size_t)B * Ho * Wo;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  for (size_t base = (size_t)blockIdx.x * FC_PIX; base < total; base += (size_t)gridDim.x * FC_PIX) {
    float xv[FC_PPT][K];
    float r[FC_PPT];
    uint8_t* o[FC_PPT];
    bool live[FC_PPT];
#pragma unroll
    for (int q = 0; q < FC_PPT; ++q) {
      const int lp = threadIdx.x + q * FC_THREADS;             // pixel slot inside the block's run
      const size_t i = base + lp;
      live[q] = i < total;
      int xo = 0, yo = 0, b = 0;
      if (live[q]) {
        xo = (int)(i % Wo);
        size_t t = i / Wo;
        yo = (int)(t % Ho);
        b = (int)(t / Ho);
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* px = x + (((size_t)b * H + yo + kh) * W + xo + kw) * CIN;
          if constexpr (CIN == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(px));
            xv[q][(kh * 3 + kw) * 4 + 0] = v.x; xv[q][(kh * 3 + kw) * 4 + 1] = v.y;
            xv[q][(kh * 3 + kw) * 4 + 2] = v.z; xv[q][(kh * 3 + kw) * 4 + 3] = v.w;
          } else {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xv[q][(kh * 3 + kw) * CIN + c] = __ldg(px + c);
          }
        }
      float rr = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) rr = fmaf(xv[q][k], xv[q][k], rr);
      r[q] = rr;
      o[q] = contiguous ? stage + lp * ROWB
                        : reinterpret_cast<uint8_t*>(
                              out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0);
    }
    const int plane_b = contiguous ? COUT * 2 : dst.c * 2;
#pragma unroll
    for (int n8 = 0; n8 < COUT; n8 += 8) {
      float acc[FC_PPT][8];
#pragma unroll
      for (int q = 0; q < FC_PPT; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + k * COUT + n8);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + k * COUT + n8 + 4);
#pragma unroll
        for (int q = 0; q < FC_PPT; ++q) {
          const float xk = xv[q][k];
          // FFMA2 (fma.rn.f32x2): the kernel is bound by the 1152 FMAs per pixel, two per issue slot halves them
          fma2(acc[q][0], acc[q][1], xk, w0.x, w0.y);
          fma2(acc[q][2], acc[q][3], xk, w0.z, w0.w);
          fma2(acc[q][4], acc[q][5], xk, w1.x, w1.y);
          fma2(acc[q][6], acc[q][7], xk, w1.z, w1.w);
        }
      }
#pragma unroll
      for (int q = 0; q < FC_PPT; ++q) {
        float var[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          var[j] = ss[n8 + j] * r[q];
          if (relu) {
            var[j] = acc[q][j] > 0.f ? var[j] : 0.f;
            acc[q][j] = fmaxf(acc[q][j], 0.f);
          }
        }
        uint4 hi, lo;
        split8(acc[q], hi, lo);
        if (live[q] || contiguous) {
          *reinterpret_cast<uint4*>(o[q] + n8 * 2) = hi;
          *reinterpret_cast<uint4*>(o[q] + plane_b + n8 * 2) = lo;
          *reinterpret_cast<uint4*>(o[q] + 2 * plane_b + n8 * 2) = pack8(var);
        }
      }
    }
    if (contiguous) {
      __syncthreads();
      const size_t remain = total - base;
      const int chunks = (int)(remain < FC_PIX ? remain : FC_PIX) * 12;
      uint4* g = reinterpret_cast<uint4*>(out + base * 3 * COUT);
      for (int c = threadIdx.x; c < chunks; c += FC_THREADS) {
        const int pix = c / 12, part = c - pix * 12;
        g[c] = *reinterpret_cast<const uint4*>(stage + pix * ROWB + part * 16);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// First convolution on the tensor cores (k = 3, 32 output channels, Cin = 4 or 1; myConv_input, Brats.py:65-76).
// GEMM view: M = 128 consecutive output pixels, K = 9*Cin (<= 36, zero-padded to 16/48), N = 32.  The input is the
// fp32 image, so there is no packed tile for TMA to fetch: each thread builds its pixel's im2col row itself (nine
// 16-byte loads), splits it into bf16 hi/lo and writes both A planes into shared memory in the canonical
// SWIZZLE_128B K-major layout (128-byte rows, 16-byte chunk c of row r stored at chunk c ^ (r & 7)); the weights sit
// in the same layout as B = [W_hi ; W_lo] (64 rows).  mean = hi x [W_hi;W_lo] (one N = 64 UMMA per K step, the two
// column halves added in the epilogue) + lo x W_hi; the variance is rank-1, s_n * sum(x^2) over the patch, from the
// thread's own fp32 values.  One role per CTA (build -> UMMA -> epilogue); 5 CTAs per SM overlap the phases.
// The CUDA-core version above spends ~1500 instructions per pixel on the 1152 FMAs; here the tensor core does them.
constexpr int FT_THREADS = 128;
constexpr int FT_ROWB = 3 * 32 * 2 + 16;       // staged output pixel: 192 B + 16 B pad (conflict-free 16-B accesses)
constexpr int FT_SMEM = 1024 + 2 * 16384 + 8192 + 256;   // the output stage (26 KB) reuses the two A planes (32 KB)
constexpr int FT_CTAS_PER_SM = 5;

template <int CIN>
__global__ void __launch_bounds__(FT_THREADS) first_conv_tc_kernel(int B, int H, int W, const float* __restrict__ x,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ ws, sn_packed_view dst,
                                                                   int relu, int contiguous) {
  constexpr int K = 9 * CIN, COUT = 32;
  constexpr int KSTEPS = (K + 15) / 16;
  constexpr int CHUNKS = KSTEPS * 2;                   // 16-byte chunks (8 bf16) per A row actually read
  extern __shared__ uint8_t ft_raw[];
  const uint32_t base = (ptx::smem_u32(ft_raw) + 1023u) & ~1023u;
  uint8_t* gen = ft_raw + (base - ptx::smem_u32(ft_raw));
  uint8_t* a_hi = gen;                                 // [128 rows][128 B]
  uint8_t* a_lo = gen + 16384;
  uint8_t* b_w = gen + 32768;                          // [64 rows][128 B]: W_hi rows 0-31, W_lo rows 32-63
  uint8_t* stage = gen;                                // [128][FT_ROWB]: over the A planes, dead once the UMMAs are done
  const uint32_t bar = base + 32768 + 8192;
  const uint32_t tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 32768 + 8192 + 8);
  __shared__ float ss[COUT];
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- B operand: element (n, k) = W[k][n] (HWIO: k = (kh*3+kw)*CIN + c), hi in rows 0-31, lo in rows 32-63
  for (int i = tid; i < 64 * 8; i += FT_THREADS) {     // (row, chunk): 8 bf16 each
    const int row = i >> 3, c = i & 7;
    const int n = row & 31;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = c * 8 + e * 2 + h;
        float wv = k < K ? w[k * COUT + n] : 0.f;
        const float hi = __bfloat162float(__float2bfloat16_rn(wv));
        f[h] = row < 32 ? hi : wv - hi;
      }
      v[e] = pk2(f[0], f[1]);
    }
    *reinterpret_cast<uint4*>(b_w + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)) =
        make_uint4(v[0], v[1], v[2], v[3]);
  }
  if (tid < COUT) ss[tid] = softplus_f(ws[tid]);
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  const int Ho = H - 2, Wo = W - 2;
  const size_t total = (size_t)B * Ho * Wo;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  constexpr uint32_t idesc64 = ptx::idesc_bf16_f32(128, 64), idesc32 = ptx::idesc_bf16_f32(128, 32);
  uint32_t parity = 0;
  for (size_t tile0 = (size_t)blockIdx.x * 128; tile0 < total; tile0 += (size_t)gridDim.x * 128) {
    // ---- build this thread's im2col row (row = tid), keep r = sum x^2 for the variance
    const size_t i = tile0 + tid;
    const bool live = i < total;
    int xo = 0, yo = 0, b = 0;
    if (live) {
      xo = (int)(i % Wo);
      size_t t = i / Wo;
      yo = (int)(t % Ho);
      b = (int)(t / Ho);
    }
    float xv[KSTEPS * 16];
#pragma unroll
    for (int k = K; k < KSTEPS * 16; ++k) xv[k] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float* px = x + (((size_t)b * H + yo + kh) * W + xo + kw) * CIN;
        if constexpr (CIN == 4) {
          const float4 v = live ? __ldg(reinterpret_cast<const float4*>(px)) : make_float4(0.f, 0.f, 0.f, 0.f);
          xv[(kh * 3 + kw) * 4 + 0] = v.x; xv[(kh * 3 + kw) * 4 + 1] = v.y;
          xv[(kh * 3 + kw) * 4 + 2] = v.z; xv[(kh * 3 + kw) * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int c = 0; c < CIN; ++c) xv[(kh * 3 + kw) * CIN + c] = live ? __ldg(px + c) : 0.f;
        }
      }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) r = fmaf(xv[k], xv[k], r);
    {
      uint8_t* rh = a_hi + (tid >> 3) * 1024 + (tid & 7) * 128;
      uint8_t* rl = a_lo + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float f0 = xv[c * 8 + 2 * e], f1 = xv[c * 8 + 2 * e + 1];
          h[e] = pk2(f0, f1);
          l[e] = pk2(f0 - blo(h[e]), f1 - bhi(h[e]));
        }
        const int pc = (c ^ (tid & 7)) << 4;
        *reinterpret_cast<uint4*>(rh + pc) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(rl + pc) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    ptx::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint64_t da_hi = ptx::smem_desc_kmajor<128>(base + ks * 32);
        const uint64_t da_lo = ptx::smem_desc_kmajor<128>(base + 16384 + ks * 32);
        const uint64_t db = ptx::smem_desc_kmajor<128>(base + 32768 + ks * 32);
        ptx::umma_bf16(tmem, da_hi, db, idesc64, ks > 0 ? 1u : 0u);      // hi x [W_hi ; W_lo] -> columns [0, 64)
        ptx::umma_bf16(tmem, da_lo, db, idesc32, 1u);                     // lo x W_hi          -> columns [0, 32)
      }
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, parity);
    parity ^= 1u;
    ptx::tc_fence_after();
    // ---- epilogue: thread = pixel row = TMEM lane
    uint8_t* o = contiguous == 1 ? stage + tid * FT_ROWB
                            : reinterpret_cast<uint8_t*>(
                                  out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0);
    const int plane_b = contiguous == 1 ? COUT * 2 : dst.c * 2;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 16) {
      uint32_t a0[16], a1[16];
      ptx::tmem_ld16(lane_base + c0, a0);
      ptx::tmem_ld16(lane_base + 32 + c0, a1);
      ptx::tmem_ld_wait();
      float mu[16], var[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float m = __uint_as_float(a0[j]) + __uint_as_float(a1[j]);
        float v = ss[c0 + j] * r;
        if (relu) {
          v = m > 0.f ? v : 0.f;
          m = fmaxf(m, 0.f);
        }
        mu[j] = m;
        var[j] = v;
      }
      if (contiguous == 2) {
        // direct 256-bit stores: a lane writes its 16 channels of a plane as ONE full 32-byte sector (STG.E.ENL2.256,
        // what the halo kernel's epilogue does) -- no staging pass, no copy-out loop, one block-wide barrier less
        if (live) {
          uint32_t hi[8], lo[8], vr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = pk2(mu[2 * j], mu[2 * j + 1]);
            lo[j] = pk2(mu[2 * j] - blo(hi[j]), mu[2 * j + 1] - bhi(hi[j]));
            vr[j] = pk2(var[2 * j], var[2 * j + 1]);
          }
          ptx::st_global_v8(o + c0 * 2, hi);
          ptx::st_global_v8(o + plane_b + c0 * 2, lo);
          ptx::st_global_v8(o + 2 * plane_b + c0 * 2, vr);
        }
      } else if (live || contiguous) {
#pragma unroll
        for (int h8 = 0; h8 < 16; h8 += 8) {
          float m8[8], v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { m8[j] = mu[h8 + j]; v8[j] = var[h8 + j]; }
          uint4 hi, lo;
          split8(m8, hi, lo);
          *reinterpret_cast<uint4*>(o + (c0 + h8) * 2) = hi;
          *reinterpret_cast<uint4*>(o + plane_b + (c0 + h8) * 2) = lo;
          *reinterpret_cast<uint4*>(o + 2 * plane_b + (c0 + h8) * 2) = pack8(v8);
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();                       // TMEM reads done (next tile's UMMAs may overwrite); stage complete
    if (contiguous == 1) {
      const size_t remain = total - tile0;
      const int chunks = (int)(remain < 128 ? remain : 128) * 12;
      uint4* g = reinterpret_cast<uint4*>(out + tile0 * 3 * COUT);
      for (int c = tid; c < chunks; c += FT_THREADS) {
        const int pix = c / 12, part = c - pix * 12;
        g[c] = *reinterpret_cast<const uint4*>(stage + pix * FT_ROWB + part * 16);
      }
      __syncthreads();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------------------------------
// The same first convolution, warp-specialised (round 2, last session).  The kernel above runs one role per CTA --
// build, UMMA, epilogue, with three block-wide barriers per tile -- and relies on 5 CTAs per SM to overlap them:
// measured 42 % issue-active and 3.3 TB/s of output against the 6.45 TB/s copy peak (profiles/r02_conv_input_by_line.md).
// Here the three phases are ROLES of one persistent CTA per SM connected by mbarriers:
//   warp 0        UMMA issuer: waits for a built A stage and a free TMEM accumulator stage, issues the tile's UMMAs
//   warps 1-8     two epilogue groups (thread = TMEM lane = pixel), group e on the tiles of accumulator stage e: tcgen05.ld,
//                 + the rank-1 variance s_n * sum x^2 (r comes from the builder through shared memory), ReLU gate, bf16
//                 hi/lo/var split, stores (one group measured as THE serial stage: ~1 500 cycles of dependent work per tile)
//   warps 9-20    three builder groups of four warps (thread = pixel), group g on the CTA's tiles g, g+3, ...: nine 16-byte
//                 loads, hi/lo split, the im2col row written in the SWIZZLE_128B K-major layout of one of four A stages
// so the global-load latency of tile t+1/t+2, the UMMAs of tile t and the stores of tile t-1 overlap inside one CTA.
// Same operands, same UMMA order, same epilogue arithmetic as first_conv_tc_kernel: bit-identical output.
// Measured (B200, batch 64, L2 flushed): 0.141-0.147 ms against 0.141 ms for the kernel above -- with one epilogue group
// 0.191 ms (the epilogue, ~1 500 cycles of dependent work per tile, was the serial stage), with two or three groups
// 0.153 ms in the bench; direct 256-bit stores instead of the TMA store 0.159-0.172 ms; a linear staging image + plain bulk
// copy (4-way bank conflicts) 0.229 ms.  Not the default: see sn_first_conv_fwd_packed.
constexpr int FWS_GROUPS = 3, FWS_EGROUPS = 2;                  // builder groups, epilogue groups (one per TMEM stage)
constexpr int FWS_W_BUILD = 1 + 4 * FWS_EGROUPS;                // first builder warp
constexpr int FW_WARPS = FWS_W_BUILD + 4 * FWS_GROUPS, FW_THREADS_WS = FW_WARPS * 32;
constexpr int FWS_STAGES = 4, FWS_RSLOTS = FWS_STAGES + 4;      // r slots: a builder may run STAGES tiles ahead of the UMMAs,
                                                                // the epilogue reads r up to EGROUPS tiles (TMEM stages) later
constexpr int FWS_TILE_OUT = 128 * 3 * 32 * 2;                  // one tile's packed output: 128 pixels x 192 B, contiguous
constexpr int FWS_OFF_OUT = FWS_STAGES * 32768 + 8192 + 5120;    // staging buffers, 1 KB aligned (SWIZZLE_64B pattern: 512 B)
constexpr int FWS_SMEM = 1024 + FWS_OFF_OUT + FWS_EGROUPS * FWS_TILE_OUT;
constexpr int FWS_TMEM_COLS = FWS_EGROUPS * 64 <= 128 ? 128 : 256;

template <int CIN>
__global__ void __launch_bounds__(FW_THREADS_WS, 1) first_conv_ws_kernel(const __grid_constant__ CUtensorMap dmap, int B,
                                                                        int H, int W, const float* __restrict__ x,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ ws, sn_packed_view dst,
                                                                        int relu, int v8, int bulk) {
  // bulk: the destination is the whole buffer, so a tile's 128 pixels x 3 planes x 64 B are ONE box of the tensor
  // (channel, plane, pixel): the epilogue stages them in shared memory in the SWIZZLE_64B layout (the 192-byte pixel stride
  // is then bank-conflict free) and ONE TMA tensor store moves the 24 KB; the map clips the last, partial tile.  Why: a
  // lane-per-pixel 256-bit store puts one 32-byte sector of 32 different lines into every LSU wavefront (768 wavefronts
  // per tile), and both forms of this kernel topped out at ~3.3 TB/s of output whatever overlapped with the stores; a
  // linear staging image (4-way bank conflicts on the 192-byte stride) was measured slower still (0.229 ms).
  constexpr int K = 9 * CIN, COUT = 32;
  constexpr int KSTEPS = (K + 15) / 16;
  constexpr int CHUNKS = KSTEPS * 2;
  extern __shared__ uint8_t fws_raw[];
  const uint32_t base = (ptx::smem_u32(fws_raw) + 1023u) & ~1023u;
  uint8_t* gen = fws_raw + (base - ptx::smem_u32(fws_raw));
  constexpr int OFF_B = FWS_STAGES * 32768, OFF_R = OFF_B + 8192, OFF_BAR = OFF_R + FWS_RSLOTS * 128 * 4;
  constexpr int OFF_OUT = FWS_OFF_OUT;                          // one staging buffer of a tile's output per epilogue group
  static_assert(OFF_BAR + 256 <= OFF_OUT, "barrier block overlaps the staging buffers");
  uint8_t* b_w = gen + OFF_B;
  float* r_buf = reinterpret_cast<float*>(gen + OFF_R);
  const uint32_t bar = base + OFF_BAR;
  auto a_full = [&](int s) { return bar + 8u * s; };
  auto a_empty = [&](int s) { return bar + 8u * (FWS_STAGES + s); };
  auto acc_full = [&](int s) { return bar + 8u * (2 * FWS_STAGES + s); };
  auto acc_empty = [&](int s) { return bar + 8u * (2 * FWS_STAGES + FWS_EGROUPS + s); };
  const uint32_t tmem_slot = bar + 8u * (2 * FWS_STAGES + 2 * FWS_EGROUPS);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + 8 * (2 * FWS_STAGES + 2 * FWS_EGROUPS));
  __shared__ float ss[COUT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- B operand, once per CTA: element (n, k) = W[k][n], hi in rows 0-31, lo in rows 32-63 (as in the kernel above)
  for (int i = tid; i < 64 * 8; i += FW_THREADS_WS) {
    const int row = i >> 3, c = i & 7;
    const int n = row & 31;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = c * 8 + e * 2 + h;
        float wv = k < K ? w[k * COUT + n] : 0.f;
        const float hi = __bfloat162float(__float2bfloat16_rn(wv));
        f[h] = row < 32 ? hi : wv - hi;
      }
      v[e] = pk2(f[0], f[1]);
    }
    *reinterpret_cast<uint4*>(b_w + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)) =
        make_uint4(v[0], v[1], v[2], v[3]);
  }
  if (tid < COUT) ss[tid] = softplus_f(ws[tid]);
  // the A stages' unused K columns (chunks >= CHUNKS are never read; chunks < CHUNKS are always written) need no clearing
  if (tid == 0) {
    for (int s = 0; s < FWS_STAGES; ++s) {
      ptx::mbar_init(a_full(s), 128);       // every builder thread of the tile's group arrives (after its own proxy fence)
      ptx::mbar_init(a_empty(s), 1);        // UMMA commit
    }
    for (int s = 0; s < FWS_EGROUPS; ++s) {
      ptx::mbar_init(acc_full(s), 1);       // UMMA commit
      ptx::mbar_init(acc_empty(s), 4);      // the four warps of the stage's epilogue group
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, FWS_TMEM_COLS);      // one accumulator stage of 64 columns per epilogue group
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  const int Ho = H - 2, Wo = W - 2;
  const uint32_t total = (uint32_t)B * Ho * Wo;                 // < 2^31 (checked by the host)
  const uint32_t n_tiles = (total + 127u) / 128u;
  const int my_tiles = n_tiles > blockIdx.x ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

  if (warp == 0) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc64 = ptx::idesc_bf16_f32(128, 64), idesc32 = ptx::idesc_bf16_f32(128, 32);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t % FWS_STAGES, as = t % FWS_EGROUPS;
        ptx::mbar_wait(acc_empty(as), (((uint32_t)(t / FWS_EGROUPS)) & 1u) ^ 1u);
        ptx::mbar_wait(a_full(s), ((uint32_t)(t / FWS_STAGES)) & 1u);
        ptx::tc_fence_after();
        const uint32_t a_hi = base + s * 32768, a_lo = a_hi + 16384, acc = tmem + as * 64;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t da_hi = ptx::smem_desc_kmajor<128>(a_hi + ks * 32);
          const uint64_t da_lo = ptx::smem_desc_kmajor<128>(a_lo + ks * 32);
          const uint64_t db = ptx::smem_desc_kmajor<128>(base + OFF_B + ks * 32);
          ptx::umma_bf16(acc, da_hi, db, idesc64, ks > 0 ? 1u : 0u);      // hi x [W_hi ; W_lo] -> columns [0, 64)
          ptx::umma_bf16(acc, da_lo, db, idesc32, 1u);                     // lo x W_hi          -> columns [0, 32)
        }
        ptx::umma_commit(a_empty(s));
        ptx::umma_commit(acc_full(as));
      }
    }
  } else if (warp < FWS_W_BUILD) {
    // ===================== epilogue: thread = pixel row = TMEM lane; group eg drains accumulator stage eg =====================
    const int q = warp & 3;                                      // the TMEM lane quarter this warp may read
    const int eg = (warp - 1) >> 2;
    const bool elected = ((warp - 1) & 3) == 0 && lane == 0;
    const int row = q * 32 + lane;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
    for (int t = eg; t < my_tiles; t += FWS_EGROUPS) {
      const int as = eg;
      const uint32_t i = ((uint32_t)blockIdx.x + (uint32_t)t * gridDim.x) * 128u + (uint32_t)row;
      const bool live = i < total;
      uint32_t xo = 0, yo = 0, b = 0;
      if (live) {
        xo = i % (uint32_t)Wo;
        const uint32_t tt = i / (uint32_t)Wo;
        yo = tt % (uint32_t)Ho;
        b = tt / (uint32_t)Ho;
      }
      uint8_t* o = bulk ? nullptr
                        : reinterpret_cast<uint8_t*>(
                              out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0);
      const int plane_b = bulk ? COUT * 2 : dst.c * 2;
      ptx::mbar_wait(acc_full(as), ((uint32_t)(t / FWS_EGROUPS)) & 1u);
      ptx::tc_fence_after();
      const float r = r_buf[(t % FWS_RSLOTS) * 128 + row];
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + as * 64;
      uint32_t a0[2][16], a1[2][16];
      ptx::tmem_ld16(lane_base, a0[0]);
      ptx::tmem_ld16(lane_base + 32, a1[0]);
      ptx::tmem_ld16(lane_base + 16, a0[1]);
      ptx::tmem_ld16(lane_base + 48, a1[1]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty(as));            // the accumulator stage is in registers
      if (bulk) {
        // the group's previous store must have finished reading its staging buffer
        if (elected) ptx::bulk_wait_read0();
        ptx::named_barrier(1 + eg, 128);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = c * 16;
        float mu[16], var[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float m = __uint_as_float(a0[c][j]) + __uint_as_float(a1[c][j]);
          float v = ss[c0 + j] * r;
          if (relu) {
            v = m > 0.f ? v : 0.f;
            m = fmaxf(m, 0.f);
          }
          mu[j] = m;
          var[j] = v;
        }
        if (live || bulk) {
          uint32_t hi[8], lo[8], vr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = pk2(mu[2 * j], mu[2 * j + 1]);
            lo[j] = pk2(mu[2 * j] - blo(hi[j]), mu[2 * j + 1] - bhi(hi[j]));
            vr[j] = pk2(var[2 * j], var[2 * j + 1]);
          }
          if (bulk) {
            // SWIZZLE_64B: 16-byte chunk bits [4,6) ^= address bits [7,9) (the staging buffers are 1 KB aligned)
            const uint32_t stg = base + OFF_OUT + eg * FWS_TILE_OUT;
            auto put = [&](int pl, const uint32_t (&w8)[8]) {
              const uint32_t L = (uint32_t)row * 192u + (uint32_t)pl * 64u + (uint32_t)c0 * 2u;
              const uint32_t x4 = ((L >> 7) & 3u) << 4;
              ptx::st_shared_v4(stg + (L ^ x4), w8[0], w8[1], w8[2], w8[3]);
              ptx::st_shared_v4(stg + ((L + 16u) ^ x4), w8[4], w8[5], w8[6], w8[7]);
            };
            put(0, hi);
            put(1, lo);
            put(2, vr);
          } else if (v8) {
            ptx::st_global_v8(o + c0 * 2, hi);
            ptx::st_global_v8(o + plane_b + c0 * 2, lo);
            ptx::st_global_v8(o + 2 * plane_b + c0 * 2, vr);
          } else {
            uint4* ph = reinterpret_cast<uint4*>(o + c0 * 2);
            uint4* pl = reinterpret_cast<uint4*>(o + plane_b + c0 * 2);
            uint4* pv = reinterpret_cast<uint4*>(o + 2 * plane_b + c0 * 2);
            ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            pv[0] = make_uint4(vr[0], vr[1], vr[2], vr[3]); pv[1] = make_uint4(vr[4], vr[5], vr[6], vr[7]);
          }
        }
      }
      if (bulk) {
        ptx::fence_proxy_async();                                // generic-proxy writes of the staging buffer -> the bulk copy
        ptx::named_barrier(1 + eg, 128);
        if (elected) {
          const uint32_t tile0 = ((uint32_t)blockIdx.x + (uint32_t)t * gridDim.x) * 128u;
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&dmap)),
                       "r"(base + OFF_OUT + eg * FWS_TILE_OUT), "r"(0), "r"(0), "r"((int)tile0)
                       : "memory");
          ptx::bulk_commit_group();
        }
      }
    }
    if (bulk && elected) ptx::bulk_wait_read0();                 // the staging buffers must outlive the copies' reads
  } else {
    // ===================== builders: group g takes the CTA's tiles t = g, g + GROUPS, ... ; thread = pixel row =====================
    const int g = (warp - FWS_W_BUILD) >> 2;
    const int row = ((warp - FWS_W_BUILD) & 3) * 32 + lane;
    for (int t = g; t < my_tiles; t += FWS_GROUPS) {
      const int s = t % FWS_STAGES;
      const uint32_t i = ((uint32_t)blockIdx.x + (uint32_t)t * gridDim.x) * 128u + (uint32_t)row;
      const bool live = i < total;
      uint32_t xo = 0, yo = 0, b = 0;
      if (live) {
        xo = i % (uint32_t)Wo;
        const uint32_t tt = i / (uint32_t)Wo;
        yo = tt % (uint32_t)Ho;
        b = tt / (uint32_t)Ho;
      }
      float xv[KSTEPS * 16];
#pragma unroll
      for (int k = K; k < KSTEPS * 16; ++k) xv[k] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* px = x + (((size_t)b * H + yo + kh) * W + xo + kw) * CIN;
          if constexpr (CIN == 4) {
            const float4 v = live ? __ldg(reinterpret_cast<const float4*>(px)) : make_float4(0.f, 0.f, 0.f, 0.f);
            xv[(kh * 3 + kw) * 4 + 0] = v.x; xv[(kh * 3 + kw) * 4 + 1] = v.y;
            xv[(kh * 3 + kw) * 4 + 2] = v.z; xv[(kh * 3 + kw) * 4 + 3] = v.w;
          } else {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xv[(kh * 3 + kw) * CIN + c] = live ? __ldg(px + c) : 0.f;
          }
        }
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) r = fmaf(xv[k], xv[k], r);
      // the stage must have been drained by the UMMAs of the tile that used it FWS_STAGES tiles ago
      ptx::mbar_wait(a_empty(s), (((uint32_t)(t / FWS_STAGES)) & 1u) ^ 1u);
      uint8_t* rh = gen + s * 32768 + (row >> 3) * 1024 + (row & 7) * 128;
      uint8_t* rl = rh + 16384;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float f0 = xv[c * 8 + 2 * e], f1 = xv[c * 8 + 2 * e + 1];
          h[e] = pk2(f0, f1);
          l[e] = pk2(f0 - blo(h[e]), f1 - bhi(h[e]));
        }
        const int pc = (c ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(rh + pc) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(rl + pc) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      r_buf[(t % FWS_RSLOTS) * 128 + row] = r;
      ptx::fence_proxy_async();                                  // this thread's generic-proxy writes -> the UMMA's async proxy
      ptx::mbar_arrive(a_full(s));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, FWS_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// arg-max pooling on packed windows (Brats.py:171-174,206-216); thread = (output pixel, 8 channels)
// ---------------------------------------------------------------------------------------------------------
__global__ void maxpool_packed_kernel(sn_packed_view src, int B, int H, int W, int c, sn_packed_view dst) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int g = c / 8;
  const size_t total = (size_t)B * Ho * Wo * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.base);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % g) * 8;
    size_t t = i / g;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[8];
    uint32_t bh[8], bl[8], bv[8];      // winning (hi, lo, var) as raw bf16 bits
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bh[j] = bl[j] = bv[j] = 0; }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), xx = 2 * xo + (d & 1);
      if (y < H && xx < W) {
        const __nv_bfloat16* s =
            in + ((((size_t)b * src.h + y + src.y0) * src.w + xx + src.x0) * 3) * src.c + src.c0 + c8;
        const uint4 h = *reinterpret_cast<const uint4*>(s);
        const uint4 l = *reinterpret_cast<const uint4*>(s + src.c);
        const uint4 v = *reinterpret_cast<const uint4*>(s + 2 * src.c);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m0 = blo(hw[e]) + blo(lw[e]), m1 = bhi(hw[e]) + bhi(lw[e]);
          if (m0 > best[2 * e]) {
            best[2 * e] = m0; bh[2 * e] = hw[e] & 0xFFFFu; bl[2 * e] = lw[e] & 0xFFFFu; bv[2 * e] = vw[e] & 0xFFFFu;
          }
          if (m1 > best[2 * e + 1]) {
            best[2 * e + 1] = m1; bh[2 * e + 1] = hw[e] >> 16; bl[2 * e + 1] = lw[e] >> 16; bv[2 * e + 1] = vw[e] >> 16;
          }
        }
      }
    }
    __nv_bfloat16* o =
        out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0 + c8;
    *reinterpret_cast<uint4*>(o) = make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16), bh[4] | (bh[5] << 16),
                                              bh[6] | (bh[7] << 16));
    *reinterpret_cast<uint4*>(o + dst.c) = make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16),
                                                      bl[4] | (bl[5] << 16), bl[6] | (bl[7] << 16));
    *reinterpret_cast<uint4*>(o + 2 * dst.c) = make_uint4(bv[0] | (bv[1] << 16), bv[2] | (bv[3] << 16),
                                                          bv[4] | (bv[5] << 16), bv[6] | (bv[7] << 16));
  }
}


// ---------------------------------------------------------------------------------------------------------
// myReLU (Brats.py:233-238) / plain window copy on packed windows; thread = (pixel, 8 channels).  Inside the engines the
// gate is a conv-epilogue flag and windows are address arithmetic; this kernel serves the layer-by-layer FAST API when a
// ReLU or a pad follows something that is already in memory (fastlayers.py).
// ---------------------------------------------------------------------------------------------------------
__global__ void relu_copy_packed_kernel(sn_packed_view src, int B, int H, int W, int c, sn_packed_view dst, int gate) {
  const int g = c / 8;
  const size_t total = (size_t)B * H * W * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.base);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % g) * 8;
    size_t t = i / g;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const __nv_bfloat16* sp = in + ((((size_t)b * src.h + y + src.y0) * src.w + x + src.x0) * 3) * src.c + src.c0 + c8;
    uint4 h = *reinterpret_cast<const uint4*>(sp);
    uint4 l = *reinterpret_cast<const uint4*>(sp + src.c);
    uint4 v = *reinterpret_cast<const uint4*>(sp + 2 * src.c);
    if (gate) {
      uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // strict mean > 0 (TF ReluGrad), mean = hi + lo
        const uint32_t k0 = (blo(hw[e]) + blo(lw[e])) > 0.f ? 0x0000FFFFu : 0u;
        const uint32_t k1 = (bhi(hw[e]) + bhi(lw[e])) > 0.f ? 0xFFFF0000u : 0u;
        hw[e] &= k0 | k1; lw[e] &= k0 | k1; vw[e] &= k0 | k1;
      }
      h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      v = make_uint4(vw[0], vw[1], vw[2], vw[3]);
    }
    __nv_bfloat16* o = out + ((((size_t)b * dst.h + y + dst.y0) * dst.w + x + dst.x0) * 3) * dst.c + dst.c0 + c8;
    *reinterpret_cast<uint4*>(o) = h;
    *reinterpret_cast<uint4*>(o + dst.c) = l;
    *reinterpret_cast<uint4*>(o + 2 * dst.c) = v;
  }
}

// Same, 16 channels per thread through 256-bit accesses (one full 32-byte sector per lane and instruction): the
// variant used whenever the views are 32-byte aligned (every buffer of the engines is).
__global__ void maxpool_packed16_kernel(sn_packed_view src, int B, int H, int W, int c, sn_packed_view dst) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int g = c / 16;
  const size_t total = (size_t)B * Ho * Wo * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.base);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c16 = (int)(i % g) * 16;
    size_t t = i / g;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[16];
    uint32_t bh[8], bl[8], bv[8];      // winning (hi, lo, var) as packed bf16 pairs
#pragma unroll
    for (int j = 0; j < 16; ++j) best[j] = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) bh[j] = bl[j] = bv[j] = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), xx = 2 * xo + (d & 1);
      if (y < H && xx < W) {
        const __nv_bfloat16* sp =
            in + ((((size_t)b * src.h + y + src.y0) * src.w + xx + src.x0) * 3) * src.c + src.c0 + c16;
        uint32_t hw[8], lw[8], vw[8];
        ptx::ld_global_v8(sp, hw);
        ptx::ld_global_v8(sp + src.c, lw);
        ptx::ld_global_v8(sp + 2 * src.c, vw);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float m0 = blo(hw[e]) + blo(lw[e]), m1 = bhi(hw[e]) + bhi(lw[e]);
          if (m0 > best[2 * e]) {
            best[2 * e] = m0;
            bh[e] = (bh[e] & 0xFFFF0000u) | (hw[e] & 0xFFFFu);
            bl[e] = (bl[e] & 0xFFFF0000u) | (lw[e] & 0xFFFFu);
            bv[e] = (bv[e] & 0xFFFF0000u) | (vw[e] & 0xFFFFu);
          }
          if (m1 > best[2 * e + 1]) {
            best[2 * e + 1] = m1;
            bh[e] = (bh[e] & 0xFFFFu) | (hw[e] & 0xFFFF0000u);
            bl[e] = (bl[e] & 0xFFFFu) | (lw[e] & 0xFFFF0000u);
            bv[e] = (bv[e] & 0xFFFFu) | (vw[e] & 0xFFFF0000u);
          }
        }
      }
    }
    __nv_bfloat16* o =
        out + ((((size_t)b * dst.h + yo + dst.y0) * dst.w + xo + dst.x0) * 3) * dst.c + dst.c0 + c16;
    ptx::st_global_v8(o, bh);
    ptx::st_global_v8(o + dst.c, bl);
    ptx::st_global_v8(o + 2 * dst.c, bv);
  }
}

static bool view_v8(const sn_packed_view* v, int c) {
  return (reinterpret_cast<uintptr_t>(v->base) & 31u) == 0 && v->c % 16 == 0 && v->c0 % 16 == 0 && c % 16 == 0;
}

// ---------------------------------------------------------------------------------------------------------
// final 1x1 convolution + softmax with Jacobian variance (Brats.py:367,454,269-283); thread = pixel
// ---------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) final_conv_softmax_kernel(sn_packed_view src, int B, int H, int W, int cin,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ ws,
                                                                 float* __restrict__ p_out,
                                                                 float* __restrict__ v_out,
                                                                 float* __restrict__ pre_mu,
                                                                 float* __restrict__ pre_var, int contiguous) {
  extern __shared__ __align__(16) float sm[];   // W [cin][C], W^2 [cin][C], s [C] (padded to 4), then the pixel stage
  float* sw = sm;
  float* sw2 = sm + cin * C;
  float* ss = sm + 2 * cin * C;
  const int rowb = 6 * cin + 16;                // one pixel's (hi|lo|var) + 16 B pad: conflict-free 16-B row reads
  uint8_t* stage = reinterpret_cast<uint8_t*>(sm + ((2 * cin * C + C + 3) & ~3));
  for (int i = threadIdx.x; i < cin * C; i += blockDim.x) {
    const float v = w[i];
    sw[i] = v;
    sw2[i] = v * v;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) ss[i] = softplus_f(ws[i]);
  __syncthreads();
  const size_t total = (size_t)B * H * W;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.base);
  for (size_t base = (size_t)blockIdx.x * 128; base < total; base += (size_t)gridDim.x * 128) {
    const size_t i = base + threadIdx.x;
    const bool live = i < total;
    const uint8_t* s8;
    int plane_b;
    if (contiguous) {
      // the block's 128 pixels are one contiguous run of the source: fully coalesced 16-byte loads into smem
      const int cpp = (6 * cin) / 16;           // 16-byte chunks per pixel
      const size_t remain = total - base;
      const int chunks = (int)(remain < 128 ? remain : 128) * cpp;
      const uint4* g = reinterpret_cast<const uint4*>(in + base * 3 * cin);
      for (int c = threadIdx.x; c < chunks; c += 128) {
        const int pix = c / cpp, part = c - pix * cpp;
        *reinterpret_cast<uint4*>(stage + pix * rowb + part * 16) = __ldg(g + c);
      }
      __syncthreads();
      s8 = stage + threadIdx.x * rowb;
      plane_b = cin * 2;
    } else {
      int xx = 0, y = 0, b = 0;
      if (live) {
        xx = (int)(i % W);
        size_t t = i / W;
        y = (int)(t % H);
        b = (int)(t / H);
      }
      s8 = reinterpret_cast<const uint8_t*>(in + ((((size_t)b * src.h + y + src.y0) * src.w + xx + src.x0) * 3) * src.c +
                                            src.c0);
      plane_b = src.c * 2;
    }
    if (live) {
      float m[C], v[C];
#pragma unroll
      for (int j = 0; j < C; ++j) m[j] = v[j] = 0.f;
      float r = 0.f;
      for (int c8 = 0; c8 < cin; c8 += 8) {
        float h[8], l[8], vv[8];
        unpack8(*reinterpret_cast<const uint4*>(s8 + c8 * 2), h);
        unpack8(*reinterpret_cast<const uint4*>(s8 + plane_b + c8 * 2), l);
        unpack8(*reinterpret_cast<const uint4*>(s8 + 2 * plane_b + c8 * 2), vv);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          head_accumulate<C>(h[e] + l[e], vv[e], sw + (c8 + e) * C, sw2 + (c8 + e) * C, m, v, r);
      }
      float p[C], vo[C];
      head_finish<C>(m, v, r, ss, p, vo);
      if constexpr (C == 4) {
        reinterpret_cast<float4*>(p_out)[i] = make_float4(p[0], p[1], p[2], p[3]);
        reinterpret_cast<float4*>(v_out)[i] = make_float4(vo[0], vo[1], vo[2], vo[3]);
        if (pre_mu) {
          reinterpret_cast<float4*>(pre_mu)[i] = make_float4(m[0], m[1], m[2], m[3]);
          reinterpret_cast<float4*>(pre_var)[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
#pragma unroll
        for (int a = 0; a < C; ++a) { p_out[i * C + a] = p[a]; v_out[i * C + a] = vo[a]; }
        if (pre_mu) {
#pragma unroll
          for (int j = 0; j < C; ++j) { pre_mu[i * C + j] = m[j]; pre_var[i * C + j] = v[j]; }
        }
      }
    }
    if (contiguous) __syncthreads();
  }
}

static int check_pview(const sn_packed_view* v, int batch, int h, int w, int c, const char* who) {
  SN_REQUIRE(v && v->base && aligned16(v->base), SN_ERR_BAD_ARG, "%s: null/misaligned packed view", who);
  SN_REQUIRE(v->n >= batch && v->c % 8 == 0 && v->c0 % 8 == 0 && c % 8 == 0, SN_ERR_MISALIGNED,
             "%s: channel counts/offsets must be multiples of 8", who);
  SN_REQUIRE(v->y0 >= 0 && v->x0 >= 0 && v->c0 >= 0 && v->y0 + h <= v->h && v->x0 + w <= v->w && v->c0 + c <= v->c,
             SN_ERR_BAD_ARG, "%s: window outside the buffer", who);
  return SN_OK;
}

}  // namespace sn

using namespace sn;

extern "C" {

size_t sn_packed_bytes(int32_t n, int32_t h, int32_t w, int32_t c) {
  return (size_t)n * h * w * 3 * c * sizeof(__nv_bfloat16);
}

int sn_pack_moments(size_t pixels, int32_t c, const float* mu, const float* var, void* packed, sn_stream_t st) {
  SN_REQUIRE(mu && packed, SN_ERR_BAD_ARG, "pack: null pointer");
  SN_REQUIRE(c > 0 && c % 8 == 0, SN_ERR_MISALIGNED, "pack: channels %d must be a multiple of 8", c);
  SN_REQUIRE(aligned16(mu) && (!var || aligned16(var)) && aligned16(packed), SN_ERR_MISALIGNED, "pack: misaligned");
  if (pixels == 0) return SN_OK;
  pack_kernel<<<ew_grid(pixels * (c / 8), 256), 256, 0, as_stream(st)>>>(pixels, c, mu, var,
                                                                         reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("pack");
}

int sn_unpack_moments(size_t pixels, int32_t c, const void* packed, float* mu, float* var, sn_stream_t st) {
  SN_REQUIRE(mu && packed, SN_ERR_BAD_ARG, "unpack: null pointer");
  SN_REQUIRE(c > 0 && c % 8 == 0, SN_ERR_MISALIGNED, "unpack: channels %d must be a multiple of 8", c);
  SN_REQUIRE(aligned16(mu) && (!var || aligned16(var)) && aligned16(packed), SN_ERR_MISALIGNED, "unpack: misaligned");
  if (pixels == 0) return SN_OK;
  unpack_kernel<<<ew_grid(pixels * (c / 8), 256), 256, 0, as_stream(st)>>>(
      pixels, c, reinterpret_cast<const __nv_bfloat16*>(packed), mu, var);
  return check_launch("unpack");
}

int sn_packed_fill(void* packed, size_t pixels, int32_t c, float var_fill, sn_stream_t st) {
  SN_REQUIRE(packed && aligned16(packed), SN_ERR_BAD_ARG, "packed_fill: null/misaligned pointer");
  SN_REQUIRE(c > 0 && c % 8 == 0, SN_ERR_MISALIGNED, "packed_fill: channels %d must be a multiple of 8", c);
  if (pixels == 0) return SN_OK;
  packed_fill_kernel<<<ew_grid(pixels * (c / 8), 256), 256, 0, as_stream(st)>>>(
      pixels, c, var_fill, reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("packed_fill");
}

size_t sn_prepared_weight_bytes(int32_t ksize, int32_t cin, int32_t cout) {
  return (size_t)3 * ksize * ksize * cin * cout * sizeof(__nv_bfloat16);
}

int sn_prepare_weights(const float* w_mu, const float* w_sigma, int32_t ksize, int32_t cin, int32_t cout,
                       int32_t upconv, void* w_packed, float* s_out, sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && w_packed && s_out, SN_ERR_BAD_ARG, "prepare_weights: null pointer");
  SN_REQUIRE(ksize >= 1 && ksize <= 3 && cin > 0 && cout > 0, SN_ERR_BAD_ARG, "prepare_weights: bad sizes");
  SN_REQUIRE(!upconv || ksize == 2, SN_ERR_BAD_ARG, "prepare_weights: upconv needs ksize == 2");
  size_t n = (size_t)ksize * ksize * cin * cout;
  prepare_weights_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(
      w_mu, w_sigma, ksize, cin, cout, upconv, reinterpret_cast<__nv_bfloat16*>(w_packed), s_out);
  return check_launch("prepare_weights");
}

int sn_first_conv_fwd_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout, int32_t ksize,
                             const float* x, const float* w_mu, const float* w_sigma, const sn_packed_view* dst,
                             int32_t flags, sn_stream_t st) {
  SN_REQUIRE(x && w_mu && w_sigma && dst, SN_ERR_BAD_ARG, "first_conv: null pointer");
  SN_REQUIRE(batch > 0 && cin >= 1 && cin <= 8 && ksize >= 1 && ksize <= 3 && in_h >= ksize && in_w >= ksize,
             SN_ERR_UNSUPPORTED, "first_conv: needs cin <= 8 and k <= 3 (got cin %d, k %d)", cin, ksize);
  SN_REQUIRE(cout % 8 == 0 && cout <= 256, SN_ERR_UNSUPPORTED, "first_conv: cout %d", cout);
  const int Ho = in_h - ksize + 1, Wo = in_w - ksize + 1;
  int rc = check_pview(dst, batch, Ho, Wo, cout, "first_conv dst");
  if (rc) return rc;
  const int relu = (flags & SN_TC_RELU) ? 1 : 0;
  static const bool first_tc = [] {
    const char* e = getenv("SN_FIRST_CONV_TC");
    return e == nullptr || e[0] != '0';
  }();
  if (first_tc && !(flags & SN_TC_EXACT) && ksize == 3 && cout == 32 && (cin == 4 || cin == 1) && aligned16(x)) {
    const size_t pixels = (size_t)batch * Ho * Wo;
    static std::once_flag ft_once;
    std::call_once(ft_once, [] {
      cudaFuncSetAttribute(first_conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
      cudaFuncSetAttribute(first_conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    });
    const size_t tiles = (pixels + 127) / 128;
    const size_t cap = (size_t)num_sms() * FT_CTAS_PER_SM;
    const int grid = (int)(tiles < cap ? tiles : cap);
    int contiguous = dst->y0 == 0 && dst->x0 == 0 && dst->c0 == 0 && dst->h == Ho && dst->w == Wo && dst->c == cout;
    const bool whole_buffer = contiguous != 0;
    // 2: every (pixel, plane, 16-channel chunk) segment of the destination is 32-byte aligned -> direct 256-bit stores
    static const bool first_v8 = [] {
      const char* e = getenv("SN_FIRST_V8");
      return e == nullptr || e[0] != '0';
    }();
    if (first_v8 && (reinterpret_cast<uintptr_t>(dst->base) & 31u) == 0 && dst->c % 16 == 0 && dst->c0 % 16 == 0)
      contiguous = 2;
    // The warp-specialised form measures the SAME 0.141 ms as the one-role-per-CTA kernel at batch 64 (L2 flushed, A/B in
    // one process, profiles/r02_session3.md): neither the overlap of the phases nor the store path is what bounds this
    // layer, ~790 warp instructions per pixel row at ~43 % issue utilisation are.  It is therefore NOT the default;
    // SN_TC_ROWS in `flags` or SN_FIRST_WS=1 selects it (parity-tested bit-identical), SN_TC_IM2COL forces the kernel above.
    static const bool first_ws = [] {
      const char* e = getenv("SN_FIRST_WS");
      return e != nullptr && e[0] == '1';
    }();
    if ((first_ws || (flags & SN_TC_ROWS)) && !(flags & SN_TC_IM2COL) && pixels < (1ull << 31) - 128) {
      static std::once_flag fws_once;
      std::call_once(fws_once, [] {
        cudaFuncSetAttribute(first_conv_ws_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWS_SMEM);
        cudaFuncSetAttribute(first_conv_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWS_SMEM);
      });
      const size_t cap_ws = (size_t)num_sms();
      const int grid_ws = (int)(tiles < cap_ws ? tiles : cap_ws);
      const int v8 = contiguous == 2 ? 1 : 0;
      static const bool first_bulk = [] {
        const char* e = getenv("SN_FIRST_BULK");
        return e == nullptr || e[0] != '0';
      }();
      int bulk = first_bulk && whole_buffer && aligned16(dst->base) ? 1 : 0;
      CUtensorMap dmap{};
      if (bulk) {
        // the whole destination as (channel, plane, pixel): one box = a tile's 128 pixels x 3 planes x 32 channels
        typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        EncFn enc = reinterpret_cast<EncFn>(tensor_map_encoder());
        cuuint64_t dims[3] = {32, 3, (cuuint64_t)pixels};
        cuuint64_t strides[2] = {64, 192};
        cuuint32_t box[3] = {32, 3, 128};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc == nullptr ||
            enc(&dmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dst->base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          bulk = 0;                          // no driver entry point: direct stores
      }
      if (cin == 4)
        first_conv_ws_kernel<4><<<grid_ws, FW_THREADS_WS, FWS_SMEM, as_stream(st)>>>(dmap, batch, in_h, in_w, x, w_mu,
                                                                                    w_sigma, *dst, relu, v8, bulk);
      else
        first_conv_ws_kernel<1><<<grid_ws, FW_THREADS_WS, FWS_SMEM, as_stream(st)>>>(dmap, batch, in_h, in_w, x, w_mu,
                                                                                    w_sigma, *dst, relu, v8, bulk);
      return check_launch("first_conv_ws");
    }
    if (cin == 4)
      first_conv_tc_kernel<4><<<grid, FT_THREADS, FT_SMEM, as_stream(st)>>>(batch, in_h, in_w, x, w_mu, w_sigma, *dst,
                                                                            relu, contiguous);
    else
      first_conv_tc_kernel<1><<<grid, FT_THREADS, FT_SMEM, as_stream(st)>>>(batch, in_h, in_w, x, w_mu, w_sigma, *dst,
                                                                            relu, contiguous);
    return check_launch("first_conv_tc");
  }
  if (ksize == 3 && cout == 32 && (cin == 4 || cin == 1)) {
    const size_t pixels = (size_t)batch * Ho * Wo;
    const int fc_smem_bytes = (9 * cin * 32 + 32) * (int)sizeof(float) + FC_PIX * (3 * 32 * 2 + 16);
    static std::once_flag fc_once;
    std::call_once(fc_once, [] {
      cudaFuncSetAttribute(first_conv_k3c32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
      cudaFuncSetAttribute(first_conv_k3c32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    });
    const int grid = ew_grid(pixels, FC_PIX, 3);
    // the output is one contiguous run when the destination window is the whole buffer
    const int contiguous = dst->y0 == 0 && dst->x0 == 0 && dst->c0 == 0 && dst->h == Ho && dst->w == Wo && dst->c == cout;
    if (cin == 4)
      first_conv_k3c32_kernel<4><<<grid, FC_THREADS, fc_smem_bytes, as_stream(st)>>>(batch, in_h, in_w, x, w_mu,
                                                                                      w_sigma, *dst, relu, contiguous);
    else
      first_conv_k3c32_kernel<1><<<grid, FC_THREADS, fc_smem_bytes, as_stream(st)>>>(batch, in_h, in_w, x, w_mu,
                                                                                      w_sigma, *dst, relu, contiguous);
    return check_launch("first_conv_k3c32");
  }
  const size_t smem = ((size_t)ksize * ksize * cin * cout + cout) * sizeof(float);
  const size_t total = (size_t)batch * Ho * Wo * (cout / 8);
  first_conv_packed_kernel<<<ew_grid(total, 256, 4), 256, smem, as_stream(st)>>>(
      batch, in_h, in_w, cin, cout, ksize, x, w_mu, w_sigma, *dst, relu);
  return check_launch("first_conv_packed");
}

int sn_maxpool2_packed(const sn_packed_view* src, int32_t batch, int32_t in_h, int32_t in_w, int32_t c,
                       const sn_packed_view* dst, sn_stream_t st) {
  SN_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && c > 0, SN_ERR_BAD_ARG, "maxpool_packed: bad shape");
  int rc = check_pview(src, batch, in_h, in_w, c, "maxpool_packed src");
  if (rc) return rc;
  const int Ho = (in_h + 1) / 2, Wo = (in_w + 1) / 2;
  if ((rc = check_pview(dst, batch, Ho, Wo, c, "maxpool_packed dst"))) return rc;
  if (view_v8(src, c) && view_v8(dst, c)) {
    const size_t total16 = (size_t)batch * Ho * Wo * (c / 16);
    maxpool_packed16_kernel<<<ew_grid(total16, 256), 256, 0, as_stream(st)>>>(*src, batch, in_h, in_w, c, *dst);
    return check_launch("maxpool_packed16");
  }
  const size_t total = (size_t)batch * Ho * Wo * (c / 8);
  maxpool_packed_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(*src, batch, in_h, in_w, c, *dst);
  return check_launch("maxpool_packed");
}

int sn_relu_packed(const sn_packed_view* src, int32_t batch, int32_t h, int32_t w, int32_t c,
                   const sn_packed_view* dst, int32_t gate, sn_stream_t st) {
  SN_REQUIRE(batch > 0 && h > 0 && w > 0 && c > 0, SN_ERR_BAD_ARG, "relu_packed: bad shape");
  int rc = check_pview(src, batch, h, w, c, "relu_packed src");
  if (rc) return rc;
  if ((rc = check_pview(dst, batch, h, w, c, "relu_packed dst"))) return rc;
  const size_t total = (size_t)batch * h * w * (c / 8);
  relu_copy_packed_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(*src, batch, h, w, c, *dst, gate ? 1 : 0);
  return check_launch("relu_packed");
}

int sn_final_conv_softmax_packed(const sn_packed_view* src, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                 int32_t n_labels, const float* w_mu, const float* w_sigma, float* p_out,
                                 float* var_out, float* presoftmax_mu, float* presoftmax_var, sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && p_out && var_out, SN_ERR_BAD_ARG, "final_conv: null pointer");
  SN_REQUIRE((presoftmax_mu == nullptr) == (presoftmax_var == nullptr), SN_ERR_BAD_ARG,
             "final_conv: pass both pre-softmax outputs or neither");
  SN_REQUIRE(n_labels >= 1 && n_labels <= 8, SN_ERR_UNSUPPORTED, "final_conv: %d classes (max 8)", n_labels);
  SN_REQUIRE(cin > 0 && cin % 8 == 0 && cin <= 256, SN_ERR_UNSUPPORTED, "final_conv: cin %d", cin);
  int rc = check_pview(src, batch, in_h, in_w, cin, "final_conv src");
  if (rc) return rc;
  const size_t total = (size_t)batch * in_h * in_w;
  SN_REQUIRE(aligned16(p_out) && aligned16(var_out) && (!presoftmax_mu || (aligned16(presoftmax_mu) &&
                                                                            aligned16(presoftmax_var))),
             SN_ERR_MISALIGNED, "final_conv: outputs must be 16-byte aligned");
  int contiguous = src->y0 == 0 && src->x0 == 0 && src->c0 == 0 && src->h == in_h && src->w == in_w && src->c == cin;
  const size_t smem_w = (((size_t)2 * cin * n_labels + n_labels + 3) & ~(size_t)3) * sizeof(float);
  size_t smem = smem_w + (contiguous ? (size_t)128 * (6 * cin + 16) : 0);
  if (smem > 48 * 1024) {          // wide inputs: skip the coalescing stage, read the pixel rows straight from global
    contiguous = 0;
    smem = smem_w;
  }
  SN_REQUIRE(smem <= 48 * 1024, SN_ERR_UNSUPPORTED, "final_conv: cin %d too large", cin);
  const int grid = ew_grid((total + 127) / 128 * 128, 128, 6);
#define SN_FINAL(CC)                                                                                           \
  case CC:                                                                                                     \
    final_conv_softmax_kernel<CC><<<grid, 128, smem, as_stream(st)>>>(*src, batch, in_h, in_w, cin, w_mu, w_sigma, \
                                                                       p_out, var_out, presoftmax_mu,          \
                                                                       presoftmax_var, contiguous);            \
    break;
  switch (n_labels) {
    SN_FINAL(1) SN_FINAL(2) SN_FINAL(3) SN_FINAL(4) SN_FINAL(5) SN_FINAL(6) SN_FINAL(7) SN_FINAL(8)
  }
#undef SN_FINAL
  return check_launch("final_conv_softmax");
}

}  // extern "C"
