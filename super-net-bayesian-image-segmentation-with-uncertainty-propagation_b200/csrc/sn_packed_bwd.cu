// FAST mode backward, bandwidth-bound kernels around the tensor-core data gradient (sn_tc_halo.cu, DGRAD):
// transposed weight preparation, arg-max pooling adjoint, the fused head backward (NLL -> softmax Jacobian ->
// 1x1 conv -> ReLU gate) and the input gradient of the first convolution.  Gradients use the packed layout
// [n][h][w][3][c] bf16 (planes g_mean_hi, g_mean_lo, g_variance); gates and arg-max routing are recomputed from
// the saved forward activations.  Formulas: SURVEY.md A.3 (what tf.GradientTape derives at Brats.py:578,593).
#include "sn_common.cuh"
#include "sn_sm100.cuh"

#include <mutex>

namespace sn {

__device__ __forceinline__ float b_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float b_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t b_pk2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void b_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = b_lo(u.x); f[1] = b_hi(u.x); f[2] = b_lo(u.y); f[3] = b_hi(u.y);
  f[4] = b_lo(u.z); f[5] = b_hi(u.z); f[6] = b_lo(u.w); f[7] = b_hi(u.w);
}
__device__ __forceinline__ void b_split8(const float (&m)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = b_pk2(m[2 * j], m[2 * j + 1]);
    l[j] = b_pk2(m[2 * j] - b_lo(h[j]), m[2 * j + 1] - b_hi(h[j]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 b_pack8(const float (&v)[8]) {
  return make_uint4(b_pk2(v[0], v[1]), b_pk2(v[2], v[3]), b_pk2(v[4], v[5]), b_pk2(v[6], v[7]));
}
__device__ __forceinline__ size_t pv_off(const sn_packed_view& v, int b, int y, int x) {
  return ((((size_t)b * v.h + y + v.y0) * v.w + x + v.x0) * 3) * v.c + v.c0;
}

// ---------------------------------------------------------------------------------------------------------
// transposed weights for the data gradient: out[pl][tap'][ci][kcol]
// ---------------------------------------------------------------------------------------------------------
__global__ void prepare_weights_bwd_kernel(const float* __restrict__ w, int k, int cin, int cout, int upconv,
                                           __nv_bfloat16* __restrict__ out) {
  const int taps = upconv ? 1 : k * k;
  const int K = upconv ? 4 * cout : cout;
  const size_t plane = (size_t)taps * cin * K;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += stride) {
    const int kc = (int)(i % K);
    size_t t = i / K;
    const int ci = (int)(t % cin);
    const int tap = (int)(t / cin);
    int kh, kw, n;
    if (upconv) {
      const int par = kc / cout;
      n = kc - par * cout;
      kh = 1 - (par >> 1); kw = 1 - (par & 1);           // parity (a,b) <- W[1-a, 1-b]
    } else {
      n = kc;
      kh = k - 1 - tap / k; kw = k - 1 - tap % k;        // full correlation: flipped filter
    }
    const float v = w[(((size_t)kh * k + kw) * cin + ci) * cout + n];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    out[i] = hi;
    out[plane + i] = lo;
    out[2 * plane + i] = __float2bfloat16_rn(v * v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// arg-max pooling adjoint; thread = (output pixel, 8 channels)
// ---------------------------------------------------------------------------------------------------------
__global__ void maxpool_bwd_packed_kernel(sn_packed_view in, int B, int H, int W, int c, sn_packed_view gout,
                                          sn_packed_view gin, int ky0, int kx0, int kh, int kw) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int g = c / 8;
  const size_t total = (size_t)B * Ho * Wo * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in.base);
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(gout.base);
  __nv_bfloat16* gi = reinterpret_cast<__nv_bfloat16*>(gin.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % g) * 8;
    size_t t = i / g;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), x = 2 * xo + (d & 1);
      if (y < H && x < W) {
        const __nv_bfloat16* s = src + pv_off(in, b, y, x) + c8;
        float h[8], l[8];
        b_unpack8(*reinterpret_cast<const uint4*>(s), h);
        b_unpack8(*reinterpret_cast<const uint4*>(s + in.c), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float m = h[j] + l[j];
          if (m > best[j]) { best[j] = m; arg[j] = d; }     // first maximum wins, like the forward
        }
      }
    }
    const __nv_bfloat16* gp = go + pv_off(gout, b, yo, xo) + c8;
    float gh[8], gl[8], gv[8];
    b_unpack8(*reinterpret_cast<const uint4*>(gp), gh);
    b_unpack8(*reinterpret_cast<const uint4*>(gp + gout.c), gl);
    b_unpack8(*reinterpret_cast<const uint4*>(gp + 2 * gout.c), gv);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), x = 2 * xo + (d & 1);
      if (y < H && x < W) {
        __nv_bfloat16* o = gi + pv_off(gin, b, y, x) + c8;
        float m[8], v[8];
        const bool keep = y >= ky0 && y < ky0 + kh && x >= kx0 && x < kx0 + kw;
        if (keep) {
          float h[8], l[8];
          b_unpack8(*reinterpret_cast<const uint4*>(o), h);
          b_unpack8(*reinterpret_cast<const uint4*>(o + gin.c), l);
          b_unpack8(*reinterpret_cast<const uint4*>(o + 2 * gin.c), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = h[j] + l[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (arg[j] == d) { m[j] += gh[j] + gl[j]; v[j] += gv[j]; }
        uint4 hi, lo;
        b_split8(m, hi, lo);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + gin.c) = lo;
        *reinterpret_cast<uint4*>(o + 2 * gin.c) = b_pack8(v);
      }
    }
  }
}

// 16 channels per thread through 256-bit accesses (full 32-byte sectors); used when the views are 32-byte aligned.
__device__ __forceinline__ void b_unpack16(const uint32_t (&u)[8], float (&f)[16]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) { f[2 * e] = b_lo(u[e]); f[2 * e + 1] = b_hi(u[e]); }
}
__device__ __forceinline__ void b_split16(const float (&m)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = b_pk2(m[2 * j], m[2 * j + 1]);
    lo[j] = b_pk2(m[2 * j] - b_lo(hi[j]), m[2 * j + 1] - b_hi(hi[j]));
  }
}

__global__ void maxpool_bwd_packed16_kernel(sn_packed_view in, int B, int H, int W, int c, sn_packed_view gout,
                                            sn_packed_view gin, int ky0, int kx0, int kh, int kw) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int g = c / 16;
  const size_t total = (size_t)B * Ho * Wo * g;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in.base);
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(gout.base);
  __nv_bfloat16* gi = reinterpret_cast<__nv_bfloat16*>(gin.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c16 = (int)(i % g) * 16;
    size_t t = i / g;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[16];
    int arg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { best[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), x = 2 * xo + (d & 1);
      if (y < H && x < W) {
        const __nv_bfloat16* sp = src + pv_off(in, b, y, x) + c16;
        uint32_t hw[8], lw[8];
        ptx::ld_global_v8(sp, hw);
        ptx::ld_global_v8(sp + in.c, lw);
        float h[16], l[16];
        b_unpack16(hw, h);
        b_unpack16(lw, l);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float m = h[j] + l[j];
          if (m > best[j]) { best[j] = m; arg[j] = d; }
        }
      }
    }
    const __nv_bfloat16* gp = go + pv_off(gout, b, yo, xo) + c16;
    float gm[16], gv[16];
    {
      uint32_t a[8], bq[8], cq[8];
      ptx::ld_global_v8(gp, a);
      ptx::ld_global_v8(gp + gout.c, bq);
      ptx::ld_global_v8(gp + 2 * gout.c, cq);
      float gh[16], gl[16];
      b_unpack16(a, gh);
      b_unpack16(bq, gl);
      b_unpack16(cq, gv);
#pragma unroll
      for (int j = 0; j < 16; ++j) gm[j] = gh[j] + gl[j];
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * yo + (d >> 1), x = 2 * xo + (d & 1);
      if (y < H && x < W) {
        __nv_bfloat16* o = gi + pv_off(gin, b, y, x) + c16;
        float m[16], v[16];
        const bool keep = y >= ky0 && y < ky0 + kh && x >= kx0 && x < kx0 + kw;
        if (keep) {
          uint32_t a[8], bq[8], cq[8];
          ptx::ld_global_v8(o, a);
          ptx::ld_global_v8(o + gin.c, bq);
          ptx::ld_global_v8(o + 2 * gin.c, cq);
          float h[16], l[16];
          b_unpack16(a, h);
          b_unpack16(bq, l);
          b_unpack16(cq, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = h[j] + l[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (arg[j] == d) { m[j] += gm[j]; v[j] += gv[j]; }
        uint32_t hi[8], lo[8], vr[8];
        b_split16(m, hi, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) vr[j] = b_pk2(v[2 * j], v[2 * j + 1]);
        ptx::st_global_v8(o, hi);
        ptx::st_global_v8(o + gin.c, lo);
        ptx::st_global_v8(o + 2 * gin.c, vr);
      }
    }
  }
}

static bool pv_v8(const sn_packed_view* v, int c) {
  return (reinterpret_cast<uintptr_t>(v->base) & 31u) == 0 && v->c % 16 == 0 && v->c0 % 16 == 0 && c % 16 == 0;
}

// ---------------------------------------------------------------------------------------------------------
// fused head backward; thread = pixel
// ---------------------------------------------------------------------------------------------------------
constexpr float kHeadEps = 1e-3f;     // nll_gaussian's epsilon (Brats.py:298)

template <int C>
__global__ void __launch_bounds__(128) head_bwd_kernel(sn_packed_view in, int B, int H, int W, int cin,
                                                       const float* __restrict__ w, const float* __restrict__ ws,
                                                       const float* __restrict__ y, float clip_lo, float clip_hi,
                                                       const double* __restrict__ acc, float loss_scale,
                                                       sn_packed_view gin, float* __restrict__ g_logit_mu,
                                                       float* __restrict__ g_logit_var, float* __restrict__ rsum_out,
                                                       const float* __restrict__ up_gp,
                                                       const float* __restrict__ up_gv, int v8) {
  // up_gp != nullptr: the upstream gradients w.r.t. (p, var_out) are given (saliency, Brats.py:598-609) instead of
  // being derived from the NLL; `y` and `acc` are then unused.
  extern __shared__ __align__(16) float hsm[];   // W [cin][C], W^2 [cin][C], s [C]
  float* sw = hsm;
  float* sw2 = hsm + cin * C;
  float* ss = hsm + 2 * cin * C;
  for (int i = threadIdx.x; i < cin * C; i += blockDim.x) {
    const float v = w[i];
    sw[i] = v;
    sw2[i] = v * v;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) ss[i] = softplus_f(ws[i]);
  __syncthreads();
  const size_t total = (size_t)B * H * W;
  const float qmean = up_gp ? 0.f : (float)(acc[0] / (double)total);
  const float qon = (isnan(qmean) || isinf(qmean)) ? 0.f : 1.f;      // Brats.py:304-305
  const float g = 0.5f * loss_scale / (float)total;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in.base);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(gin.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    size_t t = i / W;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    const __nv_bfloat16* s = src + pv_off(in, b, yy, xx);
    // ---- forward recompute: conv_final (k = 1) + softmax moments
    float m[C], v[C];
#pragma unroll
    for (int j = 0; j < C; ++j) m[j] = v[j] = 0.f;
    float r = 0.f;
    for (int c16 = 0; c16 < cin; c16 += 16) {
      float h[16], l[16], vv[16];
      if (v8) {                              // 256-bit loads: one full 32-byte sector per lane and instruction
        uint32_t a[8], bq[8], cq[8];
        ptx::ld_global_v8(s + c16, a);
        ptx::ld_global_v8(s + in.c + c16, bq);
        ptx::ld_global_v8(s + 2 * in.c + c16, cq);
        b_unpack16(a, h);
        b_unpack16(bq, l);
        b_unpack16(cq, vv);
      } else {
#pragma unroll
        for (int q8 = 0; q8 < 16; q8 += 8) {
          float h8[8], l8[8], v8f[8];
          const bool in_range = c16 + q8 < cin;
          if (in_range) {
            b_unpack8(*reinterpret_cast<const uint4*>(s + c16 + q8), h8);
            b_unpack8(*reinterpret_cast<const uint4*>(s + in.c + c16 + q8), l8);
            b_unpack8(*reinterpret_cast<const uint4*>(s + 2 * in.c + c16 + q8), v8f);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            h[q8 + e] = in_range ? h8[e] : 0.f;
            l[q8 + e] = in_range ? l8[e] : 0.f;
            vv[q8 + e] = in_range ? v8f[e] : 0.f;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (c16 + e < cin) {
          const float mu = h[e] + l[e];
          r += fmaf(mu, mu, vv[e]);
#pragma unroll
          for (int j = 0; j < C; ++j) {
            m[j] = fmaf(mu, sw[(c16 + e) * C + j], m[j]);
            v[j] = fmaf(vv[e], sw2[(c16 + e) * C + j], v[j]);
          }
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      v[j] = fmaxf(fmaf(ss[j], r, v[j]), 0.f);
      mx = fmaxf(mx, m[j]);
    }
    float p[C], sum = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) { p[j] = expf(m[j] - mx); sum += p[j]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < C; ++j) p[j] *= inv;
    // ---- NLL gradient w.r.t. (p, softmax variance) on the clipped variance (Brats.py:293-311)
    float gp[C], gvo[C];
#pragma unroll
    for (int a = 0; a < C; ++a) {
      float vo = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float J = p[a] * ((a == j ? 1.f : 0.f) - p[j]);
        vo = fmaf(J * J, v[j], vo);
      }
      if (up_gp) {
        gp[a] = up_gp[i * C + a];
        gvo[a] = up_gv ? up_gv[i * C + a] : 0.f;
        continue;
      }
      const float vc = fminf(fmaxf(vo, clip_lo), clip_hi) + kHeadEps;
      const float iv = 1.f / vc;
      const float d = p[a] - y[i * C + a];
      gp[a] = qon * g * 2.f * d * iv;
      const bool pass = vo >= clip_lo && vo <= clip_hi;
      gvo[a] = pass ? g * (iv - qon * d * d * iv * iv) : 0.f;
    }
    // ---- softmax-Jacobian VJP
    float gv[C], gpt[C];
#pragma unroll
    for (int j = 0; j < C; ++j) { gv[j] = 0.f; gpt[j] = gp[j]; }
#pragma unroll
    for (int a = 0; a < C; ++a) {
      float sa = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float d = (a == j ? 1.f : 0.f) - p[j];
        const float J = p[a] * d;
        gv[j] += gvo[a] * J * J;
        sa += J * v[j] * d;
        gpt[j] -= gvo[a] * 2.f * J * v[j] * p[a];
      }
      gpt[a] += gvo[a] * 2.f * sa;
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) dot += gpt[k] * p[k];
    float gm[C];
    float tt = 0.f;                         // t = sum_n g_var_n s_n
#pragma unroll
    for (int j = 0; j < C; ++j) {
      gm[j] = p[j] * (gpt[j] - dot);
      tt = fmaf(gv[j], ss[j], tt);
    }
    if (g_logit_mu) {               // what the weight gradient of conv_final needs
#pragma unroll
      for (int j = 0; j < C; ++j) { g_logit_mu[i * C + j] = gm[j]; g_logit_var[i * C + j] = gv[j]; }
      rsum_out[i] = r;
    }
    // ---- conv_final data gradient + ReLU gate of its (post-ReLU) input
    __nv_bfloat16* o = dst + pv_off(gin, b, yy, xx);
    if (v8) {
      for (int c16 = 0; c16 < cin; c16 += 16) {
        uint32_t a8[8], b8[8];
        ptx::ld_global_v8(s + c16, a8);
        ptx::ld_global_v8(s + in.c + c16, b8);
        float h[16], l[16], om[16], ov[16];
        b_unpack16(a8, h);
        b_unpack16(b8, l);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float mu = h[e] + l[e];
          float a = 2.f * mu * tt, vv = tt;
#pragma unroll
          for (int j = 0; j < C; ++j) {
            a = fmaf(gm[j], sw[(c16 + e) * C + j], a);
            vv = fmaf(gv[j], sw2[(c16 + e) * C + j], vv);
          }
          const bool on = mu > 0.f;
          om[e] = on ? a : 0.f;
          ov[e] = on ? vv : 0.f;
        }
        uint32_t hi[8], lo[8], vr[8];
        b_split16(om, hi, lo);
#pragma unroll
        for (int q = 0; q < 8; ++q) vr[q] = b_pk2(ov[2 * q], ov[2 * q + 1]);
        ptx::st_global_v8(o + c16, hi);
        ptx::st_global_v8(o + gin.c + c16, lo);
        ptx::st_global_v8(o + 2 * gin.c + c16, vr);
      }
    } else
    for (int c8 = 0; c8 < cin; c8 += 8) {
      float h[8], l[8];
      b_unpack8(*reinterpret_cast<const uint4*>(s + c8), h);
      b_unpack8(*reinterpret_cast<const uint4*>(s + in.c + c8), l);
      float om[8], ov[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float mu = h[e] + l[e];
        float a = 2.f * mu * tt, vv = tt;
#pragma unroll
        for (int j = 0; j < C; ++j) {
          a = fmaf(gm[j], sw[(c8 + e) * C + j], a);
          vv = fmaf(gv[j], sw2[(c8 + e) * C + j], vv);
        }
        const bool on = mu > 0.f;
        om[e] = on ? a : 0.f;
        ov[e] = on ? vv : 0.f;
      }
      uint4 hi, lo;
      b_split8(om, hi, lo);
      *reinterpret_cast<uint4*>(o + c8) = hi;
      *reinterpret_cast<uint4*>(o + gin.c + c8) = lo;
      *reinterpret_cast<uint4*>(o + 2 * gin.c + c8) = b_pack8(ov);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// input gradient of the first convolution; thread = input pixel
// ---------------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(128) first_conv_bwd_kernel(int B, int H, int W, int cout, int k,
                                                             const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ ws, sn_packed_view gout,
                                                             float* __restrict__ gx) {
  extern __shared__ __align__(16) float fsm[];   // W [tap][cout][CIN], s [cout]
  float* sw = fsm;
  float* ss = fsm + k * k * cout * CIN;
  for (int i = threadIdx.x; i < k * k * cout * CIN; i += blockDim.x) {
    const int ci = i % CIN;
    const int n = (i / CIN) % cout;
    const int tap = i / (CIN * cout);
    sw[i] = w[((size_t)tap * CIN + ci) * cout + n];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) ss[i] = softplus_f(ws[i]);
  __syncthreads();
  const int Ho = H - k + 1, Wo = W - k + 1;
  const size_t total = (size_t)B * H * W;
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(gout.base);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    size_t t = i / W;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    float acc[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
    float T = 0.f;
    for (int kh = 0; kh < k; ++kh) {
      const int oy = yy - kh;
      if (oy < 0 || oy >= Ho) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int ox = xx - kw;
        if (ox < 0 || ox >= Wo) continue;
        const __nv_bfloat16* gp = go + pv_off(gout, b, oy, ox);
        const float* wt = sw + (kh * k + kw) * cout * CIN;
        for (int n8 = 0; n8 < cout; n8 += 8) {
          float h[8], l[8], v[8];
          b_unpack8(*reinterpret_cast<const uint4*>(gp + n8), h);
          b_unpack8(*reinterpret_cast<const uint4*>(gp + gout.c + n8), l);
          b_unpack8(*reinterpret_cast<const uint4*>(gp + 2 * gout.c + n8), v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float gm = h[e] + l[e];
            T = fmaf(v[e], ss[n8 + e], T);
#pragma unroll
            for (int c = 0; c < CIN; ++c) acc[c] = fmaf(gm, wt[(n8 + e) * CIN + c], acc[c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CIN; ++c) gx[i * CIN + c] = fmaf(2.f * x[i * CIN + c], T, acc[c]);
  }
}

// Specialisation for the shapes the two networks use (k = 3, 32 gradient channels, Cin = 4 or 1).  A block owns a
// 32 x 8 tile of input pixels; the (32+2) x (8+2) gradient pixels it touches are read once with fully coalesced
// 16-byte loads (one pixel = 192 contiguous bytes), combined to fp32 (g_mean = hi + lo; t = sum_n g_var_n s_n)
// and staged in shared memory (rows padded to 36 floats: conflict-free 16-byte reads).  Each thread then
// accumulates 2 pixels x 9 taps x 32 channels x CIN with the transposed weights broadcast from shared memory.
constexpr int FB_TW = 32, FB_TH = 8, FB_THREADS = 128, FB_ROW = 36;
constexpr int FB_GW = FB_TW + 2, FB_GH = FB_TH + 2;

template <int CIN>
__global__ void __launch_bounds__(FB_THREADS) first_conv_bwd_k3c32_kernel(int B, int H, int W,
                                                                          const float* __restrict__ x,
                                                                          const float* __restrict__ w,
                                                                          const float* __restrict__ ws,
                                                                          sn_packed_view gout, float* __restrict__ gx,
                                                                          int tiles_x, int tiles_y) {
  constexpr int COUT = 32;
  extern __shared__ __align__(16) float bsm[];
  float* sg = bsm;                                   // [FB_GH * FB_GW][FB_ROW] g_mean
  float* st = sg + FB_GH * FB_GW * FB_ROW;           // [FB_GH * FB_GW] t
  float* sw = st + ((FB_GH * FB_GW + 3) & ~3);       // [9][COUT][CIN]
  float* ss = sw + 9 * COUT * CIN;                   // [COUT]
  for (int i = threadIdx.x; i < 9 * COUT * CIN; i += FB_THREADS) {
    const int ci = i % CIN;
    const int n = (i / CIN) % COUT;
    const int tap = i / (CIN * COUT);
    sw[i] = w[((size_t)tap * CIN + ci) * COUT + n];
  }
  if (threadIdx.x < COUT) ss[threadIdx.x] = softplus_f(ws[threadIdx.x]);
  __syncthreads();
  const int Ho = H - 2, Wo = W - 2;
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(gout.base);
  const int total_tiles = B * tiles_y * tiles_x;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int b = tile / (tiles_x * tiles_y);
    const int x0 = tx * FB_TW, y0 = ty * FB_TH;
    // ---- stage the gradient tile: chunk = (pixel, 8 channels); gradient pixel (y0 - 2 + gy, x0 - 2 + gxx)
    constexpr int CHUNKS = FB_GH * FB_GW * 4;
    for (int c = threadIdx.x; c < ((CHUNKS + FB_THREADS - 1) / FB_THREADS) * FB_THREADS; c += FB_THREADS) {
      const bool live = c < CHUNKS;                 // whole warps stay in the loop: the shuffles below need them
      const int pix = live ? c >> 2 : 0, part = c & 3;
      const int gy = pix / FB_GW, gxx = pix - gy * FB_GW;
      const int oy = y0 - 2 + gy, ox = x0 - 2 + gxx;
      float m[8], tpart = 0.f;
      if (live && oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
        const __nv_bfloat16* gp = go + pv_off(gout, b, oy, ox) + part * 8;
        float h[8], l[8], v[8];
        b_unpack8(*reinterpret_cast<const uint4*>(gp), h);
        b_unpack8(*reinterpret_cast<const uint4*>(gp + gout.c), l);
        b_unpack8(*reinterpret_cast<const uint4*>(gp + 2 * gout.c), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          m[e] = h[e] + l[e];
          tpart = fmaf(v[e], ss[part * 8 + e], tpart);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = 0.f;
      }
      if (live) {
        float* d = sg + pix * FB_ROW + part * 8;
        *reinterpret_cast<float4*>(d) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(m[4], m[5], m[6], m[7]);
      }
      // the four parts of a pixel sit in four consecutive lanes
      tpart += __shfl_xor_sync(0xffffffffu, tpart, 1);
      tpart += __shfl_xor_sync(0xffffffffu, tpart, 2);
      if (live && part == 0) st[pix] = tpart;
    }
    __syncthreads();
    // ---- thread = 2 input pixels (rows ly and ly + 4 of the tile, column lx)
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    float acc[2][CIN];
    float T[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int c = 0; c < CIN; ++c) acc[q][c] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        // input pixel (y, x) reads gradient pixel (y - kh, x - kw) = staged (ly + 2 - kh, lx + 2 - kw)
        const float* wt = sw + (kh * 3 + kw) * COUT * CIN;
        const int p0 = (ly + 2 - kh) * FB_GW + lx + 2 - kw;
        const int p1 = p0 + 4 * FB_GW;
        T[0] += st[p0];
        T[1] += st[p1];
        const float* g0 = sg + p0 * FB_ROW;
        const float* g1 = sg + p1 * FB_ROW;
#pragma unroll
        for (int n4 = 0; n4 < COUT; n4 += 4) {
          const float4 a = *reinterpret_cast<const float4*>(g0 + n4);
          const float4 bq = *reinterpret_cast<const float4*>(g1 + n4);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if constexpr (CIN == 4) {
              const float4 wv = *reinterpret_cast<const float4*>(wt + (n4 + e) * 4);
              acc[0][0] = fmaf(av[e], wv.x, acc[0][0]); acc[0][1] = fmaf(av[e], wv.y, acc[0][1]);
              acc[0][2] = fmaf(av[e], wv.z, acc[0][2]); acc[0][3] = fmaf(av[e], wv.w, acc[0][3]);
              acc[1][0] = fmaf(bv[e], wv.x, acc[1][0]); acc[1][1] = fmaf(bv[e], wv.y, acc[1][1]);
              acc[1][2] = fmaf(bv[e], wv.z, acc[1][2]); acc[1][3] = fmaf(bv[e], wv.w, acc[1][3]);
            } else {
#pragma unroll
              for (int c = 0; c < CIN; ++c) {
                const float wv = wt[(n4 + e) * CIN + c];
                acc[0][c] = fmaf(av[e], wv, acc[0][c]);
                acc[1][c] = fmaf(bv[e], wv, acc[1][c]);
              }
            }
          }
        }
      }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int yy = y0 + ly + 4 * q, xx = x0 + lx;
      if (yy < H && xx < W) {
        const size_t i = ((size_t)b * H + yy) * W + xx;
        if constexpr (CIN == 4) {
          const float4 xv = *reinterpret_cast<const float4*>(x + i * 4);
          const float t2 = 2.f * T[q];
          *reinterpret_cast<float4*>(gx + i * 4) =
              make_float4(fmaf(xv.x, t2, acc[q][0]), fmaf(xv.y, t2, acc[q][1]), fmaf(xv.z, t2, acc[q][2]),
                          fmaf(xv.w, t2, acc[q][3]));
        } else {
#pragma unroll
          for (int c = 0; c < CIN; ++c) gx[i * CIN + c] = fmaf(2.f * x[i * CIN + c], T[q], acc[q][c]);
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradients of the two layers too thin for the tensor cores
// ---------------------------------------------------------------------------------------------------------
// myConv_input (k = 3, 32 output channels, Cin = 4 or 1):  g_w[tap,c,n] = sum_p x[p+tap,c] g_mu'[p,n];
// ds[n] = sum_p g_var'[p,n] r[p], r = box_3(sum_c x^2).  A block owns 32 x 4 output pixels per step; warp w < 9 is
// filter tap w (lane = output channel n, CIN accumulators), warp 9 accumulates ds.  Operands are staged in shared
// memory as fp32 (x tile with its halo; g_mean = hi + lo; g_var); every lane reads its own bank, x is a broadcast.
constexpr int FW_TW = 32, FW_TH = 4, FW_PIX = FW_TW * FW_TH, FW_THREADS = 320;

template <int CIN>
__global__ void __launch_bounds__(FW_THREADS) first_conv_wgrad_k3c32_kernel(int B, int H, int W,
                                                                            const float* __restrict__ x,
                                                                            sn_packed_view gout, float* __restrict__ g_w,
                                                                            float* __restrict__ ds, int tiles_x,
                                                                            int tiles_y) {
  constexpr int COUT = 32, XW = FW_TW + 2, XH = FW_TH + 2;
  __shared__ __align__(16) float xs[XH * XW * CIN];
  __shared__ __align__(16) float gm[FW_PIX * COUT];
  __shared__ __align__(16) float gv[FW_PIX * COUT];
  __shared__ float rr[FW_PIX];
  const int Ho = H - 2, Wo = W - 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(gout.base);
  float acc[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
  const int total_tiles = B * tiles_y * tiles_x;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int b = tile / (tiles_x * tiles_y);
    const int x0 = tx * FW_TW, y0 = ty * FW_TH;
    for (int i = threadIdx.x; i < XH * XW; i += FW_THREADS) {
      const int yy = y0 + i / XW, xx = x0 + i % XW;
      const bool in = yy < H && xx < W;
#pragma unroll
      for (int c = 0; c < CIN; ++c) xs[i * CIN + c] = in ? x[(((size_t)b * H + yy) * W + xx) * CIN + c] : 0.f;
    }
    for (int i = threadIdx.x; i < FW_PIX * 4; i += FW_THREADS) {
      const int pix = i >> 2, part = i & 3;
      const int oy = y0 + pix / FW_TW, ox = x0 + pix % FW_TW;
      float m[8], v[8];
      if (oy < Ho && ox < Wo) {
        const __nv_bfloat16* gp = go + pv_off(gout, b, oy, ox) + part * 8;
        float h[8], l[8];
        b_unpack8(*reinterpret_cast<const uint4*>(gp), h);
        b_unpack8(*reinterpret_cast<const uint4*>(gp + gout.c), l);
        b_unpack8(*reinterpret_cast<const uint4*>(gp + 2 * gout.c), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = h[e] + l[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = v[e] = 0.f;
      }
      float* dm = gm + pix * COUT + part * 8;
      float* dv = gv + pix * COUT + part * 8;
      *reinterpret_cast<float4*>(dm) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(dm + 4) = make_float4(m[4], m[5], m[6], m[7]);
      *reinterpret_cast<float4*>(dv) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dv + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    if (threadIdx.x < FW_PIX) {
      const int py = threadIdx.x / FW_TW, px = threadIdx.x % FW_TW;
      float r = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float xv = xs[((py + kh) * XW + px + kw) * CIN + c];
            r = fmaf(xv, xv, r);
          }
      rr[threadIdx.x] = r;
    }
    __syncthreads();
    if (warp < 9) {
      const int kh = warp / 3, kw = warp - kh * 3;
#pragma unroll 4
      for (int pix = 0; pix < FW_PIX; ++pix) {
        const int py = pix / FW_TW, px = pix % FW_TW;
        const float g = gm[pix * COUT + lane];
        const float* xp = xs + ((py + kh) * XW + px + kw) * CIN;
        if constexpr (CIN == 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xp);
          fma2(acc[0], acc[1], g, xv.x, xv.y);          // FFMA2: this loop is issue-bound (62 % issue-active in ncu)
          fma2(acc[2], acc[3], g, xv.z, xv.w);
        } else {
#pragma unroll
          for (int c = 0; c < CIN; ++c) acc[c] = fmaf(xp[c], g, acc[c]);
        }
      }
    } else {
#pragma unroll 4
      for (int pix = 0; pix < FW_PIX; ++pix) acc[0] = fmaf(gv[pix * COUT + lane], rr[pix], acc[0]);
    }
    __syncthreads();
  }
  if (warp < 9) {
#pragma unroll
    for (int c = 0; c < CIN; ++c) atomicAdd(g_w + ((size_t)warp * CIN + c) * COUT + lane, acc[c]);
  } else {
    atomicAdd(ds + lane, acc[0]);
  }
}

// conv_final (k = 1, cin = 32, C <= 8 classes): P_mu[ci,j] = sum_p mu[p,ci] g_mu[p,j], P_var likewise, ds[j] =
// sum_p g_var[p,j] r[p].  lane = (pixel sub-index 0..7) x (8-channel chunk 0..3): 16-byte loads of the packed input,
// 2*8*C accumulators per thread, reduced over the pixel lanes with shuffles and over warps through atomics.
template <int C>
__global__ void __launch_bounds__(256) final_conv_wgrad_kernel(sn_packed_view in, int B, int H, int W,
                                                               const float* __restrict__ g_mu,
                                                               const float* __restrict__ g_var,
                                                               const float* __restrict__ rsum, float* __restrict__ p_mu,
                                                               float* __restrict__ p_var, float* __restrict__ ds) {
  const int lane = threadIdx.x & 31;
  const int chunk = lane & 3, psub = lane >> 2;
  const size_t total = (size_t)B * H * W;
  const size_t warp_id = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in.base);
  float am[8][C], av[8][C], ad[C];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < C; ++j) am[e][j] = av[e][j] = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) ad[j] = 0.f;
  for (size_t p0 = warp_id * 8; p0 < total; p0 += n_warps * 8) {
    const size_t i = p0 + psub;
    if (i < total) {
      const int xx = (int)(i % W);
      size_t t = i / W;
      const int yy = (int)(t % H);
      const int b = (int)(t / H);
      const __nv_bfloat16* s = src + pv_off(in, b, yy, xx) + chunk * 8;
      float h[8], l[8], v[8];
      b_unpack8(*reinterpret_cast<const uint4*>(s), h);
      b_unpack8(*reinterpret_cast<const uint4*>(s + in.c), l);
      b_unpack8(*reinterpret_cast<const uint4*>(s + 2 * in.c), v);
      float gmj[C], gvj[C];
#pragma unroll
      for (int j = 0; j < C; ++j) { gmj[j] = g_mu[i * C + j]; gvj[j] = g_var[i * C + j]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float mu = h[e] + l[e];
#pragma unroll
        for (int j = 0; j < C; ++j) {
          am[e][j] = fmaf(mu, gmj[j], am[e][j]);
          av[e][j] = fmaf(v[e], gvj[j], av[e][j]);
        }
      }
      if (chunk == 0) {
        const float r = rsum[i];
#pragma unroll
        for (int j = 0; j < C; ++j) ad[j] = fmaf(gvj[j], r, ad[j]);
      }
    }
  }
  // reduce over the 8 pixel lanes (lane bits 2..4)
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        am[e][j] += __shfl_xor_sync(0xffffffffu, am[e][j], o);
        av[e][j] += __shfl_xor_sync(0xffffffffu, av[e][j], o);
      }
#pragma unroll
    for (int j = 0; j < C; ++j) ad[j] += __shfl_xor_sync(0xffffffffu, ad[j], o);
  }
  // reduce over the block's warps in shared memory, then ONE atomic per (block, output): with one atomic per warp the
  // kernel ended in ~1.2 M atomics on nine cache lines (B200, batch 64: 485 us for 505 MB of input)
  constexpr int NOUT = 2 * 32 * C + C;
  __shared__ float red[8][NOUT];
  const int warp = threadIdx.x >> 5;
  if (psub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        red[warp][(chunk * 8 + e) * C + j] = am[e][j];
        red[warp][32 * C + (chunk * 8 + e) * C + j] = av[e][j];
      }
    if (chunk == 0) {
#pragma unroll
      for (int j = 0; j < C; ++j) red[warp][64 * C + j] = ad[j];
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < NOUT; o += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][o];
    float* dst = o < 32 * C ? p_mu + o : (o < 64 * C ? p_var + (o - 32 * C) : ds + (o - 64 * C));
    atomicAdd(dst, t);
  }
}

// g_w = P_mu + 2 W P_var (p_var may be NULL: first layer) ; g_w_sigma[n] = sigmoid(w_sigma[n]) ds[n]
__global__ void thin_wgrad_finalize_kernel(size_t n_w, int cout, const float* __restrict__ w,
                                           const float* __restrict__ ws, const float* __restrict__ p_mu,
                                           const float* __restrict__ p_var, const float* __restrict__ ds,
                                           float* __restrict__ g_w, float* __restrict__ g_ws) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n_w; i += stride) g_w[i] = p_var ? fmaf(2.f * w[i], p_var[i], p_mu[i]) : p_mu[i];
  for (size_t i = t0; i < (size_t)cout; i += stride) g_ws[i] = sigmoid_f(ws[i]) * ds[i];
}

static int check_pv(const sn_packed_view* v, int batch, int h, int w, int c, const char* who) {
  SN_REQUIRE(v && v->base && aligned16(v->base), SN_ERR_BAD_ARG, "%s: null/misaligned packed view", who);
  SN_REQUIRE(v->n >= batch && v->c % 8 == 0 && v->c0 % 8 == 0 && c % 8 == 0, SN_ERR_MISALIGNED,
             "%s: channel counts/offsets must be multiples of 8", who);
  SN_REQUIRE(v->y0 >= 0 && v->x0 >= 0 && v->c0 >= 0 && v->y0 + h <= v->h && v->x0 + w <= v->w && v->c0 + c <= v->c,
             SN_ERR_BAD_ARG, "%s: window outside the buffer", who);
  return SN_OK;
}

}  // namespace sn

using namespace sn;

namespace sn {
int sn_head_bwd_general(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                        int32_t n_labels, const float* w_mu, const float* w_sigma, const float* y, float clip_lo,
                        float clip_hi, const double* acc, float loss_scale, const sn_packed_view* g_in,
                        float* g_logit_mu, float* g_logit_var, float* rsum_out, const float* up_gp, const float* up_gv,
                        sn_stream_t st);
}

extern "C" {

int sn_prepare_weights_bwd(const float* w_mu, int32_t ksize, int32_t cin, int32_t cout, int32_t upconv,
                           void* wt_packed, sn_stream_t st) {
  SN_REQUIRE(w_mu && wt_packed, SN_ERR_BAD_ARG, "prepare_weights_bwd: null pointer");
  SN_REQUIRE(ksize >= 1 && ksize <= 3 && cin > 0 && cout > 0, SN_ERR_BAD_ARG, "prepare_weights_bwd: bad sizes");
  SN_REQUIRE(!upconv || ksize == 2, SN_ERR_BAD_ARG, "prepare_weights_bwd: upconv needs ksize == 2");
  const size_t n = (size_t)ksize * ksize * cin * cout;
  prepare_weights_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(w_mu, ksize, cin, cout, upconv,
                                                                          reinterpret_cast<__nv_bfloat16*>(wt_packed));
  return check_launch("prepare_weights_bwd");
}

int sn_maxpool2_bwd_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t c,
                           const sn_packed_view* g_out, const sn_packed_view* g_in, int32_t keep_y0, int32_t keep_x0,
                           int32_t keep_h, int32_t keep_w, sn_stream_t st) {
  SN_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && c > 0, SN_ERR_BAD_ARG, "maxpool_bwd_packed: bad shape");
  SN_REQUIRE(keep_h >= 0 && keep_w >= 0 && keep_y0 >= 0 && keep_x0 >= 0 && keep_y0 + keep_h <= in_h &&
                 keep_x0 + keep_w <= in_w, SN_ERR_BAD_ARG, "maxpool_bwd_packed: keep window outside the tensor");
  int rc = check_pv(in, batch, in_h, in_w, c, "maxpool_bwd_packed in");
  if (rc) return rc;
  const int Ho = (in_h + 1) / 2, Wo = (in_w + 1) / 2;
  if ((rc = check_pv(g_out, batch, Ho, Wo, c, "maxpool_bwd_packed g_out"))) return rc;
  if ((rc = check_pv(g_in, batch, in_h, in_w, c, "maxpool_bwd_packed g_in"))) return rc;
  if (pv_v8(in, c) && pv_v8(g_out, c) && pv_v8(g_in, c)) {
    const size_t total16 = (size_t)batch * Ho * Wo * (c / 16);
    maxpool_bwd_packed16_kernel<<<ew_grid(total16, 128), 128, 0, as_stream(st)>>>(*in, batch, in_h, in_w, c, *g_out,
                                                                                 *g_in, keep_y0, keep_x0, keep_h, keep_w);
    return check_launch("maxpool_bwd_packed16");
  }
  const size_t total = (size_t)batch * Ho * Wo * (c / 8);
  maxpool_bwd_packed_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(*in, batch, in_h, in_w, c, *g_out, *g_in,
                                                                             keep_y0, keep_x0, keep_h, keep_w);
  return check_launch("maxpool_bwd_packed");
}

int sn_head_bwd_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                       int32_t n_labels, const float* w_mu, const float* w_sigma, const float* y, float clip_lo,
                       float clip_hi, const double* acc, float loss_scale, const sn_packed_view* g_in,
                       float* g_logit_mu, float* g_logit_var, float* rsum_out, sn_stream_t st) {
  return sn_head_bwd_general(in, batch, in_h, in_w, cin, n_labels, w_mu, w_sigma, y, clip_lo, clip_hi, acc, loss_scale,
                             g_in, g_logit_mu, g_logit_var, rsum_out, nullptr, nullptr, st);
}

int sn_head_bwd_upstream_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                int32_t n_labels, const float* w_mu, const float* w_sigma, const float* g_p,
                                const float* g_var_out, const sn_packed_view* g_in, sn_stream_t st) {
  SN_REQUIRE(g_p, SN_ERR_BAD_ARG, "head_bwd_upstream: null upstream gradient");
  return sn_head_bwd_general(in, batch, in_h, in_w, cin, n_labels, w_mu, w_sigma, nullptr, 0.f, 0.f, nullptr, 1.f, g_in,
                             nullptr, nullptr, nullptr, g_p, g_var_out, st);
}

}  // extern "C"

namespace sn {
int sn_head_bwd_general(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                        int32_t n_labels, const float* w_mu, const float* w_sigma, const float* y, float clip_lo,
                        float clip_hi, const double* acc, float loss_scale, const sn_packed_view* g_in,
                        float* g_logit_mu, float* g_logit_var, float* rsum_out, const float* up_gp, const float* up_gv,
                        sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && (up_gp || (y && acc)), SN_ERR_BAD_ARG, "head_bwd: null pointer");
  SN_REQUIRE((g_logit_mu != nullptr) == (g_logit_var != nullptr) && (g_logit_mu != nullptr) == (rsum_out != nullptr),
             SN_ERR_BAD_ARG, "head_bwd: pass all three optional outputs or none");
  SN_REQUIRE(n_labels >= 1 && n_labels <= 8, SN_ERR_UNSUPPORTED, "head_bwd: %d classes (max 8)", n_labels);
  SN_REQUIRE(cin > 0 && cin % 8 == 0 && cin <= 256, SN_ERR_UNSUPPORTED, "head_bwd: cin %d", cin);
  int rc = check_pv(in, batch, in_h, in_w, cin, "head_bwd in");
  if (rc) return rc;
  if ((rc = check_pv(g_in, batch, in_h, in_w, cin, "head_bwd g_in"))) return rc;
  const size_t total = (size_t)batch * in_h * in_w;
  const size_t smem = ((size_t)2 * cin * n_labels + n_labels) * sizeof(float);
  const int grid = ew_grid(total, 128, 8);
  const int v8 = pv_v8(in, cin) && pv_v8(g_in, cin) ? 1 : 0;
#define SN_HEAD(CC)                                                                                              \
  case CC:                                                                                                       \
    head_bwd_kernel<CC><<<grid, 128, smem, as_stream(st)>>>(*in, batch, in_h, in_w, cin, w_mu, w_sigma, y, clip_lo, \
                                                             clip_hi, acc, loss_scale, *g_in, g_logit_mu,        \
                                                             g_logit_var, rsum_out, up_gp, up_gv, v8);           \
    break;
  switch (n_labels) {
    SN_HEAD(1) SN_HEAD(2) SN_HEAD(3) SN_HEAD(4) SN_HEAD(5) SN_HEAD(6) SN_HEAD(7) SN_HEAD(8)
  }
#undef SN_HEAD
  return check_launch("head_bwd");
}
}  // namespace sn

extern "C" {

int sn_first_conv_bwd_data_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout, int32_t ksize,
                                  const float* x, const float* w_mu, const float* w_sigma,
                                  const sn_packed_view* g_out, float* g_x, sn_stream_t st) {
  SN_REQUIRE(x && w_mu && w_sigma && g_x, SN_ERR_BAD_ARG, "first_conv_bwd: null pointer");
  SN_REQUIRE(batch > 0 && cin >= 1 && cin <= 8 && ksize >= 1 && ksize <= 3 && in_h >= ksize && in_w >= ksize,
             SN_ERR_UNSUPPORTED, "first_conv_bwd: needs cin <= 8 and k <= 3 (got cin %d, k %d)", cin, ksize);
  SN_REQUIRE(cout % 8 == 0 && cout <= 256, SN_ERR_UNSUPPORTED, "first_conv_bwd: cout %d", cout);
  int rc = check_pv(g_out, batch, in_h - ksize + 1, in_w - ksize + 1, cout, "first_conv_bwd g_out");
  if (rc) return rc;
  if (ksize == 3 && cout == 32 && (cin == 4 || cin == 1) && aligned16(x) && aligned16(g_x)) {
    const int tiles_x = (in_w + FB_TW - 1) / FB_TW, tiles_y = (in_h + FB_TH - 1) / FB_TH;
    const size_t fb_smem = ((size_t)FB_GH * FB_GW * FB_ROW + ((FB_GH * FB_GW + 3) & ~3) + 9 * 32 * cin + 32) * sizeof(float);
    static std::once_flag fb_once;
    std::call_once(fb_once, [] {
      cudaFuncSetAttribute(first_conv_bwd_k3c32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(first_conv_bwd_k3c32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    });
    const long long tiles = (long long)batch * tiles_x * tiles_y;
    const int grid = (int)(tiles < (long long)num_sms() * 3 ? tiles : (long long)num_sms() * 3);
    if (cin == 4)
      first_conv_bwd_k3c32_kernel<4><<<grid, FB_THREADS, fb_smem, as_stream(st)>>>(batch, in_h, in_w, x, w_mu, w_sigma,
                                                                                   *g_out, g_x, tiles_x, tiles_y);
    else
      first_conv_bwd_k3c32_kernel<1><<<grid, FB_THREADS, fb_smem, as_stream(st)>>>(batch, in_h, in_w, x, w_mu, w_sigma,
                                                                                   *g_out, g_x, tiles_x, tiles_y);
    return check_launch("first_conv_bwd_k3c32");
  }
  const size_t total = (size_t)batch * in_h * in_w;
  const size_t smem = ((size_t)ksize * ksize * cout * cin + cout) * sizeof(float);
  SN_REQUIRE(smem <= 48 * 1024, SN_ERR_UNSUPPORTED, "first_conv_bwd: weights do not fit shared memory");
  const int grid = ew_grid(total, 128, 8);
#define SN_FCB(CC)                                                                                                  \
  case CC:                                                                                                          \
    first_conv_bwd_kernel<CC><<<grid, 128, smem, as_stream(st)>>>(batch, in_h, in_w, cout, ksize, x, w_mu, w_sigma, \
                                                                   *g_out, g_x);                                    \
    break;
  switch (cin) {
    SN_FCB(1) SN_FCB(2) SN_FCB(3) SN_FCB(4) SN_FCB(5) SN_FCB(6) SN_FCB(7) SN_FCB(8)
  }
#undef SN_FCB
  return check_launch("first_conv_bwd");
}

int sn_first_conv_bwd_weight_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout,
                                    int32_t ksize, const float* x, const float* w_sigma, const sn_packed_view* g_out,
                                    void* workspace, float* g_w_mu, float* g_w_sigma, sn_stream_t st) {
  SN_REQUIRE(x && w_sigma && workspace && g_w_mu && g_w_sigma, SN_ERR_BAD_ARG, "first_conv_wgrad: null pointer");
  SN_REQUIRE(batch > 0 && ksize == 3 && cout == 32 && (cin == 4 || cin == 1) && in_h >= 3 && in_w >= 3,
             SN_ERR_UNSUPPORTED, "first_conv_wgrad: supports k = 3, 32 output channels, cin 1 or 4 (got k %d, cout %d, cin %d)",
             ksize, cout, cin);
  int rc = check_pv(g_out, batch, in_h - 2, in_w - 2, cout, "first_conv_wgrad g_out");
  if (rc) return rc;
  cudaStream_t s = as_stream(st);
  const size_t n_w = (size_t)9 * cin * cout;
  float* wsf = reinterpret_cast<float*>(workspace);      // [n_w] P_mu, then [cout] ds
  if (cudaMemsetAsync(wsf, 0, (n_w + cout) * sizeof(float), s) != cudaSuccess)
    return fail(SN_ERR_LAUNCH, "first_conv_wgrad: memset failed");
  const int tiles_x = (in_w - 2 + FW_TW - 1) / FW_TW, tiles_y = (in_h - 2 + FW_TH - 1) / FW_TH;
  const long long tiles = (long long)batch * tiles_x * tiles_y;
  const int grid = (int)(tiles < (long long)num_sms() * 2 ? tiles : (long long)num_sms() * 2);
  if (cin == 4)
    first_conv_wgrad_k3c32_kernel<4><<<grid, FW_THREADS, 0, s>>>(batch, in_h, in_w, x, *g_out, wsf, wsf + n_w, tiles_x,
                                                                 tiles_y);
  else
    first_conv_wgrad_k3c32_kernel<1><<<grid, FW_THREADS, 0, s>>>(batch, in_h, in_w, x, *g_out, wsf, wsf + n_w, tiles_x,
                                                                 tiles_y);
  if ((rc = check_launch("first_conv_wgrad"))) return rc;
  thin_wgrad_finalize_kernel<<<ew_grid(n_w, 256), 256, 0, s>>>(n_w, cout, nullptr, w_sigma, wsf, nullptr, wsf + n_w,
                                                              g_w_mu, g_w_sigma);
  return check_launch("first_conv_wgrad_finalize");
}

int sn_final_conv_bwd_weight_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                    int32_t n_labels, const float* w_mu, const float* w_sigma, const float* g_logit_mu,
                                    const float* g_logit_var, const float* rsum, void* workspace, float* g_w_mu,
                                    float* g_w_sigma, sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && g_logit_mu && g_logit_var && rsum && workspace && g_w_mu && g_w_sigma, SN_ERR_BAD_ARG,
             "final_conv_wgrad: null pointer");
  SN_REQUIRE(cin == 32, SN_ERR_UNSUPPORTED, "final_conv_wgrad: cin %d (the networks feed conv_final 32 channels)", cin);
  SN_REQUIRE(n_labels >= 1 && n_labels <= 8, SN_ERR_UNSUPPORTED, "final_conv_wgrad: %d classes (max 8)", n_labels);
  int rc = check_pv(in, batch, in_h, in_w, cin, "final_conv_wgrad in");
  if (rc) return rc;
  cudaStream_t s = as_stream(st);
  const size_t n_w = (size_t)cin * n_labels;
  float* wsf = reinterpret_cast<float*>(workspace);      // P_mu [n_w], P_var [n_w], ds [n_labels]
  if (cudaMemsetAsync(wsf, 0, (2 * n_w + n_labels) * sizeof(float), s) != cudaSuccess)
    return fail(SN_ERR_LAUNCH, "final_conv_wgrad: memset failed");
  const size_t total = (size_t)batch * in_h * in_w;
  const int grid = ew_grid((total + 7) / 8 * 32, 256, 2);     // 256 threads x ~114 registers: two blocks per SM
#define SN_FCW(CC)                                                                                                  \
  case CC:                                                                                                          \
    final_conv_wgrad_kernel<CC><<<grid, 256, 0, s>>>(*in, batch, in_h, in_w, g_logit_mu, g_logit_var, rsum, wsf,    \
                                                     wsf + n_w, wsf + 2 * n_w);                                     \
    break;
  switch (n_labels) {
    SN_FCW(1) SN_FCW(2) SN_FCW(3) SN_FCW(4) SN_FCW(5) SN_FCW(6) SN_FCW(7) SN_FCW(8)
  }
#undef SN_FCW
  if ((rc = check_launch("final_conv_wgrad"))) return rc;
  thin_wgrad_finalize_kernel<<<1, 256, 0, s>>>(n_w, n_labels, w_mu, w_sigma, wsf, wsf + n_w, wsf + 2 * n_w, g_w_mu,
                                               g_w_sigma);
  return check_launch("final_conv_wgrad_finalize");
}

}  // extern "C"
