// FP32-mode bandwidth-bound kernels of the moment path: ReLU gate, arg-max pooling, window copies
// (unpool / pad / crop / concat and adjoints), softmax-Jacobian variance, Gaussian NLL, KL regulariser.
// Every kernel is coalesced along the NHWC channel axis and uses 128-bit accesses when alignment allows.
#include "sn_common.cuh"

#include <math.h>
#include <string.h>

namespace sn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SN_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return SN_OK;
}
int num_sms() {
  static int cached = 0;  // immutable after first query (per-process device properties cache)
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// ---------------------------------------------------------------------------------------------
// ReLU moment gate
// ---------------------------------------------------------------------------------------------
__global__ void relu_fwd_kernel(size_t n4, size_t n, const float* __restrict__ mu, const float* __restrict__ var,
                                float* __restrict__ mu_out, float* __restrict__ var_out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = i; v < n4; v += stride) {
    float4 m = reinterpret_cast<const float4*>(mu)[v];
    float4 s = reinterpret_cast<const float4*>(var)[v];
    float4 mo, so;
    mo.x = fmaxf(m.x, 0.f); so.x = m.x > 0.f ? s.x : 0.f;
    mo.y = fmaxf(m.y, 0.f); so.y = m.y > 0.f ? s.y : 0.f;
    mo.z = fmaxf(m.z, 0.f); so.z = m.z > 0.f ? s.z : 0.f;
    mo.w = fmaxf(m.w, 0.f); so.w = m.w > 0.f ? s.w : 0.f;
    reinterpret_cast<float4*>(mu_out)[v] = mo;
    reinterpret_cast<float4*>(var_out)[v] = so;
  }
  for (size_t e = n4 * 4 + i; e < n; e += stride) {
    float m = mu[e], s = var[e];
    mu_out[e] = fmaxf(m, 0.f);
    var_out[e] = m > 0.f ? s : 0.f;
  }
}

__global__ void relu_bwd_kernel(size_t n, const float* __restrict__ mu_in, const float* __restrict__ gm,
                                const float* __restrict__ gv, float* __restrict__ gm_in, float* __restrict__ gv_in) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = i; e < n; e += stride) {
    bool on = mu_in[e] > 0.f;
    gm_in[e] = on ? gm[e] : 0.f;
    gv_in[e] = on ? gv[e] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// 2x2/2 SAME max-pool of the mean, variance taken at the arg-max (first max in row-major order)
// ---------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(int B, int H, int W, int C, int Ho, int Wo, const float* __restrict__ mu,
                                   const float* __restrict__ var, float* __restrict__ mu_out,
                                   float* __restrict__ var_out, uint8_t* __restrict__ amax) {
  size_t total = (size_t)B * Ho * Wo * C;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += stride) {
    int c = (int)(o % C);
    size_t t = o / C;
    int xo = (int)(t % Wo);
    t /= Wo;
    int yo = (int)(t % Ho);
    int b = (int)(t / Ho);
    float best = -INFINITY, bv = 0.f;
    int bi = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      int y = yo * 2 + (d >> 1), x = xo * 2 + (d & 1);
      if (y < H && x < W) {
        size_t idx = (((size_t)b * H + y) * W + x) * C + c;
        float m = mu[idx];
        if (m > best) { best = m; bv = var[idx]; bi = d; }
      }
    }
    mu_out[o] = best;
    var_out[o] = bv;
    if (amax) amax[o] = (uint8_t)bi;
  }
}

__global__ void maxpool_bwd_kernel(int B, int H, int W, int C, int Ho, int Wo, const uint8_t* __restrict__ amax,
                                   const float* __restrict__ gm, const float* __restrict__ gv,
                                   float* __restrict__ gm_in, float* __restrict__ gv_in) {
  // one thread per INPUT element: gathers from its window's output, so the write is coalesced and complete
  size_t total = (size_t)B * H * W * C;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    size_t t = i / C;
    int x = (int)(t % W);
    t /= W;
    int y = (int)(t % H);
    int b = (int)(t / H);
    size_t o = (((size_t)b * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c;
    bool hit = amax[o] == (uint8_t)(((y & 1) << 1) | (x & 1));
    gm_in[i] = hit ? gm[o] : 0.f;
    gv_in[i] = hit ? gv[o] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// window copy / fill
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void window_copy_kernel(sn_window w, const float* __restrict__ src, float* __restrict__ dst) {
  int cv = w.c / VEC;
  size_t total = (size_t)w.batch * w.h * w.w * cv;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % cv) * VEC;
    size_t t = i / cv;
    int x = (int)(t % w.w);
    t /= w.w;
    int y = (int)(t % w.h);
    int b = (int)(t / w.h);
    const int ss = w.src_step > 1 ? w.src_step : 1;
    size_t s = (((size_t)b * w.src_h + w.src_y0 + (size_t)ss * y) * w.src_w + w.src_x0 + (size_t)ss * x) * w.src_c +
               w.src_c0 + c;
    size_t d = (((size_t)b * w.dst_h + w.dst_y0 + (size_t)w.dst_step * y) * w.dst_w + w.dst_x0 +
                (size_t)w.dst_step * x) * w.dst_c + w.dst_c0 + c;
    if (VEC == 4)
      *reinterpret_cast<float4*>(dst + d) = *reinterpret_cast<const float4*>(src + s);
    else
      dst[d] = src[s];
  }
}

__global__ void fill_kernel(float* __restrict__ dst, size_t n, float v) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

// ---------------------------------------------------------------------------------------------
// softmax with Jacobian-propagated variance; one thread per pixel row, C <= 8 in registers
// ---------------------------------------------------------------------------------------------
constexpr int kMaxC = 8;

template <int C>
__device__ __forceinline__ void softmax_row(const float* m, float* p) {
  float mx = m[0];
#pragma unroll
  for (int i = 1; i < C; ++i) mx = fmaxf(mx, m[i]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) { p[i] = expf(m[i] - mx); sum += p[i]; }
  float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < C; ++i) p[i] *= inv;
}

// var_out_i = sum_j J_ij^2 v_j, J_ij = p_i (delta_ij - p_j).  The direct sum of non-negative terms
// (no (1-2p) cancellation) keeps the variance >= 0.
template <int C>
__device__ __forceinline__ void softmax_var(const float* p, const float* v, float* vo) {
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      float J = p[i] * ((i == j ? 1.f : 0.f) - p[j]);
      a = fmaf(J * J, v[j], a);
    }
    vo[i] = a;
  }
}

template <int C>
__global__ void softmax_fwd_kernel(size_t rows, const float* __restrict__ mu, const float* __restrict__ var,
                                   float* __restrict__ p_out, float* __restrict__ v_out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += stride) {
    float m[C], v[C], p[C], vo[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { m[i] = mu[r * C + i]; v[i] = var[r * C + i]; }
    softmax_row<C>(m, p);
    softmax_var<C>(p, v, vo);
#pragma unroll
    for (int i = 0; i < C; ++i) { p_out[r * C + i] = p[i]; v_out[r * C + i] = vo[i]; }
  }
}

template <int C>
__global__ void softmax_bwd_kernel(size_t rows, const float* __restrict__ p_in, const float* __restrict__ var_in,
                                   const float* __restrict__ g_p, const float* __restrict__ g_vo,
                                   float* __restrict__ g_mu, float* __restrict__ g_var) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += stride) {
    float p[C], v[C], gp[C], gvo[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
      p[i] = p_in[r * C + i]; v[i] = var_in[r * C + i];
      gp[i] = g_p[r * C + i]; gvo[i] = g_vo[r * C + i];
    }
    float gv[C], gpt[C];
#pragma unroll
    for (int j = 0; j < C; ++j) { gv[j] = 0.f; gpt[j] = gp[j]; }
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float a = 0.f;  // sum_j J_ij v_j (delta_ij - p_j)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        float d = (i == j ? 1.f : 0.f) - p[j];
        float J = p[i] * d;
        gv[j] += gvo[i] * J * J;
        a += J * v[j] * d;
        gpt[j] -= gvo[i] * 2.f * J * v[j] * p[i];   // - 2 J_ik v_k p_i
      }
      gpt[i] += gvo[i] * 2.f * a;
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) dot += gpt[k] * p[k];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      g_mu[r * C + j] = p[j] * (gpt[j] - dot);   // sum_k gpt_k J_kj
      g_var[r * C + j] = gv[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian NLL
// ---------------------------------------------------------------------------------------------
constexpr float kNllEps = 1e-3f;

__global__ void nll_fwd_kernel(size_t rows, int C, const float* __restrict__ y, const float* __restrict__ p,
                               const float* __restrict__ var, float lo, float hi, double* __restrict__ acc) {
  double q = 0.0, l = 0.0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += stride) {
    float qr = 0.f, prod = 1.f;
    for (int c = 0; c < C; ++c) {
      float v = fminf(fmaxf(var[r * C + c], lo), hi) + kNllEps;
      float d = p[r * C + c] - y[r * C + c];
      qr += d * d * (1.f / v);
      prod *= v;
    }
    q += (double)qr;
    l += (double)logf(prod);
  }
  q = warp_sum_d(q);
  l = warp_sum_d(l);
  __shared__ double sq[32], sl[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sq[wid] = q; sl[wid] = l; }
  __syncthreads();
  if (wid == 0) {
    int nw = blockDim.x >> 5;
    q = lane < nw ? sq[lane] : 0.0;
    l = lane < nw ? sl[lane] : 0.0;
    q = warp_sum_d(q);
    l = warp_sum_d(l);
    if (lane == 0) { atomicAdd(acc, q); atomicAdd(acc + 1, l); }
  }
}

__global__ void nll_finalize_kernel(size_t rows, const double* __restrict__ acc, float* __restrict__ loss) {
  float q = (float)(acc[0] / (double)rows);
  if (isnan(q) || isinf(q)) q = 0.f;  // Brats.py:304-305
  float l = (float)(acc[1] / (double)rows);
  loss[0] = 0.5f * (q + l);
}

__global__ void nll_bwd_kernel(size_t rows, int C, const float* __restrict__ y, const float* __restrict__ p,
                               const float* __restrict__ var, float lo, float hi, const double* __restrict__ acc,
                               const float* __restrict__ g_loss, float* __restrict__ g_p, float* __restrict__ g_var) {
  float qmean = (float)(acc[0] / (double)rows);
  float qon = (isnan(qmean) || isinf(qmean)) ? 0.f : 1.f;
  float g = 0.5f * g_loss[0] / (float)rows;
  size_t total = rows * (size_t)C;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    float vr = var[i];
    float v = fminf(fmaxf(vr, lo), hi) + kNllEps;
    float inv = 1.f / v;
    float d = p[i] - y[i];
    g_p[i] = qon * g * 2.f * d * inv;
    bool pass = vr >= lo && vr <= hi;
    g_var[i] = pass ? g * (inv - qon * d * d * inv * inv) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// KL regulariser: sum w^2  -  k^2 * mean_n(1 + log s_n - s_n)
// ---------------------------------------------------------------------------------------------
__global__ void kl_fwd_kernel(const float* __restrict__ w, size_t n_w, const float* __restrict__ ws, int cout,
                              int ksize, double* __restrict__ acc) {
  double a = 0.0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n_w; i += stride) { float x = w[i]; a += (double)(x * x); }
  for (size_t i = t0; i < (size_t)cout; i += stride) {
    float s = softplus_f(ws[i]);
    a += -(double)(ksize * ksize) * (double)(1.f + logf(s) - s) / (double)cout;
  }
  a = warp_sum_d(a);
  __shared__ double sh[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = a;
  __syncthreads();
  if (wid == 0) {
    int nw = blockDim.x >> 5;
    a = lane < nw ? sh[lane] : 0.0;
    a = warp_sum_d(a);
    if (lane == 0) atomicAdd(acc, a);
  }
}

__global__ void kl_bwd_kernel(const float* __restrict__ w, size_t n_w, const float* __restrict__ ws, int cout,
                              int ksize, float scale, float* __restrict__ g_w, float* __restrict__ g_ws) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n_w; i += stride) g_w[i] += scale * 2.f * w[i];
  for (size_t i = t0; i < (size_t)cout; i += stride) {
    float x = ws[i];
    float s = softplus_f(x);
    g_ws[i] += scale * (-(float)(ksize * ksize) / (float)cout) * (1.f / s - 1.f) * sigmoid_f(x);
  }
}

}  // namespace sn

using namespace sn;

extern "C" {

int sn_version(void) { return SN_ABI_VERSION; }
const char* sn_last_error(void) { return sn::g_err; }

int sn_device_check(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(SN_ERR_ARCH, "no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  SN_REQUIRE(major == 10, SN_ERR_ARCH, "device compute capability %d.x is not sm_100", major);
  return SN_OK;
}

int sn_relu_moments_fwd(size_t n, const float* mu, const float* var, float* mu_out, float* var_out, sn_stream_t st) {
  SN_REQUIRE(mu && var && mu_out && var_out, SN_ERR_BAD_ARG, "relu_fwd: null pointer");
  if (n == 0) return SN_OK;
  bool vec = aligned16(mu) && aligned16(var) && aligned16(mu_out) && aligned16(var_out);
  size_t n4 = vec ? n / 4 : 0;
  relu_fwd_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, as_stream(st)>>>(n4, n, mu, var, mu_out, var_out);
  return check_launch("relu_fwd");
}

int sn_relu_moments_bwd(size_t n, const float* mu_in, const float* g_mu_out, const float* g_var_out, float* g_mu_in,
                        float* g_var_in, sn_stream_t st) {
  SN_REQUIRE(mu_in && g_mu_out && g_var_out && g_mu_in && g_var_in, SN_ERR_BAD_ARG, "relu_bwd: null pointer");
  if (n == 0) return SN_OK;
  relu_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(n, mu_in, g_mu_out, g_var_out, g_mu_in, g_var_in);
  return check_launch("relu_bwd");
}

int sn_maxpool2_moments_fwd(int32_t B, int32_t H, int32_t W, int32_t C, const float* mu, const float* var,
                            float* mu_out, float* var_out, uint8_t* argmax_out, sn_stream_t st) {
  SN_REQUIRE(mu && var && mu_out && var_out, SN_ERR_BAD_ARG, "maxpool_fwd: null pointer");
  SN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, SN_ERR_BAD_ARG, "maxpool_fwd: bad shape %d %d %d %d", B, H, W, C);
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  size_t total = (size_t)B * Ho * Wo * C;
  maxpool_fwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(B, H, W, C, Ho, Wo, mu, var, mu_out, var_out,
                                                                    argmax_out);
  return check_launch("maxpool_fwd");
}

int sn_maxpool2_moments_bwd(int32_t B, int32_t H, int32_t W, int32_t C, const uint8_t* argmax, const float* g_mu_out,
                            const float* g_var_out, float* g_mu_in, float* g_var_in, sn_stream_t st) {
  SN_REQUIRE(argmax && g_mu_out && g_var_out && g_mu_in && g_var_in, SN_ERR_BAD_ARG, "maxpool_bwd: null pointer");
  SN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, SN_ERR_BAD_ARG, "maxpool_bwd: bad shape");
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  size_t total = (size_t)B * H * W * C;
  maxpool_bwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(B, H, W, C, Ho, Wo, argmax, g_mu_out, g_var_out,
                                                                    g_mu_in, g_var_in);
  return check_launch("maxpool_bwd");
}

int sn_window_copy(const sn_window* w, const float* src, float* dst, sn_stream_t st) {
  SN_REQUIRE(w && src && dst, SN_ERR_BAD_ARG, "window_copy: null pointer");
  SN_REQUIRE(w->batch > 0 && w->h > 0 && w->w > 0 && w->c > 0, SN_ERR_BAD_ARG, "window_copy: empty window");
  SN_REQUIRE(w->dst_step == 1 || w->dst_step == 2, SN_ERR_UNSUPPORTED, "window_copy: dst_step %d", w->dst_step);
  SN_REQUIRE(w->src_step >= 0 && w->src_step <= 2, SN_ERR_UNSUPPORTED, "window_copy: src_step %d", w->src_step);
  const int sstep = w->src_step > 1 ? w->src_step : 1;
  SN_REQUIRE(w->src_y0 >= 0 && w->src_x0 >= 0 && w->src_c0 >= 0 && w->src_y0 + sstep * (w->h - 1) < w->src_h &&
                 w->src_x0 + sstep * (w->w - 1) < w->src_w && w->src_c0 + w->c <= w->src_c,
             SN_ERR_BAD_ARG, "window_copy: source window out of bounds");
  SN_REQUIRE(w->dst_y0 >= 0 && w->dst_x0 >= 0 && w->dst_c0 >= 0 &&
                 w->dst_y0 + w->dst_step * (w->h - 1) < w->dst_h && w->dst_x0 + w->dst_step * (w->w - 1) < w->dst_w &&
                 w->dst_c0 + w->c <= w->dst_c,
             SN_ERR_BAD_ARG, "window_copy: destination window out of bounds");
  bool vec = (w->c % 4 == 0) && (w->src_c % 4 == 0) && (w->dst_c % 4 == 0) && (w->src_c0 % 4 == 0) &&
             (w->dst_c0 % 4 == 0) && aligned16(src) && aligned16(dst);
  size_t total = (size_t)w->batch * w->h * w->w * (vec ? w->c / 4 : w->c);
  if (vec)
    window_copy_kernel<4><<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(*w, src, dst);
  else
    window_copy_kernel<1><<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(*w, src, dst);
  return check_launch("window_copy");
}

int sn_fill(float* dst, size_t n, float value, sn_stream_t st) {
  SN_REQUIRE(dst, SN_ERR_BAD_ARG, "fill: null pointer");
  if (n == 0) return SN_OK;
  fill_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(dst, n, value);
  return check_launch("fill");
}

#define SN_DISPATCH_C(C, CALL)                  \
  switch (C) {                                  \
    case 1: { constexpr int kC = 1; CALL; } break; \
    case 2: { constexpr int kC = 2; CALL; } break; \
    case 3: { constexpr int kC = 3; CALL; } break; \
    case 4: { constexpr int kC = 4; CALL; } break; \
    case 5: { constexpr int kC = 5; CALL; } break; \
    case 6: { constexpr int kC = 6; CALL; } break; \
    case 7: { constexpr int kC = 7; CALL; } break; \
    case 8: { constexpr int kC = 8; CALL; } break; \
  }

int sn_softmax_moments_fwd(size_t rows, int32_t c, const float* mu, const float* var, float* p_out, float* var_out,
                           sn_stream_t st) {
  SN_REQUIRE(mu && var && p_out && var_out, SN_ERR_BAD_ARG, "softmax_fwd: null pointer");
  SN_REQUIRE(c >= 1 && c <= kMaxC, SN_ERR_UNSUPPORTED, "softmax_fwd: %d classes (max %d)", c, kMaxC);
  if (rows == 0) return SN_OK;
  int grid = ew_grid(rows, 128);
  SN_DISPATCH_C(c, (softmax_fwd_kernel<kC><<<grid, 128, 0, as_stream(st)>>>(rows, mu, var, p_out, var_out)));
  return check_launch("softmax_fwd");
}

int sn_softmax_moments_bwd(size_t rows, int32_t c, const float* p, const float* var_in, const float* g_p,
                           const float* g_var_out, float* g_mu, float* g_var_in, sn_stream_t st) {
  SN_REQUIRE(p && var_in && g_p && g_var_out && g_mu && g_var_in, SN_ERR_BAD_ARG, "softmax_bwd: null pointer");
  SN_REQUIRE(c >= 1 && c <= kMaxC, SN_ERR_UNSUPPORTED, "softmax_bwd: %d classes (max %d)", c, kMaxC);
  if (rows == 0) return SN_OK;
  int grid = ew_grid(rows, 128);
  SN_DISPATCH_C(c, (softmax_bwd_kernel<kC><<<grid, 128, 0, as_stream(st)>>>(rows, p, var_in, g_p, g_var_out, g_mu,
                                                                           g_var_in)));
  return check_launch("softmax_bwd");
}

int sn_nll_gaussian_fwd(size_t rows, int32_t c, const float* y, const float* p, const float* var, float clip_lo,
                        float clip_hi, double* acc, float* loss_out, sn_stream_t st) {
  SN_REQUIRE(y && p && var && acc && loss_out, SN_ERR_BAD_ARG, "nll_fwd: null pointer");
  SN_REQUIRE(rows > 0 && c >= 1, SN_ERR_BAD_ARG, "nll_fwd: empty input");
  cudaError_t e = cudaMemsetAsync(acc, 0, 2 * sizeof(double), as_stream(st));
  if (e != cudaSuccess) return fail(SN_ERR_LAUNCH, "nll_fwd memset: %s", cudaGetErrorString(e));
  nll_fwd_kernel<<<ew_grid(rows, 256, 4), 256, 0, as_stream(st)>>>(rows, c, y, p, var, clip_lo, clip_hi, acc);
  nll_finalize_kernel<<<1, 1, 0, as_stream(st)>>>(rows, acc, loss_out);
  return check_launch("nll_fwd");
}

int sn_nll_gaussian_bwd(size_t rows, int32_t c, const float* y, const float* p, const float* var, float clip_lo,
                        float clip_hi, const double* acc, const float* g_loss, float* g_p, float* g_var,
                        sn_stream_t st) {
  SN_REQUIRE(y && p && var && acc && g_loss && g_p && g_var, SN_ERR_BAD_ARG, "nll_bwd: null pointer");
  SN_REQUIRE(rows > 0 && c >= 1, SN_ERR_BAD_ARG, "nll_bwd: empty input");
  nll_bwd_kernel<<<ew_grid(rows * c, 256), 256, 0, as_stream(st)>>>(rows, c, y, p, var, clip_lo, clip_hi, acc, g_loss,
                                                                   g_p, g_var);
  return check_launch("nll_bwd");
}

int sn_kl_regularizer_fwd(const float* w_mu, size_t n_w, const float* w_sigma, int32_t cout, int32_t ksize,
                          double* acc, sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && acc, SN_ERR_BAD_ARG, "kl_fwd: null pointer");
  SN_REQUIRE(n_w > 0 && cout > 0 && ksize > 0, SN_ERR_BAD_ARG, "kl_fwd: bad sizes");
  kl_fwd_kernel<<<ew_grid(n_w, 256, 1), 256, 0, as_stream(st)>>>(w_mu, n_w, w_sigma, cout, ksize, acc);
  return check_launch("kl_fwd");
}

int sn_kl_regularizer_bwd(const float* w_mu, size_t n_w, const float* w_sigma, int32_t cout, int32_t ksize,
                          float scale, float* g_w_mu, float* g_w_sigma, sn_stream_t st) {
  SN_REQUIRE(w_mu && w_sigma && g_w_mu && g_w_sigma, SN_ERR_BAD_ARG, "kl_bwd: null pointer");
  SN_REQUIRE(n_w > 0 && cout > 0 && ksize > 0, SN_ERR_BAD_ARG, "kl_bwd: bad sizes");
  kl_bwd_kernel<<<ew_grid(n_w, 256), 256, 0, as_stream(st)>>>(w_mu, n_w, w_sigma, cout, ksize, scale, g_w_mu,
                                                             g_w_sigma);
  return check_launch("kl_bwd");
}

}  // extern "C"
