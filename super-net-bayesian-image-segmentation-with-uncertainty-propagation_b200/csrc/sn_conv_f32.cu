// FP32 mode of the moment convolution (CUDA-core implicit GEMM, fp32 operands + fp32 accumulation).
//
// One kernel computes two GEMMs that share every A tile plus a rank-1 side reduction:
//   forward : o1 = a1 (*) W,  o2 = a2 (*) W^2 + s[n] * r,   r[m] = sum_K (a1^2 + a2)      (Brats.py:118-137)
//   dgrad   : o1 = g1 (*)^T W + 2 mu . r,  o2 = g2 (*)^T W^2 + r,  r[m] = sum_{tap,n} g2 s[n]  
// The data gradient is the same implicit GEMM run over the (k-1)-padded output gradient with the
// weight tensor addressed flipped/transposed, so forward and backward mirror one design.
// The weight gradient is a second kernel (reduction over pixels, split across CTAs + atomics).
#include "sn_common.cuh"

namespace sn {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, CONV_THREADS = 256;

struct ConvP {
  const float* a1;
  const float* a2;      // may be null (first layer / deterministic input)
  const float* w;       // HWIO base
  const float* ws;      // raw w_sigma
  float* o1;
  float* o2;            // may be null
  float* rsum;          // optional (forward)
  const float* mu_in;   // dgrad: conv input mean for the 2 mu r term
  int B, H, W, Cin;     // GEMM-K side tensor (a1/a2): dims and channels
  int Ho, Wo, N;        // output dims and GEMM-N
  int k, pad;
  int w_tap_stride, w_kstride, w_nstride, flip;
  int mode;             // 0 forward, 1 dgrad
  int relu;
};

__global__ void __launch_bounds__(CONV_THREADS, 2) conv_moments_f32_kernel(ConvP p) {
  __shared__ __align__(16) float As1[BK][BM];
  __shared__ __align__(16) float As2[BK][BM];
  __shared__ __align__(16) float Bs1[BK][BN];
  __shared__ __align__(16) float Bs2[BK][BN];
  __shared__ float Ss[BK];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long M = (long long)p.B * p.Ho * p.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = p.k * p.k * p.Cin;
  const bool vec4 = (p.Cin % 4 == 0);

  // loader mapping: row lm of the tile, k-quad lq
  const int lm = tid % BM, lq = tid / BM;
  long long gm = m0 + lm;
  bool m_ok = gm < M;
  int pb = 0, py = 0, px = 0;
  if (m_ok) {
    px = (int)(gm % p.Wo);
    long long t = gm / p.Wo;
    py = (int)(t % p.Ho);
    pb = (int)(t / p.Ho);
  }
  const int ln = tid % BN;  // weight loader: column ln, k-quad lq

  float acc1[TM][TN], acc2[TM][TN], r[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    r[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) acc1[i][j] = acc2[i][j] = 0.f;
  }

  float ra1[4], ra2[4], rb[4], rs[4];

  auto load_tile = [&](int k0) {
    int kk0 = k0 + lq * 4;
    // ---- A operands
#pragma unroll
    for (int j = 0; j < 4; ++j) { ra1[j] = 0.f; ra2[j] = 0.f; }
    if (m_ok) {
      if (vec4) {
        if (kk0 < Ktot) {
          int tap = kk0 / p.Cin, ci = kk0 - tap * p.Cin;
          int kh = tap / p.k, kw = tap - kh * p.k;
          int iy = py - p.pad + kh, ix = px - p.pad + kw;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            size_t off = (((size_t)pb * p.H + iy) * p.W + ix) * p.Cin + ci;
            float4 v = *reinterpret_cast<const float4*>(p.a1 + off);
            ra1[0] = v.x; ra1[1] = v.y; ra1[2] = v.z; ra1[3] = v.w;
            if (p.a2) {
              float4 u = *reinterpret_cast<const float4*>(p.a2 + off);
              ra2[0] = u.x; ra2[1] = u.y; ra2[2] = u.z; ra2[3] = u.w;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int kk = kk0 + j;
          if (kk < Ktot) {
            int tap = kk / p.Cin, ci = kk - tap * p.Cin;
            int kh = tap / p.k, kw = tap - kh * p.k;
            int iy = py - p.pad + kh, ix = px - p.pad + kw;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
              size_t off = (((size_t)pb * p.H + iy) * p.W + ix) * p.Cin + ci;
              ra1[j] = p.a1[off];
              if (p.a2) ra2[j] = p.a2[off];
            }
          }
        }
      }
    }
    // ---- B operand (W; W^2 is formed at the smem store) + per-K-channel softplus for dgrad
    int n = n0 + ln;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = kk0 + j;
      rb[j] = 0.f;
      rs[j] = 0.f;
      if (kk < Ktot) {
        int tap = kk / p.Cin, kc = kk - tap * p.Cin;
        if (p.flip) tap = p.k * p.k - 1 - tap;
        if (n < p.N) rb[j] = p.w[(size_t)tap * p.w_tap_stride + (size_t)kc * p.w_kstride + (size_t)n * p.w_nstride];
        if (p.mode == 1 && ln == 0) rs[j] = softplus_f(p.ws[kc]);
      }
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As1[lq * 4 + j][lm] = ra1[j];
      As2[lq * 4 + j][lm] = ra2[j];
      Bs1[lq * 4 + j][ln] = rb[j];
      Bs2[lq * 4 + j][ln] = rb[j] * rb[j];
      if (p.mode == 1 && ln == 0) Ss[lq * 4 + j] = rs[j];
    }
  };

  load_tile(0);
  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    store_tile();
    __syncthreads();
    if (k0 + BK < Ktot) load_tile(k0 + BK);  // register prefetch of the next tile overlaps the FMAs
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a1 = *reinterpret_cast<const float4*>(&As1[kk][ty * TM]);
      float4 a2 = *reinterpret_cast<const float4*>(&As2[kk][ty * TM]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs1[kk][tx * TN]);
      float4 b2 = *reinterpret_cast<const float4*>(&Bs2[kk][tx * TN]);
      float av1[4] = {a1.x, a1.y, a1.z, a1.w}, av2[4] = {a2.x, a2.y, a2.z, a2.w};
      float bv1[4] = {b1.x, b1.y, b1.z, b1.w}, bv2[4] = {b2.x, b2.y, b2.z, b2.w};
      if (p.mode == 0) {
#pragma unroll
        for (int i = 0; i < TM; ++i) r[i] += fmaf(av1[i], av1[i], av2[i]);
      } else {
        float s = Ss[kk];
#pragma unroll
        for (int i = 0; i < TM; ++i) r[i] = fmaf(av2[i], s, r[i]);
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc1[i][j] = fmaf(av1[i], bv1[j], acc1[i][j]);
          acc2[i][j] = fmaf(av2[i], bv2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }

  // ---- epilogue
  float sn_[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int n = n0 + tx * TN + j;
    sn_[j] = (p.mode == 0 && n < p.N) ? softplus_f(p.ws[n]) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    if (p.mode == 0 && p.rsum && tx == 0 && blockIdx.y == 0) p.rsum[m] = r[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      size_t o = (size_t)m * p.N + n;
      float v1 = acc1[i][j], v2 = acc2[i][j];
      if (p.mode == 0) {
        v2 = fmaxf(fmaf(sn_[j], r[i], v2), 0.f);   // every term is >= 0; clamp guards rounding only
        if (p.relu) { v2 = v1 > 0.f ? v2 : 0.f; v1 = fmaxf(v1, 0.f); }
      } else {
        v1 = fmaf(2.f * p.mu_in[o], r[i], v1);
        v2 += r[i];
      }
      p.o1[o] = v1;
      if (p.o2) p.o2[o] = v2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------
constexpr int WG_T = 32, WG_PK = 32, WG_THREADS = 256;

struct WgradP {
  const float* a1;
  const float* a2;  // may be null
  const float* g1;
  const float* g2;
  const float* w;
  float* gw;
  int B, H, W, Cin, Ho, Wo, N, k;
  int ci_tiles, n_tiles;
  long long pix_per_split;
};

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_f32_kernel(WgradP p) {
  __shared__ float A1[WG_PK][WG_T + 1], A2[WG_PK][WG_T + 1], G1[WG_PK][WG_T + 1], G2[WG_PK][WG_T + 1];
  const int tid = threadIdx.x;
  const int ct = blockIdx.x % p.ci_tiles, nt = blockIdx.x / p.ci_tiles;
  const int tap = blockIdx.y;
  const int kh = tap / p.k, kw = tap - kh * p.k;
  const int c0 = ct * WG_T, n0 = nt * WG_T;
  const long long P = (long long)p.B * p.Ho * p.Wo;
  const long long pbeg = (long long)blockIdx.z * p.pix_per_split;
  const long long pend = pbeg + p.pix_per_split < P ? pbeg + p.pix_per_split : P;

  const int tn = tid % 32, tg = tid / 32;  // outputs: n = n0+tn, ci = c0 + tg*4 .. +3
  float acc1[4] = {0, 0, 0, 0}, acc2[4] = {0, 0, 0, 0};

  for (long long pc = pbeg; pc < pend; pc += WG_PK) {
    // load: thread -> (pixel row tid/32 + 8*i, column tid%32)
#pragma unroll
    for (int i = 0; i < WG_PK / 8; ++i) {
      int pr = tg + 8 * i;
      long long pix = pc + pr;
      float va1 = 0.f, va2 = 0.f, vg1 = 0.f, vg2 = 0.f;
      if (pix < pend) {
        int x = (int)(pix % p.Wo);
        long long t = pix / p.Wo;
        int y = (int)(t % p.Ho);
        int b = (int)(t / p.Ho);
        int c = c0 + tn, n = n0 + tn;
        if (c < p.Cin) {
          size_t off = (((size_t)b * p.H + y + kh) * p.W + x + kw) * p.Cin + c;
          va1 = p.a1[off];
          if (p.a2) va2 = p.a2[off];
        }
        if (n < p.N) {
          size_t off = (size_t)pix * p.N + n;
          vg1 = p.g1[off];
          vg2 = p.g2[off];
        }
      }
      A1[pr][tn] = va1; A2[pr][tn] = va2; G1[pr][tn] = vg1; G2[pr][tn] = vg2;
    }
    __syncthreads();
#pragma unroll 8
    for (int pr = 0; pr < WG_PK; ++pr) {
      float g1 = G1[pr][tn], g2 = G2[pr][tn];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc1[j] = fmaf(A1[pr][tg * 4 + j], g1, acc1[j]);
        acc2[j] = fmaf(A2[pr][tg * 4 + j], g2, acc2[j]);
      }
    }
    __syncthreads();
  }
  int n = n0 + tn;
  if (n < p.N) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c0 + tg * 4 + j;
      if (c < p.Cin) {
        size_t o = ((size_t)tap * p.Cin + c) * p.N + n;
        atomicAdd(p.gw + o, acc1[j] + 2.f * p.w[o] * acc2[j]);
      }
    }
  }
}

// g_w_sigma[n] = sigmoid(w_sigma[n]) * sum_p g2[p,n] * rsum[p]
__global__ void conv_wsigma_grad_kernel(long long P, int N, const float* __restrict__ g2,
                                        const float* __restrict__ rsum, const float* __restrict__ ws,
                                        float* __restrict__ gws, long long pix_per_block) {
  long long pbeg = (long long)blockIdx.x * pix_per_block;
  long long pend = pbeg + pix_per_block < P ? pbeg + pix_per_block : P;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float a = 0.f;
    for (long long pix = pbeg; pix < pend; ++pix) a = fmaf(g2[(size_t)pix * N + n], rsum[pix], a);
    atomicAdd(gws + n, a * sigmoid_f(ws[n]));
  }
}

static int check_desc(const sn_conv_desc* d, const char* who) {
  SN_REQUIRE(d, SN_ERR_BAD_ARG, "%s: null descriptor", who);
  SN_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0 && d->cin > 0 && d->cout > 0, SN_ERR_BAD_ARG,
             "%s: non-positive dimension", who);
  SN_REQUIRE(d->ksize >= 1 && d->ksize <= 7, SN_ERR_UNSUPPORTED, "%s: kernel size %d", who, d->ksize);
  SN_REQUIRE(d->in_h >= d->ksize && d->in_w >= d->ksize, SN_ERR_BAD_ARG, "%s: input smaller than the kernel", who);
  return SN_OK;
}

}  // namespace sn

using namespace sn;

extern "C" {

int sn_conv_moments_fwd(const sn_conv_desc* d, const float* mu_in, const float* var_in, const float* w_mu,
                        const float* w_sigma, float* mu_out, float* var_out, float* rsum_out, sn_stream_t st) {
  int rc = check_desc(d, "conv_fwd");
  if (rc) return rc;
  SN_REQUIRE(mu_in && w_mu && w_sigma && mu_out && var_out, SN_ERR_BAD_ARG, "conv_fwd: null pointer");
  SN_REQUIRE(aligned16(mu_in) && (!var_in || aligned16(var_in)), SN_ERR_MISALIGNED,
             "conv_fwd: inputs must be 16-byte aligned");
  ConvP p{};
  p.a1 = mu_in; p.a2 = var_in; p.w = w_mu; p.ws = w_sigma; p.o1 = mu_out; p.o2 = var_out; p.rsum = rsum_out;
  p.B = d->batch; p.H = d->in_h; p.W = d->in_w; p.Cin = d->cin;
  p.Ho = d->in_h - d->ksize + 1; p.Wo = d->in_w - d->ksize + 1; p.N = d->cout;
  p.k = d->ksize; p.pad = 0;
  p.w_tap_stride = d->cin * d->cout; p.w_kstride = d->cout; p.w_nstride = 1; p.flip = 0;
  p.mode = 0; p.relu = (d->flags & SN_CONV_RELU) ? 1 : 0;
  long long M = (long long)p.B * p.Ho * p.Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
  conv_moments_f32_kernel<<<grid, CONV_THREADS, 0, as_stream(st)>>>(p);
  return check_launch("conv_moments_f32(fwd)");
}

int sn_conv_moments_bwd_data(const sn_conv_desc* d, const float* g_mu_out, const float* g_var_out,
                             const float* mu_in, const float* w_mu, const float* w_sigma, float* g_mu_in,
                             float* g_var_in, sn_stream_t st) {
  int rc = check_desc(d, "conv_bwd_data");
  if (rc) return rc;
  SN_REQUIRE(g_mu_out && g_var_out && mu_in && w_mu && w_sigma && g_mu_in, SN_ERR_BAD_ARG,
             "conv_bwd_data: null pointer");
  SN_REQUIRE(!(d->flags & SN_CONV_RELU), SN_ERR_UNSUPPORTED, "conv_bwd_data: apply sn_relu_moments_bwd first");
  SN_REQUIRE(aligned16(g_mu_out) && aligned16(g_var_out), SN_ERR_MISALIGNED, "conv_bwd_data: misaligned gradient");
  ConvP p{};
  p.a1 = g_mu_out; p.a2 = g_var_out; p.w = w_mu; p.ws = w_sigma; p.o1 = g_mu_in; p.o2 = g_var_in; p.mu_in = mu_in;
  p.B = d->batch; p.H = d->in_h - d->ksize + 1; p.W = d->in_w - d->ksize + 1; p.Cin = d->cout;
  p.Ho = d->in_h; p.Wo = d->in_w; p.N = d->cin;
  p.k = d->ksize; p.pad = d->ksize - 1;
  p.w_tap_stride = d->cin * d->cout; p.w_kstride = 1; p.w_nstride = d->cout; p.flip = 1;
  p.mode = 1; p.relu = 0;
  long long M = (long long)p.B * p.Ho * p.Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
  conv_moments_f32_kernel<<<grid, CONV_THREADS, 0, as_stream(st)>>>(p);
  return check_launch("conv_moments_f32(dgrad)");
}

int sn_conv_moments_bwd_weight(const sn_conv_desc* d, const float* mu_in, const float* var_in,
                               const float* g_mu_out, const float* g_var_out, const float* rsum, const float* w_mu,
                               const float* w_sigma, float* g_w_mu, float* g_w_sigma, sn_stream_t st) {
  int rc = check_desc(d, "conv_bwd_weight");
  if (rc) return rc;
  SN_REQUIRE(mu_in && g_mu_out && g_var_out && rsum && w_mu && w_sigma && g_w_mu && g_w_sigma, SN_ERR_BAD_ARG,
             "conv_bwd_weight: null pointer");
  cudaStream_t s = as_stream(st);
  size_t nw = (size_t)d->ksize * d->ksize * d->cin * d->cout;
  if (cudaMemsetAsync(g_w_mu, 0, nw * sizeof(float), s) != cudaSuccess ||
      cudaMemsetAsync(g_w_sigma, 0, (size_t)d->cout * sizeof(float), s) != cudaSuccess)
    return fail(SN_ERR_LAUNCH, "conv_bwd_weight: memset failed");
  WgradP p{};
  p.a1 = mu_in; p.a2 = var_in; p.g1 = g_mu_out; p.g2 = g_var_out; p.w = w_mu; p.gw = g_w_mu;
  p.B = d->batch; p.H = d->in_h; p.W = d->in_w; p.Cin = d->cin;
  p.Ho = d->in_h - d->ksize + 1; p.Wo = d->in_w - d->ksize + 1; p.N = d->cout; p.k = d->ksize;
  p.ci_tiles = (d->cin + WG_T - 1) / WG_T;
  p.n_tiles = (d->cout + WG_T - 1) / WG_T;
  long long P = (long long)p.B * p.Ho * p.Wo;
  long long tiles = (long long)p.ci_tiles * p.n_tiles * d->ksize * d->ksize;
  long long want = (4LL * num_sms() + tiles - 1) / tiles;          // ~4 CTAs per SM overall
  long long max_split = (P + 4 * WG_PK - 1) / (4 * WG_PK);
  long long splits = want < 1 ? 1 : (want > max_split ? max_split : want);
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.pix_per_split = ((P + splits - 1) / splits + WG_PK - 1) / WG_PK * WG_PK;
  splits = (P + p.pix_per_split - 1) / p.pix_per_split;
  dim3 grid((unsigned)(p.ci_tiles * p.n_tiles), (unsigned)(d->ksize * d->ksize), (unsigned)splits);
  conv_wgrad_f32_kernel<<<grid, WG_THREADS, 0, s>>>(p);
  rc = check_launch("conv_wgrad_f32");
  if (rc) return rc;
  long long ppb = (P + 2LL * num_sms() - 1) / (2LL * num_sms());
  if (ppb < 64) ppb = 64;
  unsigned nb = (unsigned)((P + ppb - 1) / ppb);
  conv_wsigma_grad_kernel<<<nb, 128, 0, s>>>(P, d->cout, g_var_out, rsum, w_sigma, g_w_sigma, ppb);
  return check_launch("conv_wsigma_grad");
}

}  // extern "C"
