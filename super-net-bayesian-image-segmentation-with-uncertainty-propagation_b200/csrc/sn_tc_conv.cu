// FAST mode moment convolution: one warp-specialised tcgen05 / TMEM implicit-GEMM kernel per layer.
//
//   mean  = mu (*) W            -> TMEM accumulator 0, bf16x3: mu_hi*W_hi + mu_lo*W_hi + mu_hi*W_lo
//   var   = var (*) W^2 + s_n r -> TMEM accumulator 1 (one bf16 UMMA) + rank-1 term in the epilogue,
//   r[m]  = sum_K (mu^2 + var)  -> side reduction over the very A tiles TMA put in shared memory
//
// (Brats.py:118-137; the reference's three patch-matmuls collapse to this because vect_sigma has equal rows,
// Brats.py:121.)  GEMM view: M = output pixels (128 per CTA, linear over (b, y, x): TMA im2col walks rows and
// images), N = output channels (NT per CTA), K = (tap, 32-channel block).  Warp roles (192 threads):
//   warp 0      TMA producer: per K step 3 im2col loads (A planes hi, lo, var) + 3 tiled loads (W_hi, W_lo, W^2)
//   warp 1      TMEM allocation, single-thread UMMA issue (8 tcgen05.mma per K step), commits to mbarriers
//   warps 2-5   side reduction r[m] during the main loop (thread <-> pixel row), then the epilogue:
//               tcgen05.ld -> +s_n r, clamp >= 0, ReLU gate -> bf16 hi/lo/var -> packed window (or fp32) store
// The same kernel runs the transposed-conv up-sampling (unpool + 2x2 conv, Brats.py:178-203,414-415) as four
// parity GEMMs with K = Cin: the N tiles enumerate (parity, channel block) and the epilogue scatters to
// (2y+a, 2x+b).  Padding, crop and concat are address arithmetic on the source/destination windows.
#include "sn_common.cuh"
#include "sn_sm100.cuh"

#include <cuda.h>
#include <mutex>

namespace sn {

constexpr int TC_BM = 128;           // pixels per CTA (UMMA M)
constexpr int TC_KC = 32;            // channels per K step = one 64-byte swizzle row of bf16
constexpr int TC_A_PLANE = TC_BM * TC_KC * 2;   // 8192 B
constexpr int TC_THREADS = 192;

struct TcMaps {
  CUtensorMap a[2][3];   // [source][plane] im2col maps over the activation windows
  CUtensorMap w;         // prepared weights [3*taps][cout][cin]
};

struct TcP {
  int m_total, Ho, Wo;   // GEMM rows and the traversal extents (regular: output h,w; upconv: input h,w)
  int ksize;             // K-side taps per dimension (1 for upconv)
  int taps_w;            // taps in the prepared weights (ksize^2, or 4 for upconv)
  int cblk0, cblk1;      // 32-channel blocks of source 0 / 1
  int cout;              // output channels (per parity group)
  int relu, upconv, dst_f32;
  __nv_bfloat16* dst;
  int dh, dw, dc, dy0, dx0, dc0;
  float* dst_mu;
  float* dst_var;
  int out_h, out_w;
  const float* s;
};

template <int NT>
struct TcCfg {
  static constexpr int B_PLANE = NT * TC_KC * 2;
  static constexpr int STAGE = 3 * TC_A_PLANE + 3 * B_PLANE;
  static constexpr int STAGES = NT >= 128 ? 4 : 3;
  static constexpr int TMEM_COLS = 2 * NT;                       // 64 / 128 / 256: powers of two >= 32
  static constexpr int SMEM = STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_moments_tc_kernel(const __grid_constant__ TcMaps maps,
                                                                        const TcP p) {
  using Cfg = TcCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE;
  // barrier block: full[STAGES], empty[STAGES], tmem_full, then the TMEM base-address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE + 8 * (2 * Cfg::STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = p.cblk0 + p.cblk1;
  const int kiters = p.ksize * p.ksize * cblk;
  const int m0 = blockIdx.x * TC_BM;
  const int ncol0 = blockIdx.y * NT;          // column in the (group, channel) space
  const int group = ncol0 / p.cout;           // parity (a,b) = (group >> 1, group & 1) for upconv, else 0
  const int n0 = ncol0 - group * p.cout;      // first output channel of this tile

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s)
      for (int pl = 0; pl < 3; ++pl) ptx::prefetch_tensormap(&maps.a[s][pl]);
    ptx::prefetch_tensormap(&maps.w);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 5);       // UMMA commit + one arrive per reducer warp
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int x_start = m0 % p.Wo;
      const int t = m0 / p.Wo;
      const int y_start = t % p.Ho;
      const int b_start = t / p.Ho;
      const int wtap_base = p.upconv ? group : 0;
      int it = 0;
      for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
        const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
        for (int cbt = 0; cbt < cblk; ++cbt, ++it) {
          const int stage = it % Cfg::STAGES;
          const uint32_t parity = (uint32_t)(it / Cfg::STAGES) & 1u;
          ptx::mbar_wait(empty_bar(stage), parity ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), (uint32_t)Cfg::STAGE);
          const uint32_t sa = smem_base + stage * Cfg::STAGE;
          const uint32_t sb = sa + 3 * TC_A_PLANE;
          const int src = cbt >= p.cblk0 ? 1 : 0;
          const int cb = src ? cbt - p.cblk0 : cbt;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            ptx::tma_load_im2col_4d(sa + pl * TC_A_PLANE, &maps.a[src][pl], full_bar(stage), cb * TC_KC, x_start,
                                    y_start, b_start, (uint16_t)kw, (uint16_t)kh);
          const int wtap = p.upconv ? wtap_base : tap;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            ptx::tma_load_3d(sb + pl * Cfg::B_PLANE, &maps.w, full_bar(stage), cbt * TC_KC, n0,
                             pl * p.taps_w + wtap);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(TC_BM, NT);
      const uint32_t acc_mu = tmem_base, acc_var = tmem_base + NT;
      for (int it = 0; it < kiters; ++it) {
        const int stage = it % Cfg::STAGES;
        const uint32_t parity = (uint32_t)(it / Cfg::STAGES) & 1u;
        ptx::mbar_wait(full_bar(stage), parity);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::STAGE;
        const uint32_t sb = sa + 3 * TC_A_PLANE;
#pragma unroll
        for (int ks = 0; ks < TC_KC / 16; ++ks) {
          const uint32_t koff = ks * 32;      // 16 bf16 = 32 bytes along K inside the swizzled row
          const uint64_t a_hi = ptx::smem_desc_kmajor<64>(sa + koff);
          const uint64_t a_lo = ptx::smem_desc_kmajor<64>(sa + TC_A_PLANE + koff);
          const uint64_t a_vr = ptx::smem_desc_kmajor<64>(sa + 2 * TC_A_PLANE + koff);
          const uint64_t b_hi = ptx::smem_desc_kmajor<64>(sb + koff);
          const uint64_t b_lo = ptx::smem_desc_kmajor<64>(sb + Cfg::B_PLANE + koff);
          const uint64_t b_sq = ptx::smem_desc_kmajor<64>(sb + 2 * Cfg::B_PLANE + koff);
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          ptx::umma_bf16(acc_mu, a_hi, b_hi, idesc, acc);
          ptx::umma_bf16(acc_mu, a_lo, b_hi, idesc, 1u);
          ptx::umma_bf16(acc_mu, a_hi, b_lo, idesc, 1u);
          ptx::umma_bf16(acc_var, a_vr, b_sq, idesc, acc);
        }
        ptx::umma_commit(empty_bar(stage));   // frees the smem slot once these UMMAs have read it
      }
      ptx::umma_commit(tmem_full_bar);        // accumulators complete
    }
  } else {
    // ===================== side reduction, then epilogue =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;            // pixel row of the tile == TMEM lane
    float r = 0.f;
    for (int it = 0; it < kiters; ++it) {
      const int stage = it % Cfg::STAGES;
      const uint32_t parity = (uint32_t)(it / Cfg::STAGES) & 1u;
      ptx::mbar_wait(full_bar(stage), parity);
      const uint8_t* a = smem_gen + stage * Cfg::STAGE + row * 64;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // the 64-byte swizzle only permutes the four 16-byte chunks inside a row, and a sum does not care;
        // rotating the chunk by row/2 spreads the eight threads of a quarter-warp over all 32 banks
        const int ch = ((j + (row >> 1)) & 3) * 16;
        const uint4 h = *reinterpret_cast<const uint4*>(a + ch);
        const uint4 l = *reinterpret_cast<const uint4*>(a + TC_A_PLANE + ch);
        const uint4 v = *reinterpret_cast<const uint4*>(a + 2 * TC_A_PLANE + ch);
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m_a = bf16lo(hh[e]) + bf16lo(ll[e]);
          const float m_b = bf16hi(hh[e]) + bf16hi(ll[e]);
          acc = fmaf(m_a, m_a, acc);
          acc = fmaf(m_b, m_b, acc);
          acc += bf16lo(vv[e]) + bf16hi(vv[e]);
        }
      }
      r += acc;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(empty_bar(stage));
    }

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();

    const int m = m0 + row;
    const bool valid = m < p.m_total;
    int b = 0, y = 0, x = 0;
    if (valid) {
      x = m % p.Wo;
      const int t = m / p.Wo;
      y = t % p.Ho;
      b = t / p.Ho;
    }
    int oy = y, ox = x;
    if (p.upconv) { oy = 2 * y + (group >> 1); ox = 2 * x + (group & 1); }
    __nv_bfloat16* d_hi = nullptr;
    float *f_mu = nullptr, *f_var = nullptr;
    if (p.dst_f32) {
      const size_t o = (((size_t)b * p.out_h + oy) * p.out_w + ox) * p.cout + n0;
      f_mu = p.dst_mu + o;
      f_var = p.dst_var + o;
    } else {
      d_hi = p.dst + ((((size_t)b * p.dh + oy + p.dy0) * p.dw + ox + p.dx0) * 3) * p.dc + p.dc0 + n0;
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      uint32_t am[16], av[16];
      ptx::tmem_ld16(lane_base + c0, am);
      ptx::tmem_ld16(lane_base + NT + c0, av);
      ptx::tmem_ld_wait();
      float mu[16], var[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float sn = __ldg(p.s + n0 + c0 + j);
        float mj = __uint_as_float(am[j]);
        float vj = fmaxf(fmaf(sn, r, __uint_as_float(av[j])), 0.f);   // every term is >= 0: the clamp guards rounding
        if (p.relu) {
          vj = mj > 0.f ? vj : 0.f;
          mj = fmaxf(mj, 0.f);
        }
        mu[j] = mj;
        var[j] = vj;
      }
      if (valid) {
        if (p.dst_f32) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            *reinterpret_cast<float4*>(f_mu + c0 + j) = make_float4(mu[j], mu[j + 1], mu[j + 2], mu[j + 3]);
            *reinterpret_cast<float4*>(f_var + c0 + j) = make_float4(var[j], var[j + 1], var[j + 2], var[j + 3]);
          }
        } else {
          uint32_t hi[8], lo[8], vr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a0 = mu[2 * j], a1 = mu[2 * j + 1];
            const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
            hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[j] = pack_bf16x2(a0 - __bfloat162float(h0), a1 - __bfloat162float(h1));
            vr[j] = pack_bf16x2(var[2 * j], var[2 * j + 1]);
          }
          uint4* ph = reinterpret_cast<uint4*>(d_hi + c0);
          uint4* pl = reinterpret_cast<uint4*>(d_hi + p.dc + c0);
          uint4* pv = reinterpret_cast<uint4*>(d_hi + 2 * p.dc + c0);
          ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          pv[0] = make_uint4(vr[0], vr[1], vr[2], vr[3]);
          pv[1] = make_uint4(vr[4], vr[5], vr[6], vr[7]);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry points (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DriverApi {
  EncodeTiledFn tiled = nullptr;
  EncodeIm2colFn im2col = nullptr;
  int driver_version = 0;
  bool ok = false;
};

static const DriverApi& driver_api() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1) != cudaSuccess ||
        q1 != cudaDriverEntryPointSuccess || !f1)
      return;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2) != cudaSuccess ||
        q2 != cudaDriverEntryPointSuccess || !f2)
      return;
    api.tiled = reinterpret_cast<EncodeTiledFn>(f1);
    api.im2col = reinterpret_cast<EncodeIm2colFn>(f2);
    cudaDriverGetVersion(&api.driver_version);
    api.ok = true;
  });
  return api;
}

// im2col map over one plane of a packed window: dims (c, w, h, n), 32 channels x 128 pixels per load
static int make_act_map(CUtensorMap* out, const sn_packed_view& v, int plane, int src_c, int batch, int in_h,
                        int in_w, int ksize) {
  const DriverApi& api = driver_api();
  const size_t pix = (size_t)3 * v.c;   // elements per pixel
  char* base = reinterpret_cast<char*>(v.base) +
               ((((size_t)v.y0 * v.w + v.x0) * 3 + plane) * v.c + v.c0) * sizeof(__nv_bfloat16);
  cuuint64_t dims[4] = {(cuuint64_t)src_c, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)batch};
  cuuint64_t strides[3] = {pix * 2, (cuuint64_t)v.w * pix * 2, (cuuint64_t)v.h * v.w * pix * 2};
  int lower[2] = {0, 0};
  int upper[2] = {-(ksize - 1), -(ksize - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = api.im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, lower, upper, TC_KC, TC_BM,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeIm2col failed (%d)", (int)r);
  // Driver quirk (CUDA <= 13.1) for im2col maps over tensors smaller than 128 KiB: bit 21 of the second
  // descriptor word must be cleared (same workaround CUTLASS applies in make_im2col_tma_copy_desc).
  const size_t span = ((size_t)(batch - 1) * strides[2] + (size_t)(in_h - 1) * strides[1] +
                       (size_t)(in_w - 1) * strides[0] + (size_t)src_c * 2);
  if (api.driver_version <= 13010 && span < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return SN_OK;
}

static int make_weight_map(CUtensorMap* out, const void* w_packed, int taps, int cout, int cin, int nt) {
  const DriverApi& api = driver_api();
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)(3 * taps)};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * cout * 2};
  cuuint32_t box[3] = {TC_KC, (cuuint32_t)nt, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SN_OK;
}

static int check_view(const sn_packed_view& v, int batch, int h, int w, int c, const char* who) {
  SN_REQUIRE(v.base && aligned16(v.base), SN_ERR_MISALIGNED, "%s: packed buffer must be 16-byte aligned", who);
  SN_REQUIRE(v.n >= batch && v.h > 0 && v.w > 0 && v.c > 0 && v.c % 8 == 0, SN_ERR_BAD_ARG,
             "%s: bad packed buffer dims [%d,%d,%d,%d]", who, v.n, v.h, v.w, v.c);
  SN_REQUIRE(v.y0 >= 0 && v.x0 >= 0 && v.c0 >= 0 && v.y0 + h <= v.h && v.x0 + w <= v.w && v.c0 + c <= v.c,
             SN_ERR_BAD_ARG, "%s: window [%d+%d, %d+%d, %d+%d] outside buffer [%d,%d,%d]", who, v.y0, h, v.x0, w,
             v.c0, c, v.h, v.w, v.c);
  SN_REQUIRE(v.c0 % 8 == 0, SN_ERR_MISALIGNED, "%s: channel offset %d must be a multiple of 8", who, v.c0);
  return SN_OK;
}

template <int NT>
static int launch_tc(const TcMaps& maps, const TcP& p, int n_tiles, cudaStream_t st) {
  using Cfg = TcCfg<NT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_moments_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM);
  });
  if (attr_err != cudaSuccess) return fail(SN_ERR_LAUNCH, "conv_tc: cannot reserve %d B of shared memory", Cfg::SMEM);
  dim3 grid((unsigned)((p.m_total + TC_BM - 1) / TC_BM), (unsigned)n_tiles);
  conv_moments_tc_kernel<NT><<<grid, TC_THREADS, Cfg::SMEM, st>>>(maps, p);
  return check_launch("conv_moments_tc");
}

int conv_moments_halo_dispatch(const sn_tc_conv_desc* d, cudaStream_t stream, const sn_tc_head_desc* head);   // sn_tc_halo.cu

}  // namespace sn

using namespace sn;

// Argument checks shared by sn_conv_moments_fwd_tc and sn_conv_moments_fwd_tc_head (dst_optional: the fused head may
// run without a packed destination for the 32-channel tensor).
static int conv_tc_validate(const sn_tc_conv_desc* d, bool dst_optional) {
  SN_REQUIRE(d, SN_ERR_BAD_ARG, "conv_tc: null descriptor");
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  const bool dst_f32 = (d->flags & SN_TC_DST_F32) != 0;
  SN_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0 && d->cout > 0, SN_ERR_BAD_ARG, "conv_tc: bad geometry");
  SN_REQUIRE(d->ksize >= 1 && d->ksize <= 3, SN_ERR_UNSUPPORTED, "conv_tc: kernel size %d", d->ksize);
  SN_REQUIRE(!upconv || d->ksize == 2, SN_ERR_BAD_ARG, "conv_tc: SN_TC_UPCONV needs ksize == 2");
  SN_REQUIRE(d->src_c[0] > 0 && d->src_c[0] % TC_KC == 0 && d->src_c[1] >= 0 && d->src_c[1] % TC_KC == 0,
             SN_ERR_UNSUPPORTED, "conv_tc: source channels (%d, %d) must be multiples of %d", d->src_c[0],
             d->src_c[1], TC_KC);
  SN_REQUIRE(d->cout % 32 == 0, SN_ERR_UNSUPPORTED, "conv_tc: cout %d must be a multiple of 32", d->cout);
  SN_REQUIRE(d->w_packed && d->s && aligned16(d->w_packed), SN_ERR_BAD_ARG, "conv_tc: weights missing/misaligned");
  const int keff = upconv ? 1 : d->ksize;
  SN_REQUIRE(d->in_h >= keff && d->in_w >= keff, SN_ERR_BAD_ARG, "conv_tc: input smaller than the kernel");
  const DriverApi& api = driver_api();
  SN_REQUIRE(api.ok, SN_ERR_DRIVER, "conv_tc: cuTensorMapEncode* driver entry points unavailable");

  const int Ho = d->in_h - keff + 1, Wo = d->in_w - keff + 1;
  const int out_h = upconv ? 2 * d->in_h : Ho, out_w = upconv ? 2 * d->in_w : Wo;
  const long long m_total = (long long)d->batch * Ho * Wo;
  SN_REQUIRE(m_total < (1ll << 31) - TC_BM, SN_ERR_UNSUPPORTED, "conv_tc: too many pixels");
  int rc;
  for (int s = 0; s < 2; ++s) {
    if (d->src_c[s] == 0) continue;
    if ((rc = check_view(d->src[s], d->batch, d->in_h, d->in_w, d->src_c[s], "conv_tc src"))) return rc;
  }
  if (dst_f32) {
    SN_REQUIRE(d->dst_mu && d->dst_var && aligned16(d->dst_mu) && aligned16(d->dst_var), SN_ERR_BAD_ARG,
               "conv_tc: fp32 destinations missing/misaligned");
  } else if (!(dst_optional && d->dst.base == nullptr)) {
    if ((rc = check_view(d->dst, d->batch, out_h, out_w, d->cout, "conv_tc dst"))) return rc;
  }
  return SN_OK;
}

extern "C" int sn_conv_moments_fwd_tc_head(const sn_tc_conv_desc* d, const sn_tc_head_desc* h, sn_stream_t st) {
  int rc;
  if ((rc = conv_tc_validate(d, true))) return rc;
  SN_REQUIRE(h, SN_ERR_BAD_ARG, "conv_tc_head: null head descriptor");
  SN_REQUIRE(sn_tc_head_fusable(d, h->n_labels), SN_ERR_UNSUPPORTED,
             "conv_tc_head: this layer / label count cannot end in the fused head (sn_tc_head_fusable)");
  SN_REQUIRE(h->w_mu && h->w_sigma && h->p_out && h->var_out, SN_ERR_BAD_ARG, "conv_tc_head: null weights / outputs");
  SN_REQUIRE((h->presoftmax_mu == nullptr) == (h->presoftmax_var == nullptr), SN_ERR_BAD_ARG,
             "conv_tc_head: presoftmax_mu and presoftmax_var go together");
  SN_REQUIRE(aligned16(h->p_out) && aligned16(h->var_out) && aligned16(h->presoftmax_mu) && aligned16(h->presoftmax_var),
             SN_ERR_BAD_ARG, "conv_tc_head: outputs must be 16-byte aligned");
  return conv_moments_halo_dispatch(d, as_stream(st), h);
}

extern "C" int sn_conv_moments_fwd_tc(const sn_tc_conv_desc* d, sn_stream_t st) {
  int rc;
  if ((rc = conv_tc_validate(d, false))) return rc;
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  const bool dst_f32 = (d->flags & SN_TC_DST_F32) != 0;
  const int keff = upconv ? 1 : d->ksize;
  const int Ho = d->in_h - keff + 1, Wo = d->in_w - keff + 1;
  const int out_h = upconv ? 2 * d->in_h : Ho, out_w = upconv ? 2 * d->in_w : Wo;
  const long long m_total = (long long)d->batch * Ho * Wo;

  if (!(d->flags & SN_TC_IM2COL)) return conv_moments_halo_dispatch(d, as_stream(st), nullptr);
  SN_REQUIRE(d->rsum_out == nullptr, SN_ERR_UNSUPPORTED, "conv_tc: rsum_out needs the halo kernel (no SN_TC_IM2COL)");

  const int nt = d->cout % 128 == 0 ? 128 : (d->cout % 64 == 0 ? 64 : 32);
  const int groups = upconv ? 4 : 1;
  const int taps_w = upconv ? 4 : d->ksize * d->ksize;
  const int cin = d->src_c[0] + d->src_c[1];

  TcMaps maps;
  for (int s = 0; s < 2; ++s) {
    const int srcs = d->src_c[s] ? s : 0;   // unused second source: alias the first (never dereferenced)
    for (int pl = 0; pl < 3; ++pl)
      if ((rc = make_act_map(&maps.a[s][pl], d->src[srcs], pl, d->src_c[srcs], d->batch, d->in_h, d->in_w, keff)))
        return rc;
  }
  if ((rc = make_weight_map(&maps.w, d->w_packed, taps_w, d->cout, cin, nt))) return rc;

  TcP p{};
  p.m_total = (int)m_total; p.Ho = Ho; p.Wo = Wo;
  p.ksize = keff; p.taps_w = taps_w;
  p.cblk0 = d->src_c[0] / TC_KC; p.cblk1 = d->src_c[1] / TC_KC;
  p.cout = d->cout;
  p.relu = (d->flags & SN_TC_RELU) ? 1 : 0; p.upconv = upconv ? 1 : 0; p.dst_f32 = dst_f32 ? 1 : 0;
  p.dst = reinterpret_cast<__nv_bfloat16*>(d->dst.base);
  p.dh = d->dst.h; p.dw = d->dst.w; p.dc = d->dst.c; p.dy0 = d->dst.y0; p.dx0 = d->dst.x0; p.dc0 = d->dst.c0;
  p.dst_mu = d->dst_mu; p.dst_var = d->dst_var; p.out_h = out_h; p.out_w = out_w;
  p.s = d->s;
  const int n_tiles = groups * d->cout / nt;
  cudaStream_t stream = as_stream(st);
  switch (nt) {
    case 128: return launch_tc<128>(maps, p, n_tiles, stream);
    case 64: return launch_tc<64>(maps, p, n_tiles, stream);
    default: return launch_tc<32>(maps, p, n_tiles, stream);
  }
}

namespace sn {
int conv_moments_halo_dgrad_dispatch(const sn_tc_dgrad_desc* d, cudaStream_t stream);   // sn_tc_halo.cu
}

// Data gradient of the fused moment convolution (the chain tf.GradientTape builds at Brats.py:578,593).
extern "C" int sn_conv_moments_bwd_data_tc(const sn_tc_dgrad_desc* d, sn_stream_t st) {
  SN_REQUIRE(d, SN_ERR_BAD_ARG, "dgrad_tc: null descriptor");
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  SN_REQUIRE((d->flags & ~(SN_TC_UPCONV | SN_TC_KWC | SN_TC_NO_KWC | SN_TC_CTA2 | SN_TC_NO_CTA2)) == 0, SN_ERR_BAD_ARG,
             "dgrad_tc: only SN_TC_UPCONV / SN_TC_KWC / SN_TC_NO_KWC / SN_TC_CTA2 / SN_TC_NO_CTA2 are valid flags");
  SN_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0 && d->cout > 0, SN_ERR_BAD_ARG, "dgrad_tc: bad geometry");
  SN_REQUIRE(d->ksize >= 1 && d->ksize <= 3, SN_ERR_UNSUPPORTED, "dgrad_tc: kernel size %d", d->ksize);
  SN_REQUIRE(!upconv || d->ksize == 2, SN_ERR_BAD_ARG, "dgrad_tc: SN_TC_UPCONV needs ksize == 2");
  SN_REQUIRE(d->in_c[0] > 0 && d->in_c[0] % TC_KC == 0 && d->in_c[1] >= 0 && d->in_c[1] % TC_KC == 0,
             SN_ERR_UNSUPPORTED, "dgrad_tc: source channels (%d, %d) must be multiples of %d", d->in_c[0], d->in_c[1],
             TC_KC);
  SN_REQUIRE(d->cout % TC_KC == 0, SN_ERR_UNSUPPORTED, "dgrad_tc: cout %d must be a multiple of %d", d->cout, TC_KC);
  SN_REQUIRE(d->wt_packed && d->s && aligned16(d->wt_packed), SN_ERR_BAD_ARG, "dgrad_tc: weights missing/misaligned");
  const int keff = upconv ? 1 : d->ksize;
  SN_REQUIRE(d->in_h >= keff && d->in_w >= keff, SN_ERR_BAD_ARG, "dgrad_tc: input smaller than the kernel");
  const int Ho = d->in_h - keff + 1, Wo = d->in_w - keff + 1;
  const int out_h = upconv ? 2 * d->in_h : Ho, out_w = upconv ? 2 * d->in_w : Wo;
  int rc;
  if ((rc = check_view(d->g_out, d->batch, out_h, out_w, d->cout, "dgrad_tc g_out"))) return rc;
  for (int s = 0; s < 2; ++s) {
    if (d->in_c[s] == 0) continue;
    if ((rc = check_view(d->in[s], d->batch, d->in_h, d->in_w, d->in_c[s], "dgrad_tc in"))) return rc;
    if ((rc = check_view(d->g_in[s], d->batch, d->in_h, d->in_w, d->in_c[s], "dgrad_tc g_in"))) return rc;
  }
  return conv_moments_halo_dgrad_dispatch(d, as_stream(st));
}
