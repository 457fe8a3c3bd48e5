// FAST mode weight gradient of the moment convolution on the tensor cores (what tf.GradientTape
// derives for train_on_batch, Brats.py:569-580):
//
//   dL/dW[kh,kw,ci,n] = sum_p mu[p + (kh,kw), ci] g_mu'[p, n]  +  2 W[kh,kw,ci,n] sum_p var[p + (kh,kw), ci] g_var'[p, n]
//
// Both sums are GEMMs whose K axis is the PIXEL axis: D[(tap,ci), n] = A^T G with A = activations [pixels x ci]
// and G = output gradient [pixels x n], both stored pixel-major (NHWC packed planes).  That is the "MN-major"
// operand layout of tcgen05.mma: a TMA box {32 channels x 64 pixels} lands as 64-byte swizzled rows (row = pixel),
// which is exactly the canonical SWIZZLE_64B MN-major tile ((4,n),(8,k)):((1,LBO),(4,SBO)) in 16-byte units --
// 32 channels contiguous, 8-pixel groups SBO = 512 B apart, 32-channel blocks LBO = 4 KB apart -- so the same
// NHWC bytes feed the forward (K = channels, K-major descriptors) and the weight gradient (K = pixels, MN-major
// descriptors) without a transpose pass.
//
//   M tile = 4 blocks of 32 rows in the (tap, ci-block) space (128 UMMA rows); TMA im2col loads shift each block by
//            its tap, handle row/image wrap and zero-fill past the last pixel;
//   N tile = 1/2/4 blocks of 32 output channels;   K tile = 64 pixels per pipeline stage (3 stages);
//   split-K over persistent CTAs; partial sums leave through fp32 atomics into P_mu / P_var [taps][cin][cout]
//   (the HWIO layout of w_mu), and sn_wgrad_finalize forms P_mu + 2 W P_var.
// The up-conv (unpool + 2x2 conv, Brats.py:178-203,414-415) is four parity GEMMs sharing A: its N axis enumerates
// (parity, channel block) over four stride-2 views of the output gradient.
// Operands are single bf16 (mean_hi / g_mean_hi and the variance planes): the K axis is millions of pixels, so the
// independent operand roundings average out (measured in tests/test_gpu_tc_bwd.py against the fp64 oracle).
#include "sn_common.cuh"
#include "sn_sm100.cuh"

#include <cuda.h>
#include <mutex>

namespace sn {

constexpr int WG_KT = 64;                  // pixels per K tile
constexpr int WG_BLK = WG_KT * 64;         // bytes of one 32-channel block of one plane (4 KB)
constexpr int WG_A_PLANE = 4 * WG_BLK;     // 4 M blocks
constexpr int WG_STAGE = 2 * WG_A_PLANE + 2 * 4 * WG_BLK;   // A (mean, var) + G (mean, var; up to 4 N blocks) = 64 KB
constexpr int WG_STAGES = 3;
constexpr int WG_THREADS = 192;            // warp 0 TMA, warp 1 UMMA, warps 2-5 epilogue
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 256;

struct WgMaps {
  CUtensorMap a[2][2];     // [forward source][plane: mean_hi, variance] im2col maps (k x k taps)
  CUtensorMap g[4][2];     // [parity view][plane: g_mean_hi, g_variance] im2col maps (k = 1)
};

struct WgP {
  int m_blocks, n_blocks;          // valid 32-row blocks of the (tap, ci) axis / 32-column blocks of the N axis
  int nb_n;                        // N blocks per tile
  int m_tiles, n_tiles, splits, total_jobs;
  int k_tiles, tiles_per_split;
  int m_total, Ho, Wo;             // traversal: pixels of the gradient grid (up-conv: the input grid)
  int ksize, cblk0, cblk1, cin, cout, upconv;
  float* p_mu;
  float* p_var;
};

// Instruction descriptor for kind::f16, fp32 accumulate, bf16 operands, BOTH operands MN-major (bits 15 and 16).
__host__ __device__ constexpr uint32_t wg_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_64B operand: LBO = distance between 32-channel blocks, SBO = distance between 8-pixel groups.
__device__ __forceinline__ uint64_t wg_desc(uint32_t smem_addr) {
  constexpr uint64_t lbo = WG_BLK >> 4, sbo = 512 >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_moments_wgrad_kernel(const __grid_constant__ WgMaps maps,
                                                                           const WgP p) {
  constexpr int TMEM_COLS = 2 * NT < 32 ? 32 : 2 * NT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + WG_STAGES * WG_STAGE;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * WG_STAGES);
  const uint32_t acc_empty = bar_base + 8u * (2 * WG_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WG_STAGES + 2);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + WG_STAGES * WG_STAGE + 8 * (2 * WG_STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = p.cblk0 + p.cblk1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s)
      for (int pl = 0; pl < 2; ++pl) ptx::prefetch_tensormap(&maps.a[s][pl]);
    for (int s = 0; s < (p.upconv ? 4 : 1); ++s)
      for (int pl = 0; pl < 2; ++pl) ptx::prefetch_tensormap(&maps.g[s][pl]);
    for (int s = 0; s < WG_STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
        const int mt = job % p.m_tiles;
        const int nt = (job / p.m_tiles) % p.n_tiles;
        const int sp = job / (p.m_tiles * p.n_tiles);
        int mvalid = p.m_blocks - mt * 4;
        if (mvalid > 4) mvalid = 4;
        const uint32_t bytes = (uint32_t)((mvalid + p.nb_n) * 2 * WG_BLK);
        const int kt0 = sp * p.tiles_per_split;
        int kt1 = kt0 + p.tiles_per_split;
        if (kt1 > p.k_tiles) kt1 = p.k_tiles;
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
          const int stage = it % WG_STAGES;
          const uint32_t parity = (uint32_t)(it / WG_STAGES) & 1u;
          ptx::mbar_wait(empty_bar(stage), parity ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), bytes);
          const int m0 = kt * WG_KT;
          const int x_start = m0 % p.Wo;
          const int t = m0 / p.Wo;
          const int y_start = t % p.Ho;
          const int b_start = t / p.Ho;
          const uint32_t sa = smem_base + stage * WG_STAGE;
          const uint32_t sg = sa + 2 * WG_A_PLANE;
          for (int j = 0; j < mvalid; ++j) {
            const int mb = mt * 4 + j;
            const int tap = mb / cblk, cbt = mb - tap * cblk;
            const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
            const int src = cbt >= p.cblk0 ? 1 : 0;
            const int cb = src ? cbt - p.cblk0 : cbt;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              ptx::tma_load_im2col_4d(sa + pl * WG_A_PLANE + j * WG_BLK, &maps.a[src][pl], full_bar(stage),
                                      cb * 32, x_start, y_start, b_start, (uint16_t)kw, (uint16_t)kh);
          }
          const int cob = p.cout / 32;
          for (int j = 0; j < p.nb_n; ++j) {
            const int nb = nt * p.nb_n + j;
            const int view = p.upconv ? nb / cob : 0;
            const int ch = (p.upconv ? nb - view * cob : nb) * 32;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              ptx::tma_load_im2col_4d(sg + pl * 4 * WG_BLK + j * WG_BLK, &maps.g[view][pl], full_bar(stage), ch,
                                      x_start, y_start, b_start, 0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = wg_idesc(128, NT);
      const uint32_t acc_mu = tmem_base, acc_var = tmem_base + NT;
      int it = 0, jobi = 0;
      for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++jobi) {
        const int sp = job / (p.m_tiles * p.n_tiles);
        const int kt0 = sp * p.tiles_per_split;
        int kt1 = kt0 + p.tiles_per_split;
        if (kt1 > p.k_tiles) kt1 = p.k_tiles;
        ptx::mbar_wait(acc_empty, ((uint32_t)jobi & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
          const int stage = it % WG_STAGES;
          const uint32_t parity = (uint32_t)(it / WG_STAGES) & 1u;
          ptx::mbar_wait(full_bar(stage), parity);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * WG_STAGE;
          const uint32_t sg = sa + 2 * WG_A_PLANE;
#pragma unroll
          for (int ks = 0; ks < WG_KT / 16; ++ks) {
            const uint32_t koff = ks * 16 * 64;      // 16 pixels = 16 rows of 64 B
            const uint32_t acc = (kt > kt0 || ks > 0) ? 1u : 0u;
            ptx::umma_bf16(acc_mu, wg_desc(sa + koff), wg_desc(sg + koff), idesc, acc);
            ptx::umma_bf16(acc_var, wg_desc(sa + WG_A_PLANE + koff), wg_desc(sg + 4 * WG_BLK + koff), idesc, acc);
          }
          ptx::umma_commit(empty_bar(stage));
        }
        ptx::umma_commit(acc_full);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 atomics into the HWIO-shaped partial sums =====================
    const int q = warp & 3;                   // TMEM lane quarter == M block of the tile
    int jobi = 0;
    for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++jobi) {
      const int mt = job % p.m_tiles;
      const int nt = (job / p.m_tiles) % p.n_tiles;
      const int sp = job / (p.m_tiles * p.n_tiles);
      const bool has_k = sp * p.tiles_per_split < p.k_tiles;
      ptx::mbar_wait(acc_full, (uint32_t)jobi & 1u);
      ptx::tc_fence_after();
      const int mb = mt * 4 + q;
      if (mb < p.m_blocks && has_k) {
        const int tap = mb / cblk, cbt = mb - tap * cblk;
        const int ci = cbt * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int cob = p.cout / 32;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t am[16], av[16];
          ptx::tmem_ld16(lane_base + c0, am);
          ptx::tmem_ld16(lane_base + NT + c0, av);
          ptx::tmem_ld_wait();
          const int nb = nt * p.nb_n + (c0 >> 5);
          int wtap = tap, n = nb * 32 + (c0 & 31);
          if (p.upconv) {
            const int par = nb / cob;
            n = (nb - par * cob) * 32 + (c0 & 31);
            wtap = (1 - (par >> 1)) * 2 + (1 - (par & 1));      // parity (a,b) <- W[1-a, 1-b]
          }
          const size_t o = ((size_t)wtap * p.cin + ci) * p.cout + n;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            atomicAdd(p.p_mu + o + j, __uint_as_float(am[j]));
            atomicAdd(p.p_var + o + j, __uint_as_float(av[j]));
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// Row-halo form for k = 3 (the forward's tap-as-row-shift trick on the pixel axis).
//
// TMA im2col re-fetches every activation block once per tap and, with 64-byte rows, tops out near 16 B/clk per SM
// (measured: the kernel above runs at 4-15 % tensor pipe on every layer).  Here ONE tiled TMA box brings a band of
// full-width input rows {32 ch, R = in_w, THb rows, TN images} (row = pixel, 64-byte swizzled) and all nine taps
// read it in place: in an MN-major operand the pixel is the K index, so tap (kh, kw) is the same tile with the
// descriptor start advanced by (kh*R + kw) rows.  The three kw shifts of one kh even share one UMMA: M = 128 =
// 4 blocks of 32 channels with LBO = 64 B, i.e. block i is the tile shifted by i pixels (block 3 is junk and never
// stored).  The gradient tile uses the same row pitch: its box is {32 ch, R, rows, TN} over a tensor map whose
// extents are the OUTPUT grid (Wo x Ho), so the k-1 junk columns (and, for whole-image tiles, junk rows) arrive as
// TMA out-of-bounds zeros and contribute nothing.  Rows past the box (K padding to a multiple of 16) are zeroed once.
//   job = (32-channel input block, N tile of 32/64 output channels, split of the row bands); TMEM holds the
//   3 (kh) x 2 (mean, var) accumulators of 128 x NT.
constexpr int WR_STAGES = 2;
constexpr int WR_STAGE_MAX = 110 * 1024;
constexpr int WR_SMEM = WR_STAGES * WR_STAGE_MAX + 1024 + 256;

struct WrMaps {
  CUtensorMap a[2][2];     // [forward source][plane: mean_hi, variance] tiled boxes {32, R, THb, TN}
  CUtensorMap g[2];        // [plane: g_mean_hi, g_variance] tiled boxes {32, R, THg, TN}
};

struct WrP {
  int cblk0, cblk1, nb_n, n_tiles, splits, total_jobs;
  int R, THo, THb, THg, TN, tiles_y, tiles_b, tiles_total, tiles_per_split;
  int rows_a, rows_g, ksteps;      // box rows of A / G, K steps of 16 pixels per tile
  int a_plane, g_blk, stage;       // bytes: one A plane, one 32-channel G block, one pipeline stage
  int cin, cout;
  float* p_mu;
  float* p_var;
};

__device__ __forceinline__ uint64_t wr_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_moments_wgrad_rows_kernel(const __grid_constant__ WrMaps maps,
                                                                                const WrP p) {
  constexpr int TMEM_COLS = NT == 32 ? 256 : 512;      // 6 * NT accumulator columns, rounded up to a power of two
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + WR_STAGES * p.stage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WR_STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * WR_STAGES);
  const uint32_t acc_empty = bar_base + 8u * (2 * WR_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WR_STAGES + 2);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + WR_STAGES * p.stage + 8 * (2 * WR_STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = p.cblk0 + p.cblk1;

  // rows the TMA boxes never write are read by the UMMAs as K padding / shifted tails: zero everything once
  for (int i = threadIdx.x; i < WR_STAGES * p.stage / 16; i += WG_THREADS)
    reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s)
      for (int pl = 0; pl < 2; ++pl) ptx::prefetch_tensormap(&maps.a[s][pl]);
    for (int pl = 0; pl < 2; ++pl) ptx::prefetch_tensormap(&maps.g[pl]);
    for (int s = 0; s < WR_STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int jobs_cn = cblk * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)((2 * p.rows_a + 2 * p.nb_n * p.rows_g) * 64);
      int it = 0;
      for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
        const int cbt = job % cblk;
        const int nt = (job / cblk) % p.n_tiles;
        const int sp = job / jobs_cn;
        const int src = cbt >= p.cblk0 ? 1 : 0;
        const int cb = src ? cbt - p.cblk0 : cbt;
        const int t0 = sp * p.tiles_per_split;
        int t1 = t0 + p.tiles_per_split;
        if (t1 > p.tiles_total) t1 = p.tiles_total;
        for (int t = t0; t < t1; ++t, ++it) {
          const int stage = it % WR_STAGES;
          const uint32_t parity = (uint32_t)(it / WR_STAGES) & 1u;
          ptx::mbar_wait(empty_bar(stage), parity ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), bytes);
          const int ty = t % p.tiles_y, tb = t / p.tiles_y;
          const int y0 = ty * p.THo, b0 = tb * p.TN;
          const uint32_t sa = smem_base + stage * p.stage;
          const uint32_t sg = sa + 2 * p.a_plane;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5, %6}], [%2];"
                :
                : "r"(sa + pl * p.a_plane), "l"(reinterpret_cast<uint64_t>(&maps.a[src][pl])), "r"(full_bar(stage)),
                  "r"(cb * 32), "r"(0), "r"(y0), "r"(b0)
                : "memory");
          for (int j = 0; j < p.nb_n; ++j)
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                  " [%0], [%1, {%3, %4, %5, %6}], [%2];"
                  :
                  : "r"(sg + (pl * p.nb_n + j) * p.g_blk), "l"(reinterpret_cast<uint64_t>(&maps.g[pl])),
                    "r"(full_bar(stage)), "r"((nt * p.nb_n + j) * 32), "r"(0), "r"(y0), "r"(b0)
                  : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = wg_idesc(128, NT);
      int it = 0, jobi = 0;
      for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++jobi) {
        const int sp = job / jobs_cn;
        const int t0 = sp * p.tiles_per_split;
        int t1 = t0 + p.tiles_per_split;
        if (t1 > p.tiles_total) t1 = p.tiles_total;
        ptx::mbar_wait(acc_empty, ((uint32_t)jobi & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int t = t0; t < t1; ++t, ++it) {
          const int stage = it % WR_STAGES;
          const uint32_t parity = (uint32_t)(it / WR_STAGES) & 1u;
          ptx::mbar_wait(full_bar(stage), parity);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage;
          const uint32_t sg = sa + 2 * p.a_plane;
          const uint32_t gv_off = (uint32_t)(p.nb_n * p.g_blk);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t acc_mu = tmem_base + (uint32_t)(kh * 2 * NT), acc_var = acc_mu + NT;
            const uint32_t a_kh = sa + (uint32_t)(kh * p.R) * 64u;
#pragma unroll 1
            for (int ks = 0; ks < p.ksteps; ++ks) {
              const uint32_t koff = (uint32_t)ks * 1024u;       // 16 pixels = 16 rows of 64 B
              const uint32_t acc = (t > t0 || ks > 0) ? 1u : 0u;
              ptx::umma_bf16(acc_mu, wr_desc(a_kh + koff, 64u), wr_desc(sg + koff, (uint32_t)p.g_blk), idesc, acc);
              ptx::umma_bf16(acc_var, wr_desc(a_kh + p.a_plane + koff, 64u),
                             wr_desc(sg + gv_off + koff, (uint32_t)p.g_blk), idesc, acc);
            }
          }
          ptx::umma_commit(empty_bar(stage));
        }
        ptx::umma_commit(acc_full);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                   // TMEM lane quarter == kw shift block (3 = junk)
    int jobi = 0;
    for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++jobi) {
      const int cbt = job % cblk;
      const int nt = (job / cblk) % p.n_tiles;
      ptx::mbar_wait(acc_full, (uint32_t)jobi & 1u);
      ptx::tc_fence_after();
      if (q < 3) {
        const int ci = cbt * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll 1
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t am[16], av[16];
            ptx::tmem_ld16(lane_base + kh * 2 * NT + c0, am);
            ptx::tmem_ld16(lane_base + kh * 2 * NT + NT + c0, av);
            ptx::tmem_ld_wait();
            const size_t o = ((size_t)(kh * 3 + q) * p.cin + ci) * p.cout + nt * NT + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              atomicAdd(p.p_mu + o + j, __uint_as_float(am[j]));
              atomicAdd(p.p_var + o + j, __uint_as_float(av[j]));
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// finalize: g_w = P_mu + 2 W P_var ;  g_w_sigma[n] = sigmoid(w_sigma[n]) * dsig[n]
// ---------------------------------------------------------------------------------------------------------
__global__ void wgrad_finalize_kernel(size_t n_w, int cout, const float* __restrict__ w, const float* __restrict__ ws,
                                      const float* __restrict__ p_mu, const float* __restrict__ p_var,
                                      const float* __restrict__ dsig, float* __restrict__ g_w,
                                      float* __restrict__ g_ws) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n_w; i += stride) g_w[i] = fmaf(2.f * w[i], p_var[i], p_mu[i]);
  for (size_t i = t0; i < (size_t)cout; i += stride) g_ws[i] = sigmoid_f(ws[i]) * dsig[i];
}

// dsig[n] += sum_p g_var'[p, n] r[p]  (dL/ds_n); thread = (pixel lane, 8 channels)
__global__ void __launch_bounds__(256) wgrad_dsigma_kernel(sn_packed_view g, int B, int H, int W, int cout, int upconv,
                                                           const float* __restrict__ r, float* __restrict__ dsig) {
  __shared__ float red[256][9];
  const int groups = cout / 8;
  const int lanes = 256 / groups;                 // pixel lanes per block (cout <= 512 -> groups <= 64)
  const int gidx = threadIdx.x % groups, pl = threadIdx.x / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const size_t total = (size_t)B * H * W;
  const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(g.base);
  const int rh = upconv ? H / 2 : H, rw = upconv ? W / 2 : W;
  if (pl < lanes) {
    for (size_t i = (size_t)blockIdx.x * lanes + pl; i < total; i += (size_t)gridDim.x * lanes) {
      const int x = (int)(i % W);
      size_t t = i / W;
      const int y = (int)(t % H);
      const int b = (int)(t / H);
      const float rv = upconv ? r[((size_t)b * rh + (y >> 1)) * rw + (x >> 1)] : r[i];
      const uint4 v = *reinterpret_cast<const uint4*>(
          gp + ((((size_t)b * g.h + y + g.y0) * g.w + x + g.x0) * 3 + 2) * g.c + g.c0 + gidx * 8);
      const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(__uint_as_float(vw[e] << 16), rv, acc[2 * e]);
        acc[2 * e + 1] = fmaf(__uint_as_float(vw[e] & 0xFFFF0000u), rv, acc[2 * e + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int idx = threadIdx.x; idx < cout; idx += 256) {
    const int gi = idx / 8, j = idx % 8;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * groups + gi][j];
    atomicAdd(dsig + idx, s);
  }
}

// r[b,y,x] = sum over the k x k window and channels of x^2: the rank-1 statistic of myConv_input (Brats.py:69-73)
__global__ void first_conv_rsum_kernel(int B, int H, int W, int cin, int k, const float* __restrict__ x,
                                       float* __restrict__ r) {
  const int Ho = H - k + 1, Wo = W - k + 1;
  const size_t total = (size_t)B * Ho * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo);
    size_t t = i / Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float s = 0.f;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const float* px = x + (((size_t)b * H + yo + kh) * W + xo + kw) * cin;
        for (int c = 0; c < cin; ++c) s = fmaf(px[c], px[c], s);
      }
    r[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*WgEncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct WgDriver {
  WgEncodeIm2colFn im2col = nullptr;
  int driver_version = 0;
};
static const WgDriver& wg_driver() {
  static WgDriver d;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      d.im2col = reinterpret_cast<WgEncodeIm2colFn>(f);
    cudaDriverGetVersion(&d.driver_version);
  });
  return d;
}

// im2col map over one plane of a packed window: 64 pixels x 32 channels per load.  `step` = 2 with origin (oy, ox)
// addresses the parity view (2y+oy, 2x+ox) of the window.
static int wg_make_map(CUtensorMap* out, const sn_packed_view& v, int plane, int src_c, int batch, int in_h, int in_w,
                       int ksize, int step = 1, int oy = 0, int ox = 0) {
  const WgDriver& api = wg_driver();
  const size_t pix = (size_t)3 * v.c;
  char* base = reinterpret_cast<char*>(v.base) +
               ((((size_t)(v.y0 + oy) * v.w + v.x0 + ox) * 3 + plane) * v.c + v.c0) * sizeof(__nv_bfloat16);
  cuuint64_t dims[4] = {(cuuint64_t)src_c, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)batch};
  cuuint64_t strides[3] = {pix * 2 * step, (cuuint64_t)v.w * pix * 2 * step, (cuuint64_t)v.h * v.w * pix * 2};
  int lower[2] = {0, 0};
  int upper[2] = {-(ksize - 1), -(ksize - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = api.im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, lower, upper, 32, WG_KT, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeIm2col(wgrad) failed (%d)", (int)r);
  // same driver quirk as sn_tc_conv.cu::make_act_map (im2col maps over tensors smaller than 128 KiB)
  const size_t span = ((size_t)(batch - 1) * strides[2] + (size_t)(in_h - 1) * strides[1] +
                       (size_t)(in_w - 1) * strides[0] + (size_t)src_c * 2);
  if (api.driver_version <= 13010 && span < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return SN_OK;
}

static int wg_check_view(const sn_packed_view& v, int batch, int h, int w, int c, const char* who) {
  SN_REQUIRE(v.base && aligned16(v.base), SN_ERR_MISALIGNED, "%s: packed buffer must be 16-byte aligned", who);
  SN_REQUIRE(v.n >= batch && v.h > 0 && v.w > 0 && v.c > 0 && v.c % 8 == 0 && v.c0 % 8 == 0, SN_ERR_BAD_ARG,
             "%s: bad packed buffer dims [%d,%d,%d,%d]", who, v.n, v.h, v.w, v.c);
  SN_REQUIRE(v.y0 >= 0 && v.x0 >= 0 && v.c0 >= 0 && v.y0 + h <= v.h && v.x0 + w <= v.w && v.c0 + c <= v.c,
             SN_ERR_BAD_ARG, "%s: window outside the buffer", who);
  return SN_OK;
}

template <int NT>
static int wg_launch(const WgMaps& maps, const WgP& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_moments_wgrad_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(SN_ERR_LAUNCH, "wgrad: cannot reserve %d B of shared memory", WG_SMEM);
  const int grid = p.total_jobs < num_sms() ? p.total_jobs : num_sms();
  conv_moments_wgrad_kernel<NT><<<grid, WG_THREADS, WG_SMEM, st>>>(maps, p);
  return check_launch("conv_moments_wgrad");
}

typedef CUresult (*WrEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WrEncodeTiledFn wr_encode_tiled() {
  static WrEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WrEncodeTiledFn>(f);
  });
  return fn;
}

// Tiled map over one plane of a packed window with extents (w x h): a box may overhang them (zero fill).
static int wr_make_map(CUtensorMap* out, const sn_packed_view& v, int plane, int c, int batch, int h, int w, int box_w,
                       int box_h, int box_n) {
  const size_t pix = (size_t)3 * v.c;
  char* base = reinterpret_cast<char*>(v.base) +
               ((((size_t)v.y0 * v.w + v.x0) * 3 + plane) * v.c + v.c0) * sizeof(__nv_bfloat16);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  cuuint64_t strides[3] = {pix * 2, (cuuint64_t)v.w * pix * 2, (cuuint64_t)v.h * v.w * pix * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = wr_encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeTiled(wgrad rows) failed (%d)", (int)r);
  return SN_OK;
}

// Tile plan of the row-halo kernel; returns false when the layer does not fit (the im2col kernel takes it).
static bool wr_plan(WrP& p, int batch, int in_h, int in_w, int cin0, int cin1, int cout) {
  if (in_w > 256 || wr_encode_tiled() == nullptr) return false;
  const int R = in_w, Ho = in_h - 2;
  p.R = R;
  p.cblk0 = cin0 / 32; p.cblk1 = cin1 / 32;
  p.cin = cin0 + cin1; p.cout = cout;
  p.nb_n = (cout / 32) % 2 == 0 ? 2 : 1;
  p.n_tiles = cout / 32 / p.nb_n;
  auto fits = [&](int rows_a, int rows_g) {
    const int ks = (rows_g + 15) / 16;
    const int need_a = 16 * ks + 2 * R + 3 > rows_a ? 16 * ks + 2 * R + 3 : rows_a;
    p.ksteps = ks;
    p.a_plane = ((need_a * 64 + 1023) / 1024) * 1024;
    p.g_blk = ((16 * ks * 64 + 1023) / 1024) * 1024;
    p.stage = 2 * p.a_plane + 2 * p.nb_n * p.g_blk;
    // UMMA descriptor fields: LBO (g_blk) and every start address must stay below the 14-bit x 16 B range
    return p.stage <= WR_STAGE_MAX && p.g_blk < (1 << 18);
  };
  // whole (small) images, several per tile: gradient box = input box, junk rows/columns are out-of-bounds zeros
  int tn = 0;
  for (int t = 1; t <= batch && t <= 256 && in_h <= 256; ++t) {
    if (!fits(t * in_h * R, t * in_h * R)) break;
    tn = t;
  }
  if (tn >= 1) {
    fits(tn * in_h * R, tn * in_h * R);
    p.TN = tn; p.THo = Ho; p.THb = in_h; p.THg = in_h;
    p.tiles_y = 1; p.tiles_b = (batch + tn - 1) / tn;
  } else {
    int th = 0;
    for (int t = 1; t <= Ho && t + 2 <= 256; ++t) {
      if (!fits((t + 2) * R, t * R)) break;
      th = t;
    }
    if (th < 1) return false;
    fits((th + 2) * R, th * R);
    p.TN = 1; p.THo = th; p.THb = th + 2; p.THg = th;
    p.tiles_y = (Ho + th - 1) / th; p.tiles_b = batch;
  }
  p.rows_a = p.TN * p.THb * R;
  p.rows_g = p.TN * p.THg * R;
  p.tiles_total = p.tiles_y * p.tiles_b;
  const int jobs_cn = (p.cblk0 + p.cblk1) * p.n_tiles;
  int splits = (2 * num_sms() + jobs_cn - 1) / jobs_cn;
  if (splits > p.tiles_total) splits = p.tiles_total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles_total + splits - 1) / splits;
  p.splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  p.total_jobs = jobs_cn * p.splits;
  return true;
}

template <int NT>
static int wr_launch(const WrMaps& maps, const WrP& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_moments_wgrad_rows_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    WR_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(SN_ERR_LAUNCH, "wgrad rows: cannot reserve %d B of shared memory", WR_SMEM);
  const int grid = p.total_jobs < num_sms() ? p.total_jobs : num_sms();
  conv_moments_wgrad_rows_kernel<NT><<<grid, WG_THREADS, WR_STAGES * p.stage + 1024 + 256, st>>>(maps, p);
  return check_launch("conv_moments_wgrad_rows");
}

}  // namespace sn

using namespace sn;

extern "C" {

size_t sn_wgrad_workspace_bytes(int32_t ksize, int32_t cin, int32_t cout) {
  return ((size_t)2 * ksize * ksize * cin * cout + (size_t)cout) * sizeof(float);
}

int sn_conv_moments_bwd_weight_tc(const sn_tc_wgrad_desc* d, sn_stream_t st) {
  SN_REQUIRE(d, SN_ERR_BAD_ARG, "wgrad_tc: null descriptor");
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  SN_REQUIRE((d->flags & ~(SN_TC_UPCONV | SN_TC_IM2COL | SN_TC_ROWS)) == 0, SN_ERR_BAD_ARG,
             "wgrad_tc: valid flags are SN_TC_UPCONV, SN_TC_IM2COL and SN_TC_ROWS");
  SN_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0 && d->cout > 0, SN_ERR_BAD_ARG, "wgrad_tc: bad geometry");
  SN_REQUIRE(d->ksize >= 1 && d->ksize <= 3, SN_ERR_UNSUPPORTED, "wgrad_tc: kernel size %d", d->ksize);
  SN_REQUIRE(!upconv || d->ksize == 2, SN_ERR_BAD_ARG, "wgrad_tc: SN_TC_UPCONV needs ksize == 2");
  SN_REQUIRE(d->in_c[0] > 0 && d->in_c[0] % 32 == 0 && d->in_c[1] >= 0 && d->in_c[1] % 32 == 0, SN_ERR_UNSUPPORTED,
             "wgrad_tc: source channels (%d, %d) must be multiples of 32", d->in_c[0], d->in_c[1]);
  SN_REQUIRE(d->cout % 32 == 0 && d->cout <= 512, SN_ERR_UNSUPPORTED, "wgrad_tc: cout %d", d->cout);
  SN_REQUIRE(d->w_mu && d->w_sigma && d->rsum && d->workspace && d->g_w_mu && d->g_w_sigma, SN_ERR_BAD_ARG,
             "wgrad_tc: null pointer");
  SN_REQUIRE(wg_driver().im2col != nullptr, SN_ERR_DRIVER, "wgrad_tc: cuTensorMapEncodeIm2col unavailable");
  const int keff = upconv ? 1 : d->ksize;
  SN_REQUIRE(d->in_h >= keff && d->in_w >= keff, SN_ERR_BAD_ARG, "wgrad_tc: input smaller than the kernel");
  const int Ho = d->in_h - keff + 1, Wo = d->in_w - keff + 1;          // traversal grid (up-conv: the input grid)
  const int out_h = upconv ? 2 * d->in_h : Ho, out_w = upconv ? 2 * d->in_w : Wo;
  const int cin = d->in_c[0] + d->in_c[1];
  int rc;
  if ((rc = wg_check_view(d->g_out, d->batch, out_h, out_w, d->cout, "wgrad_tc g_out"))) return rc;
  for (int s = 0; s < 2; ++s)
    if (d->in_c[s] && (rc = wg_check_view(d->in[s], d->batch, d->in_h, d->in_w, d->in_c[s], "wgrad_tc in"))) return rc;
  const long long m_total = (long long)d->batch * Ho * Wo;
  SN_REQUIRE(m_total < (1ll << 31) - WG_KT, SN_ERR_UNSUPPORTED, "wgrad_tc: too many pixels");

  WgP p{};
  const int cblk = cin / 32;
  p.cblk0 = d->in_c[0] / 32; p.cblk1 = d->in_c[1] / 32;
  p.m_blocks = keff * keff * cblk;
  p.n_blocks = (upconv ? 4 : 1) * d->cout / 32;
  p.nb_n = p.n_blocks % 4 == 0 ? 4 : (p.n_blocks % 2 == 0 ? 2 : 1);
  p.m_tiles = (p.m_blocks + 3) / 4;
  p.n_tiles = p.n_blocks / p.nb_n;
  p.m_total = (int)m_total; p.Ho = Ho; p.Wo = Wo;
  p.k_tiles = (int)((m_total + WG_KT - 1) / WG_KT);
  const int out_tiles = p.m_tiles * p.n_tiles;
  int splits = (2 * num_sms() + out_tiles - 1) / out_tiles;
  if (splits > p.k_tiles) splits = p.k_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.total_jobs = out_tiles * p.splits;
  p.ksize = keff; p.cin = cin; p.cout = d->cout; p.upconv = upconv ? 1 : 0;
  const size_t n_w = (size_t)d->ksize * d->ksize * cin * d->cout;
  float* ws_f = reinterpret_cast<float*>(d->workspace);
  p.p_mu = ws_f;
  p.p_var = ws_f + n_w;
  float* dsig = ws_f + 2 * n_w;
  cudaStream_t stream = as_stream(st);
  cudaError_t e = cudaMemsetAsync(d->workspace, 0, (2 * n_w + d->cout) * sizeof(float), stream);
  if (e != cudaSuccess) return fail(SN_ERR_LAUNCH, "wgrad_tc memset: %s", cudaGetErrorString(e));

  const int planes[2] = {0, 2};
  WrP rp{};
  // deep layers (few pixels, >= 256 channels on both sides) are faster through the im2col kernel (measured, batch 64)
  const bool deep = cin >= 256 && d->cout >= 256;
  if (!upconv && d->ksize == 3 && !(d->flags & SN_TC_IM2COL) && !(deep && !(d->flags & SN_TC_ROWS)) && wr_plan(rp, d->batch, d->in_h, d->in_w, d->in_c[0], d->in_c[1], d->cout)) {
    rp.p_mu = p.p_mu; rp.p_var = p.p_var;
    WrMaps rmaps;
    for (int s = 0; s < 2; ++s) {
      const int ss = d->in_c[s] ? s : 0;
      for (int pl = 0; pl < 2; ++pl)
        if ((rc = wr_make_map(&rmaps.a[s][pl], d->in[ss], planes[pl], d->in_c[ss], d->batch, d->in_h, d->in_w, rp.R,
                              rp.THb, rp.TN)))
          return rc;
    }
    for (int pl = 0; pl < 2; ++pl)
      if ((rc = wr_make_map(&rmaps.g[pl], d->g_out, planes[pl], d->cout, d->batch, Ho, Wo, rp.R, rp.THg, rp.TN)))
        return rc;
    rc = rp.nb_n == 2 ? wr_launch<64>(rmaps, rp, stream) : wr_launch<32>(rmaps, rp, stream);
    if (rc) return rc;
  } else {
  WgMaps maps;
  for (int s = 0; s < 2; ++s) {
    const int ss = d->in_c[s] ? s : 0;
    for (int pl = 0; pl < 2; ++pl)
      if ((rc = wg_make_map(&maps.a[s][pl], d->in[ss], planes[pl], d->in_c[ss], d->batch, d->in_h, d->in_w, keff)))
        return rc;
  }
  for (int v = 0; v < 4; ++v) {
    const int oy = upconv ? (v >> 1) : 0, ox = upconv ? (v & 1) : 0;
    for (int pl = 0; pl < 2; ++pl)
      if ((rc = wg_make_map(&maps.g[v][pl], d->g_out, planes[pl], d->cout, d->batch, Ho, Wo, 1, upconv ? 2 : 1, oy, ox)))
        return rc;
  }
  switch (p.nb_n) {
    case 4: rc = wg_launch<128>(maps, p, stream); break;
    case 2: rc = wg_launch<64>(maps, p, stream); break;
    default: rc = wg_launch<32>(maps, p, stream); break;
  }
  if (rc) return rc;
  }
  const size_t pixels = (size_t)d->batch * out_h * out_w;
  const int lanes = 256 / (d->cout / 8);
  size_t need = (pixels + lanes - 1) / lanes;
  const size_t cap = (size_t)num_sms() * 4;
  wgrad_dsigma_kernel<<<(int)(need < cap ? need : cap), 256, 0, stream>>>(d->g_out, d->batch, out_h, out_w, d->cout,
                                                                           upconv ? 1 : 0, d->rsum, dsig);
  wgrad_finalize_kernel<<<ew_grid(n_w, 256), 256, 0, stream>>>(n_w, d->cout, d->w_mu, d->w_sigma, p.p_mu, p.p_var, dsig,
                                                                d->g_w_mu, d->g_w_sigma);
  return check_launch("wgrad_finalize");
}

int sn_first_conv_rsum(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t ksize, const float* x,
                       float* rsum, sn_stream_t st) {
  SN_REQUIRE(x && rsum, SN_ERR_BAD_ARG, "first_conv_rsum: null pointer");
  SN_REQUIRE(batch > 0 && cin >= 1 && ksize >= 1 && in_h >= ksize && in_w >= ksize, SN_ERR_BAD_ARG,
             "first_conv_rsum: bad geometry");
  const size_t total = (size_t)batch * (in_h - ksize + 1) * (in_w - ksize + 1);
  first_conv_rsum_kernel<<<ew_grid(total, 256), 256, 0, as_stream(st)>>>(batch, in_h, in_w, cin, ksize, x, rsum);
  return check_launch("first_conv_rsum");
}

}  // extern "C"
