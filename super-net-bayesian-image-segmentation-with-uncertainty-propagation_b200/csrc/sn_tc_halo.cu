// FAST mode moment convolution, second generation: persistent, halo-tiled tcgen05 / TMEM kernel.
//
// The im2col kernel (sn_tc_conv.cu) pulls every input pixel k*k times from L2 -- measured on B200 the
// 32/64-channel BraTS layers sit at the L2->SM feed limit (lts 57 %, tensor pipe 13 %, DRAM 15 %).  Here a
// CTA loads one HALO box of the input per 32-channel block (TMA tiled 4-D box {32 ch, R cols, TH+k-1 rows, TN
// images}; smem row = pixel, 64-byte swizzled rows) and all k*k filter taps read it IN PLACE: the A operand of
// tap (kh,kw) is the same tile with its UMMA descriptor start address advanced by (kh*R + kw) rows.  (Probed
// on B200: the swizzle XOR is applied to absolute smem address bits, so any row shift is legal with
// base_offset = 0; tools/probe_shift.cu.)  GEMM row j of the 128-row tile is halo pixel j = (n*THb + y)*R + x;
// rows with x >= R-(k-1) or y >= THb-(k-1) are junk and never stored.  Consequences:
//   * L2->smem bytes per output pixel drop ~k*k-fold for A; for small layers the prepared weights stay
//     RESIDENT in smem for the CTA's whole life (persistent tile loop), so only activations stream;
//   * the rank-1 term needs q[pixel] = sum_c(mu^2+var) once per halo pixel (not once per tap):
//     r[j] = sum_taps q[j + kh*R + kw] from a 1 KB smem array;
//   * two TMEM accumulator stages: the epilogue of tile i overlaps the TMA/UMMA main loop of tile i+1.
// Warp roles (576 threads): warp 0 TMA producer, warp 1 UMMA issuer (+TMEM alloc), warps 2-9 q reduction over
// every A stage (thread = halo pixel), warps 10-17 epilogue: two warps per TMEM lane quarter, each taking half of
// the tile's columns
// (q arrives through a double-buffered smem array guarded by mbarriers).  Profiling the first version (4
// epilogue warps, one per SM sub-partition) showed the epilogue, a long dependent instruction chain per pixel
// row, as the critical path of the 32/64-channel layers; hence the second set of warps.
//
// What bounds it now (measured, round 1): shared-memory bandwidth.  A per-role clock64 trace of conv1 gives a tile
// period of ~3650 cycles for 54 UMMAs; every UMMA streams its 4 KB A operand (+1-4 KB of B) from smem, the TMA fill
// writes 40 KB and the q reduction re-reads 40 KB per tile: ~380 KB per tile = 104 B/clk of the SM's 128 B/clk.  More
// epilogue warps, 4 accumulator stages and two alternating issuer warps (all tried, all parity-green) left the
// period unchanged, which is what a bandwidth bound looks like; they were dropped again.  The levers that remain are
// the ones that remove smem bytes: q computed by the producing layer's epilogue, and cta_group::2 UMMAs that halve
// the B-operand traffic per SM for the NT = 128 layers.  Same math and data layout as sn_tc_conv.cu
// (Brats.py:118-137 incl. ReLU :233-238, pad :159-163, concat :247-261, unpool+2x2 conv :178-203,414-415).
#include "sn_common.cuh"
#include "sn_sm100.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <mutex>

namespace sn {

// SN_HL_DBG knob experiments (profiles/r02_kwc_knobs.md) are compiled in only with -DSN_HL_KNOBS: the shipped kernels
// carry none of the checks (they sit in the UMMA issue loop and the epilogue).
#ifdef SN_HL_KNOBS
#define HL_DBG(p) ((p).dbg)
#else
#define HL_DBG(p) 0
#endif

constexpr int HL_BM = 128;
constexpr int HL_KC = 32;
constexpr int HL_THREADS = 576;       // 18 warps: TMA, UMMA, 8 x q reduction, 8 x epilogue
constexpr int HL_THREADS_2SETS = 832; // 26 warps: a second set of 8 epilogue warps
// Two epilogue sets (on alternate tiles, set s draining TMEM stage s) where the epilogue is on the critical path: the
// kw-concatenated kernels.  Tried and measured neutral for the forward kernels with two pixel tiles per weight slot
// (G = 2; conv3 0.246 vs 0.240 ms): there the UMMA issuer does wait for the pair's epilogues (28 % of its samples on
// acc_empty, tools/ncu_waits.py), but the layer is bound by shared-memory bandwidth -- every N <= 128 UMMA streams
// both operands from shared memory at exactly 128 B/clk, and the TMA fills and the q reduction share that pipe
// (profiles/r02_kwc_knobs.md section 6) -- so a second set only moves the wait.  `g == 2 && !dgrad` re-enables it.
__host__ __device__ constexpr bool hl_two_sets(bool kwc, int g, bool dgrad) { return kwc; }
__host__ __device__ constexpr int hl_threads(bool two_sets) { return two_sets ? HL_THREADS_2SETS : HL_THREADS; }
constexpr int HL_MAX_BSLOTS = 36;
constexpr int HL_MAX_ASTAGES = 4;
constexpr int HL_CTA2_MAX_BSLOTS = 16;  // CTA pairs: the leader also keeps one "peer's half-slot landed" barrier per slot
constexpr int HL_SMEM = 232448;        // 227 KB: always requested so exactly one CTA owns an SM (and its TMEM)

struct HlMaps {
  CUtensorMap a[4][3];     // forward: up to two concatenated sources; data gradient of an up-conv: four parity views
  CUtensorMap w;
  CUtensorMap d;           // destination window (c, plane, x, y, n) for the TMA-store epilogue
};
constexpr int HL_HEAD_SMEM = 16384;   // fused head: W, W^2, s of conv_final + the partial-sum exchange of the warp pairs
constexpr int HL_STG_BUF = 6144;      // staging buffer of one (epilogue set, tile row): 30 pixels x 3 planes x 32 channels

struct HlView {            // a window of a packed buffer, device side
  __nv_bfloat16* base;
  int h, w, c, y0, x0, c0;
};

struct HlP {
  int tiles_x, tiles_y, tiles_b, tiles_n, total_tiles;
  int pix_tiles, total_groups;   // G = 2 (two pixel tiles per weight slot): pixel tiles per N tile, (N tile, tile pair) groups
  int TWo, THo;            // valid output columns / rows per tile
  int R, THb, TN;          // halo box: columns, rows, images
  int rows_box;            // R * THb * TN  (<= 254)
  int a_plane;             // bytes reserved per A plane per stage (multiple of 1024; 512 in one CTA-pair plan, see hl_plan)
  int ksize, taps_w;       // K-side taps per dim; taps in the prepared weights
  int cblk_s[4];           // 32-channel K blocks taken from each source
  int pad;                 // data gradient: the box origin is shifted by -(k-1); TMA zero-fills outside the window
  int Ho, Wo, B;           // valid output extents (upconv: the input grid)
  int cout, relu, upconv, dst_f32;
  int sa, sb, b_resident;  // A stages, B slots, weights resident?
  __nv_bfloat16* dst;
  int dh, dw, dc, dy0, dx0, dc0;
  float* dst_mu;
  float* dst_var;
  int out_h, out_w;
  const float* s;
  float* r_out;            // optional: box_k(sum_c mu^2 + var) per valid pixel, fp32 [B][Ho][Wo]
  int s_len;               // forward: cout; data gradient: channels of one gradient source
  // data gradient only: destination / saved-activation windows of the two forward sources
  HlView gdst[2], saved[2];
  int csplit, gate[2];
  int v8;                  // every packed row segment the epilogue touches is 32-byte aligned: use 256-bit accesses
  int kwc;                 // kw-concatenated variant (see the kernel): R = 32, one weight slot per filter row
  int cta2;                // CTA pair (cta_group::2) variant: M = 256 over two SMs, half of every weight slot per CTA
  int tma_store;           // forward KWC, packed destination: rows leave through shared memory + TMA tensor stores
  // fused head (template HEAD = n_labels > 0): conv_final + mysoftmax in the epilogue of the last 32-channel conv
  const float* head_w;     // conv_final w_mu, fp32 [32][n_labels]
  const float* head_ws;    // conv_final raw w_sigma [n_labels]
  float *head_p, *head_v;  // fp32 [B*Ho*Wo][n_labels] softmax probabilities / their variances
  float *head_pre_mu, *head_pre_var;   // optional pre-softmax moments
  int head_labels;         // host side only: which HEAD instantiation to launch (0: none)
  int dbg;                 // SN_HL_DBG knob experiments (profiling only; results are wrong when non-zero): 1 epilogue
                           // skips its work, 2 one UMMA group per tile, 4 reducers skip their loads, 8 no global stores
};

__device__ __forceinline__ float hl_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float hl_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t hl_pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// Mixed-precision arithmetic of sm_100 (PTX fma/add.rn.f32.bf16 = FHFMA.BF16 / FHADD.BF16 in SASS): the bf16 operands
// are read straight from a half of a 32-bit register, the accumulator is fp32 -- no shift / mask to unpack first.  The
// CUDA-core side of the 32/64-channel layers is issue-bound (profiles/r02_kwc_knobs.md), so instructions are the currency.
//   hl_fma2(a, b, c0, c1):  c0 += a.lo * b.lo,  c1 += a.hi * b.hi        (products of two bf16 are exact in fp32)
//   hl_add2(a, c0, c1):     c0 += a.lo,         c1 += a.hi
//   hl_resid2(hi, a0, a1):  bf16x2(a0 - hi.lo, a1 - hi.hi), the "lo" plane of a hi/lo split (same value as the
//                           unpack + FADD form: one rounding of an exactly representable difference)
//   hl_sum2(h, l, m0, m1):  m0 = h.lo + l.lo, m1 = h.hi + l.hi in fp32
__device__ __forceinline__ void hl_fma2(uint32_t a, uint32_t b, float& c0, float& c1) {
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %1, ah, bh, %1;\n\t}"
      : "+f"(c0), "+f"(c1) : "r"(a), "r"(b));
}
__device__ __forceinline__ void hl_add2(uint32_t a, float& c0, float& c1) {
  asm("{\n\t.reg .b16 al, ah;\n\t"
      "mov.b32 {al, ah}, %2;\n\t"
      "add.rn.f32.bf16 %0, al, %0;\n\tadd.rn.f32.bf16 %1, ah, %1;\n\t}"
      : "+f"(c0), "+f"(c1) : "r"(a));
}
__device__ __forceinline__ uint32_t hl_resid2(uint32_t hi, float a0, float a1) {
  uint32_t lo;
  asm("{\n\t.reg .b16 x, y, m1;\n\t.reg .f32 r0, r1;\n\t"
      "mov.b32 {x, y}, %1;\n\tmov.b16 m1, 0xBF80;\n\t"            // bf16(-1)
      "fma.rn.f32.bf16 r0, x, m1, %2;\n\tfma.rn.f32.bf16 r1, y, m1, %3;\n\t"
      "cvt.rn.bf16x2.f32 %0, r1, r0;\n\t}"
      : "=r"(lo) : "r"(hi), "f"(a0), "f"(a1));
  return lo;
}
__device__ __forceinline__ void hl_sum2(uint32_t h, uint32_t l, float& m0, float& m1) {
  m0 = hl_lo(h);
  m1 = hl_hi(h);
  hl_add2(l, m0, m1);
}

// Mixed-radix walk over (n-tile, x-tile, y-tile, image-tile) in steps of gridDim.x: no divisions per tile.
struct TileIt {
  int nt, tx, ty, tb, dn, dx, dy, db;
  __device__ __forceinline__ void init(int tile0, int step, const HlP& p) {
    int t = tile0;
    nt = t % p.tiles_n; t /= p.tiles_n;
    tx = t % p.tiles_x; t /= p.tiles_x;
    ty = t % p.tiles_y; tb = t / p.tiles_y;
    t = step;
    dn = t % p.tiles_n; t /= p.tiles_n;
    dx = t % p.tiles_x; t /= p.tiles_x;
    dy = t % p.tiles_y; db = t / p.tiles_y;
  }
  __device__ __forceinline__ void next(const HlP& p) {
    nt += dn;
    int c = nt >= p.tiles_n; nt -= c ? p.tiles_n : 0;
    tx += dx + c;
    c = tx >= p.tiles_x; tx -= c ? p.tiles_x : 0;
    ty += dy + c;
    c = ty >= p.tiles_y; ty -= c ? p.tiles_y : 0;
    tb += db + c;
  }
};

// G = 2: every weight slot (tap, 32-channel block) serves TWO pixel tiles before it is released -- the A stage holds
// both tiles, the two TMEM accumulator stages are the two tiles of the pair -- which halves the L2->SM weight traffic of
// the layers whose weights do not fit in shared memory (measured 38-47 B/clk/SM of weights: the UMMA pipe of those
// layers idled 20-35 % waiting for them).
struct TileCoord {
  int nt, tx, ty, tb;
  bool store;
};
__device__ __forceinline__ TileCoord decode_group_tile(int grp, int t, const HlP& p) {
  TileCoord c;
  c.nt = grp % p.tiles_n;
  int pt = (grp / p.tiles_n) * 2 + t;
  c.store = pt < p.pix_tiles;
  if (!c.store) pt = p.pix_tiles - 1;          // odd tile count: the pair's second half repeats the last tile, unsaved
  c.tx = pt % p.tiles_x;
  pt /= p.tiles_x;
  c.ty = pt % p.tiles_y;
  c.tb = pt / p.tiles_y;
  return c;
}

// KWC ("kw-concatenated", NT = 32, k = 3): the three kw taps of a filter row become N columns of ONE UMMA,
//   D[j, kw*NT + n] = sum_{kh, c} A[j + kh*R, c] * W[kh, kw, c, n]      (N = 3*NT = 96, A fetched once per kh)
//   out[j, n]       = sum_kw D[j + kw, kw*NT + n]                        (epilogue: warp shuffles across TMEM lanes)
// Why: a UMMA streams its operands from shared memory at 128 B/clk (tools/probe_mma_rate.cu: 44.8 / 48.1 / 64.1 cycles
// at N = 32 / 64 / 128 = (4 KB of A + N*32 B of B) / 128), so at N = 32 the 4 KB A tile costs 45 cycles for 16 cycles
// of math.  One A fetch for three taps cuts the UMMAs per (kh, K step) from 9 (414 cycles) to 4 of N = 96 (~224).
// The lane shift must stay inside one 32-lane TMEM quarter (a warp only reads its own quarter), so the halo box is
// R = 32 columns wide: quarter = one output row of the tile, lanes 30 / 31 are the halo margin that is junk anyway.
// CTA2 (NT = 128, streamed weights): two CTAs of a cluster -- the two SMs of a TPC -- work on two pixel tiles that
// share one N tile, and every UMMA is a cta_group::2 instruction of M = 256: rows 0-127 are the leader CTA's halo tile,
// rows 128-255 the peer's, and each CTA holds (and TMA-loads) only HALF of the weight slot (64 of the 128 rows of
// B).  Why: at N = 128 a single-CTA UMMA streams 4 KB of A + 4 KB of B per 64 cycles = exactly the 128 B/clk of shared
// memory, so the weight TMA fills (216 KB per channel block and tile) push the UMMA pipe down to ~70 % busy
// (profiles/r02_kwc_knobs.md section 6).  A pair halves the B bytes each SM reads AND writes.
// Protocol on top of the single-CTA one: the leader's issuer also waits for the peer's operands (pa_full / pb_full,
// relayed by the peer's otherwise idle warp 1 with remote mbarrier arrives), its commits are multicast to both CTAs'
// a_empty / b_empty / acc_full, and the peer's epilogue warps release the accumulator stage on the LEADER's acc_empty.
// HEAD = n_labels > 0 (NT = 32, resident weights, one tile per accumulator stage; the network's LAST 3x3 conv): the
// epilogue goes on with conv_final (k = 1, Brats.py:367,454) and mysoftmax (Brats.py:269-283) on the 32 channels it has
// just produced, so the forward ends here -- no final_conv_softmax launch, and the 32-channel tensor is not written
// unless a destination is given.  The two warps of a TMEM lane quarter own channels 0-15 / 16-31 of the same pixels:
// the first runs the per-pixel chain (sn_common.cuh: head_accumulate) over its channels, hands its 2*HEAD+1 partial sums
// to the second through shared memory, the second continues the SAME chain over channels 16-31 and finishes.  Channel
// order and arithmetic are those of final_conv_softmax_kernel on the bf16-rounded values the unfused path would have
// stored, so fused and unfused forwards agree bit for bit.
template <int NT, int KS, bool RESIDENT, bool DGRAD, int G = 1, bool KWC = false, bool CTA2 = false, int HEAD = 0>
__global__ void __launch_bounds__(hl_threads(hl_two_sets(KWC, G, DGRAD)), 1) conv_moments_halo_kernel(const __grid_constant__ HlMaps maps,
                                                                          const HlP p) {
  static_assert(HEAD == 0 || (NT == 32 && RESIDENT && !DGRAD && G == 1 && !KWC && !CTA2 && HEAD <= 5),
                "fused head: last 32-channel conv, tap-shift kernel");
  static_assert(!KWC || (NT == 32 && KS == 3 && G == 1), "kw-concatenation: 32-column tiles of a 3x3 conv");
  static_assert(!CTA2 || (((NT == 128 && !RESIDENT) || (NT == 64 && RESIDENT)) && G == 1 && !KWC),
                "CTA pairs: 128-column tiles with streamed weights, 64-column tiles with resident half-slots");
  constexpr int NB = KWC ? KS * NT : NT;                  // rows (GEMM N) of one weight plane of a slot
  constexpr int SLOTS_PER_CB = KWC ? KS : KS * KS;        // weight slots per 32-channel block: filter rows / taps
  constexpr bool CONCAT = !KWC && NT <= 64;               // hi x [W_hi ; W_lo] as one UMMA of N = 2*NT
  // CTA pair + CONCAT (NT = 64): a cta_group::2 UMMA takes the first half of its N columns from the leader's shared
  // memory and the second half from the peer's, at the SAME offsets.  Slot layout per CTA (rank r), 2*NT rows of 64 B:
  //   X [NT rows]   : r = 0 W_hi[0:NT], r = 1 W_lo[0:NT]        -> hi x [W_hi ; W_lo], N = 2*NT
  //   Y [NT/2 rows] : W_hi[r*NT/2 : (r+1)*NT/2]                  -> lo x W_hi,         N = NT
  //   Z [NT/2 rows] : W^2 [r*NT/2 : (r+1)*NT/2]                  -> var x W^2,         N = NT
  // Same instruction order per accumulator column as the single-CTA CONCAT kernel: bit-identical results.
  constexpr bool PAIR_CONCAT = CTA2 && CONCAT;
  constexpr int B_PLANE = (CTA2 ? NB / 2 : NB) * HL_KC * 2;   // CTA pair: this CTA's half of the plane
  constexpr int B_SLOT = PAIR_CONCAT ? 2 * NT * HL_KC * 2 : 3 * B_PLANE;
  constexpr int B_OFF_Y = NT * HL_KC * 2, B_OFF_Z = B_OFF_Y + (NT / 2) * HL_KC * 2;      // PAIR_CONCAT regions
  constexpr int ACC_STAGE = KWC ? 2 * NB : (CONCAT ? 3 * NT : 2 * NT);     // TMEM columns per accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_STAGE <= 256 ? 256 : 512;              // 2 stages, rounded up to a power of two
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  static_assert(G == 1 || !RESIDENT, "two tiles per weight slot only make sense for streamed weights");
  const int a_stage = G * 3 * p.a_plane;
  const uint32_t b_base = smem_base + p.sa * a_stage;
  const uint32_t bar_base = b_base + p.sb * B_SLOT;           // 1 KB barrier block, then 2 KB q buffers
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (HL_MAX_ASTAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + HL_MAX_BSLOTS + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 2 + s); };
  auto q_full = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 4 + s); };
  auto q_empty = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 6 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 8);
  // CTA pair, leader only: "the peer's A stage has landed" (its weight half-slots signal the leader's b_full directly)
  auto pa_full = [&](int s) { return bar_base + 8u * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 9 + s); };
  const uint32_t crank = CTA2 ? ptx::cluster_ctarank() : 0u;
  // CTA pairs poll (barriers there are completed from the other CTA; see mbar_wait_poll), everything else may park
  auto WAIT = [](uint32_t bar, uint32_t parity) {
    if constexpr (CTA2) ptx::mbar_wait_poll(bar, parity);
    else ptx::mbar_wait(bar, parity);
  };
  // persistent walk: tiles (or tile pairs: G = 2 inside one CTA, CTA2 across the pair) u0, u0 + ustep, ...
  constexpr bool GROUPED = G == 2 || CTA2;
  const int u0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_units = GROUPED ? p.total_groups : p.total_tiles;
  const int bar_off = p.sa * a_stage + p.sb * B_SLOT;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * HL_MAX_ASTAGES + 2 * HL_MAX_BSLOTS + 8));
  float* qbuf = reinterpret_cast<float*>(smem_gen + bar_off + 1024);     // [2 stages][256]
  float* s_sm = qbuf + 512;                                              // softplus(w_sigma) [cout <= 512]
  // fused head: W [32][HEAD], W^2 [32][HEAD], s [8], then the exchange array [2*HEAD+1][128 rows]
  float* head_w = s_sm + 512;
  float* head_w2 = head_w + 32 * (HEAD > 0 ? HEAD : 1);
  float* head_s = head_w2 + 32 * (HEAD > 0 ? HEAD : 1);
  float* head_ex = head_s + 8;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int cblk = p.cblk_s[0] + p.cblk_s[1] + p.cblk_s[2] + p.cblk_s[3];
  // Programmatic dependent launch: let the next kernel of the stream start its prologue (barriers, TMEM, resident
  // weights) on SMs this grid has already left; it blocks in griddepcontrol.wait until this grid has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int taps = KS * KS;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 4; ++s)
      if (p.cblk_s[s])
        for (int pl = 0; pl < 3; ++pl) ptx::prefetch_tensormap(&maps.a[s][pl]);
    ptx::prefetch_tensormap(&maps.w);
    if ((KWC && !DGRAD && p.tma_store) || PAIR_CONCAT) ptx::prefetch_tensormap(&maps.d);
    for (int s = 0; s < p.sa; ++s) {
      ptx::mbar_init(a_full(s), 1);
      ptx::mbar_init(a_empty(s), 9);          // UMMA commit + the 8 reducer warps
    }
    for (int s = 0; s < p.sb; ++s) {
      ptx::mbar_init(b_full(s), 1);
      ptx::mbar_init(b_empty(s), 1);
    }
    if constexpr (CTA2) {
      for (int s = 0; s < p.sa; ++s) ptx::mbar_init(pa_full(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc_full(s), 1);
      ptx::mbar_init(acc_empty(s), CTA2 ? 16 : 8);        // 8 epilogue warps (CTA pair: of both CTAs, on the leader)
      ptx::mbar_init(q_full(s), 8);           // 8 reducer warps
      ptx::mbar_init(q_empty(s), 8);          // 8 epilogue warps
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      ptx::tmem_alloc2(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  constexpr bool TWO_SETS = hl_two_sets(KWC, G, DGRAD);
  for (int i = threadIdx.x; i < p.s_len; i += hl_threads(TWO_SETS)) s_sm[i] = p.s[i];
  if constexpr (HEAD > 0) {
    for (int i = threadIdx.x; i < 32 * HEAD; i += hl_threads(TWO_SETS)) {
      const float v = p.head_w[i];
      head_w[i] = v;
      head_w2[i] = v * v;
    }
    if (threadIdx.x < HEAD) head_s[threadIdx.x] = softplus_f(p.head_ws[threadIdx.x]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) ptx::cluster_sync_all();      // both CTAs' barriers exist before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Everything below that depends on the previous kernel of the stream -- activation tiles (TMA), saved activations
  // (data-gradient epilogue) and the destination buffers -- comes after this wait.  The producer warp is the exception:
  // it issues the RESIDENT weight loads (constants) first and waits right before its first activation load.
  if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // one weight slot = ONE TMA request: the weight tensor is mapped as (cin, n, tap, plane) and the box takes all three
      // operand planes of tap `tap` (KWC: the three kw taps of filter row `tap` as well), landing as [plane][(kw)][n] rows
      // -- exactly the slot layout.  (Three / nine requests per slot made the weight stream request-rate bound: ~250
      // cycles per request whatever its size, profiles/r02_cta2.md.)
      auto load_b_slot = [&](uint32_t sb_addr, uint32_t bar, int cbt, int tap, int ncol0, int n0) {
        const int row = p.upconv ? ncol0 : n0;                      // up-conv: the (parity, channel) axis, one "tap"
        const int t0 = p.upconv ? 0 : (KWC ? tap * KS : tap);
        if constexpr (PAIR_CONCAT) {
          // three requests (maps.w: NT rows of one plane, maps.d: NT/2 rows of one plane), all signalled on the LEADER's
          // b_full; used with resident weights only, so the request count is a one-off
          const uint32_t lb = ptx::map_to_rank(bar, 0);
          ptx::tma_load_4d_pair(sb_addr, &maps.w, lb, cbt * HL_KC, row, t0, (int)crank);                        // X
          ptx::tma_load_4d_pair(sb_addr + B_OFF_Y, &maps.d, lb, cbt * HL_KC, row + (int)crank * (NT / 2), t0, 0);   // Y
          ptx::tma_load_4d_pair(sb_addr + B_OFF_Z, &maps.d, lb, cbt * HL_KC, row + (int)crank * (NT / 2), t0, 2);   // Z
        } else if constexpr (CTA2) {
          // CTA pair: this CTA's 64 rows of the 128-row N tile, landing in its own shared memory but signalled on the
          // LEADER's b_full, which expects both halves
          ptx::tma_load_4d_pair(sb_addr, &maps.w, ptx::map_to_rank(bar, 0), cbt * HL_KC, row + (int)crank * (NT / 2), t0, 0);
        } else {
          ptx::tma_load_4d(sb_addr, &maps.w, bar, cbt * HL_KC, row, t0, 0);
        }
      };
      bool dep_waited = false;
      int ai = 0, bi = 0;     // running A-stage / B-slot fill counters
      TileIt it;
      it.init(blockIdx.x, gridDim.x, p);
      if (RESIDENT && u0 < n_units) {
        // resident weights: the whole layer's operands for this CTA's N tile, before the dependency wait (CTA pair:
        // tiles_n == 1, each CTA keeps ITS half of every slot and both halves complete the leader's barrier)
        const int ncol0 = CTA2 ? 0 : it.nt * NT;
        const int group = ncol0 / p.cout;
        const int n0 = ncol0 - group * p.cout;
        for (int cbt = 0; cbt < cblk; ++cbt)
          for (int tap = 0; tap < SLOTS_PER_CB; ++tap) {
            const int slot = cbt * SLOTS_PER_CB + tap;
            if (!CTA2 || crank == 0) ptx::mbar_arrive_expect_tx(b_full(slot), (uint32_t)(CTA2 ? 2 * B_SLOT : B_SLOT));
            load_b_slot(b_base + slot * B_SLOT, b_full(slot), cbt, tap, ncol0, n0);
          }
      }
      for (int tile = u0, titer = 0; tile < n_units; tile += ustep, ++titer, it.next(p)) {
        int nt_i, x0s[G], y0s[G], b0s[G];
        if constexpr (CTA2) {
          const TileCoord c = decode_group_tile(tile, (int)crank, p);      // this CTA's half of the pair
          nt_i = c.nt;
          x0s[0] = c.tx * p.TWo - p.pad; y0s[0] = c.ty * p.THo - p.pad; b0s[0] = c.tb * p.TN;
        } else if constexpr (G == 1) {
          nt_i = it.nt;
          x0s[0] = it.tx * p.TWo - p.pad; y0s[0] = it.ty * p.THo - p.pad; b0s[0] = it.tb * p.TN;
        } else {
#pragma unroll
          for (int t = 0; t < G; ++t) {
            const TileCoord c = decode_group_tile(tile, t, p);
            nt_i = c.nt;
            x0s[t] = c.tx * p.TWo - p.pad; y0s[t] = c.ty * p.THo - p.pad; b0s[t] = c.tb * p.TN;
          }
        }
        const int ncol0 = nt_i * NT;
        const int group = ncol0 / p.cout;
        const int n0 = ncol0 - group * p.cout;
        int src = 0, cb = 0;
        for (int cbt = 0; cbt < cblk; ++cbt, ++cb) {
          while (cb >= p.cblk_s[src]) { cb = 0; ++src; }
          if (!dep_waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            dep_waited = true;
          }
          {
            const int stage = ai % p.sa;
            const uint32_t parity = (uint32_t)(ai / p.sa) & 1u;
            ++ai;
            WAIT(a_empty(stage), parity ^ 1u);
            ptx::mbar_arrive_expect_tx(a_full(stage), (uint32_t)(G * 3 * p.rows_box * 64));
            const uint32_t sa_addr = smem_base + stage * a_stage;
#pragma unroll
            for (int t = 0; t < G; ++t)
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                    " [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    :
                    : "r"(sa_addr + (t * 3 + pl) * p.a_plane), "l"(reinterpret_cast<uint64_t>(&maps.a[src][pl])),
                      "r"(a_full(stage)), "r"(cb * HL_KC), "r"(x0s[t]), "r"(y0s[t]), "r"(b0s[t])
                    : "memory");
              }
          }
          if (!RESIDENT) {
            for (int tap = 0; tap < SLOTS_PER_CB; ++tap) {
              int slot;
              {
                slot = bi % p.sb;
                const uint32_t parity = (uint32_t)(bi / p.sb) & 1u;
                ++bi;
                WAIT(b_empty(slot), parity ^ 1u);
              }
              if (!CTA2 || crank == 0) ptx::mbar_arrive_expect_tx(b_full(slot), (uint32_t)(CTA2 ? 2 * B_SLOT : B_SLOT));
              load_b_slot(b_base + slot * B_SLOT, b_full(slot), cbt, tap, ncol0, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    // UMMA cost measured on B200 (tools/probe_mma_rate.cu): M=128,K=16 takes 44.8 / 48.1 / 64.1 / 128.1 cycles at
    // N = 32 / 64 / 128 / 256 -- a ~45-cycle floor per instruction (the 4 KB A-operand fetch).  For NT <= 64 the
    // mean therefore uses TWO instructions instead of three: A_hi x [W_hi ; W_lo] as one N = 2*NT UMMA (the two
    // weight planes are adjacent in the B slot) into columns [0, 2NT), and A_lo x W_hi into columns [0, NT);
    // the epilogue adds the two column halves.  NT = 128 is already math-bound and keeps 3 + 1 N = 128 UMMAs.
    {
      // All 32 lanes run this loop (uniform control flow keeps the descriptor arithmetic in the uniform datapath);
      // one elected lane issues the UMMAs and commits.  Taps and K steps are fully unrolled so every descriptor is
      // "per-stage base + immediate": the issuing thread must sustain one UMMA per ~45 cycles.
      constexpr uint32_t idesc_n = ptx::idesc_bf16_f32(CTA2 ? 2 * HL_BM : HL_BM, NB);
      constexpr uint32_t idesc_2n = ptx::idesc_bf16_f32(CTA2 ? 2 * HL_BM : HL_BM, CONCAT ? 2 * NT : NT);
      // descriptor = constant high word | (1 << 16 | address >> 4) in the low word (smem < 256 KB: 14 bits);
      // high word of smem_desc_kmajor<64>: SBO = 512 B >> 4 at bits [32,46), version 1 at bit 46, SW64 (4) at [61,64)
      constexpr uint32_t DESC_HI = 32u | (1u << 14) | (4u << 29);
      auto desc = [&](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
      const bool leader = ptx::elect_one();
      const uint32_t a_lo_base = 0x10000u | (smem_base >> 4);
      const uint32_t b_lo_base = 0x10000u | (b_base >> 4);
      const uint32_t stage16 = (uint32_t)a_stage >> 4, plane16 = (uint32_t)p.a_plane >> 4;
      const uint32_t row_shift16 = (uint32_t)p.R * 4u;            // one halo row = R pixels x 64 B, in 16-byte units
      constexpr uint32_t SLOT16 = B_SLOT >> 4, BPLANE16 = B_PLANE >> 4;
      int a_stage_i = 0, b_slot_i = 0;
      uint32_t a_par = 0, b_par = 0;
      if (CTA2 && crank != 0) {
        // ---- peer CTA of a pair: no UMMA issue here; relay "my operands have landed" to the leader's barriers, in the
        // order the leader consumes them (a refill of a stage / slot needs the leader's multicast commit first, so no
        // phase can be skipped or signalled twice)
        for (int tile = u0; tile < n_units; tile += ustep) {
          for (int cbt = 0; cbt < cblk; ++cbt) {
            WAIT(a_full(a_stage_i), a_par);
            if (leader) ptx::mbar_arrive_remote(pa_full(a_stage_i), 0);
            if (++a_stage_i == p.sa) { a_stage_i = 0; a_par ^= 1u; }
          }
        }
      } else
      for (int tile = u0, uiter = 0; tile < n_units; tile += ustep, ++uiter) {
        // unit = one tile (G = 1, accumulator stage alternates) or a pair of tiles (G = 2, stage = tile of the pair)
        const int titer = uiter * G;
#pragma unroll
        for (int t = 0; t < G; ++t)
          WAIT(acc_empty((titer + t) & 1), (((uint32_t)(titer + t) >> 1) & 1u) ^ 1u);
        const int as = titer & 1;
        const uint32_t acc_mu0 = tmem_base + as * ACC_STAGE;
        for (int cbt = 0; cbt < cblk; ++cbt) {
          WAIT(a_full(a_stage_i), a_par);
          if constexpr (CTA2) WAIT(pa_full(a_stage_i), a_par);
          if (RESIDENT && titer == 0) {
            for (int tap = 0; tap < SLOTS_PER_CB; ++tap) WAIT(b_full(cbt * SLOTS_PER_CB + tap), 0);
          }
          ptx::tc_fence_after();
          const uint32_t a_st = a_lo_base + a_stage_i * stage16;
          uint32_t b_st = b_lo_base + (RESIDENT ? (uint32_t)(cbt * SLOTS_PER_CB) * SLOT16 : 0u);
#pragma unroll
          for (int tap = 0; tap < SLOTS_PER_CB; ++tap) {
            // compile-time after unrolling; KWC: one slot per filter row, the kw shift happens in the epilogue
            const int kh = KWC ? tap : tap / KS, kw = KWC ? 0 : tap % KS;
            uint32_t b0;
            int slot = 0;
            if (RESIDENT) {
              b0 = b_st + (uint32_t)tap * SLOT16;
            } else {
              slot = b_slot_i;
              WAIT(b_full(slot), b_par);
              ptx::tc_fence_after();
              if (++b_slot_i == p.sb) { b_slot_i = 0; b_par ^= 1u; }
              b0 = b_st + (uint32_t)slot * SLOT16;
            }
            const uint32_t a00 = a_st + (uint32_t)kh * row_shift16 + (uint32_t)kw * 4u;  // tap = row offset in the halo
            if (leader && (!(HL_DBG(p) & 2) || tap == 0)) {
#pragma unroll
              for (int t = 0; t < G; ++t) {
                const uint32_t a0 = a00 + (uint32_t)t * 3u * plane16;             // tile t of the pair
                const uint32_t acc_mu = G == 1 ? acc_mu0 : tmem_base + (uint32_t)t * ACC_STAGE;
                const uint32_t acc_var = acc_mu + (CONCAT ? 2 * NT : NB);
#pragma unroll
                for (int ks = 0; ks < HL_KC / 16; ++ks) {
                  const uint32_t k16 = ks * 2;                          // 16 bf16 = 32 B along K
                  const uint32_t acc = (cbt > 0 || tap > 0 || ks > 0) ? 1u : 0u;
                  if constexpr (PAIR_CONCAT) {
                    ptx::umma2_bf16(acc_mu, desc(a0 + k16), desc(b0 + k16), idesc_2n, acc);                           // X
                    ptx::umma2_bf16(acc_mu, desc(a0 + plane16 + k16), desc(b0 + (B_OFF_Y >> 4) + k16), idesc_n, 1u);  // Y
                  } else if constexpr (CONCAT) {
                    ptx::umma_bf16(acc_mu, desc(a0 + k16), desc(b0 + k16), idesc_2n, acc);               // hi x [Whi;Wlo]
                    ptx::umma_bf16(acc_mu, desc(a0 + plane16 + k16), desc(b0 + k16), idesc_n, 1u);      // lo x Whi
                  } else if constexpr (CTA2) {
                    ptx::umma2_bf16(acc_mu, desc(a0 + k16), desc(b0 + k16), idesc_n, acc);
                    ptx::umma2_bf16(acc_mu, desc(a0 + plane16 + k16), desc(b0 + k16), idesc_n, 1u);
                    ptx::umma2_bf16(acc_mu, desc(a0 + k16), desc(b0 + BPLANE16 + k16), idesc_n, 1u);
                  } else {
                    ptx::umma_bf16(acc_mu, desc(a0 + k16), desc(b0 + k16), idesc_n, acc);
                    ptx::umma_bf16(acc_mu, desc(a0 + plane16 + k16), desc(b0 + k16), idesc_n, 1u);
                    ptx::umma_bf16(acc_mu, desc(a0 + k16), desc(b0 + BPLANE16 + k16), idesc_n, 1u);
                  }
                  if constexpr (PAIR_CONCAT)
                    ptx::umma2_bf16(acc_var, desc(a0 + 2 * plane16 + k16), desc(b0 + (B_OFF_Z >> 4) + k16), idesc_n, acc);
                  else if constexpr (CTA2)
                    ptx::umma2_bf16(acc_var, desc(a0 + 2 * plane16 + k16), desc(b0 + 2 * BPLANE16 + k16), idesc_n, acc);
                  else
                    ptx::umma_bf16(acc_var, desc(a0 + 2 * plane16 + k16), desc(b0 + 2 * BPLANE16 + k16), idesc_n, acc);
                }
              }
            }
            if (leader && !RESIDENT) {
              if constexpr (CTA2) ptx::umma2_commit(b_empty(slot)); else ptx::umma_commit(b_empty(slot));
            }
          }
          if (leader) {
            if constexpr (CTA2) ptx::umma2_commit(a_empty(a_stage_i)); else ptx::umma_commit(a_empty(a_stage_i));
          }
          if (++a_stage_i == p.sa) { a_stage_i = 0; a_par ^= 1u; }
        }
        if (leader) {
#pragma unroll
          for (int t = 0; t < G; ++t) {
            if constexpr (CTA2) ptx::umma2_commit(acc_full((titer + t) & 1));
            else ptx::umma_commit(acc_full((titer + t) & 1));
          }
        }
      }
    }
  } else {
    if (warp < 10) {
      // ===================== q reduction (warps 2-9): q[halo pixel] = sum_c (mu^2 + var) =====================
      // These eight warps (thread = halo pixel row) consume EVERY A stage in order, so the stage barriers see one
      // consistent consumer, and hand the per-tile q array to the epilogue warps through a double-buffered smem
      // array.  (With four warps this role was the critical path of the 32-channel layers.)
      const int row = (warp - 2) * 32 + lane;
      int a_stage_i = 0;
      uint32_t a_par = 0;
      for (int tile = u0, uiter = 0; tile < n_units; tile += ustep, ++uiter) {
        float qsum[G];
#pragma unroll
        for (int t = 0; t < G; ++t) qsum[t] = 0.f;
        int src = 0, cb = 0;
        for (int cbt = 0; cbt < cblk; ++cbt, ++cb) {
          if constexpr (DGRAD) {
            while (cb >= p.cblk_s[src]) { cb = 0; ++src; }
          }
          WAIT(a_full(a_stage_i), a_par);
          if (row < p.rows_box && !(HL_DBG(p) & 4)) {
#pragma unroll
           for (int t = 0; t < G; ++t) {
            float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;         // independent chains for ILP
            float qe = 0.f, qf = 0.f;                             // forward only: sum of hi * lo
            const uint8_t* ar = smem_gen + a_stage_i * a_stage + t * 3 * p.a_plane + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = ((j + (row >> 1)) & 3) * 16;   // chunk rotation: conflict-free, order-insensitive sum
              if constexpr (DGRAD) {
                // t[pixel] = sum_n g_var[pixel, n] * s_n: only the variance-gradient plane
                const uint4 vv4 = *reinterpret_cast<const uint4*>(ar + 2 * p.a_plane + ch);
                // SWIZZLE_64B: physical 16-byte chunk c of smem row r holds logical chunk c ^ ((r >> 1) & 3)
                const float* sp = s_sm + cb * HL_KC + (((ch >> 4) ^ ((row >> 1) & 3)) << 3);
                const float4 s0 = *reinterpret_cast<const float4*>(sp), s1 = *reinterpret_cast<const float4*>(sp + 4);
                qa = fmaf(hl_lo(vv4.x), s0.x, qa); qb = fmaf(hl_hi(vv4.x), s0.y, qb);
                qc = fmaf(hl_lo(vv4.y), s0.z, qc); qd = fmaf(hl_hi(vv4.y), s0.w, qd);
                qa = fmaf(hl_lo(vv4.z), s1.x, qa); qb = fmaf(hl_hi(vv4.z), s1.y, qb);
                qc = fmaf(hl_lo(vv4.w), s1.z, qc); qd = fmaf(hl_hi(vv4.w), s1.w, qd);
              } else {
                const uint4 hh4 = *reinterpret_cast<const uint4*>(ar + ch);
                const uint4 ll4 = *reinterpret_cast<const uint4*>(ar + p.a_plane + ch);
                const uint4 vv4 = *reinterpret_cast<const uint4*>(ar + 2 * p.a_plane + ch);
                const uint32_t hh[4] = {hh4.x, hh4.y, hh4.z, hh4.w}, ll[4] = {ll4.x, ll4.y, ll4.z, ll4.w},
                               vv[4] = {vv4.x, vv4.y, vv4.z, vv4.w};
                // mu^2 = (hi + lo)^2 = hi^2 + 2 hi lo (+ lo^2 <= 2^-18 hi^2, dropped: far below the bf16 variance operands
                // this term is added to); six mixed-precision instructions per channel pair instead of fourteen
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hl_fma2(hh[e], hh[e], qa, qb);
                  hl_fma2(hh[e], ll[e], qe, qf);
                  hl_add2(vv[e], qc, qd);
                }
              }
            }
            qsum[t] += ((qa + qb) + (qc + qd)) + 2.f * (qe + qf);
           }
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(a_empty(a_stage_i));
          if (++a_stage_i == p.sa) { a_stage_i = 0; a_par ^= 1u; }
        }
#pragma unroll
        for (int t = 0; t < G; ++t) {
          const int titer = uiter * G + t;
          const int qs = titer & 1;
          WAIT(q_empty(qs), (((uint32_t)titer >> 1) & 1u) ^ 1u);
          qbuf[qs * 256 + row] = qsum[t];                      // rows >= rows_box hold 0
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(q_full(qs));
        }
      }
    } else {
      // ===================== epilogue (warps 10-17; kw-concatenated kernels: 10-25 in two sets) =====================
      // With the UMMA count per tile halved (KWC) this role's per-warp dependent chain (barrier waits, TMEM loads,
      // shuffles, converts, stores: ~500 instructions per tile at ~0.17 IPC) is the longest path per tile (ncu: the
      // issuer waits on acc_empty, the reduction warps on q_empty).  Two sets of eight warps take alternate tiles, set s
      // always draining accumulator stage s.
      constexpr int ESETS = TWO_SETS ? 2 : 1;
      const int eset = (warp - 10) >> 3;
      const int q = warp & 3;                   // TMEM lane quarter this warp may access
      const int half = ((warp - 10) >> 2) & 1;  // which half of the tile's NT columns this warp converts
      constexpr int NH = NT / 2;
      const int row = q * 32 + lane;            // GEMM row == TMEM lane == halo pixel index
      const int x = row % p.R;
      const int yy = row / p.R;
      const int y = yy % p.THb;
      const int n = yy / p.THb;
      TileIt it;
      it.init(blockIdx.x, gridDim.x, p);
      // G = 2: the "tiles" of this CTA are the halves of its tile pairs, in order (stage = half of the pair)
      const int n_units_cta = n_units > u0 ? (n_units - 1 - u0) / ustep + 1 : 0;
      const int n_tiles_cta = G * n_units_cta;
      for (int titer = 0; titer < n_tiles_cta; ++titer) {
        TileCoord tc;
        if constexpr (CTA2) {
          tc = decode_group_tile(u0 + titer * ustep, (int)crank, p);       // this CTA's half of the pair
        } else if constexpr (G == 1) {
          tc.nt = it.nt; tc.tx = it.tx; tc.ty = it.ty; tc.tb = it.tb; tc.store = true;
          it.next(p);
          if (ESETS == 2 && (titer & 1) != eset) continue;     // the other set's tile
        } else {
          if (ESETS == 2 && (titer & 1) != eset) continue;     // the other set's half of the pair
          tc = decode_group_tile((int)blockIdx.x + (titer >> 1) * (int)gridDim.x, titer & 1, p);
        }
        const int as = titer & 1;
        const uint32_t par = ((uint32_t)titer >> 1) & 1u;
        // this warp's TMEM reads of stage `as` are done (CTA pair: the issuer lives in the leader CTA)
        auto release_acc = [&]() {
          if (CTA2 && crank != 0) ptx::mbar_arrive_remote(acc_empty(as), 0);
          else ptx::mbar_arrive(acc_empty(as));
        };
        // The accumulator stage goes back to the issuer as soon as this warp's LAST TMEM load has returned -- the math,
        // conversions and stores of that chunk run on registers.  (With two pixel tiles per weight slot the issuer needs
        // BOTH stages back before the next pair: 28 % of its stall samples were this wait, tools/ncu_waits.py.)
        constexpr bool EARLY_REL = DGRAD || !TWO_SETS;
        auto early_release = [&]() {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc();
        };
        WAIT(q_full(as), par);
        const float* myq = qbuf + as * 256;
        float qv[taps];
#pragma unroll
        for (int kh = 0; kh < KS; ++kh)
#pragma unroll
          for (int kw = 0; kw < KS; ++kw) {
            const int idx = row + kh * p.R + kw;        // independent loads first, then one short add chain
            qv[kh * KS + kw] = idx < 256 ? myq[idx] : 0.f;
          }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < taps; ++i) r += qv[i];
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(q_empty(as));

        // ---- tile coordinates
        const int nt_i = tc.nt, tx = tc.tx, ty = tc.ty, tb = tc.tb;
        const int gcol0 = nt_i * NT + half * NH;               // first column (parity group, channel) of this warp
        const int ox_i = tx * p.TWo + x, oy_i = ty * p.THo + y, ob = tb * p.TN + n;
        const bool valid = tc.store && x < p.TWo && y < p.THo && n < p.TN && ox_i < p.Wo && oy_i < p.Ho && ob < p.B &&
                           !(HL_DBG(p) & 8);
        if (!DGRAD && p.r_out != nullptr && half == 0 && nt_i == 0 && valid)
          p.r_out[((size_t)ob * p.Ho + oy_i) * p.Wo + ox_i] = r;
        WAIT(acc_full(as), par);
        ptx::tc_fence_after();
        if (HL_DBG(p) & 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc();
          continue;
        }
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STAGE + half * NH;
        // 16 mean and 16 variance accumulator columns (c0 .. c0+15 of this warp's half) of this thread's GEMM row
        // TMA-store epilogue (KWC forward): the two warps of a (set, quarter) pair stage their tile row -- 30 pixels x
        // [hi | lo | var] x 32 channels = 5760 contiguous bytes, 64-byte swizzled so the 192-byte pixel stride is
        // bank-conflict free -- and one lane hands it to the TMA.  Why: a lane-per-pixel st.global.v8 touches 32 cache
        // lines per instruction = 32 LSU wavefronts (758 of the 2185 LSU wavefronts per tile, ncu); the LSU data pipe,
        // not the UMMA pipe, bounded the kw-concatenated kernel.
        const bool tma_st = KWC && !DGRAD && p.tma_store;
        const uint32_t stg = bar_base + 5120u + (uint32_t)(eset * 4 + q) * HL_STG_BUF;
        const uint32_t pair_bar = 1u + (uint32_t)(eset * 4 + q);
        auto stage32 = [&](int pl, const uint32_t (&w8)[8]) {      // this lane's 16 channels of plane pl
          if (HL_DBG(p) & 16) return;
          const uint32_t L = (uint32_t)lane * 192u + (uint32_t)pl * 64u + (uint32_t)half * 32u;
          const uint32_t x4 = ((L >> 7) & 3u) << 4;                 // SWIZZLE_64B: bits [4,6) ^= bits [7,9)
          ptx::st_shared_v4(stg + (L ^ x4), w8[0], w8[1], w8[2], w8[3]);
          ptx::st_shared_v4(stg + ((L + 16u) ^ x4), w8[4], w8[5], w8[6], w8[7]);
        };
        if (tma_st) {
          if (half == 0 && lane == 0) ptx::bulk_wait_read0();       // the pair's previous store has left the buffer
          if (!(HL_DBG(p) & 128)) ptx::named_barrier(pair_bar, 64);
        }
        // KWC: out[j] = D[j][kw = 0] + D[j+1][kw = 1] + D[j+2][kw = 2]; rows j+1, j+2 are the next TMEM lanes = the next
        // threads of this warp (R = 32: a quarter is one tile row; lanes 30, 31 wrap into junk and are never stored)
        auto load_kwc16 = [&](uint32_t taddr, uint32_t (&a)[16]) {
          uint32_t t1[16], t2[16];
          ptx::tmem_ld16(taddr, a);
          ptx::tmem_ld16(taddr + NT, t1);
          ptx::tmem_ld16(taddr + 2 * NT, t2);
          ptx::tmem_ld_wait();
          if (HL_DBG(p) & 32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(t1[j]) + __uint_as_float(t2[j]));
            return;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            a[j] = __float_as_uint(__uint_as_float(a[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(t1[j]), 1) +
                                   __shfl_down_sync(0xffffffffu, __uint_as_float(t2[j]), 2));
        };
        auto load_acc16 = [&](int c0, uint32_t (&am)[16], uint32_t (&av)[16]) {
          if constexpr (KWC) {
            load_kwc16(lane_base + c0, am);
            load_kwc16(lane_base + NB + c0, av);
          } else {
            ptx::tmem_ld16(lane_base + c0, am);
            ptx::tmem_ld16(lane_base + (CONCAT ? 2 * NT : NT) + c0, av);
            if constexpr (CONCAT) {
              uint32_t am2[16];
              ptx::tmem_ld16(lane_base + NT + c0, am2);     // the hi x W_lo half of the mean
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) am[j] = __float_as_uint(__uint_as_float(am[j]) + __uint_as_float(am2[j]));
            } else {
              ptx::tmem_ld_wait();
            }
          }
        };
        if constexpr (DGRAD) {
          // ---- data gradient: g_mu = acc_mu + 2 mu_saved T, g_var = acc_var + T, T = r; then the
          // ReLU gate of the layer that produced the forward input (Brats.py:233-238: the gate is mu_saved > 0)
#pragma unroll 1
          for (int c0 = 0; c0 < NH; c0 += 16) {
            const int gcol = gcol0 + c0;
            const int seg = gcol >= p.csplit ? 1 : 0;
            const int nch = gcol - (seg ? p.csplit : 0);
            const HlView& dv = p.gdst[seg];
            const HlView& sv = p.saved[seg];
            uint32_t shw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, slw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (valid) {
              const __nv_bfloat16* sp =
                  sv.base + ((((size_t)ob * sv.h + oy_i + sv.y0) * sv.w + ox_i + sv.x0) * 3) * sv.c + sv.c0 + nch;
              if (p.v8) {
                ptx::ld_global_v8(sp, shw);
                ptx::ld_global_v8(sp + sv.c, slw);
              } else {
                const uint4 sh0 = *reinterpret_cast<const uint4*>(sp), sh1 = *reinterpret_cast<const uint4*>(sp + 8);
                const uint4 sl0 = *reinterpret_cast<const uint4*>(sp + sv.c);
                const uint4 sl1 = *reinterpret_cast<const uint4*>(sp + sv.c + 8);
                shw[0] = sh0.x; shw[1] = sh0.y; shw[2] = sh0.z; shw[3] = sh0.w;
                shw[4] = sh1.x; shw[5] = sh1.y; shw[6] = sh1.z; shw[7] = sh1.w;
                slw[0] = sl0.x; slw[1] = sl0.y; slw[2] = sl0.z; slw[3] = sl0.w;
                slw[4] = sl1.x; slw[5] = sl1.y; slw[6] = sl1.z; slw[7] = sl1.w;
              }
            }
            uint32_t am[16], av[16];
            load_acc16(c0, am, av);
            if (c0 + 16 >= NH) early_release();
            const bool gate = p.gate[seg] != 0;
            const float r2 = 2.f * r;
            uint32_t hi[8], lo[8], vr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float m0, m1;
              hl_sum2(shw[j], slw[j], m0, m1);
              float a0 = fmaf(m0, r2, __uint_as_float(am[2 * j])), a1 = fmaf(m1, r2, __uint_as_float(am[2 * j + 1]));
              float v0 = __uint_as_float(av[2 * j]) + r, v1 = __uint_as_float(av[2 * j + 1]) + r;
              if (gate) {
                if (!(m0 > 0.f)) { a0 = 0.f; v0 = 0.f; }
                if (!(m1 > 0.f)) { a1 = 0.f; v1 = 0.f; }
              }
              hi[j] = hl_pack2(a0, a1);
              lo[j] = hl_resid2(hi[j], a0, a1);
              vr[j] = hl_pack2(v0, v1);
            }
            if (valid) {
              __nv_bfloat16* d_hi =
                  dv.base + ((((size_t)ob * dv.h + oy_i + dv.y0) * dv.w + ox_i + dv.x0) * 3) * dv.c + dv.c0 + nch;
              if (p.v8) {
                ptx::st_global_v8(d_hi, hi);
                ptx::st_global_v8(d_hi + dv.c, lo);
                ptx::st_global_v8(d_hi + 2 * dv.c, vr);
              } else {
                uint4* ph = reinterpret_cast<uint4*>(d_hi);
                uint4* pl = reinterpret_cast<uint4*>(d_hi + dv.c);
                uint4* pv = reinterpret_cast<uint4*>(d_hi + 2 * dv.c);
                ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                pv[0] = make_uint4(vr[0], vr[1], vr[2], vr[3]);
                pv[1] = make_uint4(vr[4], vr[5], vr[6], vr[7]);
              }
            }
          }
        } else {
#pragma unroll 1
        for (int c0 = 0; c0 < NH; c0 += 16) {
          // destination of this 16-channel chunk (an up-conv tile may span several parity groups; cout % 32 == 0,
          // so a chunk never straddles two)
          const int gcol = gcol0 + c0;
          const int group = p.upconv ? gcol / p.cout : 0;
          const int nch = gcol - group * p.cout;               // first output channel of the chunk
          int oy = oy_i, ox = ox_i;
          if (p.upconv) { oy = 2 * oy_i + (group >> 1); ox = 2 * ox_i + (group & 1); }
          __nv_bfloat16* d_hi = nullptr;
          float *f_mu = nullptr, *f_var = nullptr;
          if (valid) {
            if (p.dst_f32) {
              const size_t o = (((size_t)ob * p.out_h + oy) * p.out_w + ox) * p.cout + nch;
              f_mu = p.dst_mu + o;
              f_var = p.dst_var + o;
            } else if (HEAD == 0 || p.dst != nullptr) {
              d_hi = p.dst + ((((size_t)ob * p.dh + oy + p.dy0) * p.dw + ox + p.dx0) * 3) * p.dc + p.dc0 + nch;
            }
          }
          if constexpr (TWO_SETS) {
            // mean first, then variance (one 16-bit gate mask in between): half the live registers of the joint form,
            // which matters at the 72 registers per thread the two epilogue sets leave
            uint32_t gate = 0xFFFFu;
            {
              uint32_t am[16];
              if constexpr (KWC) {
                load_kwc16(lane_base + c0, am);
              } else {
                ptx::tmem_ld16(lane_base + c0, am);
                if constexpr (CONCAT) {
                  uint32_t am2[16];
                  ptx::tmem_ld16(lane_base + NT + c0, am2);     // the hi x W_lo half of the mean
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 16; ++j) am[j] = __float_as_uint(__uint_as_float(am[j]) + __uint_as_float(am2[j]));
                } else {
                  ptx::tmem_ld_wait();
                }
              }
              if (p.relu) {
                gate = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float mj = __uint_as_float(am[j]);
                  gate |= (mj > 0.f ? 1u : 0u) << j;
                  am[j] = __float_as_uint(fmaxf(mj, 0.f));
                }
              }
              if (tma_st ? x < p.TWo : valid) {
                if (p.dst_f32) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<uint4*>(f_mu + j) = make_uint4(am[j], am[j + 1], am[j + 2], am[j + 3]);
                } else {
                  uint32_t hi[8], lo[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float a0 = __uint_as_float(am[2 * j]), a1 = __uint_as_float(am[2 * j + 1]);
                    hi[j] = hl_pack2(a0, a1);
                    lo[j] = hl_resid2(hi[j], a0, a1);
                  }
                  if (tma_st) {
                    stage32(0, hi);
                    stage32(1, lo);
                  } else if (p.v8) {
                    ptx::st_global_v8(d_hi, hi);
                    ptx::st_global_v8(d_hi + p.dc, lo);
                  } else {
                    uint4* ph = reinterpret_cast<uint4*>(d_hi);
                    uint4* pl = reinterpret_cast<uint4*>(d_hi + p.dc);
                    ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                  }
                }
              }
            }
            {
              uint32_t av[16];
              if constexpr (KWC) {
                load_kwc16(lane_base + NB + c0, av);
              } else {
                ptx::tmem_ld16(lane_base + (CONCAT ? 2 * NT : NT) + c0, av);
                ptx::tmem_ld_wait();
              }
#pragma unroll
              for (int j4 = 0; j4 < 16; j4 += 4) {
                const float4 s4 = *reinterpret_cast<const float4*>(s_sm + nch + j4);   // warp-uniform: broadcast
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = j4 + e;
                  const float vj = fmaxf(fmaf(sv[e], r, __uint_as_float(av[j])), 0.f);   // all terms >= 0
                  av[j] = (gate >> j) & 1u ? __float_as_uint(vj) : 0u;
                }
              }
              if (tma_st ? x < p.TWo : valid) {
                if (p.dst_f32) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<uint4*>(f_var + j) = make_uint4(av[j], av[j + 1], av[j + 2], av[j + 3]);
                } else {
                  uint32_t vr[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) vr[j] = hl_pack2(__uint_as_float(av[2 * j]), __uint_as_float(av[2 * j + 1]));
                  if (tma_st) {
                    stage32(2, vr);
                  } else if (p.v8) {
                    ptx::st_global_v8(d_hi + 2 * p.dc, vr);
                  } else {
                    uint4* pv = reinterpret_cast<uint4*>(d_hi + 2 * p.dc);
                    pv[0] = make_uint4(vr[0], vr[1], vr[2], vr[3]);
                    pv[1] = make_uint4(vr[4], vr[5], vr[6], vr[7]);
                  }
                }
              }
            }
            if (tma_st) {
              // both TMEM phases of this warp are in registers / staged: release the accumulator stage to the issuer
              // before the store handshake
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc();
              if (!(HL_DBG(p) & 64)) ptx::fence_proxy_async();        // generic-proxy writes -> visible to the TMA
              if (!(HL_DBG(p) & 128)) ptx::named_barrier(pair_bar, 64);
              if (half == 0 && lane == 0 && tc.store && oy_i < p.Ho && ob < p.B && !(HL_DBG(p) & 8)) {
                ptx::tma_store_5d(&maps.d, stg, nt_i * NT, 0, tx * p.TWo, oy_i, ob);
                ptx::bulk_commit_group();
              }
            }
            continue;
          }
          uint32_t am[16], av[16];
          load_acc16(c0, am, av);
          if (c0 + 16 >= NH) early_release();
          float mu[16], var[16];
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 s4 = *reinterpret_cast<const float4*>(s_sm + nch + j4);   // warp-uniform: broadcast
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = j4 + e;
              float mj = __uint_as_float(am[j]);
              float vj = fmaxf(fmaf(sv[e], r, __uint_as_float(av[j])), 0.f);   // all terms >= 0: clamp guards rounding
              if (p.relu) {
                vj = mj > 0.f ? vj : 0.f;
                mj = fmaxf(mj, 0.f);
              }
              mu[j] = mj;
              var[j] = vj;
            }
          }
          if constexpr (HEAD > 0) {
            // ---- fused conv_final + softmax (see the kernel's header comment); NH == 16: the only chunk of this warp
            uint32_t hi[8], lo[8], vr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a0 = mu[2 * j], a1 = mu[2 * j + 1];
              hi[j] = hl_pack2(a0, a1);
              lo[j] = hl_resid2(hi[j], a0, a1);
              vr[j] = hl_pack2(var[2 * j], var[2 * j + 1]);
            }
            if (valid && d_hi != nullptr) {          // optional: the 32-channel tensor itself (packed destination)
              ptx::st_global_v8(d_hi, hi);
              ptx::st_global_v8(d_hi + p.dc, lo);
              ptx::st_global_v8(d_hi + 2 * p.dc, vr);
            }
            // hand-over of the channel 0-15 partial sums: ONE exchange row per pixel, strict ping-pong between the two
            // warps of a quarter (named barriers `full` / `free`), so the first warp runs at most one tile ahead and
            // works on tile t+1 while the second finishes tile t.  (Both have released the accumulator stage already.)
            float* ex = head_ex + row;                                // [k][row]: consecutive lanes, consecutive words
            const uint32_t bar_full = 1u + (uint32_t)q, bar_free = 5u + (uint32_t)q;
            float hm[HEAD], hv[HEAD], hr;
            if (half == 0) {
#pragma unroll
              for (int k = 0; k < HEAD; ++k) hm[k] = hv[k] = 0.f;
              hr = 0.f;
            } else {
              ptx::named_barrier(bar_full, 64);                       // the channel 0-15 partial sums have landed
#pragma unroll
              for (int k = 0; k < HEAD; ++k) { hm[k] = ex[k * HL_BM]; hv[k] = ex[(HEAD + k) * HL_BM]; }
              hr = ex[2 * HEAD * HL_BM];
              if (titer + 1 < n_tiles_cta) {                          // (the last tile has no successor to admit)
                __threadfence_block();
                ptx::named_barrier_arrive(bar_free, 64);
              }
            }
            const float* hw = head_w + (half * NH) * HEAD;
            const float* hw2 = head_w2 + (half * NH) * HEAD;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // the values final_conv_softmax_kernel would read back: mean = hi + lo, variance = bf16
              float m0, m1;
              hl_sum2(hi[j], lo[j], m0, m1);
              head_accumulate<HEAD>(m0, hl_lo(vr[j]), hw + (2 * j) * HEAD, hw2 + (2 * j) * HEAD, hm, hv, hr);
              head_accumulate<HEAD>(m1, hl_hi(vr[j]), hw + (2 * j + 1) * HEAD, hw2 + (2 * j + 1) * HEAD, hm, hv, hr);
            }
            if (half == 0) {
              if (titer > 0) ptx::named_barrier(bar_free, 64);        // the previous tile's sums have been taken
#pragma unroll
              for (int k = 0; k < HEAD; ++k) { ex[k * HL_BM] = hm[k]; ex[(HEAD + k) * HL_BM] = hv[k]; }
              ex[2 * HEAD * HL_BM] = hr;
              __threadfence_block();
              ptx::named_barrier_arrive(bar_full, 64);
            } else {
              float pj[HEAD], vo[HEAD];
              head_finish<HEAD>(hm, hv, hr, head_s, pj, vo);
              if (valid) {
                const size_t i = ((size_t)ob * p.Ho + oy_i) * p.Wo + ox_i;
                if constexpr (HEAD == 4) {
                  reinterpret_cast<float4*>(p.head_p)[i] = make_float4(pj[0], pj[1], pj[2], pj[3]);
                  reinterpret_cast<float4*>(p.head_v)[i] = make_float4(vo[0], vo[1], vo[2], vo[3]);
                  if (p.head_pre_mu) {
                    reinterpret_cast<float4*>(p.head_pre_mu)[i] = make_float4(hm[0], hm[1], hm[2], hm[3]);
                    reinterpret_cast<float4*>(p.head_pre_var)[i] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                  }
                } else {
#pragma unroll
                  for (int a = 0; a < HEAD; ++a) { p.head_p[i * HEAD + a] = pj[a]; p.head_v[i * HEAD + a] = vo[a]; }
                  if (p.head_pre_mu) {
#pragma unroll
                    for (int a = 0; a < HEAD; ++a) { p.head_pre_mu[i * HEAD + a] = hm[a]; p.head_pre_var[i * HEAD + a] = hv[a]; }
                  }
                }
              }
            }
            continue;
          }
          if (valid) {
            if (p.dst_f32) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<float4*>(f_mu + j) = make_float4(mu[j], mu[j + 1], mu[j + 2], mu[j + 3]);
                *reinterpret_cast<float4*>(f_var + j) = make_float4(var[j], var[j + 1], var[j + 2], var[j + 3]);
              }
            } else {
              uint32_t hi[8], lo[8], vr[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a0 = mu[2 * j], a1 = mu[2 * j + 1];
                hi[j] = hl_pack2(a0, a1);                                   // one packed convert
                lo[j] = hl_resid2(hi[j], a0, a1);
                vr[j] = hl_pack2(var[2 * j], var[2 * j + 1]);
              }
              if (p.v8) {
                ptx::st_global_v8(d_hi, hi);
                ptx::st_global_v8(d_hi + p.dc, lo);
                ptx::st_global_v8(d_hi + 2 * p.dc, vr);
              } else {
                uint4* ph = reinterpret_cast<uint4*>(d_hi);
                uint4* pl = reinterpret_cast<uint4*>(d_hi + p.dc);
                uint4* pv = reinterpret_cast<uint4*>(d_hi + 2 * p.dc);
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
        }
        if (!tma_st && !EARLY_REL) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc();
        }
      }
      if (KWC && !DGRAD && p.tma_store && half == 0 && lane == 0) ptx::bulk_wait_read0();   // staging must outlive the reads
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) {
    ptx::cluster_sync_all();          // the peer's shared memory / TMEM must outlive the leader's last UMMA and remote arrive
    if (warp == 1) ptx::tmem_dealloc2(tmem_base, TMEM_COLS);
  } else {
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*HlEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static HlEncodeTiledFn hl_encode_tiled() {
  static HlEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<HlEncodeTiledFn>(f);
  });
  return fn;
}

// cuTensorMapEncodeTiled for the other translation units (sn_packed.cu: the first convolution's TMA store); nullptr when
// the driver entry point is unavailable
void* tensor_map_encoder() { return reinterpret_cast<void*>(hl_encode_tiled()); }

struct HaloTiling {
  int R, THo, THb, TN, TWo, tiles_x, tiles_y, tiles_b;
  double eff;
};

// Pick the halo box (R columns x THb rows x TN images, <= 128 GEMM rows of which THo*TWo*TN are useful) that
// wastes the fewest GEMM rows over the whole layer.
static HaloTiling choose_tiling(int batch, int in_h, int in_w, int k) {
  const int Ho = in_h - k + 1, Wo = in_w - k + 1;
  HaloTiling best{};
  best.eff = -1.0;
  // whole (small) images, several per tile
  if (in_h * in_w <= HL_BM && in_w <= 62) {
    HaloTiling t{};
    t.R = in_w; t.THb = in_h; t.THo = Ho; t.TWo = Wo;
    t.TN = HL_BM / (in_h * in_w);
    if (t.TN > batch) t.TN = batch;
    t.tiles_x = t.tiles_y = 1;
    t.tiles_b = (batch + t.TN - 1) / t.TN;
    t.eff = (double)batch * Ho * Wo / ((double)t.tiles_b * HL_BM);
    best = t;
  }
  const int rmax = in_w < 62 ? in_w : 62;
  for (int R = k; R <= rmax; ++R) {
    HaloTiling t{};
    t.R = R; t.TWo = R - (k - 1); t.TN = 1;
    t.THo = HL_BM / R;
    if (t.THo > Ho) t.THo = Ho;
    if (t.THo < 1) continue;
    t.THb = t.THo + k - 1;
    t.tiles_x = (Wo + t.TWo - 1) / t.TWo;
    t.tiles_y = (Ho + t.THo - 1) / t.THo;
    t.tiles_b = batch;
    t.eff = (double)Ho * Wo / ((double)t.tiles_x * t.tiles_y * HL_BM);
    if (t.eff > best.eff + 1e-9) best = t;
  }
  return best;
}

// kw-concatenated variant: the halo box is exactly 32 columns wide (30 output columns), so that the epilogue's +1 / +2
// row shifts stay inside one 32-lane TMEM quarter; a tile is 4 output rows (quarter = row).
static HaloTiling kwc_tiling(int batch, int in_h, int in_w) {
  const int Ho = in_h - 2, Wo = in_w - 2;
  HaloTiling t{};
  t.R = 32; t.TWo = 30; t.TN = 1;
  t.THo = Ho < 4 ? Ho : 4;
  t.THb = t.THo + 2;
  t.tiles_x = (Wo + t.TWo - 1) / t.TWo;
  t.tiles_y = (Ho + t.THo - 1) / t.THo;
  t.tiles_b = batch;
  t.eff = (double)Ho * Wo / ((double)t.tiles_x * t.tiles_y * HL_BM);
  return t;
}

// Use it for 32-column tiles of 3x3 convs with >= 2 channel blocks, when the fixed 30-column tile width does not waste
// much more of the GEMM rows than the free choice would.  Measured at batch 64 (profiles/r02_kwc_knobs.md): the UMMA pipe
// time per tile halves (83 % -> 46 % busy), but the kernel then runs into the CUDA-core side -- ~10 k warp instructions
// per 128-pixel tile between the epilogue (64 extra shuffles per thread) and the q-reduction warps, LSU data pipe at
// 64-69 % -- so with ONE channel block (conv1, up4_conv2: 24 UMMAs per tile) it only ties with the tap-shift kernel
// (0.263 vs 0.252 ms), while with two (up4_conv1: 48 vs 108 UMMAs) it wins (0.355 vs 0.44 ms).  SN_KWC=2 forces it for
// every eligible layer, SN_KWC=0 disables it.
static bool kwc_wanted(int flags, int nt, int keff, int in_w, int cblk, const HaloTiling& free_choice,
                       const HaloTiling& kwc) {
  static const int env_mode = [] {
    const char* e = getenv("SN_KWC");
    return e == nullptr ? 1 : atoi(e);
  }();
  const int mode = (flags & SN_TC_NO_KWC) ? 0 : ((flags & SN_TC_KWC) ? 2 : env_mode);
  if (mode == 0 || nt != 32 || keff != 3 || in_w < 32) return false;
  if (mode == 2) return true;
  return cblk >= 2 && kwc.eff >= 0.65 * free_choice.eff;
}

// Tensor map of one plane of a packed window.  `step` = 2 with origin (oy, ox) addresses the pixels (2y+oy, 2x+ox):
// the four parity views the data gradient of an up-conv reads.  Everything outside the window reads as zero
// (TMA out-of-bounds fill), which is how the data gradient gets its (k-1)-wide zero border.
static int hl_make_act_map(CUtensorMap* out, const sn_packed_view& v, int plane, int src_c, int batch, int in_h,
                           int in_w, const HaloTiling& t, int step = 1, int oy = 0, int ox = 0) {
  const size_t pix = (size_t)3 * v.c;
  char* base = reinterpret_cast<char*>(v.base) +
               ((((size_t)(v.y0 + oy) * v.w + v.x0 + ox) * 3 + plane) * v.c + v.c0) * sizeof(__nv_bfloat16);
  cuuint64_t dims[4] = {(cuuint64_t)src_c, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)batch};
  cuuint64_t strides[3] = {pix * 2 * step, (cuuint64_t)v.w * pix * 2 * step, (cuuint64_t)v.h * v.w * pix * 2};
  cuuint32_t box[4] = {HL_KC, (cuuint32_t)t.R, (cuuint32_t)t.THb, (cuuint32_t)t.TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = hl_encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeTiled(activation) failed (%d)", (int)r);
  return SN_OK;
}

// Destination window as a 5-D tensor (channel, plane, x, y, image) for the TMA-store epilogue: the box is one tile row
// (30 pixels x 3 planes x 32 channels, 64-byte swizzled in shared memory); whatever lies outside the window is clipped.
static int hl_make_dst_map(CUtensorMap* out, const sn_packed_view& v, int cout, int out_h, int out_w, int batch,
                           int box_w) {
  const size_t pix = (size_t)3 * v.c * 2;
  char* base = reinterpret_cast<char*>(v.base) + ((((size_t)v.y0 * v.w + v.x0) * 3) * v.c + v.c0) * sizeof(__nv_bfloat16);
  cuuint64_t dims[5] = {(cuuint64_t)cout, 3, (cuuint64_t)out_w, (cuuint64_t)out_h, (cuuint64_t)batch};
  cuuint64_t strides[4] = {(cuuint64_t)v.c * 2, pix, (cuuint64_t)v.w * pix, (cuuint64_t)v.h * v.w * pix};
  cuuint32_t box[5] = {32, 3, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = hl_encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeTiled(destination) failed (%d)", (int)r);
  return SN_OK;
}

// Prepared weights [3 planes][taps][n][cin] as a 4-D tensor (cin, n, tap, plane); one box = the three planes of
// `box_taps` consecutive taps x `box_n` rows x 32 channels.
static int hl_make_weight_map(CUtensorMap* out, const void* w_packed, int taps, int cout, int cin, int box_n,
                              int box_taps = 1, int box_planes = 3) {
  cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps, 3};
  cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * cout * 2, (cuuint64_t)taps * cin * cout * 2};
  cuuint32_t box[4] = {HL_KC, (cuuint32_t)box_n, (cuuint32_t)box_taps, (cuuint32_t)box_planes};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = hl_encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(w_packed), dims, strides,
                                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SN_ERR_DRIVER, "cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
  return SN_OK;
}

template <int NT, int KS, bool RESIDENT, bool DGRAD, int G = 1, bool KWC = false, bool CTA2 = false, int HEAD = 0>
static int hl_launch3(const HlMaps& maps, const HlP& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_moments_halo_kernel<NT, KS, RESIDENT, DGRAD, G, KWC, CTA2, HEAD>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM);
  });
  if (attr_err != cudaSuccess) return fail(SN_ERR_LAUNCH, "conv_halo: cannot reserve %d B of shared memory", HL_SMEM);
  const int units = (G == 1 && !CTA2) ? p.total_tiles : p.total_groups;
  int grid = units < num_sms() ? units : num_sms();
  if (CTA2) {
    // one cluster of two CTAs per tile pair; the persistent walk assumes every cluster is resident at once, so the grid
    // is what the device can co-schedule (GPC boundaries can leave it below num_sms / 2)
    static int max_clusters = 0;
    static std::once_flag once_c;
    std::call_once(once_c, [] {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3((unsigned)(num_sms() & ~1));
      q.blockDim = dim3(hl_threads(hl_two_sets(KWC, G, DGRAD)));
      q.dynamicSmemBytes = HL_SMEM;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = 2;
      a[0].val.clusterDim.y = 1;
      a[0].val.clusterDim.z = 1;
      q.attrs = a;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_moments_halo_kernel<NT, KS, RESIDENT, DGRAD, G, KWC, CTA2, HEAD>, &q) ==
              cudaSuccess && n > 0)
        max_clusters = n;
      else
        max_clusters = num_sms() / 2;
      if (getenv("SN_CTA2_VERBOSE")) fprintf(stderr, "conv_halo CTA pairs: %d clusters co-resident\n", max_clusters);
    });
    grid = 2 * (units < max_clusters ? units : max_clusters);
  }
  static const bool pdl = [] {
    const char* e = getenv("SN_PDL");
    return e == nullptr || e[0] != '0';
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(hl_threads(hl_two_sets(KWC, G, DGRAD)));
  cfg.dynamicSmemBytes = HL_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
#ifdef SN_HL_KNOBS
  static const bool fake_cluster = getenv("SN_FAKE_CLUSTER") != nullptr;      // single-CTA kernel launched as clusters of 2
#else
  constexpr bool fake_cluster = false;
#endif
  if (CTA2 || (fake_cluster && grid % 2 == 0)) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_moments_halo_kernel<NT, KS, RESIDENT, DGRAD, G, KWC, CTA2, HEAD>, maps, p);
  if (e != cudaSuccess) return fail(SN_ERR_LAUNCH, "conv_halo launch: %s", cudaGetErrorString(e));
  return check_launch(DGRAD ? "conv_moments_halo_dgrad" : "conv_moments_halo");
}

template <int NT>
static int hl_launch(const HlMaps& maps, const HlP& p, cudaStream_t st) {
  if constexpr (NT == 128) {
    if (p.cta2) {
      switch (p.ksize) {
        case 1: return hl_launch3<128, 1, false, false, 1, false, true>(maps, p, st);
        case 2: return hl_launch3<128, 2, false, false, 1, false, true>(maps, p, st);
        default: return hl_launch3<128, 3, false, false, 1, false, true>(maps, p, st);
      }
    }
  }
  if constexpr (NT == 64) {
    if (p.cta2) return hl_launch3<64, 3, true, false, 1, false, true>(maps, p, st);   // k = 3, resident half-slots (hl_plan)
  }
  if constexpr (NT == 32) {
    if (p.head_labels) {               // fused conv_final + softmax (hl_head_fusable() has checked the variant)
      switch (p.head_labels) {
        case 2: return hl_launch3<32, 3, true, false, 1, false, false, 2>(maps, p, st);
        case 3: return hl_launch3<32, 3, true, false, 1, false, false, 3>(maps, p, st);
        case 4: return hl_launch3<32, 3, true, false, 1, false, false, 4>(maps, p, st);
        default: return hl_launch3<32, 3, true, false, 1, false, false, 5>(maps, p, st);
      }
    }
    if (p.kwc) {
      return p.b_resident ? hl_launch3<32, 3, true, false, 1, true>(maps, p, st)
                          : hl_launch3<32, 3, false, false, 1, true>(maps, p, st);
    }
  }
  if (p.total_groups > 0) {            // two pixel tiles per weight slot (streamed weights, k = 3 or 1)
    if (p.ksize == 3) return hl_launch3<NT, 3, false, false, 2>(maps, p, st);
    if (p.ksize == 1) return hl_launch3<NT, 1, false, false, 2>(maps, p, st);
  }
  switch (p.ksize * 2 + (p.b_resident ? 1 : 0)) {
    case 2: return hl_launch3<NT, 1, false, false>(maps, p, st);
    case 3: return hl_launch3<NT, 1, true, false>(maps, p, st);
    case 4: return hl_launch3<NT, 2, false, false>(maps, p, st);
    case 5: return hl_launch3<NT, 2, true, false>(maps, p, st);
    case 6: return hl_launch3<NT, 3, false, false>(maps, p, st);
    default: return hl_launch3<NT, 3, true, false>(maps, p, st);
  }
}

// The data gradient only ever needs k = 3 / 2 (regular convs) and k = 1 (1x1 convs and the up-conv, whose four
// output parities become four K-concatenated sources).
template <int NT>
static int hl_launch_dgrad(const HlMaps& maps, const HlP& p, cudaStream_t st) {
  if constexpr (NT == 128) {
    if (p.cta2) {
      switch (p.ksize) {
        case 1: return hl_launch3<128, 1, false, true, 1, false, true>(maps, p, st);
        case 2: return hl_launch3<128, 2, false, true, 1, false, true>(maps, p, st);
        default: return hl_launch3<128, 3, false, true, 1, false, true>(maps, p, st);
      }
    }
  }
  if constexpr (NT == 64) {
    if (p.cta2) return hl_launch3<64, 3, true, true, 1, false, true>(maps, p, st);
  }
  if constexpr (NT == 32) {
    if (p.kwc) {
      return p.b_resident ? hl_launch3<32, 3, true, true, 1, true>(maps, p, st)
                          : hl_launch3<32, 3, false, true, 1, true>(maps, p, st);
    }
  }
  if (p.total_groups > 0) {
    if (p.ksize == 3) return hl_launch3<NT, 3, false, true, 2>(maps, p, st);
    if (p.ksize == 1) return hl_launch3<NT, 1, false, true, 2>(maps, p, st);
  }
  switch (p.ksize * 2 + (p.b_resident ? 1 : 0)) {
    case 2: return hl_launch3<NT, 1, false, true>(maps, p, st);
    case 3: return hl_launch3<NT, 1, true, true>(maps, p, st);
    case 4: return hl_launch3<NT, 2, false, true>(maps, p, st);
    case 5: return hl_launch3<NT, 2, true, true>(maps, p, st);
    case 6: return hl_launch3<NT, 3, false, true>(maps, p, st);
    default: return hl_launch3<NT, 3, true, true>(maps, p, st);
  }
}

// 256-bit epilogue accesses need every (pixel, plane, 16-channel chunk) segment 32-byte aligned
static bool hl_view_v8(const sn_packed_view& v) {
  return (reinterpret_cast<uintptr_t>(v.base) & 31u) == 0 && v.c % 16 == 0 && v.c0 % 16 == 0;
}

// Tile geometry + shared-memory plan shared by the forward and the data-gradient dispatch.
static int hl_plan(HlP& p, const HaloTiling& t, int keff, int ncols, int nt, int cblk, bool kwc = false,
                   bool tma_store = false, int cta2_mode = 1, bool head = false) {
  p.kwc = kwc ? 1 : 0;
  p.tma_store = tma_store ? 1 : 0;
  p.cta2 = 0;
  {
    static const int dbg = [] {
      const char* e = getenv("SN_HL_DBG");
      return e ? atoi(e) : 0;
    }();
    p.dbg = dbg;
  }
  p.tiles_x = t.tiles_x; p.tiles_y = t.tiles_y; p.tiles_b = t.tiles_b;
  p.tiles_n = ncols / nt;
  const long long total = (long long)p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_n;
  SN_REQUIRE(total < (1ll << 30), SN_ERR_UNSUPPORTED, "conv_halo: too many tiles");
  p.total_tiles = (int)total;
  p.TWo = t.TWo; p.THo = t.THo; p.R = t.R; p.THb = t.THb; p.TN = t.TN;
  p.rows_box = t.R * t.THb * t.TN;
  // rows the shifted A operands reach: taps shift by kh*R + kw (KWC: by kh*R only, the kw shift is in the epilogue)
  const int rows_alloc = HL_BM + (keff - 1) * (kwc ? t.R : t.R + 1);
  const int rows_need = rows_alloc > p.rows_box ? rows_alloc : p.rows_box;
  SN_REQUIRE(rows_need <= 256, SN_ERR_UNSUPPORTED, "conv_halo: halo tile of %d rows", rows_need);
  p.a_plane = ((rows_need * 64 + 1023) / 1024) * 1024;
  p.ksize = keff;
  // shared-memory plan: [A stages][B slots][1 KB barriers][2 KB q buffers][2 KB s], 1 KB alignment slack
  // (+ 8 staging buffers behind s when the epilogue stores through TMA, or the fused head's 16 KB)
  const int avail = HL_SMEM - 1024 - 1024 - 2048 - 2048 - (tma_store ? 8 * HL_STG_BUF : 0) - (head ? HL_HEAD_SMEM : 0);
  const int a_stage = 3 * p.a_plane;
  const int b_slot = 3 * (kwc ? keff * nt : nt) * HL_KC * 2;
  const int resident_slots = cblk * (kwc ? keff : keff * keff);
  if (p.tiles_n == 1 && resident_slots <= HL_MAX_BSLOTS && resident_slots * b_slot + 2 * a_stage <= avail) {
    p.b_resident = 1;
    p.sb = resident_slots;
    p.sa = (avail - resident_slots * b_slot) / a_stage;
  } else {
    p.b_resident = 0;
    p.sa = 3;
    if (3 * a_stage + 3 * b_slot > avail) p.sa = 2;
    p.sb = (avail - p.sa * a_stage) / b_slot;
    if (p.sb > HL_MAX_BSLOTS) p.sb = HL_MAX_BSLOTS;
    SN_REQUIRE(p.sb >= 2, SN_ERR_UNSUPPORTED, "conv_halo: shared memory plan failed");
  }
  if (p.sa > HL_MAX_ASTAGES) p.sa = HL_MAX_ASTAGES;
  // streamed weights: let every weight slot serve two pixel tiles (G = 2) when there are enough tile pairs to keep
  // every SM busy and the doubled A stages still leave >= 3 weight slots
  p.pix_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  p.total_groups = 0;
  // CTA pairs for the 128-column layers with streamed weights: two pixel tiles of one N tile on the two SMs of a TPC,
  // M = 256 UMMAs, half of every weight slot per SM.  Needs enough tile pairs to fill the clusters (74 on B200).
  // cta2_mode: 0 never, 1 the library's choice, 2 whenever the shape allows (tests)
  {
    static const int env_mode = [] {
      const char* e = getenv("SN_CTA2");
      return e == nullptr ? 1 : atoi(e);
    }();
    const int mode = cta2_mode == 1 ? env_mode : cta2_mode;
    const long long pairs = (long long)p.tiles_n * ((p.pix_tiles + 1) / 2);
    // (not for k = 1 / up-convs by default: one tap per channel block is too little UMMA work per handshake -- measured
    //  +15...20 % on the three 2x2 up-convs -- while the 3x3 layers gain 3...12 %, profiles/r02_cta2.md)
    // 64-column tiles (k = 3, tiles_n == 1: conv2/3, up3_conv2, up4_conv1's data gradient ...): a pair keeps HALF of every
    // weight slot per SM -- 8 KB instead of 12 (the X / Y / Z regions of the kernel) -- so layers whose weights do not fit
    // one SM (conv3, up3_conv2: 18 slots) become resident: no weight stream at all and, unlike two tiles per weight slot
    // (G = 2), every tile keeps its own double-buffered accumulator stage.  conv3 at batch 64: 0.216 -> 0.161 ms.
    // Streamed half-slots (up3_conv1: 36 slots) measured SLOWER than G = 2 (0.353 vs 0.294 ms) and are not built.
    // SN_CTA2_64: 0 never, 1 (default) layers whose weights do not fit one SM, 2 every eligible layer.
    {
      static const int mode64_env = [] {
        const char* e = getenv("SN_CTA2_64");
        return e == nullptr ? 1 : atoi(e);
      }();
      const int slot_pair = 2 * nt * HL_KC * 2;
      // Activation planes may start on multiples of 512 B here (the period of the SWIZZLE_64B pattern: address bits
      // [4,6) ^= bits [7,9), so absolute and tile-relative swizzles agree) instead of 1 KB: that is what lets up3_conv2
      // (18 slots of 8 KB + two stages of 3 x 194 rows) fit.  (512 B everywhere was measured: conv1 +9 %, others +-3 %.)
      const int a_plane512 = ((rows_need * 64 + 511) / 512) * 512;
      int a_stage_p = a_stage;
      if (resident_slots * slot_pair + 2 * a_stage_p > avail) a_stage_p = 3 * a_plane512;
      if (mode != 0 && mode64_env != 0 && nt == 64 && keff == 3 && p.tiles_n == 1 && !kwc &&
          (mode == 2 || ((!p.b_resident || mode64_env == 2) && 4 * pairs >= 3 * (num_sms() / 2))) &&
          resident_slots <= HL_MAX_BSLOTS && resident_slots * slot_pair + 2 * a_stage_p <= avail) {
        p.cta2 = 1;
        p.b_resident = 1;
        p.a_plane = a_stage_p / 3;
        p.sb = resident_slots;
        p.sa = (avail - resident_slots * slot_pair) / a_stage_p;
        if (p.sa > HL_MAX_ASTAGES) p.sa = HL_MAX_ASTAGES;
        p.total_groups = (int)pairs;
        return SN_OK;
      }
    }
    if (mode != 0 && nt == 128 && !p.b_resident && !kwc &&
        (mode == 2 || (keff >= 2 && 4 * pairs >= 3 * (num_sms() / 2)))) {
      const int b_half = b_slot / 2;
      int sa = 3;
      if (3 * a_stage + 4 * b_half > avail) sa = 2;
      int sb = (avail - sa * a_stage) / b_half;
      if (sb > HL_CTA2_MAX_BSLOTS) sb = HL_CTA2_MAX_BSLOTS;
#ifdef SN_HL_KNOBS
      if (const char* e = getenv("SN_CTA2_SA")) { sa = atoi(e); sb = (avail - sa * a_stage) / b_half; if (sb > HL_CTA2_MAX_BSLOTS) sb = HL_CTA2_MAX_BSLOTS; }
      if (const char* e = getenv("SN_CTA2_SB")) { if (atoi(e) < sb) sb = atoi(e); }
#endif
      if (sb >= 3) {
        p.cta2 = 1;
        p.sa = sa;
        p.sb = sb;
        p.total_groups = (int)pairs;
        return SN_OK;
      }
    }
  }
  static const int dual = [] {          // 0: off, 1: 64-column tiles (default), 2: 128-column tiles too (A/B)
    const char* e = getenv("SN_DUAL");
    return e == nullptr ? 1 : atoi(e);
  }();
  // (measured at batch 64: +11 % on the 64-column layers conv3 / up3_conv1, whose UMMA pipe idled waiting for weights;
  //  -10..25 % on the 128-column layers, where the pair's two epilogues can no longer hide behind the next tile's
  //  UMMAs -- so only N tiles of 64 columns use it)
  if (dual && (nt == 64 || (dual >= 2 && nt == 128)) && !p.b_resident && (keff == 3 || keff == 1)) {
    const long long groups = (long long)p.tiles_n * ((p.pix_tiles + 1) / 2);
    const int a_stage2 = 2 * a_stage;
    const int sb2 = (avail - 2 * a_stage2) / b_slot;
    if (groups >= num_sms() && sb2 >= 3) {
      p.total_groups = (int)groups;
      p.sa = 2;
      p.sb = sb2 > HL_MAX_BSLOTS ? HL_MAX_BSLOTS : sb2;
    }
  }
  return SN_OK;
}

// Which layers can end in the fused head: 3x3, ReLU, one 32-channel source, 32 output channels, packed (or no)
// destination, 2-5 labels -- the shape of up4_conv2 -> conv_final in both reference graphs (Brats.py:453-455).
static bool hl_head_shape_ok(const sn_tc_conv_desc* d, int n_labels) {
  return d->ksize == 3 && d->cout == 32 && d->src_c[0] == 32 && d->src_c[1] == 0 && n_labels >= 2 && n_labels <= 5 &&
         (d->flags & SN_TC_RELU) && !(d->flags & (SN_TC_UPCONV | SN_TC_DST_F32 | SN_TC_IM2COL | SN_TC_KWC)) &&
         d->rsum_out == nullptr;
}

// Called by sn_conv_moments_fwd_tc / sn_conv_moments_fwd_tc_head (sn_tc_conv.cu) after argument validation.
int conv_moments_halo_dispatch(const sn_tc_conv_desc* d, cudaStream_t stream, const sn_tc_head_desc* head) {
  SN_REQUIRE(hl_encode_tiled() != nullptr, SN_ERR_DRIVER, "conv_halo: cuTensorMapEncodeTiled unavailable");
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  const bool dst_f32 = (d->flags & SN_TC_DST_F32) != 0;
  const int keff = upconv ? 1 : d->ksize;
  const int Ho = d->in_h - keff + 1, Wo = d->in_w - keff + 1;
  const int out_h = upconv ? 2 * d->in_h : Ho, out_w = upconv ? 2 * d->in_w : Wo;
  const int groups = upconv ? 4 : 1;
  const int ncols = groups * d->cout;               // GEMM N: an up-conv tile may cover several parity groups
  const int nt = ncols % 128 == 0 ? 128 : (ncols % 64 == 0 ? 64 : 32);
  const int taps_w = upconv ? 4 : d->ksize * d->ksize;
  const int cin = d->src_c[0] + d->src_c[1];
  const int cblk = cin / HL_KC;

  HaloTiling t = choose_tiling(d->batch, d->in_h, d->in_w, keff);
  SN_REQUIRE(t.eff > 0, SN_ERR_UNSUPPORTED, "conv_halo: no tiling for %dx%d k=%d", d->in_h, d->in_w, keff);
  SN_REQUIRE(d->cout <= 512, SN_ERR_UNSUPPORTED, "conv_halo: cout %d > 512", d->cout);
  bool kwc = false;
  if (head != nullptr) {
    SN_REQUIRE(hl_head_shape_ok(d, head->n_labels), SN_ERR_UNSUPPORTED,
               "conv_tc_head: needs a 3x3 ReLU conv 32 -> 32 on one packed source and 2..5 labels (sn_tc_head_fusable)");
  } else if (nt == 32 && keff == 3 && !upconv) {
    const HaloTiling tk = kwc_tiling(d->batch, d->in_h, d->in_w);
    if (kwc_wanted(d->flags, nt, keff, d->in_w, cblk, t, tk)) { t = tk; kwc = true; }
  }

  static const bool tma_store_on = [] {
    const char* e = getenv("SN_TMA_STORE");
    return e == nullptr || e[0] != '0';
  }();
  bool tma_store = kwc && !dst_f32 && tma_store_on;
  HlP p{};
  int rc;
  if (tma_store) {
    // the 48 KB of staging must leave room for resident weights and >= 2 activation stages
    HlP probe{};
    if ((rc = hl_plan(probe, t, keff, ncols, nt, cblk, kwc, true)) || !probe.b_resident || probe.sa < 2) tma_store = false;
  }
  const int cta2_mode = (d->flags & SN_TC_NO_CTA2) ? 0 : ((d->flags & SN_TC_CTA2) ? 2 : 1);
  if ((rc = hl_plan(p, t, keff, ncols, nt, cblk, kwc, tma_store, cta2_mode, head != nullptr))) return rc;
  if (head != nullptr) {
    SN_REQUIRE(p.b_resident && p.sa >= 2, SN_ERR_UNSUPPORTED, "conv_tc_head: shared-memory plan failed");
    p.head_labels = head->n_labels;
    p.head_w = head->w_mu; p.head_ws = head->w_sigma;
    p.head_p = head->p_out; p.head_v = head->var_out;
    p.head_pre_mu = head->presoftmax_mu; p.head_pre_var = head->presoftmax_var;
  }
  p.taps_w = taps_w;
  p.cblk_s[0] = d->src_c[0] / HL_KC; p.cblk_s[1] = d->src_c[1] / HL_KC;
  p.Ho = Ho; p.Wo = Wo; p.B = d->batch;
  p.cout = d->cout;
  p.relu = (d->flags & SN_TC_RELU) ? 1 : 0; p.upconv = upconv ? 1 : 0; p.dst_f32 = dst_f32 ? 1 : 0;
  p.dst = reinterpret_cast<__nv_bfloat16*>(d->dst.base);
  p.dh = d->dst.h; p.dw = d->dst.w; p.dc = d->dst.c; p.dy0 = d->dst.y0; p.dx0 = d->dst.x0; p.dc0 = d->dst.c0;
  p.dst_mu = d->dst_mu; p.dst_var = d->dst_var; p.out_h = out_h; p.out_w = out_w;
  p.s = d->s; p.s_len = d->cout;
  p.r_out = d->rsum_out;
  static const bool use_v8 = [] {
    const char* e = getenv("SN_V8");
    return e == nullptr || e[0] != '0';
  }();
  p.v8 = use_v8 && !dst_f32 && hl_view_v8(d->dst) ? 1 : 0;
  if (head != nullptr && p.dst != nullptr)
    SN_REQUIRE(p.v8, SN_ERR_UNSUPPORTED, "conv_tc_head: the optional packed destination must be 32-byte aligned");

  HlMaps maps;
  for (int s = 0; s < 2; ++s) {
    const int srcs = d->src_c[s] ? s : 0;
    for (int pl = 0; pl < 3; ++pl)
      if ((rc = hl_make_act_map(&maps.a[s][pl], d->src[srcs], pl, d->src_c[srcs], d->batch, d->in_h, d->in_w, t)))
        return rc;
  }
  for (int pl = 0; pl < 3; ++pl) maps.a[2][pl] = maps.a[3][pl] = maps.a[0][pl];     // unused
  // regular: [3*taps][cout][cin]; up-conv: [3][4*cout][cin] (same memory, parity and channel fused into one axis)
  const bool pair64 = p.cta2 && nt == 64;       // X region: NT rows of ONE plane; maps.d: NT/2 rows of one plane (Y, Z)
  if ((rc = hl_make_weight_map(&maps.w, d->w_packed, upconv ? 1 : taps_w, upconv ? ncols : d->cout, cin,
                               pair64 ? nt : (p.cta2 ? nt / 2 : nt), p.kwc ? 3 : 1, pair64 ? 1 : 3)))
    return rc;
  maps.d = maps.w;
  if (pair64 && (rc = hl_make_weight_map(&maps.d, d->w_packed, taps_w, d->cout, cin, nt / 2, 1, 1))) return rc;
  if (p.tma_store && (rc = hl_make_dst_map(&maps.d, d->dst, d->cout, out_h, out_w, d->batch, t.TWo))) return rc;
  switch (nt) {
    case 128: return hl_launch<128>(maps, p, stream);
    case 64: return hl_launch<64>(maps, p, stream);
    default: return hl_launch<32>(maps, p, stream);
  }
}

static HlView hl_view(const sn_packed_view& v) {
  return HlView{reinterpret_cast<__nv_bfloat16*>(v.base), v.h, v.w, v.c, v.y0, v.x0, v.c0};
}

// Called by sn_conv_moments_bwd_data_tc (sn_tc_conv.cu) after argument validation.  The data gradient of a VALID
// k x k moment conv is the same halo GEMM run over the gradient w.r.t. the conv's output, zero-extended by k-1
// (TMA out-of-bounds fill), against the flipped / transposed weights (sn_prepare_weights_bwd):
//   g_mu = g_mu' (*)^T W + 2 mu box^T(t),  g_var = g_var' (*)^T W^2 + box^T(t),  t = sum_n g_var'_n s_n.
// The up-conv (unpool + 2x2 conv) is pointwise per 2x2 output block: a 1x1 GEMM whose K axis concatenates the four
// parity views (2y+a, 2x+b) of the output gradient.
int conv_moments_halo_dgrad_dispatch(const sn_tc_dgrad_desc* d, cudaStream_t stream) {
  SN_REQUIRE(hl_encode_tiled() != nullptr, SN_ERR_DRIVER, "conv_halo: cuTensorMapEncodeTiled unavailable");
  const bool upconv = (d->flags & SN_TC_UPCONV) != 0;
  const int keff = upconv ? 1 : d->ksize;
  const int pad = keff - 1;
  const int Ho_f = d->in_h - keff + 1, Wo_f = d->in_w - keff + 1;     // grid of one gradient source
  const int gh = Ho_f + 2 * pad, gw = Wo_f + 2 * pad;                  // zero-extended gradient = the GEMM's "input"
  const int ncols = d->in_c[0] + d->in_c[1];                           // GEMM N = the forward's input channels
  const int nt = ncols % 128 == 0 ? 128 : (ncols % 64 == 0 ? 64 : 32);
  const int nsrc = upconv ? 4 : 1;
  const int cblk = nsrc * d->cout / HL_KC;

  HaloTiling t = choose_tiling(d->batch, gh, gw, keff);
  SN_REQUIRE(t.eff > 0, SN_ERR_UNSUPPORTED, "conv_halo dgrad: no tiling for %dx%d k=%d", gh, gw, keff);
  SN_REQUIRE(d->cout <= 512, SN_ERR_UNSUPPORTED, "conv_halo dgrad: cout %d > 512", d->cout);
  bool kwc = false;
  if (nt == 32 && keff == 3 && !upconv) {
    const HaloTiling tk = kwc_tiling(d->batch, gh, gw);
    if (kwc_wanted(d->flags, nt, keff, gw, cblk, t, tk)) { t = tk; kwc = true; }
  }

  HlP p{};
  int rc;
  const int cta2_mode = (d->flags & SN_TC_NO_CTA2) ? 0 : ((d->flags & SN_TC_CTA2) ? 2 : 1);
  if ((rc = hl_plan(p, t, keff, ncols, nt, cblk, kwc, false, cta2_mode))) return rc;
  p.taps_w = keff * keff;
  for (int s = 0; s < nsrc; ++s) p.cblk_s[s] = d->cout / HL_KC;
  p.pad = pad;
  p.Ho = d->in_h; p.Wo = d->in_w; p.B = d->batch;
  p.cout = ncols;
  p.s = d->s; p.s_len = d->cout;
  p.csplit = d->in_c[1] ? d->in_c[0] : ncols;
  p.v8 = 1;
  {
    const char* e = getenv("SN_V8");
    if (e != nullptr && e[0] == '0') p.v8 = 0;
  }
  for (int s = 0; s < 2; ++s) {
    const int ss = d->in_c[s] ? s : 0;
    p.gdst[s] = hl_view(d->g_in[ss]);
    p.saved[s] = hl_view(d->in[ss]);
    p.gate[s] = d->gate[ss];
    if (!hl_view_v8(d->g_in[ss]) || !hl_view_v8(d->in[ss]) || d->in_c[0] % 16 != 0) p.v8 = 0;
  }

  HlMaps maps;
  for (int s = 0; s < 4; ++s) {
    const int oy = upconv ? (s >> 1) : 0, ox = upconv ? (s & 1) : 0;
    for (int pl = 0; pl < 3; ++pl)
      if ((rc = hl_make_act_map(&maps.a[s][pl], d->g_out, pl, d->cout, d->batch, Ho_f, Wo_f, t, upconv ? 2 : 1, oy, ox)))
        return rc;
  }
  // transposed weights [3][taps][N = cin][K = nsrc * cout]
  const bool pair64 = p.cta2 && nt == 64;
  if ((rc = hl_make_weight_map(&maps.w, d->wt_packed, keff * keff, ncols, nsrc * d->cout,
                               pair64 ? nt : (p.cta2 ? nt / 2 : nt), p.kwc ? 3 : 1, pair64 ? 1 : 3)))
    return rc;
  maps.d = maps.w;          // unused, except by the 64-column CTA pairs
  if (pair64 && (rc = hl_make_weight_map(&maps.d, d->wt_packed, keff * keff, ncols, nsrc * d->cout, nt / 2, 1, 1)))
    return rc;
  switch (nt) {
    case 128: return hl_launch_dgrad<128>(maps, p, stream);
    case 64: return hl_launch_dgrad<64>(maps, p, stream);
    default: return hl_launch_dgrad<32>(maps, p, stream);
  }
}

}  // namespace sn

extern "C" int sn_tc_head_fusable(const sn_tc_conv_desc* d, int32_t n_labels) {
  return d != nullptr && sn::hl_head_shape_ok(d, n_labels) ? 1 : 0;
}
