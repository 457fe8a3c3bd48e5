// Shared helpers for libsupernet_b200.so (internal; the ABI is include/supernet.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "supernet.h"

namespace sn {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
// Records cudaGetLastError() (if any) after a launch. Returns SN_OK / SN_ERR_LAUNCH.
int check_launch(const char* what);
int num_sms();
void* tensor_map_encoder();      // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, or nullptr (sn_tc_halo.cu)

static inline cudaStream_t as_stream(sn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define SN_REQUIRE(cond, code, ...)                       \
  do {                                                    \
    if (!(cond)) return ::sn::fail((code), __VA_ARGS__);  \
  } while (0)

__device__ __forceinline__ float softplus_f(float x) {
  // tf.math.softplus: log(1+exp(x)), stable on both tails
  return x > 20.f ? x : (x < -20.f ? __expf(x) : log1pf(__expf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// mean -> (hi, lo) bf16 split, mean ~= hi + lo to ~2^-17 relative.
__device__ __forceinline__ void split_bf16(float m, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(m);
  lo = __float2bfloat16_rn(m - __bfloat162float(hi));
}

// conv_final (k = 1, Brats.py:367,454) + mysoftmax (Brats.py:269-283) for ONE pixel, channel by channel.  Shared by
// final_conv_softmax_kernel (sn_packed.cu) and the fused head of the halo kernel (sn_tc_halo.cu): both run this exact
// instruction sequence in channel order, so the fused and the two-kernel forward agree bit for bit.
//   head_accumulate: r += mu^2 + var;  m_j += mu W[c][j];  v_j += var W[c][j]^2
//   head_finish:     v_j = max(v_j + s_j r, 0);  p = softmax(m);  vo_a = sum_j (p_a (delta_aj - p_j))^2 v_j
// (d0, d1) += a * (b0, b1): one FFMA2 (fma.rn.f32x2, sm_100) = two IEEE fp32 FMAs in one issue slot
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
  uint64_t d, aa, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
template <int C>
__device__ __forceinline__ void head_accumulate(float mu, float vv, const float* __restrict__ wrow,
                                                const float* __restrict__ w2row, float (&m)[C], float (&v)[C], float& r) {
  r += fmaf(mu, mu, vv);
#pragma unroll
  for (int j = 0; j + 1 < C; j += 2) {
    fma2(m[j], m[j + 1], mu, wrow[j], wrow[j + 1]);
    fma2(v[j], v[j + 1], vv, w2row[j], w2row[j + 1]);
  }
  if constexpr (C & 1) {
    m[C - 1] = fmaf(mu, wrow[C - 1], m[C - 1]);
    v[C - 1] = fmaf(vv, w2row[C - 1], v[C - 1]);
  }
}
template <int C>
__device__ __forceinline__ void head_finish(const float (&m)[C], float (&v)[C], float r, const float* __restrict__ ss,
                                            float (&p)[C], float (&vo)[C]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    v[j] = fmaxf(fmaf(ss[j], r, v[j]), 0.f);
    mx = fmaxf(mx, m[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) { p[j] = expf(m[j] - mx); sum += p[j]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < C; ++j) p[j] *= inv;
#pragma unroll
  for (int a = 0; a < C; ++a) {
    float acc = 0.f;   // sum_j (p_a (delta_aj - p_j))^2 v_j : non-negative terms only
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float J = p[a] * ((a == j ? 1.f : 0.f) - p[j]);
      acc = fmaf(J * J, v[j], acc);
    }
    vo[a] = acc;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Grid size for a grid-stride elementwise kernel: a multiple of the SM count, capped by the work.
static inline int ew_grid(size_t work_items, int threads, int ctas_per_sm = 8) {
  size_t need = (work_items + threads - 1) / threads;
  size_t cap = (size_t)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace sn
