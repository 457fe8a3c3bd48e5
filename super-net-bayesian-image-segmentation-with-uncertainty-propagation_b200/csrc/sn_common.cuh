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
