// Probe: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) vs N, descriptor row shift and swizzle mode.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "sn_sm100.cuh"
using namespace sn;

template <int ROWB>
__global__ void rate(int N, int shift, int iters, int same_acc, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 400 * ROWB, bar = sb + 256 * ROWB, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(raw + (slot - ptx::smem_u32(raw)));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (400 + 256) * ROWB / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (warp == 0) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot_gen;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
    const uint64_t da0 = ptx::smem_desc_kmajor<ROWB>(sa + shift * ROWB);
    const uint64_t da1 = ptx::smem_desc_kmajor<ROWB>(sa + 128 * ROWB + shift * ROWB);
    const uint64_t db = ptx::smem_desc_kmajor<ROWB>(sb);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ptx::umma_bf16(tm, da0, db, idesc, 1);
      ptx::umma_bf16(same_acc ? tm : tm + 256, da1, db, idesc, 1);
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 2000;
  for (int rowb : {64, 128}) {
    const int smem = (400 + 256) * rowb + 2048;
    if (rowb == 64) cudaFuncSetAttribute(rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else cudaFuncSetAttribute(rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int same : {1, 0})
      for (int N : {32, 48, 64, 96, 128, 192, 256})
        for (int shift : {0, 43}) {
          if (rowb == 64) rate<64><<<1, 128, smem>>>(N, shift, iters, same, d);
          else rate<128><<<1, 128, smem>>>(N, shift, iters, same, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("rowB %3d same_acc %d N %3d shift %2d: %6.1f cycles/MMA (%s)\n", rowb, same, N, shift,
                 (double)c / (2.0 * iters), cudaGetErrorString(e));
        }
  }
  return 0;
}
