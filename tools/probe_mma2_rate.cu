// Probe: cycles per cta_group::2 tcgen05.mma (M = 256 over a CTA pair, K = 16, bf16, SS mode) vs N.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "sn_sm100.cuh"
using namespace sn;

__global__ void __cluster_dims__(2, 1, 1) rate2(int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 400 * 64, bar = sb + 256 * 64, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(raw + (slot - ptx::smem_u32(raw)));
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < (400 + 256) * 64 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (warp == 0) { ptx::tmem_alloc2(slot, 512); ptx::tmem_relinquish2(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = *slot_gen;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(256, N);
    const uint64_t da0 = ptx::smem_desc_kmajor<64>(sa);
    const uint64_t da1 = ptx::smem_desc_kmajor<64>(sa + 128 * 64);
    const uint64_t db = ptx::smem_desc_kmajor<64>(sb);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ptx::umma2_bf16(tm, da0, db, idesc, 1);
      ptx::umma2_bf16(tm + 256, da1, db, idesc, 1);
    }
    ptx::umma2_commit(bar);
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  if (threadIdx.x == 0 && rank == 1) ptx::mbar_wait(bar, 0);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc2(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 2000;
  const int smem = (400 + 256) * 64 + 2048;
  cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N : {32, 64, 128, 192, 256}) {
    rate2<<<2, 128, smem>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("cta_group::2 M 256 N %3d: %6.1f cycles/MMA (%s)\n", N, (double)c / (2.0 * iters), cudaGetErrorString(e));
  }
  return 0;
}
