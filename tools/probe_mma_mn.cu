// Probe: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) with MN-major operands (the weight-gradient layout:
// smem row = pixel = K index, 64 B = 32 channels along M/N) versus K-major, as a function of N and of the A-block
// stride LBO (4 KB = separate 32-channel blocks, 64 B = the overlapping "one pixel shift per block" trick).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "sn_sm100.cuh"
using namespace sn;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

__global__ void rate(int N, int a_mn, int b_mn, int lbo_a, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 64 * 1024, bar = sb + 64 * 1024, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(raw + (slot - ptx::smem_u32(raw)));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (warp == 0) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot_gen;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da0 = a_mn ? desc_mn(sa, lbo_a) : ptx::smem_desc_kmajor<64>(sa);
    const uint64_t da1 = a_mn ? desc_mn(sa + 16384, lbo_a) : ptx::smem_desc_kmajor<64>(sa + 16384);
    const uint64_t db = b_mn ? desc_mn(sb, 4096) : ptx::smem_desc_kmajor<64>(sb);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ptx::umma_bf16(tm, da0, db, idesc, 1);
      ptx::umma_bf16(tm + 256, da1, db, idesc, 1);
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
  const int smem = 128 * 1024 + 4096;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cfg[5][3] = {{0, 0, 4096}, {1, 1, 4096}, {1, 1, 64}, {1, 0, 4096}, {0, 1, 4096}};
  for (auto& c : cfg)
    for (int N : {32, 64, 128, 256}) {
      rate<<<1, 128, smem>>>(N, c[0], c[1], c[2], iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("A %s  B %s  LBO_A %4d  N %3d: %6.1f cycles/MMA (%s)\n", c[0] ? "MN-major" : "K-major ",
             c[1] ? "MN-major" : "K-major ", c[2], N, (double)cyc / (2.0 * iters), cudaGetErrorString(e));
    }
  return 0;
}
