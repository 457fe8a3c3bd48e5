// Probe: does a UMMA shared-memory descriptor whose start address is shifted by s rows (s not a multiple of 8)
// inside a TMA-written swizzled tile read the rows the TMA wrote?  Tests SW64 (64-byte rows) and SW128
// (128-byte rows) with three base_offset conventions.  D[128 x 32] = A[s : s+128, :] * B^T, K = 32 (SW64) or 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I<csrc> -o probe_shift probe_shift.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "sn_sm100.cuh"
using namespace sn;

template <int ROWB>
__global__ void probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, int shift,
                      int bo_mode, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  constexpr int AROWS = 256, NB = 32, KE = ROWB / 2;
  const uint32_t sa = base, sb = base + AROWS * ROWB, bar = sb + NB * ROWB, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(raw + (slot - ptx::smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(bar2, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) { ptx::tmem_alloc(slot, 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot_gen;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, AROWS * ROWB + NB * ROWB);
    // A: 2-D map {K, rows}, box {KE, 256}; B: box {KE, 32}
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&ma)), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&mb)), "r"(bar), "r"(0), "r"(0) : "memory");
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::idesc_bf16_f32(128, NB);
    for (int ks = 0; ks < KE / 16; ++ks) {
      const uint32_t a_addr = sa + shift * ROWB + ks * 32;
      uint64_t da = ptx::smem_desc_kmajor<ROWB>(a_addr);
      uint64_t bo = 0;
      if (bo_mode == 1) bo = (a_addr >> 7) & 7;
      if (bo_mode == 2) bo = (a_addr >> 7) & 3;
      da |= bo << 49;
      const uint64_t db = ptx::smem_desc_kmajor<ROWB>(sb + ks * 32);
      ptx::umma_bf16(tm, da, db, idesc, ks > 0);
    }
    ptx::umma_commit(bar2);
  }
  __syncthreads();
  if (warp < 4) {
    ptx::mbar_wait(bar2, 0);
    ptx::tc_fence_after();
    uint32_t r[16];
    for (int c0 = 0; c0 < NB; c0 += 16) {
      ptx::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * NB + c0 + j] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 32);
}

static CUtensorMap make2d(void* p, int K, int rows, int boxk, int boxr, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxk, (cuuint32_t)boxr};
  cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int ROWB>
void run_case() {
  constexpr int KE = ROWB / 2, AROWS = 256, NB = 32;
  std::vector<__nv_bfloat16> ha(AROWS * KE), hb(NB * KE);
  std::vector<float> fa(AROWS * KE), fb(NB * KE);
  srand(1);
  for (int i = 0; i < AROWS * KE; ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = v; }
  for (int i = 0; i < NB * KE; ++i) { float v = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * NB * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMapSwizzle sw = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ma = make2d(da, KE, AROWS, KE, AROWS, sw), mb = make2d(db, KE, NB, KE, NB, sw);
  const int smem = AROWS * ROWB + NB * ROWB + 1024 + 64;
  cudaFuncSetAttribute(probe<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(128 * NB);
  for (int bo = 0; bo < 3; ++bo) {
    printf("ROW_BYTES %d base_offset mode %d: ", ROWB, bo);
    for (int s = 0; s <= 19; ++s) {
      cudaMemset(dout, 0, 128 * NB * 4);
      probe<ROWB><<<1, 128, smem>>>(ma, mb, s, bo, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[s=%d CUDA error %s] ", s, cudaGetErrorString(e)); break; }
      cudaMemcpy(ho.data(), dout, 128 * NB * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
          double ref = 0;
          for (int k = 0; k < KE; ++k) ref += (double)fa[(m + s) * KE + k] * fb[n * KE + k];
          maxerr = fmax(maxerr, fabs(ref - ho[m * NB + n]));
        }
      printf("%s", maxerr < 1e-3 ? "." : "X");
    }
    printf("   (shift 0..19, . = exact)\n");
  }
  cudaFree(da); cudaFree(db); cudaFree(dout);
}

int main() {
  cuInit(0);
  cudaFree(0);
  run_case<64>();
  run_case<128>();
  return 0;
}
