// Probe: HBM read rate of the halo-box TMA load pattern, pixel-interleaved planes [B][H][W][3][C] vs planar
// [3][B][H][W][C].  Each CTA (persistent, 148) loops over tiles, 3 loads (one per plane) of box {32 ch, R, TH, 1}
// per tile into a 4-stage ring; a consumer warp just waits and releases.  C = 32, H = W = 202, B = 64.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "sn_sm100.cuh"
using namespace sn;

struct Maps { CUtensorMap m[3]; };

__global__ void __launch_bounds__(64, 1) reader(const __grid_constant__ Maps maps, int tiles_x, int tiles_y, int B,
                                                int TWo, int THo, int box_bytes, int stages) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const int plane = ((box_bytes + 1023) / 1024) * 1024;
  const uint32_t bar = base + stages * 3 * plane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(bar + 8 * s, 1); ptx::mbar_init(bar + 64 + 8 * s, 1); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int total = tiles_x * tiles_y * B;
  if (warp == 0 && lane == 0) {
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      const int s = i % stages; const uint32_t par = (i / stages) & 1;
      ptx::mbar_wait(bar + 64 + 8 * s, par ^ 1);
      ptx::mbar_arrive_expect_tx(bar + 8 * s, 3 * box_bytes);
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
      for (int pl = 0; pl < 3; ++pl)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
                     ::"r"(base + (s * 3 + pl) * plane), "l"(reinterpret_cast<uint64_t>(&maps.m[pl])), "r"(bar + 8 * s),
                       "r"(0), "r"(tx * TWo), "r"(ty * THo), "r"(b) : "memory");
    }
  } else if (warp == 1 && lane == 0) {
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      const int s = i % stages; const uint32_t par = (i / stages) & 1;
      ptx::mbar_wait(bar + 8 * s, par);
      ptx::mbar_arrive(bar + 64 + 8 * s);
    }
  }
}

int main() {
  cuInit(0); cudaFree(0);
  const int B = 64, H = 202, W = 202, C = 32, R = 42, THb = 5, TWo = 40, THo = 3;
  const size_t elems = (size_t)B * H * W * 3 * C;
  __nv_bfloat16* d; cudaMalloc(&d, elems * 2); cudaMemset(d, 0, elems * 2);
  char* flush; cudaMalloc(&flush, 512 << 20);
  for (int planar = 0; planar < 2; ++planar) {
    Maps maps;
    for (int pl = 0; pl < 3; ++pl) {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t st[3];
      char* basep;
      if (planar) { st[0] = C * 2; st[1] = (cuuint64_t)W * C * 2; st[2] = (cuuint64_t)H * W * C * 2; basep = (char*)d + (size_t)pl * B * H * W * C * 2; }
      else { st[0] = 3 * C * 2; st[1] = (cuuint64_t)W * 3 * C * 2; st[2] = (cuuint64_t)H * W * 3 * C * 2; basep = (char*)d + pl * C * 2; }
      cuuint32_t box[4] = {32, R, THb, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = cuTensorMapEncodeTiled(&maps.m[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, st, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode fail %d\n", r); return 1; }
    }
    const int box_bytes = R * THb * 64, stages = 4;
    const int smem = stages * 3 * (((box_bytes + 1023) / 1024) * 1024) + 2048;
    cudaFuncSetAttribute(reader, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int tiles_x = 5, tiles_y = 67;
    for (int it = 0; it < 3; ++it) {
      cudaMemset(flush, 1, 512 << 20);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      reader<<<148, 64, smem>>>(maps, tiles_x, tiles_y, B, TWo, THo, box_bytes, stages);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("%s: %.3f ms, %.0f GB/s (tensor bytes %.0f MB) %s\n", planar ? "planar     " : "interleaved", ms,
             elems * 2 / ms / 1e6, elems * 2 / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
