// epi_probe2.cu — development probe: cycles of the REAL main epilogue (dan_stack_epi.cuh) per read-layer, 8 warps per slot,
// one or both slots active, against a non-pipelined variant of the same chunk code.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
#include "../../dl4vc_b200/csrc/dan_stack_epi.cuh"
using namespace ptx;
constexpr int kPlane = 3392, kLeadR = 2, kBuf = 16 * kPlane;

template <int MODE>
__device__ __forceinline__ void simple_main(uint32_t tbase, uint32_t saddr0, int lane, int P, int g_begin, int g_end, const EpiConsts& k) {
#pragma unroll 1
  for (int g0 = g_begin; g0 < g_end; g0 += 2) {
    uint32_t a0[8], a1[8];
    stack_epi_load(tbase, g0, a0, a1);
    stack_epi_wait(a0, a1);
    stack_epi_do<MODE>(a0, a1, tbase, saddr0, g0, lane, P, k);
  }
  if constexpr (MODE == kEpiPreRes) tmem_st_wait();
}

// variant: 0 = pipelined (kernel), 1 = simple, 2 = pool (16 warps on one accumulator); mode: kEpi*; slots: 1 or 2
__global__ void __launch_bounds__(512, 1) probe(int variant, int mode, int slots, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const int s = warp >> 3, h = (warp >> 2) & 1, q = warp & 3;
  if (variant == 2) {
    const int w4 = warp >> 2, hh = w4 & 1;
    const uint32_t taddr0 = tm + ((uint32_t)(32 * q + 16 * hh) << 16);
    const uint32_t saddr0 = smem_u32(smem) + (4 * q + 2 * hh + ((lane >> 3) & 1)) * kPlane + (kLeadR + 8 * (lane >> 4) + (lane & 7)) * 16;
    PoolConsts k;
    for (int j = 0; j < 2; ++j) { k.nb[j] = -0.1f * (lane + j); k.scale[j] = 1.01f + j; k.c[j] = 0.5f; k.rbias[j] = 0.25f; }
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == kEpiFinal) pool_epi<kEpiFinal>(taddr0, saddr0, w4 >> 1, lane, 201, k);
      else if (mode == kEpiPreRes) pool_epi<kEpiPreRes>(taddr0, saddr0, w4 >> 1, lane, 201, k);
      else pool_epi<kEpiPostRes>(taddr0, saddr0, w4 >> 1, lane, 201, k);
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) / iters;
  } else if (s < slots) {
    const uint32_t tbase = tm + s * 256 + ((uint32_t)(32 * q) << 16);
    const uint32_t buf = smem_u32(smem) + s * kBuf;
    const uint32_t saddr0 = buf + (4 * q + (lane >> 3)) * kPlane + (kLeadR + (lane & 7)) * 16;
    const int g_begin = h == 0 ? 0 : 14, g_end = h == 0 ? 14 : 26;
    EpiConsts k;
    for (int j = 0; j < 4; ++j) { k.nb[j] = -0.1f * (lane + j); k.scale[j] = 1.01f + j; k.c[j] = 0.5f; k.rbias[j] = 0.25f; }
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (variant == 0) {
        if (mode == kEpiFinal) stack_epi_main_pipelined<kEpiFinal>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else if (mode == kEpiPreRes) stack_epi_main_pipelined<kEpiPreRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else stack_epi_main_pipelined<kEpiPostRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
      } else {
        if (mode == kEpiFinal) simple_main<kEpiFinal>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else if (mode == kEpiPreRes) simple_main<kEpiPreRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else simple_main<kEpiPostRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) / iters;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 128 * 8); cudaMemset(d, 0, 1024);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* mn[] = {"final", "pre-res", "post-res"};
  for (int variant = 0; variant < 3; ++variant)
    for (int mode = 0; mode < 3; ++mode)
      for (int slots = 1; slots <= (variant == 2 ? 1 : 2); ++slots) {
        probe<<<1, 512, 120 * 1024>>>(variant, mode, slots, 200, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < (variant == 2 ? 16 : 8 * slots); ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-10s %-8s slots active %d: %5lld cycles per epilogue call (slowest warp)\n", variant == 2 ? "pool16" : variant ? "simple" : "pipelined", mn[mode], slots, mx); fflush(stdout);
      }
  return 0;
}
