// overlap_probe.cu — development probe: does the main epilogue (TMEM -> registers -> stmatrix) slow down a concurrent tcgen05.mma
// stream (and vice versa)? Warp 16 issues back-to-back 128x208x16 MMAs into accumulator 0; warps 0..15 (pool epilogue) or
// 0..7 run the real epilogue code on accumulator 1 (TMEM columns 256..463) and a read buffer in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
#include "../../dl4vc_b200/csrc/dan_stack_epi.cuh"
using namespace ptx;
constexpr int kPlane = 3392, kLeadR = 2, kBuf = 16 * kPlane;

// what: bit 0 = run MMAs, bit 1 = run epilogue; epi_warps 8 or 16; mode kEpi*
__global__ void __launch_bounds__(544, 1) probe(int what, int epi_warps, int mode, int n_mma, int epi_iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  uint8_t* act = smem;                   // epilogue buffer (54 KB) + MMA B operand (another read buffer)
  uint8_t* actB = smem + kBuf;
  uint8_t* wts = smem + 2 * kBuf;        // 64 KB of "weights" (A operand blocks of 4 KB)
  if (warp == 16) {
    if (what & 1) {
      const uint32_t idesc = make_idesc_bf16(128, 208);
      const uint64_t ad0 = make_smem_desc(smem_u32(wts), 2048, 128), bd0 = make_smem_desc(smem_u32(actB) + 32, kPlane, 128);
      __syncwarp();
      const long long t0 = clock64();
      for (int g = 0; g < n_mma / 8; ++g) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16(tm, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 2 * (kPlane >> 4)), idesc, 1);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (lane == 0) out[16] = (t1 - t0);
    }
  } else if ((what & 2) && warp < epi_warps) {
    const int q = warp & 3;
    long long t0, t1;
    if (epi_warps == 16) {
      const int w4 = warp >> 2, hh = w4 & 1;
      const uint32_t taddr0 = tm + 256 + ((uint32_t)(32 * q + 16 * hh) << 16);
      const uint32_t saddr0 = smem_u32(act) + (4 * q + 2 * hh + ((lane >> 3) & 1)) * kPlane + (kLeadR + 8 * (lane >> 4) + (lane & 7)) * 16;
      PoolConsts k;
      for (int j = 0; j < 2; ++j) { k.nb[j] = -0.1f * (lane + j); k.scale[j] = 1.01f + j; k.c[j] = 0.5f; k.rbias[j] = 0.25f; }
      __syncwarp();
      t0 = clock64();
      for (int it = 0; it < epi_iters; ++it) {
        if (mode == kEpiFinal) pool_epi<kEpiFinal>(taddr0, saddr0, w4 >> 1, lane, 201, k);
        else if (mode == kEpiPreRes) pool_epi<kEpiPreRes>(taddr0, saddr0, w4 >> 1, lane, 201, k);
        else pool_epi<kEpiPostRes>(taddr0, saddr0, w4 >> 1, lane, 201, k);
      }
      t1 = clock64();
    } else {
      const int h = (warp >> 2) & 1;
      const uint32_t tbase = tm + 256 + ((uint32_t)(32 * q) << 16);
      const uint32_t saddr0 = smem_u32(act) + (4 * q + (lane >> 3)) * kPlane + (kLeadR + (lane & 7)) * 16;
      const int g_begin = h == 0 ? 0 : 14, g_end = h == 0 ? 14 : 26;
      EpiConsts k;
      for (int j = 0; j < 4; ++j) { k.nb[j] = -0.1f * (lane + j); k.scale[j] = 1.01f + j; k.c[j] = 0.5f; k.rbias[j] = 0.25f; }
      __syncwarp();
      t0 = clock64();
      for (int it = 0; it < epi_iters; ++it) {
        if (mode == kEpiFinal) stack_epi_main_pipelined<kEpiFinal>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else if (mode == kEpiPreRes) stack_epi_main_pipelined<kEpiPreRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
        else stack_epi_main_pipelined<kEpiPostRes>(tbase, saddr0, lane, 201, g_begin, g_end, k);
      }
      t1 = clock64();
    }
    if (lane == 0) out[warp] = (t1 - t0) / epi_iters;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 128 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* mn[] = {"final", "pre-res", "post-res"};
  const int n_mma = 4000;
  for (int epi_warps : {8, 16})
    for (int mode = 0; mode < 3; ++mode)
      for (int what : {1, 2, 3}) {
        if (what == 1 && (mode != 0 || epi_warps != 8)) continue;
        cudaMemset(d, 0, 1024);
        // epilogue iterations sized so that both activities span a similar time when overlapped
        const int epi_iters = 300;
        probe<<<148, 544, 200 * 1024>>>(what, epi_warps, mode, n_mma, epi_iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < epi_warps; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("epi warps %2d %-8s %s%s: ", epi_warps, mn[mode], (what & 1) ? "MMA " : "    ", (what & 2) ? "EPI" : "   ");
        if (what & 1) printf("%.1f cycles/MMA (ideal 104)  ", (double)h[16] / n_mma);
        if (what & 2) printf("%lld cycles/epilogue", mx);
        printf("\n"); fflush(stdout);
      }
  return 0;
}
