// commit_probe.cu — development probe: cost (to the issuing warp) of tcgen05.commit, elect.sync regions and mbarrier try_wait,
// with and without a concurrent tcgen05.mma stream from another warp of the CTA.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392;

__global__ void __launch_bounds__(128, 1) probe(int with_mma, int what, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8], done;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); stop = 0; fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 1 && with_mma) {
    const uint32_t idesc = make_idesc_bf16(128, 208);
    const uint64_t ad0 = make_smem_desc(smem_u32(smem) + 65536, 2048, 128), bd0 = make_smem_desc(smem_u32(smem) + 32, kPlane, 128);
    while (!stop) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tm + 256, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 2 * (kPlane >> 4)), idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
  } else if (warp == 0) {
    __syncwarp();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      if (what == 0) { if (elect_one()) umma_commit(&bars[it & 7]); __syncwarp(); }               // commit only
      else if (what == 1) { if (elect_one()) acc += it; __syncwarp(); }                              // elect region only
      else if (what == 2) { acc += mbar_try_wait(&bars[it & 7], 1) ? 1 : 0; }                       // try_wait on a completed phase (parity 1 of a fresh barrier)
      else if (what == 3) {                                                                          // 4 MMAs + 2 commits per elect (kernel's conv iteration)
        const uint32_t idesc = make_idesc_bf16(128, 208);
        const uint64_t ad0 = make_smem_desc(smem_u32(smem) + 65536 + (it & 3) * 8192, 2048, 128), bd0 = make_smem_desc(smem_u32(smem) + 32, kPlane, 128);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tm, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 2 * (kPlane >> 4)), idesc, 1);
          umma_commit(&bars[it & 7]); umma_commit(&bars[(it + 1) & 7]);
        }
        __syncwarp();
      } else if (what == 4) {                                                                        // 4 MMAs per elect, no commits
        const uint32_t idesc = make_idesc_bf16(128, 208);
        const uint64_t ad0 = make_smem_desc(smem_u32(smem) + 65536 + (it & 3) * 8192, 2048, 128), bd0 = make_smem_desc(smem_u32(smem) + 32, kPlane, 128);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tm, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 2 * (kPlane >> 4)), idesc, 1);
        }
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (t1 - t0) / iters; out[1] = acc; }
    stop = 1;
    if (what >= 3) { if (elect_one()) umma_commit(&done); __syncwarp(); if (!with_mma) mbar_wait(&done, 0); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[] = {"elect + tcgen05.commit", "elect region only", "mbarrier try_wait (complete)", "elect + 4 MMA + 2 commits", "elect + 4 MMA"};
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int what = 0; what < 5; ++what) {
      if (with_mma && what >= 3) continue;
      probe<<<1, 128, 160 * 1024>>>(with_mma, what, 500, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-32s %s: %lld cycles per iteration\n", names[what], with_mma ? "(other warp streams MMAs)" : "(tensor pipe idle)        ", h[0]); fflush(stdout);
    }
  return 0;
}
