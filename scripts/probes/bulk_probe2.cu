// bulk_probe2.cu — development probe: per-instruction latency of mbarrier.arrive.expect_tx and cp.async.bulk in one thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
__device__ __forceinline__ void mbar_expect_only(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* w, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long ta = 0, tb = 0, tc = 0, td = 0;
    uint32_t par = 0;
    for (int it = 0; it < 200; ++it) {
      for (int idx = 0; idx < 8; ++idx) {
        long long t0 = clock64();
        mbar_expect_tx(&full[idx], 8192);
        long long t1 = clock64();
        bulk_g2s(smem + idx * 8192, w + ((it * 8 + idx) % 64) * 8192, 8192, &full[idx]);
        long long t2 = clock64();
        ta += t1 - t0; tb += t2 - t1;
      }
      for (int idx = 0; idx < 8; ++idx) { long long t0 = clock64(); mbar_wait(&full[idx], par); tc += clock64() - t0; }
      par ^= 1;
    }
    // variant: arm with plain expect_tx (no arrive) + explicit arrive afterwards
    for (int it = 0; it < 200; ++it) {
      for (int idx = 0; idx < 8; ++idx) {
        long long t0 = clock64();
        mbar_expect_only(&full[idx], 8192);
        bulk_g2s(smem + idx * 8192, w + ((it * 8 + idx) % 64) * 8192, 8192, &full[idx]);
        mbar_arrive(&full[idx]);
        td += clock64() - t0;
      }
      for (int idx = 0; idx < 8; ++idx) mbar_wait(&full[idx], par);
      par ^= 1;
    }
    out[0] = ta / 1600; out[1] = tb / 1600; out[2] = tc / 1600; out[3] = td / 1600;
  }
}
int main() {
  uint8_t* w; cudaMalloc(&w, 64 * 8192); cudaMemset(w, 1, 64 * 8192);
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<<<1, 128, 100 * 1024>>>(w, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("arrive.expect_tx %lld cycles, cp.async.bulk issue %lld cycles, wait (8 in flight) %lld cycles, expect_tx+bulk+arrive %lld cycles\n", h[0], h[1], h[2], h[3]);
  return 0;
}
