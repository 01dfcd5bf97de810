// mbar_probe.cu — development probe: cost of mbarrier polling / tcgen05.commit / elect in a single warp.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_ptr;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<32>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&bar[0]);   // phase 0 of bar[0] complete; bar[1] stays incomplete
  __syncthreads();
  const int N = 1000;
  if (threadIdx.x < 32) {
    long long t0, t1; int acc = 0;
    t0 = clock64(); for (int i = 0; i < N; ++i) acc += mbar_test_wait(&bar[0], 0); t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) acc += mbar_test_wait(&bar[1], 0); t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) acc += mbar_try_wait(&bar[0], 0); t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { if (threadIdx.x == 0) acc += mbar_test_wait(&bar[0], 0); } t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { bool ok = false; if (threadIdx.x == 0) ok = mbar_test_wait(&bar[0], 0); acc += __any_sync(0xffffffffu, ok); } t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { if (elect_one()) umma_commit(&bar[2]); __syncwarp(); } t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < N; ++i) { tc_fence_after(); } t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0);
    // commit then wait for its arrival (round trip), all lanes
    uint32_t ph = 0;
    t0 = clock64(); for (int i = 0; i < N; ++i) { if (elect_one()) umma_commit(&bar[3]); __syncwarp(); mbar_wait(&bar[3], ph); ph ^= 1; } t1 = clock64();
    if (threadIdx.x == 0) out[7] = (t1 - t0);
    // incomplete try_wait (suspends?)
    t0 = clock64(); for (int i = 0; i < 100; ++i) acc += mbar_try_wait(&bar[1], 0); t1 = clock64();
    if (threadIdx.x == 0) out[8] = (t1 - t0) * 10;
    if (acc == -1) out[9] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<32>(tmem_ptr);
}
int main() {
  long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"test_wait complete, all lanes", "test_wait incomplete, all lanes", "try_wait complete, all lanes", "test_wait complete, lane 0", "test_wait lane 0 + any_sync",
                         "elect + tcgen05.commit + syncwarp", "tcgen05.fence::after", "commit + wait round trip", "try_wait incomplete (x10 of 100)"};
  for (int i = 0; i < 9; ++i) printf("%-40s %.1f cycles/iter\n", names[i], h[i] / 1000.0);
  return 0;
}
