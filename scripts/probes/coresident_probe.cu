// coresident_probe.cu — development probe: can a small kernel of another stream become resident on an SM that already runs one big
// persistent CTA (640 threads, 96 registers, ~221 KB dynamic shared memory, optionally all 512 TMEM columns)?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

__global__ void __launch_bounds__(640, 1) big(int use_tmem, long long spin_cycles, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  if (use_tmem && threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  __syncthreads();
  const long long t0 = clock64();
  int acc = 0;
  while (clock64() - t0 < spin_cycles) acc += (int)(clock64() & 1);
  if (acc == 123456789) sink[0] = acc;
  __syncthreads();
  if (use_tmem && threadIdx.x < 32) tmem_dealloc<512>(tmem_ptr);
}
__global__ void __launch_bounds__(64) small_k(long long spin_cycles, int* sink) {
  const long long t0 = clock64();
  int acc = 0;
  while (clock64() - t0 < spin_cycles) acc++;
  if (acc == 123456789) sink[1] = acc;
}
int main() {
  int* sink; cudaMalloc(&sink, 64);
  cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  struct C { int smem_kb, tmem, carve; const char* name; } cs[] = {
    {0, 0, 0, "big: no smem, no TMEM"}, {221, 0, 0, "big: 221 KB smem"}, {221, 0, 1, "big: 221 KB smem, small kernel carveout=max shared"},
    {0, 1, 0, "big: all TMEM columns, no smem"}, {221, 1, 1, "big: 221 KB smem + TMEM, small carveout=max shared"}, {100, 1, 1, "big: 100 KB smem + TMEM"},
  };
  for (auto& c : cs) {
    cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_kb * 1024);
    cudaFuncSetAttribute(small_k, cudaFuncAttributePreferredSharedMemoryCarveout, c.carve ? cudaSharedmemCarveoutMaxShared : cudaSharedmemCarveoutDefault);
    for (int rep = 0; rep < 2; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, a);
      big<<<148, 640, c.smem_kb * 1024, a>>>(c.tmem, 4000000, sink);      // ~2 ms
      cudaEventRecord(e1, a);
      small_k<<<148, 64, 0, b>>>(200000, sink);                           // ~0.1 ms if it can start right away
      cudaEventRecord(e2, b);
      cudaDeviceSynchronize();
      float t_big, t_small_end;
      cudaEventElapsedTime(&t_big, e0, e1);
      cudaEventElapsedTime(&t_small_end, e0, e2);
      if (rep) printf("%-58s big %.2f ms, small kernel finished at %.2f ms -> %s\n", c.name, t_big, t_small_end, t_small_end < 0.7f * t_big ? "CO-RESIDENT" : "serialised");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  }
  return 0;
}
