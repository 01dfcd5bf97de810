// queue_probe.cu — development probe: how far can one thread run ahead of the tensor pipe? Time to ISSUE a burst of K
// 128x208x16 MMAs from an idle pipe (clock right after the last tcgen05.mma) versus the time until they have all executed.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392;
template <int K>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 208);
    const uint64_t ad0 = make_smem_desc(smem_u32(smem) + 65536, 2048, 128), bd0 = make_smem_desc(smem_u32(smem) + 32, kPlane, 128);
    long long t_issue = 0, t_done = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 20; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < K; ++k) umma_bf16(tm, ad0 + (uint64_t)((k & 7) * 256), bd0 + (uint64_t)((k & 7) * 2 * (kPlane >> 4)), idesc, 1);
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, ph); ph ^= 1;
      const long long t2 = clock64();
      if (rep >= 4) { t_issue += t1 - t0; t_done += t2 - t0; }
    }
    if (lane == 0) { out[0] = t_issue / 16; out[1] = t_done / 16; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
template <int K> void run(long long* d) {
  cudaFuncSetAttribute(probe<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  probe<K><<<1, 128, 160 * 1024>>>(d);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("burst of %2d MMAs: issued after %5lld cycles, executed after %5lld cycles (pipe time %d)\n", K, h[0], h[1], K * 104);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<1>(d); run<2>(d); run<4>(d); run<6>(d); run<8>(d); run<12>(d); run<16>(d); run<24>(d);
  return 0;
}
