// twostream_probe.cu — development probe: aggregate tcgen05.mma throughput when two warps issue to different accumulators
// (the stack kernel's two slots), for conv-shaped (N = 208) and bottleneck-shaped (N = 32) streams, in blocks of `blk` MMAs per
// elected region.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392;

template <int BLK>
__device__ __forceinline__ long long stream(uint32_t d, uint32_t idesc, uint64_t ad0, uint64_t bd0, uint32_t bstep, int n, uint64_t* bar) {
  __syncwarp();
  const long long t0 = clock64();
  for (int g = 0; g < n / BLK; ++g) {
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < BLK; ++k) umma_bf16(d, ad0 + (uint64_t)((k & 7) * 256), bd0 + (uint64_t)((k & 7) * bstep), idesc, 1);
    }
    __syncwarp();
  }
  if (elect_one()) umma_commit(bar);
  __syncwarp();
  mbar_wait(bar, 0);
  return clock64() - t0;
}

// shapeY: 0 = none, 1 = N 208, 2 = N 32 (positions-as-M bottleneck orientation)
template <int BLK>
__global__ void __launch_bounds__(128, 1) probe(int shapeY, int nX, int nY, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const uint64_t w0 = make_smem_desc(smem_u32(smem) + 2 * 16 * kPlane, 2048, 128);            // weights as A (conv) ...
  const uint64_t x0 = make_smem_desc(smem_u32(smem) + 32, kPlane, 128), x1 = make_smem_desc(smem_u32(smem) + 16 * kPlane + 32, kPlane, 128);
  const uint64_t wb = make_smem_desc(smem_u32(smem) + 2 * 16 * kPlane + 65536, 512, 128);   // ... bottleneck weights as B (N = 32)
  if (warp == 0) {
    const long long t = stream<BLK>(tm, make_idesc_bf16(128, 208), w0, x0, 2 * (kPlane >> 4), nX, &bar[0]);
    if (lane == 0) out[0] = t;
  } else if (warp == 1 && shapeY) {
    long long t;
    if (shapeY == 1) t = stream<BLK>(tm + 256, make_idesc_bf16(128, 208), w0, x1, 2 * (kPlane >> 4), nY, &bar[1]);
    else t = stream<BLK>(tm + 256, make_idesc_bf16(128, 32), x1, wb, 64u >> 4, nY, &bar[1]);
    if (lane == 0) out[1] = t;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
template <int BLK>
void run(long long* d) {
  cudaFuncSetAttribute(probe<BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct C { int shapeY, nX, nY; const char* name; } cs[] = {
    {0, 4800, 0, "X alone (N=208)"}, {1, 2400, 2400, "X (N=208) + Y (N=208)"}, {2, 2400, 4800, "X (N=208) + Y (N=32, 2x count)"}, {2, 48, 4800, "Y alone (N=32)"},
  };
  for (auto& c : cs) {
    cudaMemset(d, 0, 16);
    probe<BLK><<<148, 128, 200 * 1024>>>(c.shapeY, c.nX, c.nY, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double ideal = c.nX * 104.0 + (c.shapeY == 1 ? c.nY * 104.0 : c.shapeY == 2 ? c.nY * 50.0 : 0.0);
    printf("blk %2d  %-34s X done at %8lld, Y done at %8lld cycles  (sum of stand-alone pipe times %.0f)\n", BLK, c.name, h[0], h[1], ideal); fflush(stdout);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<2>(d); run<4>(d); run<12>(d);
  return 0;
}
