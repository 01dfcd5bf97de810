// issue_probe3.cu — development probe: what makes tcgen05.mma issue slow in a loop whose descriptors change every iteration?
// One warp, groups of `per` MMAs (128x208x16, one accumulator); the group's descriptor base advances by a run-time amount.
//   bit 0: the base is derived from a per-thread value (warp index via threadIdx) -> R2UR in the loop; else provably uniform (kernel parameter math)
//   bit 1: tcgen05.commit to a dummy barrier after every group
//   bit 2: mbarrier try_wait on a completed barrier before every group
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392;

template <int PER>
__global__ void __launch_bounds__(640, 1) probe(int what, int groups, int step, long long* out, int busy_warps, int chains, float* sink, int iw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done, dummy, ready;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(&dummy, 1); mbar_init(&ready, 1); fence_mbar_init(); stop = 0; }
  fence_proxy_async_smem();
  if (warp == iw) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) mbar_arrive(&ready);
  __syncthreads();
  if (warp == iw) {
    const uint32_t idesc = make_idesc_bf16(128, 208);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = (2048u >> 4) << 16, b_lbo = ((uint32_t)kPlane >> 4) << 16;
    const uint32_t ring_lo = (smem_u32(smem) + 65536) >> 4, x_lo = (smem_u32(smem) >> 4) + 2;
    auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
    uint32_t wi = (what & 1) ? (uint32_t)(warp - iw) : 0u;      // 0 either way; bit 0: the compiler cannot know
    __syncwarp();
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (what & 4) { mbar_wait(&ready, 0); tc_fence_after(); }
      const uint32_t a = (ring_lo + wi * (uint32_t)step) | a_lbo;
      const uint32_t b = x_lo | b_lbo;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < PER; ++k) umma_bf16(tm, desc(a + (k & 1) * 256u), desc(b + (k & 3) * 2 * (kPlane >> 4)), idesc, 1);
        if (what & 2) umma_commit(&dummy);
      }
      __syncwarp();
      if (++wi == 14) wi = 0;
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if ((warp & 3) == 0 && (iw == 0 ? (warp >> 2) <= busy_warps : (warp >> 2) < busy_warps)) {
    // warps 4, 8, 12, 16 share warp 0's scheduler: `chains` independent FFMA chains each (1 chain = 1 issue slot in 4)
    float a0 = lane, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    while (!stop) {
      if (chains == 1) {
#pragma unroll
        for (int k = 0; k < 128; ++k) a0 = fmaf(a0, 1.0001f, 0.5f);
      } else if (chains == 3) {
#pragma unroll
        for (int k = 0; k < 40; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); }
      } else if (chains == 2) {
#pragma unroll
        for (int k = 0; k < 64; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f); }
      }
    }
    if (a0 + a1 + a2 + a3 == 12345.f) sink[0] = a0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == iw) tmem_dealloc<512>(tm);
}
template <int PER> void run(long long* d, float* sink, int bw, int ch, int iw) {
  cudaFuncSetAttribute(probe<PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int what : {7}) {
    const int groups = 1400;
    probe<PER><<<8, 640, 200 * 1024>>>(what, groups, 512, d, bw, ch, sink, iw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 8; ++i) m += (double)h[i] / 8;
    printf("issuer warp %2d | busy %d warps x %d chains | %d MMAs per group : %.0f cycles/group, %.1f cycles/MMA (pipe floor 104)\n", iw, bw, ch, PER, m / groups, m / groups / PER); fflush(stdout);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  float* sink; cudaMalloc(&sink, 4);
  for (int iw : {0, 16}) for (int bw : {1, 4}) for (int ch : {2, 3, 4}) { run<4>(d, sink, bw, ch, iw); }
  return 0;
}
