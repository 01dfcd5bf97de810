// issue_probe.cu — development probe: cost of the per-stage bookkeeping around tcgen05.mma in the stack kernel's issuer loop.
// One warp issues 128x208x16 MMAs, `per` MMAs per stage; variants add a tcgen05.commit per stage, an mbarrier try_wait on an
// already-completed barrier per stage, and the real thing: a producer thread refilling a ring of 8 KB stages from global memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392, kBuf = 16 * kPlane, kStage = 8192, kNS = 14;

// what: bit 0 commit per stage (to a dummy barrier), bit 1 try_wait per stage (completed barrier), bit 2 real ring with producer
__global__ void __launch_bounds__(128, 1) probe(int what, int per, int stages, const uint8_t* w, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kNS], empty[kNS], done, dummy, ready;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1); mbar_init(&dummy, 1); mbar_init(&ready, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  uint8_t* ring = smem + kBuf;
  if (threadIdx.x == 0) mbar_arrive(&ready);          // phase 0 of `ready` is complete from now on
  __syncthreads();
  if (warp == 1 && (what & 4)) {
    if (lane == 0) {
      uint32_t idx = 0, par = 1;
      for (int st = 0; st < stages; ++st) {
        mbar_wait(&empty[idx], par);
        mbar_expect_tx(&full[idx], kStage);
        bulk_g2s(ring + idx * kStage, w + (size_t)((st * 7 + blockIdx.x) % 160) * kStage, kStage, &full[idx]);
        if (++idx == kNS) { idx = 0; par ^= 1; }
      }
    }
  } else if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 208);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = (2048u >> 4) << 16, b_lbo = ((uint32_t)kPlane >> 4) << 16;
    const uint32_t ring_lo = smem_u32(ring) >> 4, x_lo = (smem_u32(smem) >> 4) + 2;
    uint32_t wi = 0, wp = 0;
    __syncwarp();
    const long long t0 = clock64();
    uint32_t bd = x_lo | b_lbo;
    for (int st = 0; st < stages; ++st) {
      if (what & 4) { mbar_wait(&full[wi], wp); tc_fence_after(); }
      else if (what & 2) { mbar_wait(&ready, 0); tc_fence_after(); }
      const uint32_t a_lo = (ring_lo + wi * (kStage >> 4)) | a_lbo;
      if (elect_one()) {
        for (int k = 0; k < per; ++k)
          umma_bf16(tm, ((uint64_t)desc_hi << 32) | (a_lo + (k & 1) * 256u), ((uint64_t)desc_hi << 32) | (bd + (k & 3) * 2 * (kPlane >> 4)), idesc, 1);
      }
      __syncwarp();
      if (what & 4) { if (elect_one()) umma_commit(&empty[wi]); __syncwarp(); }
      else if (what & 1) { if (elect_one()) umma_commit(&dummy); __syncwarp(); }
      if (++wi == kNS) { wi = 0; wp ^= 1; }
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  uint8_t* w; cudaMalloc(&w, 160 * kStage); cudaMemset(w, 0, 160 * kStage);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct C { int what, per; const char* name; } cs[] = {
    {0, 2, "MMAs only, 2 per stage"}, {1, 2, "+ commit per stage"}, {2, 2, "+ try_wait per stage"}, {3, 2, "+ commit + try_wait"},
    {4, 2, "real ring (producer, 8 KB stages), 2 MMAs per stage"}, {4, 4, "real ring, 4 MMAs per stage (same 8 KB)"},
    {3, 4, "commit + try_wait, 4 per stage"}, {4, 1, "real ring, 1 MMA per stage"},
  };
  for (auto& c : cs) {
    const int stages = 1200;
    probe<<<148, 128, 200 * 1024>>>(c.what, c.per, stages, w, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i] / 148;
    printf("%-55s: %.0f cycles/stage, %.1f cycles/MMA (pipe floor 104)\n", c.name, m / stages, m / stages / c.per); fflush(stdout);
  }
  return 0;
}
