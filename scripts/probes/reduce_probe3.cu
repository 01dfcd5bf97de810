// reduce_probe3.cu (variant of reduce_probe2.cu: ONE bulk instruction per read image of 54 272 B instead of 16 plane-sized ones)
// reduce_probe2.cu — development probe: LATENCY (issue -> shared-memory source free) of TMA bulk stores / bulk reductions of one
// read's activation planes under the stack kernel's real duty cycle (one read per slot every ~70 k cycles, 148 CTAs), as opposed to
// reduce_probe.cu which measures saturated throughput.
//   0 bulk store 16 x 3216 B (the H write-back of round 1)      1 bulk reduce max.bf16 16 x 3216 B
//   2 bulk reduce add.noftz.f16 16 x 3216 B                        3 bulk reduce add.f32 2 x (8 x 6432 B) with a wait in between
//   4 max.bf16 followed by add.f16 on the same planes (one commit group)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

constexpr int kPlane = 3392, kPlanes = 16, kPlaneBytes = 201 * 16, kReads = 40;

#define BULK_RED(name, op) \
  __device__ __forceinline__ void name(void* gdst, const void* ssrc, uint32_t bytes) { \
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group" op " [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory"); }
BULK_RED(red_max_bf16, ".max.bf16")
BULK_RED(red_add_f16, ".add.noftz.f16")
BULK_RED(red_add_f32, ".add.f32")

__global__ void __launch_bounds__(512, 1) probe(uint8_t* g0, uint8_t* g1, long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slot = warp >> 3, wl = warp & 7;
  for (int i = threadIdx.x; i < 2 * kPlanes * kPlane / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = mode == 3 ? 0x3F800000u : (mode == 2 ? 0x3C003C00u : 0x3F803F80u);
  fence_proxy_async_smem();
  __syncthreads();
  uint8_t* a = g0 + ((size_t)blockIdx.x * 2 + slot) * 2 * kPlanes * kPlane;
  uint8_t* b = g1 + ((size_t)blockIdx.x * 2 + slot) * 2 * kPlanes * kPlane;
  const uint8_t* src = smem + (size_t)slot * kPlanes * kPlane;
  long long t_sum = 0, t_max = 0;
  for (int r = 0; r < kReads; ++r) {
    // duty cycle: slot s works in the s-th half of a ~70 k cycle period
    const long long until = clock64() + 35000;
    while (clock64() < until) __nanosleep(200);
    if ((r & 1) != slot) continue;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
    const long long c0 = clock64();
    if (lane == 0 && wl == 0) {
      const uint32_t bytes = kPlanes * kPlane;
      if (mode == 0) bulk_s2g(a, src, bytes);
      else if (mode == 1 || mode == 4) red_max_bf16(a, src, bytes);
      else if (mode == 2) red_add_f16(a, src, bytes);
      else red_add_f32(a, src, bytes);
      if (mode == 4) red_add_f16(b, src, bytes);
      bulk_commit(); bulk_wait_read0();
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
    const long long dt = clock64() - c0;
    t_sum += dt; if (dt > t_max) t_max = dt;
  }
  if (lane == 0) bulk_wait0();
  if ((threadIdx.x & 255) == 0) { out[(blockIdx.x * 2 + slot) * 2] = t_sum / (kReads / 2); out[(blockIdx.x * 2 + slot) * 2 + 1] = t_max; }
}

int main() {
  uint8_t *g0, *g1; long long* d;
  const size_t bytes = (size_t)148 * 2 * 2 * kPlanes * kPlane;
  cudaMalloc(&g0, bytes); cudaMemset(g0, 0, bytes);
  cudaMalloc(&g1, bytes); cudaMemset(g1, 0, bytes);
  cudaMalloc(&d, 148 * 4 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPlanes * kPlane);
  const char* names[5] = {"bulk store 51 KB", "bulk reduce max.bf16 51 KB", "bulk reduce add.noftz.f16 51 KB", "bulk reduce add.f32 2 x 51 KB (sequential halves)", "max.bf16 + add.f16 (same planes)"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<148, 512, 2 * kPlanes * kPlane>>>(g0, g1, d, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d error %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h[148 * 4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long sm = 0, mx = 0;
    for (int i = 0; i < 296; ++i) { sm += h[2 * i]; if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; }
    printf("%-55s: issue -> source free %lld cycles (mean over slots), worst %lld\n", names[mode], sm / 296, mx);
  }
  return 0;
}
