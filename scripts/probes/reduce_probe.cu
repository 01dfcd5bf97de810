// reduce_probe.cu — development probe for the fused read-axis pooling (round 2): what does it cost an SM to fold one read's
// activations (128 channels x 201 positions) into per-candidate accumulators that live in L2?
//   (a) cp.reduce.async.bulk .max.bf16 from the bf16 activation planes in shared memory (running max, 16 planes x 3216 B)
//   (b) red.global.add.v4.f32 from registers, thread-private addresses in fragment order (running sum, 28 x 16 B per thread)
//   (c) the same sum as a plain load-add-store of thread-private words (ld.global.cg / st.global)
// All 148 CTAs run at once, each on its own accumulators (like the stack kernel: one candidate per CTA); cycles per read are
// reported for back-to-back issue (throughput) from one warp's point of view, plus the drain time of the last read.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

constexpr int kPlane = 3392, kPlanes = 16, kPlaneBytes = 201 * 16, kReads = 50;

__device__ __forceinline__ void bulk_reduce_max_bf16(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.max.bf16 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(512, 1) probe(uint8_t* gmax, float* gsum, long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slot = warp >> 3, wl = warp & 7;
  for (int i = threadIdx.x; i < 2 * kPlanes * kPlane / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u + (i & 7);
  fence_proxy_async_smem();
  __syncthreads();
  uint8_t* my_max = gmax + (size_t)blockIdx.x * kPlanes * kPlaneBytes;
  float* my_sum = gsum + ((size_t)blockIdx.x * 2 + slot) * (8 * 7 * 32 * 16) + ((size_t)wl * 7 * 32 + lane) * 16;   // [role][chunk][lane][16]
  long long t0 = clock64(), t_issue = 0;
  for (int r = 0; r < kReads; ++r) {
    long long c0 = clock64();
    if (mode == 0) {
      if (lane == 0) {
        for (int kc = 2 * wl; kc < 2 * wl + 2; ++kc) bulk_reduce_max_bf16(my_max + (size_t)kc * kPlaneBytes, smem + (size_t)slot * kPlanes * kPlane + (size_t)kc * kPlane + 32, kPlaneBytes);
        bulk_commit();
        bulk_wait_read0();
      }
      __syncwarp();
    } else if (mode == 1) {
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        float* p = my_sum + (size_t)c * 32 * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add_v4(p + 4 * k, 1.f + r, 2.f, 3.f, 4.f + k);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        float4* p = reinterpret_cast<float4*>(my_sum + (size_t)c * 32 * 16);
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcg(p + k);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k].x += 1.f + r; v[k].y += 2.f; v[k].z += 3.f; v[k].w += 4.f + k; p[k] = v[k]; }
      }
    }
    t_issue += clock64() - c0;
  }
  if (mode == 0 && lane == 0) bulk_wait0();
  __threadfence();
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (t1 - t0) / kReads; out[blockIdx.x * 2 + 1] = t_issue / kReads; }
}

int main() {
  uint8_t* gmax; float* gsum; long long* d;
  const size_t max_b = (size_t)148 * kPlanes * kPlaneBytes, sum_b = (size_t)148 * 2 * 8 * 7 * 32 * 16 * 4;
  cudaMalloc(&gmax, max_b); cudaMemset(gmax, 0, max_b);
  cudaMalloc(&gsum, sum_b); cudaMemset(gsum, 0, sum_b);
  cudaMalloc(&d, 148 * 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPlanes * kPlane);
  const char* names[3] = {"bulk reduce max.bf16 (2 slots x 16 planes x 3216 B per read)", "red.global.add.v4.f32 (2 slots x 114 KB per read)", "ld.cg + add + st (2 slots x 114 KB per read)"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<148, 512, 2 * kPlanes * kPlane>>>(gmax, gsum, d, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d error %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0, sm = 0, is = 0;
    for (int i = 0; i < 148; ++i) { sm += h[2 * i]; is += h[2 * i + 1]; if (h[2 * i] > mx) mx = h[2 * i]; }
    printf("%-70s: %lld cycles per read pair (mean over CTAs; max %lld), issue-side %lld\n", names[mode], sm / 148, mx, is / 148);
  }
  float hs[16]; cudaMemcpy(hs, gsum, sizeof(hs), cudaMemcpyDeviceToHost);
  unsigned short hm[8]; cudaMemcpy(hm, gmax, sizeof(hm), cudaMemcpyDeviceToHost);
  printf("check: sum[0..3] = %g %g %g %g (two kernels x two modes x 50 reads), max[0..3] = %04x %04x %04x %04x\n", hs[0], hs[1], hs[2], hs[3], hm[0], hm[1], hm[2], hm[3]);
  return 0;
}
