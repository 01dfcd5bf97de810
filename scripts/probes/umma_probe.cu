// umma_probe.cu — development probe: cycles per tcgen05.mma for the operand layouts used by dan_stack.cuh.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

struct Cfg { int M, N, a_lbo, b_lbo, b_off, commit_every, iters, mode, acc_stride = 256, a_tiles = 8; };

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (c.mode == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(c.M, c.N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 65536 + c.b_off;
    uint32_t phase = 0;
    for (int i = 0; i < 8; ++i) umma_bf16(tm, make_smem_desc(a0, c.a_lbo, 128), make_smem_desc(b0, c.b_lbo, 128), idesc, 1);
    umma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1;
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      umma_bf16(tm + (i & 1) * 256, make_smem_desc(a0 + (i & 7) * 4096, c.a_lbo, 128), make_smem_desc(b0 + (i & 3) * 2 * c.b_lbo, c.b_lbo, 128), idesc, 1);
      if (c.commit_every && (i % c.commit_every) == c.commit_every - 1) { umma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; }
    }
    umma_commit(&bar); mbar_wait(&bar, phase);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  if (c.mode >= 1 && threadIdx.x < 32) {     // whole warp runs the loop, one elected lane issues
    const uint32_t idesc = make_idesc_bf16(c.M, c.N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 65536 + c.b_off;
    uint32_t phase = 0;
    const uint64_t ad0 = make_smem_desc(a0, c.a_lbo, 128), bd0 = make_smem_desc(b0, c.b_lbo, 128);
    if (elect_one()) { for (int i = 0; i < 8; ++i) umma_bf16(tm, ad0, bd0, idesc, 1); umma_commit(&bar); }
    __syncwarp();
    mbar_wait(&bar, phase); phase ^= 1;
    const long long t0 = clock64();
    const int groups = c.iters / 8;
    for (int g = 0; g < groups; ++g) {
      if (c.mode == 1) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16(tm + (k & 1) * c.acc_stride, ad0 + (uint64_t)((k % c.a_tiles) * 256), bd0 + (uint64_t)((k & 3) * 2 * (c.b_lbo >> 4)), idesc, 1);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) { if (elect_one()) umma_bf16(tm + (k & 1) * c.acc_stride, ad0 + (uint64_t)((k % c.a_tiles) * 256), bd0 + (uint64_t)((k & 3) * 2 * (c.b_lbo >> 4)), idesc, 1); __syncwarp(); }
      }
      if (c.commit_every) { if (elect_one()) umma_commit(&bar); __syncwarp(); mbar_wait(&bar, phase); phase ^= 1; }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, phase);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Cfg cfgs[] = {
    {128, 208, 2048, 3392, 32, 0, 2000, 0},   // single-thread branch (old style)
    {128, 208, 2048, 3392, 32, 0, 2000, 2},   // elect per MMA
    {128, 208, 2048, 3392, 32, 0, 2000, 1},   // stack kernel conv: A=weights, B=activations (tap offset 2 rows)
    {128, 208, 2048, 3392, 0, 0, 2000, 1},    // aligned start
    {128, 208, 2048, 3328, 0, 0, 2000, 1},    // plane stride multiple of 128
    {128, 128, 2048, 2048, 0, 0, 2000, 1},    // N=128
    {128, 256, 2048, 4096, 0, 0, 2000, 1},    // N=256
    {128, 64, 2048, 1024, 0, 0, 2000, 1},     // N=64
    {128, 32, 3392, 512, 0, 0, 2000, 1},      // bottleneck orientation: A=activations, N=32
    {128, 208, 2048, 3392, 32, 8, 2000, 1},   // commit + wait every 8 MMAs (serialised)
    {128, 208, 2048, 3392, 32, 0, 2000, 1, 0, 8},   // ONE accumulator
    {128, 208, 2048, 3392, 32, 0, 2000, 2, 0, 8},   // ONE accumulator, elect per MMA
    {128, 208, 2048, 3392, 32, 0, 2000, 1, 0, 2},   // ONE accumulator, 2 A tiles
    {128, 208, 2048, 3392, 32, 0, 2000, 1, 256, 2}, // two accumulators, 2 A tiles
  };
  for (int grid : {1, 148}) {
    for (auto& c : cfgs) {
      probe<<<grid, 128, 200 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < grid; ++i) m += (double)h[i] / grid;
      printf("mode %d grid %3d  M %3d N %3d a_lbo %4d b_lbo %4d b_off %2d commit_every %d acc_stride %d a_tiles %d : %.1f cycles/MMA (ideal %.0f)\n", c.mode, grid, c.M, c.N, c.a_lbo, c.b_lbo,
             c.b_off, c.commit_every, c.acc_stride, c.a_tiles, m / c.iters, 128.0 * c.N / 256.0);
    }
  }
  return 0;
}
