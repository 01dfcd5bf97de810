// bulk_probe.cu — development probe: L2 -> shared-memory streaming rate of cp.async.bulk (1-D TMA) when every CTA streams
// the same weight image, as dan_stack.cuh does.   args: stage_bytes stages replicas region_kb producers
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* w, size_t replica_stride, int replicas, int region, int stage_bytes, int stages,
                                                int producers, int iters, int mode, int split, int nowait, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4][32];
  if (threadIdx.x == 0) { for (int p = 0; p < 4; ++p) for (int i = 0; i < 32; ++i) mbar_init(&full[p][i], 1); fence_mbar_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < producers && (mode == 1 || lane == 0)) {
    const uint8_t* src = w + (size_t)(blockIdx.x % replicas) * replica_stride;
    uint8_t* ring = smem + (size_t)warp * stages * stage_bytes;
    const long long t0 = clock64();
    size_t off = ((size_t)warp * 7919 * stage_bytes) % region;
    int idx = 0; uint32_t par = 1;
    for (int i = 0; i < iters + stages; ++i) {
      if (i >= stages && !nowait) mbar_wait(&full[warp][idx], par);     // previous fill of this stage landed
      if (i < iters) {
        if (mode == 0 || elect_one()) {
          mbar_expect_tx(&full[warp][idx], stage_bytes);
          for (int c = 0; c < split; ++c)
            bulk_g2s(ring + (size_t)idx * stage_bytes + c * (stage_bytes / split), src + off + c * (stage_bytes / split), stage_bytes / split, &full[warp][idx]);
        }
        if (mode == 1) __syncwarp();
        off += stage_bytes; if (off + stage_bytes > (size_t)region) off = 0;
      }
      if (++idx == stages) { idx = 0; if (i >= stages) par ^= 1; else par = 0; }
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
  }
}

int main(int argc, char** argv) {
  const int region = 640 * 1024;
  uint8_t* w; cudaMalloc(&w, (size_t)region * 64); cudaMemset(w, 1, (size_t)region * 64);
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct C { int stage_bytes, stages, replicas, producers, mode, split, nowait; } cs[] = {
    {8192, 7, 1, 2, 0, 1, 0}, {8192, 7, 1, 2, 0, 1, 1}, {16384, 7, 1, 1, 0, 1, 0}, {16384, 3, 1, 2, 0, 1, 0}, {16384, 3, 4, 2, 0, 1, 0}, {16384, 3, 1, 2, 0, 4, 0}, {32768, 3, 1, 2, 0, 1, 0}, {8192, 5, 1, 4, 0, 1, 0},
  };
  for (int grid : {1, 148}) for (auto& c : cs) {
    const int iters = 2000;
    probe<<<grid, 128, 200 * 1024>>>(w, region, c.replicas, region, c.stage_bytes, c.stages, c.producers, iters, c.mode, c.split, c.nowait, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 4]; cudaMemcpy(h, d, grid * 4 * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < grid; ++i) m += (double)h[i * 4] / grid;
    printf("split %d nowait %d mode %d grid %3d stage %5d B x %2d stages, %2d replicas, %d producers: %.1f B/cycle/SM (%.0f cycles per stage per producer)\n", c.split, c.nowait, c.mode, grid, c.stage_bytes, c.stages,
           c.replicas, c.producers, (double)c.producers * iters * c.stage_bytes / m, m / iters); fflush(stdout);
  }
  return 0;
}
