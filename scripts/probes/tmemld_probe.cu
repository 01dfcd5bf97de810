// tmemld_probe.cu — development probe: TMEM -> register read throughput for the shapes the stack-kernel epilogue can use.
// One "epilogue" = 128 lanes x 208 fp32 columns. warps = epilogue warps running concurrently (warp w uses lane quadrant w%4).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;

__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// mode 0: 16x256b.x2 pairs, wait after each pair (current kernel) | 1: .x4 pairs | 2: .x8 pairs | 3: 32x32b.x32
// mode +16: software pipelined (next load issued before consuming the current registers)
__global__ void __launch_bounds__(1024, 1) probe(int mode, int warps, int iters, long long* out) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp < warps) {
    const int q = warp & 3, part = warp >> 2, parts = warps / 4;      // `parts` warps share a lane quadrant, splitting the columns
    const uint32_t tbase = tm + ((uint32_t)(32 * q) << 16);
    const int cols = 416 / parts, c0 = part * cols;                    // two accumulators' worth of columns split between the parts
    float acc = 0.f;
    __syncwarp();
    const long long t0 = clock64();
    const int shape = mode & 15; const bool pipe = mode & 16;
    for (int it = 0; it < iters; ++it) {
      if (shape == 0) {
        for (int c = c0; c < c0 + cols; c += 16) {
          uint32_t a[8], b[8];
          tmem_ld_16x256b_x2(tbase + c, a); tmem_ld_16x256b_x2(tbase + (16u << 16) + c, b); tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
      } else if (shape == 1) {
        if (!pipe) {
          for (int c = c0; c < c0 + cols; c += 32) {
            uint32_t a[16], b[16];
            tmem_ld_16x256b_x4(tbase + c, a); tmem_ld_16x256b_x4(tbase + (16u << 16) + c, b); tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]);
          }
        } else {
          uint32_t a[16], b[16], a2[16], b2[16];
          tmem_ld_16x256b_x4(tbase + c0, a); tmem_ld_16x256b_x4(tbase + (16u << 16) + c0, b); tmem_ld_wait();
          for (int c = c0; c < c0 + cols; c += 64) {
            tmem_ld_16x256b_x4(tbase + c + 32, a2); tmem_ld_16x256b_x4(tbase + (16u << 16) + c + 32, b2);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]);
            tmem_ld_wait();
            if (c + 64 < c0 + cols) { tmem_ld_16x256b_x4(tbase + c + 64, a); tmem_ld_16x256b_x4(tbase + (16u << 16) + c + 64, b); }
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += __uint_as_float(a2[j]) + __uint_as_float(b2[j]);
            tmem_ld_wait();
          }
        }
      } else if (shape == 2) {
        for (int c = c0; c < c0 + cols; c += 64) {
          uint32_t a[32], b[32];
          ld_16x256b_x8(tbase + c, a); ld_16x256b_x8(tbase + (16u << 16) + c, b); tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
      } else {
        for (int c = c0; c < c0 + cols; c += 32) {
          uint32_t a[32];
          tmem_ld32(tbase + c, a); tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += __uint_as_float(a[j]);
        }
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) / iters;
    if (acc == 12345.f) out[100] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 128 * 8); cudaMemset(d, 0, 1024);
  struct C { int mode, warps; const char* name; } cs[] = {
    {0, 4, "16x256b.x2 pairs"}, {0, 8, "16x256b.x2 pairs"}, {0, 16, "16x256b.x2 pairs"},
    {1, 4, "16x256b.x4 pairs"}, {1, 8, "16x256b.x4 pairs"}, {1, 16, "16x256b.x4 pairs"},
    {17, 4, "16x256b.x4 pairs pipelined"}, {17, 8, "16x256b.x4 pairs pipelined"},
    {2, 4, "16x256b.x8 pairs"}, {2, 8, "16x256b.x8 pairs"},
    {3, 4, "32x32b.x32"}, {3, 8, "32x32b.x32"}, {3, 16, "32x32b.x32"},
  };
  for (auto& c : cs) {
    probe<<<1, 1024>>>(c.mode, c.warps, 200, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < c.warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-30s %2d warps: %6lld cycles to read 2 accumulators (128 x 416 fp32 = 208 KB) -> %.1f B/cycle/SM\n", c.name, c.warps, mx, 212992.0 / mx); fflush(stdout);
  }
  return 0;
}
