// hmma_probe.cu — development probe: throughput of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16, fp32 accumulate)
// on sm_100a: `warps` warps per SM, each with `ACC` independent accumulator tiles, back-to-back MMAs from registers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ACC>
__global__ void probe(int iters, float* out, long long* cyc) {
  float d[ACC][4] = {};
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) mma16816(d[j], a, b);
  }
  const long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < ACC; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ACC> void run(int warps, float* out, long long* cyc) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<ACC><<<148, warps * 32>>>(100, out, cyc);
  cudaEventRecord(e0);
  probe<ACC><<<148, warps * 32>>>(iters, out, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double flops = 148.0 * warps * iters * ACC * 16 * 8 * 16 * 2;
  printf("warps/SM %2d acc tiles %2d : %.1f TFLOP/s, %.2f cycles per MMA per warp, %.1f MAC/clk/SM\n", warps, ACC, flops / ms / 1e9, (double)h[0] / iters / ACC,
         (double)warps * iters * ACC * 2048 / h[0]);
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int w : {4, 8, 16, 32}) { run<4>(w, out, cyc); run<8>(w, out, cyc); }
  return 0;
}
