// issue_probe2.cu — development probe: the stack kernel's conv issue loop (two 8 KB ring stages = 4 MMAs of 128x208x16 per iteration,
// full-barrier wait, 2 commits to the empty barriers, real producer) against variants that take the barrier probe off the path
// between one iteration's commits and the next iteration's first MMA.
//   0 kernel loop: wait2(full) -> 4 MMA -> 2 commit
//   1 probe-ahead: test_wait of iteration i+1's stages BEFORE iteration i's MMAs (blocking wait only when that probe failed)
//   2 as 1, four stages (8 MMAs, 4 commits) per iteration
//   3 as 0, four stages per iteration
//   4 as 1 but ONE commit per iteration (second stage's barrier only; both stages released by it — needs count-1 empties per pair)
// `busy` > 0: that many extra warps on the issuer's scheduler run an FFMA loop (the epilogue warps of the real kernel)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dl4vc_b200/csrc/tcgen05_ptx.cuh"
using namespace ptx;
constexpr int kPlane = 3392, kBuf = 16 * kPlane, kStage = 8192, kNS = 14;

__global__ void __launch_bounds__(640, 1) probe(int mode, int iters, int busy, int reuse, const uint8_t* w, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kNS], empty[kNS], done;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    stop = 0;
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  uint8_t* ring = smem + kBuf;
  const int per = (mode == 2 || mode == 3) ? 4 : 2;        // stages per iteration
  const int stages = iters * per;
  if (warp == 1 || warp == 2) {          // two producer threads take alternate stages, as in the kernel
    if (lane == 0) {
      uint32_t idx = 0, par = 1;
      for (int st = 0; st < stages; ++st) {
        if ((st & 1) == warp - 1) {
          mbar_wait(&empty[mode == 4 ? (idx | 1) : idx], par);   // mode 4: the pair's second barrier releases both
          mbar_expect_tx(&full[idx], kStage);
          bulk_g2s(ring + idx * kStage, w + (size_t)((st * 7 + blockIdx.x) % 160) * kStage, kStage, &full[idx]);
        }
        if (++idx == kNS) { idx = 0; par ^= 1; }
      }
    }
  } else if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 208);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = (2048u >> 4) << 16, b_lbo = ((uint32_t)kPlane >> 4) << 16;
    const uint32_t ring_lo = smem_u32(ring) >> 4, x_lo = (smem_u32(smem) >> 4) + 2;
    constexpr uint32_t kStep = 2 * (kPlane >> 4);
    uint32_t wi = 0, wp = 0;
    auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
    auto probe_ahead = [&](uint32_t i0, uint32_t p0, int n) {     // non-blocking: are the n stages from (i0, p0) full?
      bool ok = true;
      for (int k = 0; k < n; ++k) { ok = ok && mbar_test_wait(&full[i0], p0); if (++i0 == kNS) { i0 = 0; p0 ^= 1; } }
      return ok;
    };
    auto wait_n = [&](uint32_t i0, uint32_t p0, int n) { for (int k = 0; k < n; ++k) { mbar_wait(&full[i0], p0); if (++i0 == kNS) { i0 = 0; p0 ^= 1; } } };
    __syncwarp();
    const long long t0 = clock64();
    const uint32_t bd = x_lo | b_lbo;
    bool ok = false;
    for (int it = 0; it < iters; ++it) {
      if (mode == 0 || mode == 3) { wait_n(wi, wp, per); tc_fence_after(); }
      else {
        if (!ok) wait_n(wi, wp, per);
        tc_fence_after();
        uint32_t ni = wi + per, np = wp; if (ni >= kNS) { ni -= kNS; np ^= 1; }
        ok = it + 1 < iters ? probe_ahead(ni, np, per) : true;
        ok = __all_sync(0xffffffffu, ok);
      }
      if (elect_one()) {
        uint32_t i0 = wi;
        for (int k = 0; k < per; ++k) {
          const uint32_t a = (ring_lo + i0 * (kStage >> 4)) | a_lbo;
          for (int r = 0; r < reuse; ++r) {      // reuse = 2: every stage feeds two reads (the kernel's two slots)
            umma_bf16(tm + r * 256, desc(a), desc(bd + (2 * k) * kStep), idesc, 1);
            umma_bf16(tm + r * 256, desc(a + 256u), desc(bd + (2 * k + 1) * kStep), idesc, 1);
          }
          if (mode != 4 || k == per - 1) umma_commit(&empty[i0]);
          if (++i0 == kNS) i0 = 0;
        }
      }
      __syncwarp();
      wi += per; if (wi >= kNS) { wi -= kNS; wp ^= 1; }
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp >= 4 && (warp & 3) == 0 && (warp >> 2) <= busy) {
    // warps 4, 8, 12, 16 share warp 0's scheduler: dependent-free FFMA streams until the issuer is done
    float a0 = lane, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    while (!stop) {
#pragma unroll
      for (int k = 0; k < 64; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f); }
    }
    if (a0 + a1 + a2 + a3 == 12345.f) sink[0] = a0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148, reuse = argc > 2 ? atoi(argv[2]) : 1;
  long long* d; cudaMalloc(&d, 148 * 8);
  float* sink; cudaMalloc(&sink, 4);
  uint8_t* w; cudaMalloc(&w, 160 * kStage); cudaMemset(w, 0, 160 * kStage);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"kernel loop (wait2, 4 MMA, 2 commit)", "probe-ahead, 4 MMA", "probe-ahead, 8 MMA / 4 stages", "kernel loop, 8 MMA / 4 stages", "probe-ahead, 4 MMA, 1 commit"};
  for (int busy : {0}) {
    for (int mode = 0; mode < 5; ++mode) {
      const int iters = 840;                 // multiple of 7: the ring position returns to 0 for every mode
      const int mmas = iters * ((mode == 2 || mode == 3) ? 8 : 4) * reuse;
      probe<<<grid, 640, 200 * 1024>>>(mode, iters, busy, reuse, w, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < grid; ++i) m += (double)h[i] / grid;
      printf("grid %d reuse %d | ", grid, reuse); printf("busy warps %d | %-40s: %.1f cycles/MMA (pipe floor 104)\n", busy, names[mode], m / mmas); fflush(stdout);
    }
  }
  return 0;
}
