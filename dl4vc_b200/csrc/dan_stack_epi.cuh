// dan_stack_epi.cuh — main-accumulator epilogue of the fused conv-stack kernel (dan_stack.cuh); also built stand-alone by
// scripts/probes/epi_probe2.cu. Needs tcgen05_ptx.cuh.
#pragma once

namespace {
using namespace ptx;

enum { kEpiFinal = 0, kEpiPreRes = 1, kEpiPostRes = 2 };

// ---- main-accumulator epilogue (see header comment). A warp owns TMEM lane quadrant q (channels 32q..32q+31) and a
// range of 8-position groups; it walks the range in chunks of two groups (16 positions) with the TMEM load of the
// next chunk in flight while the current one is converted and written. ----
// y = s*relu(z + b) + t (ReLU then BatchNorm, model.py:749-751) is evaluated as one FMNMX and one FFMA:
//   relu(z + b) = max(z, -b) + b   =>   y = s*max(z, -b) + c,   c = s*b + t
struct EpiConsts { float nb[4], scale[4], c[4], rbias[4]; };

__device__ __forceinline__ void stack_epi_load(uint32_t tbase, int g0, uint32_t (&r0)[8], uint32_t (&r1)[8]) {
  tmem_ld_16x256b_x2(tbase + g0 * 8, r0);
  tmem_ld_16x256b_x2(tbase + (16u << 16) + g0 * 8, r1);
}
// tcgen05.wait::ld with the destination registers of the awaited loads as read-write operands: nothing that consumes them
// can be scheduled above the wait
__device__ __forceinline__ void stack_epi_wait(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]) :: "memory");
}

template <int MODE, bool MASK>
__device__ __forceinline__ void stack_epi_chunk(const uint32_t (&r0)[8], const uint32_t (&r1)[8], uint32_t tbase, uint32_t saddr0, int g0, int lane,
                                                int P, const EpiConsts& k) {
  uint32_t x0[8], x1[8];
#pragma unroll
  for (int gi = 0; gi < 2; ++gi) {
    // this thread's row of the four 8x8 blocks (channel chunks 4q..4q+3) of position group g0+gi
    const uint32_t saddr = saddr0 + (uint32_t)(g0 + gi) * 128u;
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t* src = (j < 2) ? r0 : r1;
      float lo = __uint_as_float(src[4 * gi + 2 * (j & 1)]), hi = __uint_as_float(src[4 * gi + 2 * (j & 1) + 1]);
      if constexpr (MODE != kEpiPostRes) {
        lo = fmaf(k.scale[j], fmaxf(lo, k.nb[j]), k.c[j]);
        hi = fmaf(k.scale[j], fmaxf(hi, k.nb[j]), k.c[j]);
      }
      if constexpr (MASK) {                                                   // positions >= P are the zero rows behind the read
        const int pos = 8 * (g0 + gi) + 2 * (lane & 3);
        lo = pos < P ? lo : 0.f; hi = pos + 1 < P ? hi : 0.f;
      }
      pk[j] = pack_bf16x2(lo, hi);
    }
    if constexpr (MODE == kEpiPreRes) {
      uint32_t xin[4];
      ldmatrix_x4_trans(saddr, xin[0], xin[1], xin[2], xin[3]);        // layer input x (model.py:732), same fragment layout
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t* dst = (j < 2) ? x0 : x1;
        dst[4 * gi + 2 * (j & 1)] = __float_as_uint(bf16_lo(xin[j]) + k.rbias[j]);
        dst[4 * gi + 2 * (j & 1) + 1] = __float_as_uint(bf16_hi(xin[j]) + k.rbias[j]);
      }
    }
    stmatrix_x4_trans(saddr, pk[0], pk[1], pk[2], pk[3]);
  }
  if constexpr (MODE == kEpiPreRes) {     // accumulator := x + b_res; the residual 1x1 MMA accumulates on top (model.py:760-761)
    tmem_st_16x256b_x2(tbase + g0 * 8, x0);
    tmem_st_16x256b_x2(tbase + (16u << 16) + g0 * 8, x1);
  }
}

template <int MODE>
__device__ __forceinline__ void stack_epi_do(const uint32_t (&r0)[8], const uint32_t (&r1)[8], uint32_t tbase, uint32_t saddr0, int g0, int lane,
                                             int P, const EpiConsts& k) {
  if (g0 == 24) stack_epi_chunk<MODE, true>(r0, r1, tbase, saddr0, g0, lane, P, k);     // the chunk holding positions >= P
  else stack_epi_chunk<MODE, false>(r0, r1, tbase, saddr0, g0, lane, P, k);
}

// groups [g_begin, g_end), (g_end - g_begin) a multiple of 2, in chunks of two groups (16 positions). Software pipelined:
// the TMEM load of chunk i+1 is in flight while chunk i is converted and written (two register sets, ping-pong).
template <int MODE>
__device__ __forceinline__ void stack_epi_main_pipelined(uint32_t tbase, uint32_t saddr0, int lane, int P, int g_begin, int g_end, const EpiConsts& k) {
  uint32_t a0[8], a1[8], b0[8], b1[8];
  stack_epi_load(tbase, g_begin, a0, a1);
#pragma unroll 1
  for (int g0 = g_begin; g0 < g_end; g0 += 4) {
    stack_epi_wait(a0, a1);
    const bool more = g0 + 2 < g_end;
    if (more) stack_epi_load(tbase, g0 + 2, b0, b1);
    stack_epi_do<MODE>(a0, a1, tbase, saddr0, g0, lane, P, k);
    if (more) {
      stack_epi_wait(b0, b1);
      if (g0 + 4 < g_end) stack_epi_load(tbase, g0 + 4, a0, a1);
      stack_epi_do<MODE>(b0, b1, tbase, saddr0, g0 + 2, lane, P, k);
    }
  }
  if constexpr (MODE == kEpiPreRes) tmem_st_wait();
}


// plain loop (measured: no slower than the pipelined one — the epilogue is bound by ALU-pipe and TMEM-read throughput, not latency)
template <int MODE>
__device__ __forceinline__ void stack_epi_main(uint32_t tbase, uint32_t saddr0, int lane, int P, int g_begin, int g_end, const EpiConsts& k) {
#pragma unroll 1
  for (int g0 = g_begin; g0 < g_end; g0 += 2) {
    uint32_t a0[8], a1[8];
    stack_epi_load(tbase, g0, a0, a1);
    stack_epi_wait(a0, a1);
    stack_epi_do<MODE>(a0, a1, tbase, saddr0, g0, lane, P, k);
  }
  if constexpr (MODE == kEpiPreRes) tmem_st_wait();
}

// Final epilogue with the per-candidate pool bias map (dan_bf16.cu bmap_pack_kernel) instead of the per-channel conv bias:
// y = s * relu(z + B[c][p]) + t, B = conv(pool) + b as bf16 pairs in this thread's fragment order (word gi*4 + j of chunk c).
// `pre` holds chunks 0..3 on entry; chunk c + 4 is requested as soon as chunk c has been read out of its registers.
__device__ __forceinline__ void stack_epi_bmap(uint32_t tbase, uint32_t saddr0, int lane, int P, int g_begin, int nchunks, const EpiConsts& k,
                                               const uint4* bm, uint4 (&pre)[4][2]) {
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    if (c < nchunks) {
      const int g0 = g_begin + 2 * c;
      uint32_t a0[8], a1[8];
      stack_epi_load(tbase, g0, a0, a1);
      const uint4 b0 = pre[c & 3][0], b1 = pre[c & 3][1];
      if (c + 4 < nchunks) { pre[c & 3][0] = __ldg(bm + (c + 4) * kStkBmapChunk); pre[c & 3][1] = __ldg(bm + (c + 4) * kStkBmapChunk + 1); }
      stack_epi_wait(a0, a1);
      const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t* src = (j < 2) ? a0 : a1;
          const int e = 4 * gi + 2 * (j & 1);
          src[e] = __float_as_uint(__uint_as_float(src[e]) + bf16_lo(bw[gi * 4 + j]));
          src[e + 1] = __float_as_uint(__uint_as_float(src[e + 1]) + bf16_hi(bw[gi * 4 + j]));
        }
      }
      stack_epi_do<kEpiFinal>(a0, a1, tbase, saddr0, g0, lane, P, k);
    }
  }
}

// =====================================================================================================================
// Pool epilogue: ALL 16 epilogue warps work on ONE accumulator (128 channels x 208 positions). Four warps share a TMEM
// lane quadrant q (32 channels); warp w4 = 0..3 of the quadrant owns the 16-channel half hh = w4 & 1 and every second
// 16-position pair gp = (w4 >> 1), (w4 >> 1) + 2, ... < 13. One unit = one tcgen05.ld.16x256b.x2 (16 channels x 16
// positions), the epilogue math, one stmatrix.x4.trans (the [channel][position] -> [position][channel] transpose).
// Fragment of a unit: r[4i + {0,1}] = (channel t/4, positions 8i + 2(t%4) + {0,1}); r[4i + {2,3}] = (channel t/4 + 8, same)
// =====================================================================================================================
struct PoolConsts { float nb[2], scale[2], c[2], rbias[2]; };   // [cb]: channel 32q + 16hh + 8cb + lane/4

__device__ __forceinline__ void pool_wait8(uint32_t (&a)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) :: "memory");
}

// taddr0: TMEM address (lane 32q + 16hh, column 0 of the accumulator); saddr0: shared address of this thread's stmatrix row for
// gp = 0: plane 4q + 2hh + ((lane >> 3) & 1), row lead + 8 * (lane >> 4) + (lane & 7)
template <int MODE>
__device__ __forceinline__ void pool_epi(uint32_t taddr0, uint32_t saddr0, int gp0, int lane, int P, const PoolConsts& k) {
#pragma unroll 1
  for (int gp = gp0; gp < 13; gp += 2) {
    uint32_t r[8];
    tmem_ld_16x256b_x2(taddr0 + gp * 16, r);
    pool_wait8(r);
    const uint32_t saddr = saddr0 + (uint32_t)gp * 256u;
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        float lo = __uint_as_float(r[4 * i + 2 * cb]), hi = __uint_as_float(r[4 * i + 2 * cb + 1]);
        if constexpr (MODE != kEpiPostRes) {
          lo = fmaf(k.scale[cb], fmaxf(lo, k.nb[cb]), k.c[cb]);
          hi = fmaf(k.scale[cb], fmaxf(hi, k.nb[cb]), k.c[cb]);
        }
        if (gp == 12) {                                                        // positions >= P are the zero rows behind the read
          const int pos = 192 + 8 * i + 2 * (lane & 3);
          lo = pos < P ? lo : 0.f; hi = pos + 1 < P ? hi : 0.f;
        }
        pk[2 * i + cb] = pack_bf16x2(lo, hi);
      }
    }
    if constexpr (MODE == kEpiPreRes) {      // accumulator := x + b_res (layer input, model.py:732); the residual 1x1 MMA accumulates on top
      uint32_t xin[4], x[8];
      ldmatrix_x4_trans(saddr, xin[0], xin[1], xin[2], xin[3]);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          x[4 * i + 2 * cb] = __float_as_uint(bf16_lo(xin[2 * i + cb]) + k.rbias[cb]);
          x[4 * i + 2 * cb + 1] = __float_as_uint(bf16_hi(xin[2 * i + cb]) + k.rbias[cb]);
        }
      }
      stmatrix_x4_trans(saddr, pk[0], pk[1], pk[2], pk[3]);
      tmem_st_16x256b_x2(taddr0 + gp * 16, x);
    } else {
      stmatrix_x4_trans(saddr, pk[0], pk[1], pk[2], pk[3]);
    }
  }
  if constexpr (MODE == kEpiPreRes) tmem_st_wait();
}
}  // namespace
