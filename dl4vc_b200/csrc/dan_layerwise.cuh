// dan_layerwise.cuh — layer-by-layer fallback of the bf16 tcgen05 path. Included by dan_bf16.cu.
//
// dan_layer_kernel runs ONE conv layer per launch, fully fused per 128-row tile (dl4vc/model.py:749-778):
//   conv(1x3,dil) -> +bias -> ReLU -> BN  [-> 1x1 residual conv + bias + layer input]  [-> 1x1 bottleneck + bias -> ReLU];
// layer weights stay resident in shared memory, two tiles are in flight per CTA (TMEM double buffer). It serves the
// configurations the fused stack kernel (dan_stack.cuh) does not take — window != 201, a bottleneck width other than 32, a
// residual layer fed directly by a read-mean pool-add — and is never on the PROD path.
#pragma once

namespace {

// =====================================================================================================
// Fused conv-layer kernel
// =====================================================================================================
constexpr int kSlots = 2;
constexpr int kLayerThreads = 320;    // warps 0-3: epilogue slot 0, 4-7: epilogue slot 1, 8: producer, 9: MMA issuer

struct LayerParams {
  const uint4* in; long in_kstride;        // chunk-major input, rows per chunk plane
  uint4* out; long out_kstride;            // chunk-major output (C channels)
  uint4* tout; long t_reads_stride;        // bottleneck output T[p][c8][read][8]
  const uint4* resid;                      // residual source when it differs from the conv input (pool-add layers), else null
  const uint4* wconv; const uint4* wres; const uint4* wbott;   // packed weights (global), smem image
  long rows_total; int num_tiles;
  int pitch, P, gap, dil, kc_in, residual, highway, bott;
  const float* chan;                       // [4][128] conv bias, BN scale, BN shift, residual bias (device)
  const float* bbias;                      // [bott] (device)
};

struct LayerSmem {
  uint64_t w_full, a_full[kSlots], d1_full[kSlots], y_ready[kSlots], d2_full[kSlots], h_ready[kSlots], d3_full[kSlots], slot_free[kSlots];
  uint32_t tmem_base;
};

__host__ __device__ inline size_t layer_smem_bytes(int kc_in, int residual, int highway, int bott, int gap) {
  const size_t slot_rows = 128 + 2 * gap;
  size_t b = 1024;                                           // barriers + tmem pointer
  b += (size_t)3 * kc_in * kC * 16;                          // conv weights
  if (residual) b += (size_t)kKC * kC * 16;
  if (highway) b += (size_t)kKC * bott * 16;
  b += (size_t)kSlots * kKC * slot_rows * 16;                // tile slots (hold the halo'd input, then Y / H in place)
  return b;
}

__device__ __forceinline__ uint32_t word_of(const uint4& v, int w) { return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w)); }

__global__ void __launch_bounds__(kLayerThreads, 1) dan_layer_kernel(const __grid_constant__ LayerParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  LayerSmem* sm = reinterpret_cast<LayerSmem*>(smem);
  const int slot_rows = 128 + 2 * p.gap;
  uint8_t* w_conv = smem + 1024;
  uint8_t* w_res = w_conv + (size_t)3 * p.kc_in * kC * 16;
  uint8_t* w_bott = w_res + (p.residual ? (size_t)kKC * kC * 16 : 0);
  uint8_t* slots = w_bott + (p.highway ? (size_t)kKC * p.bott * 16 : 0);
  const size_t slot_bytes = (size_t)kKC * slot_rows * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&sm->w_full, 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&sm->a_full[s], 1); mbar_init(&sm->d1_full[s], 1); mbar_init(&sm->d2_full[s], 1); mbar_init(&sm->d3_full[s], 1);
      mbar_init(&sm->y_ready[s], 128); mbar_init(&sm->h_ready[s], 128); mbar_init(&sm->slot_free[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(&sm->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;
  const int iters = (p.num_tiles + kSlots * gridDim.x - 1) / (kSlots * gridDim.x);

  if (warp == 8) {
    // ===================== producer: weights once, then one halo'd input tile per (iteration, slot) ==========
    if (lane == 0) {
      const uint32_t conv_bytes = 3u * p.kc_in * kC * 16, res_bytes = p.residual ? kKC * kC * 16 : 0, bott_bytes = p.highway ? kKC * p.bott * 16 : 0;
      mbar_expect_tx(&sm->w_full, conv_bytes + res_bytes + bott_bytes);
      for (uint32_t off = 0; off < conv_bytes; off += 16384) bulk_g2s(w_conv + off, reinterpret_cast<const uint8_t*>(p.wconv) + off, min(16384u, conv_bytes - off), &sm->w_full);
      for (uint32_t off = 0; off < res_bytes; off += 16384) bulk_g2s(w_res + off, reinterpret_cast<const uint8_t*>(p.wres) + off, min(16384u, res_bytes - off), &sm->w_full);
      if (bott_bytes) bulk_g2s(w_bott, p.wbott, bott_bytes, &sm->w_full);
      for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < kSlots; ++s) {
          const int tile = (it * gridDim.x + blockIdx.x) * kSlots + s;
          if (tile >= p.num_tiles) continue;
          mbar_wait(&sm->slot_free[s], (it & 1) ^ 1);
          const uint32_t bytes = (uint32_t)slot_rows * 16;
          mbar_expect_tx(&sm->a_full[s], bytes * p.kc_in);
          const uint4* src = p.in + kLead + (long)tile * 128 - p.gap;
          uint8_t* dst = slots + s * slot_bytes;
          for (int kc = 0; kc < p.kc_in; ++kc) bulk_g2s(dst + (size_t)kc * bytes, src + kc * p.in_kstride, bytes, &sm->a_full[s]);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (one thread) ==============================================================
    if (lane == 0) {
      const uint32_t idesc_c = make_idesc_bf16(128, kC);
      const uint32_t idesc_b = make_idesc_bf16(128, p.bott > 0 ? p.bott : 16);
      const uint32_t lbo_a = (uint32_t)slot_rows * 16;
      mbar_wait(&sm->w_full, 0);
      for (int it = 0; it < iters; ++it) {
        const uint32_t ph = it & 1;
        // conv taps: D1 = sum_t A[rows + (t-1)*dil] * Wt^T
        for (int s = 0; s < kSlots; ++s) {
          const int tile = (it * gridDim.x + blockIdx.x) * kSlots + s;
          if (tile >= p.num_tiles) continue;
          mbar_wait(&sm->a_full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(slots + s * slot_bytes);
          const uint32_t d1 = tmem_base + s * 256;
          uint32_t acc = 0;
          for (int t = 0; t < 3; ++t) {
            const uint32_t a_t = a0 + (uint32_t)(p.gap + (t - 1) * p.dil) * 16;
            const uint32_t b_t = smem_u32(w_conv) + (uint32_t)t * p.kc_in * (kC * 16);
            for (int k2 = 0; k2 < p.kc_in; k2 += 2) {
              umma_bf16(d1, make_smem_desc(a_t + k2 * lbo_a, lbo_a, 128), make_smem_desc(b_t + k2 * (kC * 16), kC * 16, 128), idesc_c, acc);
              acc = 1;
            }
          }
          umma_commit(&sm->d1_full[s]);
        }
        if (p.residual) {
          for (int s = 0; s < kSlots; ++s) {
            const int tile = (it * gridDim.x + blockIdx.x) * kSlots + s;
            if (tile >= p.num_tiles) continue;
            mbar_wait(&sm->y_ready[s], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(slots + s * slot_bytes) + (uint32_t)p.gap * 16;
            const uint32_t d2 = tmem_base + s * 256;
            for (int k2 = 0; k2 < kKC; k2 += 2)
              umma_bf16(d2, make_smem_desc(a0 + k2 * lbo_a, lbo_a, 128), make_smem_desc(smem_u32(w_res) + k2 * (kC * 16), kC * 16, 128), idesc_c, k2 > 0);
            umma_commit(&sm->d2_full[s]);
          }
        }
        if (p.highway) {
          for (int s = 0; s < kSlots; ++s) {
            const int tile = (it * gridDim.x + blockIdx.x) * kSlots + s;
            if (tile >= p.num_tiles) continue;
            mbar_wait(&sm->h_ready[s], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(slots + s * slot_bytes) + (uint32_t)p.gap * 16;
            const uint32_t d3 = tmem_base + s * 256 + 128;
            for (int k2 = 0; k2 < kKC; k2 += 2)
              umma_bf16(d3, make_smem_desc(a0 + k2 * lbo_a, lbo_a, 128), make_smem_desc(smem_u32(w_bott) + k2 * (p.bott * 16), p.bott * 16, 128), idesc_b, k2 > 0);
            umma_commit(&sm->d3_full[s]);
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: thread = one row of the tile =========================================
    const int s = warp >> 2, q = warp & 3;
    const int i = q * 32 + lane;
    const uint32_t d_base = tmem_base + s * 256 + ((uint32_t)(q * 32) << 16);
    uint4* slot_row = reinterpret_cast<uint4*>(slots + s * slot_bytes) + (p.gap + i);
    const bool to_smem = p.residual || p.highway;
    for (int it = 0; it < iters; ++it) {
      const int tile = (it * gridDim.x + blockIdx.x) * kSlots + s;
      if (tile >= p.num_tiles) break;
      const uint32_t ph = it & 1;
      const long m = (long)tile * 128 + i;
      const int pp = (int)(m % p.pitch);
      const bool valid = (m < p.rows_total) && (pp < p.P);
      uint4* out_row = p.out + kLead + m;

      mbar_wait(&sm->d1_full[s], ph);
      tc_fence_after();
      uint4 resid[kKC];
      if (p.residual) {
#pragma unroll
        for (int kc = 0; kc < kKC; ++kc)   // layer input of this row (taken before the in-place overwrite; model.py:732)
          resid[kc] = p.resid ? __ldg(p.resid + kLead + m + kc * p.in_kstride) : slot_row[kc * slot_rows];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(d_base + c * 32, r);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(r[j]) + __ldg(p.chan + c * 32 + j);
          v = fmaxf(v, 0.f);                                             // ReLU, then BatchNorm (model.py:749-751)
          v = fmaf(v, __ldg(p.chan + kC + c * 32 + j), __ldg(p.chan + 2 * kC + c * 32 + j));
          f[j] = valid ? v : 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]); o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
          o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]); o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
          const int kc = c * 4 + g;
          if (!p.residual) out_row[kc * p.out_kstride] = o;
          if (to_smem) slot_row[kc * slot_rows] = o;
        }
      }
      if (to_smem) fence_proxy_async_smem();
      tc_fence_before();
      if (p.residual) {
        mbar_arrive(&sm->y_ready[s]);
        mbar_wait(&sm->d2_full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(d_base + c * 32, r);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t rw = word_of(resid[c * 4 + (j >> 3)], (j & 7) >> 1);
            const float rv = (j & 1) ? bf16_hi(rw) : bf16_lo(rw);
            const float v = __uint_as_float(r[j]) + __ldg(p.chan + 3 * kC + c * 32 + j) + rv;    // 1x1 conv + bias + layer input (model.py:760-761)
            f[j] = valid ? v : 0.f;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]); o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
            o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]); o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
            const int kc = c * 4 + g;
            out_row[kc * p.out_kstride] = o;
            if (p.highway) slot_row[kc * slot_rows] = o;
          }
        }
        if (p.highway) fence_proxy_async_smem();
        tc_fence_before();
      }
      if (p.highway) {
        mbar_arrive(&sm->h_ready[s]);
        mbar_wait(&sm->d3_full[s], ph);
        tc_fence_after();
        const long read = m / p.pitch;
        for (int c = 0; c < p.bott / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(d_base + 128 + c * 32, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(r[g * 8 + j]) + __ldg(p.bbias + c * 32 + g * 8 + j), 0.f);   // model.py:774
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              p.tout[((long)read * (p.bott / 8) + c * 4 + g) * p.P + pp] = o;
            }
          }
        }
        tc_fence_before();
      }
      mbar_arrive(&sm->slot_free[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

}  // namespace
