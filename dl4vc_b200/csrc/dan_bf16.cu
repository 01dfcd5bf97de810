// dan_bf16.cu — tcgen05 implementation of the DAN forward (bf16 operands, fp32 accumulation in TMEM).
//
// Activations are stored "chunk-major": for each group of 8 channels (one 16-byte piece per row) the rows of the
// pass form a dense array:   X[kc][row][8]  (bf16), rows = (candidate, read, position) with `gap` zero rows between
// reads (RowGeom). A tile of 128 consecutive rows (+ halo) of one channel chunk is therefore ONE contiguous run in
// HBM, moved by a single bulk async copy, and lands in shared memory exactly in the row-linear UMMA operand layout
// described in tcgen05_ptx.cuh — the three dilated taps of a conv layer are three start addresses on that tile.
//
// Kernels
//   dan_layer_kernel   one conv layer, fully fused per 128-row tile (dl4vc/model.py:749-778):
//                      conv(1x3,dil) -> +bias -> ReLU -> BN  [-> 1x1 residual conv + bias + layer input]
//                      [-> 1x1 bottleneck + bias -> ReLU]; layer weights stay resident in shared memory, two tiles
//                      are in flight per CTA (TMEM double buffer) so the tensor pipe runs under the epilogue.
//   stream_gemm_kernel generic split-K GEMM with both operands streamed (highway compression (1x201) conv as one
//                      K=6432 GEMM over reads, FC trunk, heads).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "dan_kernels_common.cuh"
#include "tcgen05_ptx.cuh"

namespace {

using namespace ptx;

constexpr int kC = 128;               // channels (bf16 path is specialised for the shipped width)
constexpr int kKC = kC / 8;           // 16-byte pieces per row
constexpr int kLead = 8;              // zero rows in front of every row matrix (>= largest dilation)
// Every CTA of the fused kernel streams the same few hundred KB of layer weights at about the same time; CTA c reads replica
// c % kWeightReplicas so that the reads spread over more L2 slices. (With per-slot rings 16 copies helped; with the shared ring,
// the evict-last L2 policy and two producer threads 1-4 copies measure the same and 16+ slightly worse.)
constexpr int kWeightReplicas = 4;

// =====================================================================================================
// Encoder (bf16, chunk-major output)
// =====================================================================================================
__global__ void __launch_bounds__(256) encode_rows_bf16_kernel(EncodeParams p, long cand0, uint4* __restrict__ out, long kstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long cand = cand0 + blockIdx.x;
  EncodeSmem s = encode_stage(p, cand, smem_raw);
  const int rows = p.g.R * p.g.pitch;
  const int kcs = p.CinPad / 8;
  const long row0 = (long)blockIdx.x * rows;
  for (int idx = threadIdx.x; idx < rows * kcs; idx += blockDim.x) {
    const int kc = idx / rows, row = idx - kc * rows;
    const int r = row / p.g.pitch, pp = row - r * p.g.pitch;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pp < p.g.P) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (kc * 8 + j < p.Cin) ? encode_channel(p, s, kc * 8 + j, pp, r) : 0.f;
      v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]); v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    }
    out[kc * kstride + kLead + row0 + row] = v;
  }
}

// ---- fast encoder for the shipped channel set (embed_dim 20, q-scores, strands, ref/var masks: 45 -> 48 channels) ----
// Two launches: (1) per-read agreement bits with the ref / var proposal (integer compare over all positions,
// model.py:592-593,607-608); (2) one CTA per (16-position chunk, candidate): the three byte tiles of the chunk are
// contiguous in the loader's [position][read] layout (dataset.py:672-680), the embedding + positional sums are built
// once per (position, token) in shared memory as bf16, and every 16-byte output piece is one shared-memory read.
constexpr int kEncPB = 16;
__global__ void __launch_bounds__(128) agree_bits_kernel(DevInputs in, long cand0, int P, int R, uint8_t* __restrict__ agree) {
  const long cand = cand0 + blockIdx.x;
  __shared__ uint8_t rm[512], vm[512];
  for (int i = threadIdx.x; i < P; i += blockDim.x) { rm[i] = in.ref_masks[cand * P + i]; vm[i] = in.var_masks[cand * P + i]; }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    bool okR = true, okV = true;
    const uint8_t* rd = in.reads + cand * P * R + r;
    for (int pp = 0; pp < P; ++pp) {
      const uint8_t a = rm[pp], b = vm[pp];
      if (a | b) {
        const uint8_t t = __ldg(rd + (long)pp * R);
        okR = okR && (a == 0 || t == a);
        okV = okV && (b == 0 || t == b);
      }
    }
    agree[((long)blockIdx.x * 2 + 0) * R + r] = okR;
    agree[((long)blockIdx.x * 2 + 1) * R + r] = okV;
  }
}

__global__ void __launch_bounds__(256) encode_prod_bf16_kernel(DevInputs in, const float* __restrict__ emb, const float* __restrict__ pe,
                                                               const uint8_t* __restrict__ agree, long cand0, RowGeom g,
                                                               uint4* __restrict__ out, long kstride) {
  constexpr int D = 20, PB = kEncPB;
  const int P = g.P, R = g.R;
  const int cl = blockIdx.y;
  const long cand = cand0 + cl;
  const int p0 = blockIdx.x * PB;
  const int np = min(PB, P - p0);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* tabR = reinterpret_cast<uint4*>(smem_raw);                 // [PB][10][3] pieces: ch 0-7 | 8-15 | 16-19 + ref 0-3
  uint4* tabF = tabR + PB * DAN_VOCAB * 3;                           // [PB][2] pieces: ref 4-11 | ref 12-19
  uint8_t* rd = reinterpret_cast<uint8_t*>(tabF + PB * 2);           // [PB][R]
  uint8_t* qq = rd + PB * R;
  uint8_t* ss = qq + PB * R;
  uint8_t* fl = ss + PB * R;                                         // [PB] ref token, [PB] nzR, [PB] nzV
  uint8_t* ag = fl + 3 * PB;                                         // [2][R]
  const long toff = cand * P * R + (long)p0 * R;
  for (int i = threadIdx.x; i < np * R; i += blockDim.x) { rd[i] = __ldg(in.reads + toff + i); qq[i] = __ldg(in.q + toff + i); ss[i] = __ldg(in.strands + toff + i); }
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    fl[i] = __ldg(in.ref + cand * P + p0 + i);
    fl[PB + i] = __ldg(in.ref_masks + cand * P + p0 + i) != 0;
    fl[2 * PB + i] = __ldg(in.var_masks + cand * P + p0 + i) != 0;
  }
  for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) ag[i] = agree[(long)cl * 2 * R + i];
  __syncthreads();
  __nv_bfloat16* tR = reinterpret_cast<__nv_bfloat16*>(tabR);
  for (int i = threadIdx.x; i < np * DAN_VOCAB * 24; i += blockDim.x) {
    const int c = i % 24, tok = (i / 24) % DAN_VOCAB, pl = i / (24 * DAN_VOCAB);
    const int pp = p0 + pl;
    const float v = c < D ? emb[tok * D + c] + pe[pp * D + c] : emb[fl[pl] * D + (c - D)] + pe[pp * D + (c - D)];
    tR[i] = __float2bfloat16_rn(v);
  }
  __nv_bfloat16* tF = reinterpret_cast<__nv_bfloat16*>(tabF);
  for (int i = threadIdx.x; i < np * 16; i += blockDim.x) {
    const int c = 4 + (i & 15), pl = i >> 4;
    tF[i] = __float2bfloat16_rn(emb[fl[pl] * D + c] + pe[(p0 + pl) * D + c]);
  }
  __syncthreads();
  const long row_base = kLead + (long)cl * R * g.pitch + p0;
  const int items = 6 * R * PB;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int pl = i % PB, r = (i / PB) % R, kc = i / (PB * R);
    if (pl >= np) continue;
    uint4 v;
    if (kc < 3) v = tabR[(pl * DAN_VOCAB + rd[pl * R + r]) * 3 + kc];
    else if (kc < 5) v = tabF[pl * 2 + (kc - 3)];
    else {
      const float m0 = (fl[PB + pl] && ag[r]) ? 1.f : 0.f, m1 = (fl[2 * PB + pl] && ag[R + r]) ? 1.f : 0.f, m2 = fl[PB + pl] ? 1.f : 0.f;
      v.x = pack_bf16x2((float)qq[pl * R + r] * 0.01f, (float)ss[pl * R + r] * 0.5f);   // model.py:24,16
      v.y = pack_bf16x2(m0, m1); v.z = pack_bf16x2(m2, 0.f); v.w = 0u;                 // var_length from the REF mask (model.py:579,584)
    }
    out[kc * kstride + row_base + (long)r * g.pitch + pl] = v;
  }
}
__host__ __device__ inline size_t encode_prod_smem_bytes(int R) {
  return (size_t)kEncPB * DAN_VOCAB * 3 * 16 + kEncPB * 2 * 16 + 3 * kEncPB * R + 3 * kEncPB + 2 * R + 16;
}

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
  float bias[kC], scale[kC], shift[kC], rbias[kC], bbias[64];
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
          float v = __uint_as_float(r[j]) + p.bias[c * 32 + j];
          v = fmaxf(v, 0.f);                                             // ReLU, then BatchNorm (model.py:749-751)
          v = fmaf(v, p.scale[c * 32 + j], p.shift[c * 32 + j]);
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
            const float v = __uint_as_float(r[j]) + p.rbias[c * 32 + j] + rv;    // 1x1 conv + bias + layer input (model.py:760-761)
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
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(r[g * 8 + j]) + p.bbias[c * 32 + g * 8 + j], 0.f);   // model.py:774
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
#include "dan_stack.cuh"
#include "dan_gemm.cuh"
namespace {

// =====================================================================================================
// Elementwise / reduction helpers on chunk-major bf16
// =====================================================================================================
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// loads in flight per thread of the pooling kernels: they run as one 64-thread CTA per SM beside the persistent stack kernel
// (4096 spare registers per SM), so memory-level parallelism has to come from the thread itself
constexpr int kPoolUnroll = 8;
constexpr size_t kBmapBytesPerCand = (size_t)kStkBmapPerCand * 16;

// pool[cand][c/8][p][c%8] = mean over reads (fp32)   (model.py:772)
// im2col (optional): the bf16 mean also goes out as the A operand of the pool-bias-map GEMM, row (cand*P + p'), columns tap*C + c holding
// pool[c][p' + (tap-1)*dil] (zero outside the read: Conv2d zero padding, model.py:214-229)
__global__ void __launch_bounds__(64) pool_mean_bf16_kernel(const uint4* __restrict__ h, long kstride, float* __restrict__ pool, RowGeom g,
                                                            uint4* __restrict__ im2col = nullptr, int dil = 0) {
  const int cand = blockIdx.y, kc = blockIdx.z;
  const int pp = blockIdx.x * blockDim.x + threadIdx.x;
  if (pp >= g.P) return;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* src = h + kc * kstride + kLead + (long)cand * g.R * g.pitch + pp;
  int r = 0;
  for (; r + kPoolUnroll <= g.R; r += kPoolUnroll) {
    uint4 v[kPoolUnroll];
#pragma unroll
    for (int u = 0; u < kPoolUnroll; ++u) v[u] = __ldg(src + (long)(r + u) * g.pitch);
#pragma unroll
    for (int u = 0; u < kPoolUnroll; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
  }
  for (; r < g.R; ++r) {
    float f[8];
    unpack8(__ldg(src + (long)r * g.pitch), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
  float4* dst = reinterpret_cast<float4*>(pool + (((long)cand * kKC + kc) * g.P + pp) * 8);
  const float inv = 1.f / (float)g.R;
  dst[0] = make_float4(s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv);
  dst[1] = make_float4(s[4] * inv, s[5] * inv, s[6] * inv, s[7] * inv);
  if (im2col) {
    const uint4 v = make_uint4(pack_bf16x2(s[0] * inv, s[1] * inv), pack_bf16x2(s[2] * inv, s[3] * inv), pack_bf16x2(s[4] * inv, s[5] * inv), pack_bf16x2(s[6] * inv, s[7] * inv));
    uint4* row0 = im2col + (long)cand * g.P * (3 * kKC) + kc;          // 3*kKC 16-byte pieces per row
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      const int pd = pp - (tap - 1) * dil;                             // the output position that reads this value through `tap`
      if (pd >= 0 && pd < g.P) row0[(long)pd * (3 * kKC) + tap * kKC] = v;
    }
    if (pp < dil) row0[(long)pp * (3 * kKC)] = make_uint4(0, 0, 0, 0);                          // tap 0 reaches in front of the read
    if (pp >= g.P - dil) row0[(long)pp * (3 * kKC) + 2 * kKC] = make_uint4(0, 0, 0, 0);         // tap 2 reaches behind it
  }
}

// ---- pool bias map (fused stack path). The read-mean pool-add in front of a conv layer (model.py:734-742) is linear:
//   conv(x + pool) + b = conv(x) + [conv(pool) + b],
// and the bracket is the same for all reads of a candidate. It is computed once per candidate on the tensor cores (tma_gemm over the
// im2col rows above) and stored as bf16 pairs in exactly the order in which the stack kernel's epilogue threads hold the accumulator
// (tcgen05.ld 16x256b fragments, dan_stack_epi.cuh): per candidate 8 roles (position half h, lane quadrant q) x 7 chunks of 16
// positions x 32 lanes x 8 words; word gi*4 + j = channel 32q + 8j + lane/4, positions 8*(g0 + gi) + 2*(lane%4) + {0,1} with
// g0 = (h ? 14 : 0) + 2*chunk. The layer's epilogue adds it instead of the per-channel bias; the per-read pool-add (103 KB of L2
// reads and a 54 KB shared-memory read-modify-write on every read boundary) disappears.
__global__ void bmap_pack_kernel(const float* __restrict__ gmap, const float* __restrict__ bias, uint4* __restrict__ out, int cands, int P) {
  const long total = (long)cands * 8 * 7 * 32 * 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int half = (int)(i & 1), lane = (int)((i >> 1) & 31);
    long t = i >> 6;
    const int chunk = (int)(t % 7); t /= 7;
    const int role = (int)(t % 8); const long cand = t / 8;
    const int h = role >> 2, q = role & 3, g0 = (h ? 14 : 0) + 2 * chunk, gi = half;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 32 * q + 8 * j + (lane >> 2), pos = 8 * (g0 + gi) + 2 * (lane & 3);
      const float b = bias[c];
      const float lo = pos < P ? gmap[(cand * P + pos) * kC + c] + b : 0.f;
      const float hi = pos + 1 < P ? gmap[(cand * P + pos + 1) * kC + c] + b : 0.f;
      w[j] = pack_bf16x2(lo, hi);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// out = bf16(h + pool) on data rows, 0 on gap rows   (model.py:742)
__global__ void add_pool_bf16_kernel(const uint4* __restrict__ h, const float* __restrict__ pool, uint4* __restrict__ out,
                                     long kstride, long rows, RowGeom g) {
  const long total = rows * kKC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kc = (int)(i / rows); const long row = i - (long)kc * rows;
    const int pp = (int)(row % g.pitch); const long cand = row / ((long)g.R * g.pitch);
    uint4 v = h[kc * kstride + kLead + row];
    if (pp < g.P) {
      float f[8];
      unpack8(v, f);
      const float4* a = reinterpret_cast<const float4*>(pool + ((cand * kKC + kc) * g.P + pp) * 8);
      const float4 a0 = a[0], a1 = a[1];
      f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w; f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
      v = pack8(f);
    }
    out[kc * kstride + kLead + row] = v;
  }
}

// final max ‖ mean over reads -> FC input pieces (bf16 feature order: max p*C+c | mean P*C+p*C+c)
__global__ void __launch_bounds__(64) pool_final_bf16_kernel(const uint4* __restrict__ h, long kstride, uint4* __restrict__ fcin, long fc_kstride,
                                       int cand0, RowGeom g, int skip_max) {
  const int cand = blockIdx.y, kc = blockIdx.z;
  const int pp = blockIdx.x * blockDim.x + threadIdx.x;
  if (pp >= g.P) return;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
  const uint4* src = h + kc * kstride + kLead + (long)cand * g.R * g.pitch + pp;
  int r = 0;
  for (; r + kPoolUnroll <= g.R; r += kPoolUnroll) {
    uint4 v[kPoolUnroll];
#pragma unroll
    for (int u = 0; u < kPoolUnroll; ++u) v[u] = __ldg(src + (long)(r + u) * g.pitch);
#pragma unroll
    for (int u = 0; u < kPoolUnroll; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; mx[j] = fmaxf(mx[j], f[j]); }
    }
  }
  for (; r < g.R; ++r) {
    float f[8];
    unpack8(__ldg(src + (long)r * g.pitch), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; mx[j] = fmaxf(mx[j], f[j]); }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] /= (float)g.R;
  uint4* row = fcin + (long)(cand0 + cand) * fc_kstride;          // fc_kstride = pieces per FC-input row
  if (skip_max) {
    row[(long)pp * kKC + kc] = pack8(s);
  } else {
    row[(long)pp * kKC + kc] = pack8(mx);
    row[(long)(g.P + pp) * kKC + kc] = pack8(s);
  }
}

// highway: hw fp32 partial [l][read][bott] + bias -> relu -> FC input pieces   (model.py:853-859)
__global__ void highway_finish_bf16_kernel(const float* __restrict__ hw, long layer_stride, const float* const* __restrict__ bias_by_layer,
                                           int L, int bott, int R, int concat, uint4* __restrict__ fcin, long fc_kstride,
                                           long base_kc, int cand0, int cands) {
  const int o8n = bott / 8;
  const long total = (long)cands * (concat ? L : 1) * R * o8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int o8 = (int)(i % o8n); long t = i / o8n;
    const int r = (int)(t % R); t /= R;
    const int l = concat ? (int)(t % L) : 0; const long cand = concat ? t / L : t;
    float f[8];
    if (concat) {
      const float* src = hw + l * layer_stride + (cand * R + r) * bott + o8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(src[j] + bias_by_layer[l][o8 * 8 + j], 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      for (int k = 0; k < L; ++k) {
        const float* src = hw + k * layer_stride + (cand * R + r) * bott + o8 * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += src[j] + bias_by_layer[k][o8 * 8 + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j] / (float)L, 0.f);
    }
    fcin[(long)(cand0 + cand) * fc_kstride + base_kc + ((long)l * R + r) * o8n + o8] = pack8(f);
  }
}

// =====================================================================================================
// Weight packing (fp32 state_dict tensors -> bf16 operand images)
// =====================================================================================================
__device__ __forceinline__ void store_bf16(uint4* base, long piece, int j, float v) {
  reinterpret_cast<__nv_bfloat16*>(base + piece)[j] = __float2bfloat16_rn(v);
}
// conv (Cout, Cin, 1, 3) -> [tap][kc][n][8]
__global__ void pack_conv_bf16_kernel(const float* __restrict__ w, uint4* __restrict__ out, int Cin, int kc_in) {
  const long total = (long)3 * kc_in * kC * 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7); long t = i >> 3;
    const int n = (int)(t % kC); t /= kC;
    const int kc = (int)(t % kc_in); const int tap = (int)(t / kc_in);
    const int c = kc * 8 + j;
    store_bf16(out, ((long)tap * kc_in + kc) * kC + n, j, c < Cin ? w[((long)n * Cin + c) * 3 + tap] : 0.f);
  }
}
// linear-like (N, K) fp32 with optional source-column map -> bf16 row-major [Npad][Kpad] (mode 3: UMMA chunk image [kc][Npad][8])
__global__ void pack_linear_bf16_kernel(const float* __restrict__ w, uint4* __restrict__ out, int N, int Npad, int K, int Kpad,
                                        int mode, int P, int C, int R, int bott, int L, int pooled, int skip_max) {
  // mode 0: identity columns. mode 1: FC1 feature permutation (see dan_bf16_forward). mode 2: compression (O, Cb, 1, P): k = ((c/8)*P + p)*8 + c%8
  // mode 4: conv weight (Cout, C, 1, 3) with k = tap*C + cin
  const long total = (long)(Kpad / 8) * Npad * 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7); long t = i >> 3;
    const int n = (int)(t % Npad); const long kc = t / Npad;
    const long k = kc * 8 + j;
    float v = 0.f;
    if (n < N && k < K) {
      long src = k;
      if (mode == 1) {
        if (k < pooled) {
          const long blk = k / ((long)P * C);            // 0 = max block (or mean when skip_max), 1 = mean block
          const long rem = k - blk * P * C;
          const int pp = (int)(rem / C), c = (int)(rem % C);
          src = (blk * C + c) * (long)P + pp;
        } else {
          const long h = k - pooled;
          const int o = (int)(h % bott); const long lr = h / bott;
          const int r = (int)(lr % R); const int l = (int)(lr / R);
          src = pooled + (long)l * bott * R + (long)o * R + r;
        }
      } else if (mode == 2) {                             // K order of the T matrix: (c / 8, p, c % 8)
        const int g = (int)(k / (P * 8)), rem = (int)(k % (P * 8));
        const int pp = rem / 8, c = g * 8 + rem % 8;
        src = (long)c * P + pp;                           // within row n: (c, p)
      } else if (mode == 4) {
        src = (k % C) * 3 + k / C;
      }
      v = w[(long)n * K + src];
    }
    if (mode == 3) store_bf16(out, kc * Npad + n, j, v);
    else reinterpret_cast<__nv_bfloat16*>(out)[(long)n * Kpad + k] = __float2bfloat16_rn(v);
  }
}

struct Bf16Weights {
  uint4* wconv[DAN_MAX_LAYERS]; uint4* wres[DAN_MAX_LAYERS]; uint4* wbott[DAN_MAX_LAYERS]; uint4* wcomp[DAN_MAX_LAYERS];
  uint4* wcomp_all;                   // all layers' compression weights, [L][bott][P*bott] bf16 (one batched GEMM operand)
  uint4* fcw[DAN_MAX_FC]; uint4* headw;
  uint4* wbmap[DAN_MAX_LAYERS];       // layers fed by a read-mean pool-add: conv weight as a row-major GEMM operand [cout][tap*C + cin] bf16 (pool bias map)
  const float** comp_bias_ptrs;       // device array of L pointers
  uint8_t* wstream[DAN_MAX_LAYERS];   // conv | residual | bottleneck operand images of a layer, contiguous (dan_stack.cuh); kWeightReplicas copies
  size_t wstream_bytes[DAN_MAX_LAYERS];   // bytes of one copy (256-byte multiple)
  float* chan_dev;                    // [L][4][128] conv bias, BN scale, BN shift, residual bias (device)
  // host copies of the per-channel epilogue constants (kernel parameters -> constant bank)
  float bias[DAN_MAX_LAYERS][kC], scale[DAN_MAX_LAYERS][kC], shift[DAN_MAX_LAYERS][kC], rbias[DAN_MAX_LAYERS][kC], bbias[DAN_MAX_LAYERS][64];
  int num_sms;
  // software pipeline across passes (dan_bf16_forward): side stream for the read-axis pooling kernels + ordering events
  cudaStream_t side; cudaEvent_t ev_seg1[2], ev_pm[2], ev_seg2[2], ev_pf[2], ev_comp[2], ev_hw[2], ev_side;
};

inline int grid_for(long total, int block = 256) {
  long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

// workspace carve-up ---------------------------------------------------------------------------------------
struct Bf16Plan {
  int S, Bc, BcPad;
  long rows, rowsPad, kstride;       // per pass; kstride = rows per chunk plane
  long readsPad;
  long hw_layer_stride;
  long t_layer_pieces;               // uint4 pieces of one layer's T matrix
  int fcKC;                          // FC input pieces
  size_t off_zero_begin, off_x0, off_h[4], off_zero_end, off_t[2], off_pool[2], off_im2col, off_bmapg, off_bmap, off_agree, off_hw[2], off_fcin, off_fcx[DAN_MAX_FC], total;
  int maxN;
};

Bf16Plan make_plan(const dan_model* m, int batch) {
  Bf16Plan pl{};
  pl.S = m->pass_candidates < batch ? m->pass_candidates : (batch > 0 ? batch : 1);
  const int fc_chunk = dan_bf16_fc_chunk(m);
  pl.Bc = batch < fc_chunk ? (batch > 0 ? batch : 1) : fc_chunk;
  pl.BcPad = round_up_i(pl.Bc, 128);
  pl.rows = m->geom.rows_of(pl.S);
  pl.rowsPad = (pl.rows + 127) / 128 * 128;
  pl.kstride = kLead + pl.rowsPad + 8;
  pl.readsPad = ((long)pl.S * m->R + 127) / 128 * 128;
  pl.fcKC = m->fcInPad / 8;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += round_up_z(bytes, 1024); return o; };
  pl.off_zero_begin = off;
  pl.off_x0 = take((size_t)(m->CinPad / 8) * pl.kstride * 16);
  for (int i = 0; i < 4; ++i) pl.off_h[i] = take((size_t)kKC * pl.kstride * 16);     // 3 rotate on the sequential path; 2 + 2 on the pipelined one
  pl.off_zero_end = off;
  const int bott = m->bott > 0 ? m->bott : 32;
  pl.t_layer_pieces = (long)m->P * (bott / 8) * pl.readsPad;
  for (int i = 0; i < 2; ++i) pl.off_t[i] = take((size_t)m->L * pl.t_layer_pieces * 16);       // two sets: passes k and k+1 are in flight together
  for (int i = 0; i < 2; ++i) pl.off_pool[i] = take((size_t)pl.S * m->P * kC * 4);
  pl.off_im2col = take((size_t)pl.S * m->P * 3 * kC * 2);      // pool bias map (see bmap_pack_kernel): im2col of the read-mean, bf16 [cand*P + p][tap*C + c]
  pl.off_bmapg = take((size_t)pl.S * m->P * kC * 4);          // conv(pool), fp32 [cand*P + p][cout]
  pl.off_bmap = take((size_t)pl.S * kBmapBytesPerCand);      // conv(pool) + bias in the stack epilogue's fragment order, bf16 pairs
  pl.off_agree = take((size_t)pl.S * 2 * m->R);
  pl.hw_layer_stride = pl.readsPad * bott;
  for (int i = 0; i < 2; ++i) pl.off_hw[i] = take((size_t)m->L * pl.hw_layer_stride * 4);
  pl.off_fcin = take((size_t)pl.fcKC * pl.BcPad * 16);           // [BcPad][fcInPad] bf16, row-major
  pl.maxN = DAN_HEAD_PAD;
  for (int i = 0; i < m->cfg.num_fc; ++i) {
    pl.off_fcx[i] = take((size_t)(m->cfg.fc_sizes[i] / 8) * pl.BcPad * 16);
    if (m->cfg.fc_sizes[i] > pl.maxN) pl.maxN = m->cfg.fc_sizes[i];
  }
  pl.total = off;
  return pl;
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------
int dan_bf16_supported(const dan_model* m) {
  if (m->C != kC) return 0;
  if (m->cfg.pool_combine_dimension != 0) return 0;
  if (m->cfg.highway && !(m->bott == 32 || m->bott == 64)) return 0;
  if (m->geom.gap > kLead) return 0;
  if (m->CinPad % 16) return 0;
  for (int i = 0; i < m->cfg.num_fc; ++i) if (m->cfg.fc_sizes[i] % 32) return 0;
  return 1;
}

size_t dan_bf16_workspace_bytes(const dan_model* m, int batch) {
  if (!dan_bf16_supported(m)) return 0;
  return make_plan(m, batch).total;
}

int dan_bf16_pack(dan_model* m, const dan_weights* w, cudaStream_t st) {
  Bf16Weights* bw = static_cast<Bf16Weights*>(m->bf16_store);
  if (!bw) {
    bw = new Bf16Weights();
    memset(bw, 0, sizeof(*bw));
    m->bf16_store = bw;
    cudaDeviceGetAttribute(&bw->num_sms, cudaDevAttrMultiProcessorCount, m->device);
  }
  const int L = m->L, bott = m->bott, P = m->P, R = m->R;
  auto alloc = [&](uint4** p, size_t pieces) -> int {
    if (*p) return DAN_OK;
    DAN_CUDA_TRY(cudaMalloc(p, pieces * 16));
    return DAN_OK;
  };
  int rc;
  for (int l = 0; l < L; ++l) {
    const int cin = l == 0 ? m->Cin : kC, kc_in = (l == 0 ? m->CinPad : kC) / 8;
    if (!bw->wstream[l]) {
      const size_t conv_b = (size_t)3 * kc_in * kC * 16, res_b = m->cfg.is_residual[l] ? (size_t)kKC * kC * 16 : 0, bott_b = m->cfg.highway ? (size_t)kKC * bott * 16 : 0;
      bw->wstream_bytes[l] = round_up_z(conv_b + res_b + bott_b, 256);
      DAN_CUDA_TRY(cudaMalloc(&bw->wstream[l], bw->wstream_bytes[l] * kWeightReplicas));
      bw->wconv[l] = reinterpret_cast<uint4*>(bw->wstream[l]);
      if (res_b) bw->wres[l] = reinterpret_cast<uint4*>(bw->wstream[l] + conv_b);
      if (bott_b) bw->wbott[l] = reinterpret_cast<uint4*>(bw->wstream[l] + conv_b + res_b);
    }
    pack_conv_bf16_kernel<<<grid_for((long)3 * kc_in * kC * 8), 256, 0, st>>>(w->conv_w[l], bw->wconv[l], cin, kc_in);
    if (l > 0 && m->cfg.pool_after[l - 1]) {
      if ((rc = alloc(&bw->wbmap[l], (size_t)kC * 3 * kC / 8))) return rc;
      pack_linear_bf16_kernel<<<grid_for((long)3 * kC * kC), 256, 0, st>>>(w->conv_w[l], bw->wbmap[l], kC, kC, 3 * kC, 3 * kC, 4, P, kC, R, bott, L, 0, 0);
    }
    if (m->cfg.is_residual[l]) {
      pack_linear_bf16_kernel<<<grid_for((long)kKC * kC * 8), 256, 0, st>>>(w->res_w[l], bw->wres[l], kC, kC, kC, kC, 3, P, kC, R, bott, L, 0, 0);
    }
    if (m->cfg.highway) {
      pack_linear_bf16_kernel<<<grid_for((long)kKC * bott * 8), 256, 0, st>>>(w->bott_w[l], bw->wbott[l], bott, bott, kC, kC, 3, P, kC, R, bott, L, 0, 0);
      const int K = P * bott;
      if ((rc = alloc(&bw->wcomp_all, (size_t)L * (K / 8) * bott))) return rc;
      bw->wcomp[l] = bw->wcomp_all + (size_t)l * (K / 8) * bott;
      pack_linear_bf16_kernel<<<grid_for((long)K * bott), 256, 0, st>>>(w->comp_w[l], bw->wcomp[l], bott, bott, K, K, 2, P, kC, R, bott, L, 0, 0);
    }
  }
  int K = m->fcIn, Kpad = m->fcInPad;
  for (int i = 0; i < m->cfg.num_fc; ++i) {
    const int N = m->cfg.fc_sizes[i];
    if ((rc = alloc(&bw->fcw[i], (size_t)(Kpad / 8) * N))) return rc;
    pack_linear_bf16_kernel<<<grid_for((long)Kpad * N), 256, 0, st>>>(w->fc_w[i], bw->fcw[i], N, N, K, Kpad, i == 0 ? 1 : 0, P, kC, R, bott,
                                                                      m->cfg.concat_hw_reads ? L : 1, m->pooled, m->cfg.skip_final_maxpool);
    K = N; Kpad = N;
  }
  if ((rc = alloc(&bw->headw, (size_t)(m->hidden / 8) * DAN_HEAD_PAD))) return rc;
  pack_linear_bf16_kernel<<<grid_for((long)m->hidden * DAN_HEAD_PAD), 256, 0, st>>>(w->head_w, bw->headw, DAN_NUM_HEAD_OUTPUTS, DAN_HEAD_PAD, m->hidden, m->hidden, 0, P, kC, R, bott, L, 0, 0);
  for (int l = 0; l < L; ++l)
    for (int r = 1; r < kWeightReplicas; ++r)
      DAN_CUDA_TRY(cudaMemcpyAsync(bw->wstream[l] + (size_t)r * bw->wstream_bytes[l], bw->wstream[l], bw->wstream_bytes[l], cudaMemcpyDeviceToDevice, st));
  DAN_CUDA_TRY(cudaGetLastError());
  // epilogue constants -> host (the fp32 packer already folded BatchNorm on this stream)
  DAN_CUDA_TRY(cudaStreamSynchronize(st));
  for (int l = 0; l < L; ++l) {
    DAN_CUDA_TRY(cudaMemcpy(bw->bias[l], m->convB[l], kC * 4, cudaMemcpyDeviceToHost));
    if (m->cfg.use_batchnorm) {
      DAN_CUDA_TRY(cudaMemcpy(bw->scale[l], m->bnScale[l], kC * 4, cudaMemcpyDeviceToHost));
      DAN_CUDA_TRY(cudaMemcpy(bw->shift[l], m->bnShift[l], kC * 4, cudaMemcpyDeviceToHost));
    } else {
      for (int c = 0; c < kC; ++c) { bw->scale[l][c] = 1.f; bw->shift[l][c] = 0.f; }
    }
    if (m->cfg.is_residual[l]) DAN_CUDA_TRY(cudaMemcpy(bw->rbias[l], m->resB[l], kC * 4, cudaMemcpyDeviceToHost));
    if (m->cfg.highway) DAN_CUDA_TRY(cudaMemcpy(bw->bbias[l], m->bottB[l], bott * 4, cudaMemcpyDeviceToHost));
  }
  if (!bw->chan_dev) DAN_CUDA_TRY(cudaMalloc(&bw->chan_dev, sizeof(float) * DAN_MAX_LAYERS * 4 * kC));
  for (int l = 0; l < L; ++l) {
    float* d = bw->chan_dev + (size_t)l * 4 * kC;
    DAN_CUDA_TRY(cudaMemcpy(d, bw->bias[l], kC * 4, cudaMemcpyHostToDevice));
    DAN_CUDA_TRY(cudaMemcpy(d + kC, bw->scale[l], kC * 4, cudaMemcpyHostToDevice));
    DAN_CUDA_TRY(cudaMemcpy(d + 2 * kC, bw->shift[l], kC * 4, cudaMemcpyHostToDevice));
    DAN_CUDA_TRY(cudaMemcpy(d + 3 * kC, bw->rbias[l], kC * 4, cudaMemcpyHostToDevice));
  }
  if (m->cfg.highway) {
    if (!bw->comp_bias_ptrs) DAN_CUDA_TRY(cudaMalloc(&bw->comp_bias_ptrs, sizeof(float*) * DAN_MAX_LAYERS));
    DAN_CUDA_TRY(cudaMemcpy(bw->comp_bias_ptrs, m->compB, sizeof(float*) * L, cudaMemcpyHostToDevice));
  }
  return DAN_OK;
}

void dan_bf16_free(dan_model* m) {
  Bf16Weights* bw = static_cast<Bf16Weights*>(m->bf16_store);
  if (!bw) return;
  for (int l = 0; l < DAN_MAX_LAYERS; ++l) cudaFree(bw->wstream[l]);
  cudaFree(bw->wcomp_all);
  cudaFree(bw->chan_dev);
  for (int i = 0; i < DAN_MAX_FC; ++i) cudaFree(bw->fcw[i]);
  cudaFree(bw->headw);
  cudaFree(bw->comp_bias_ptrs);
  if (bw->side) {
    cudaStreamDestroy(bw->side);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(bw->ev_seg1[i]); cudaEventDestroy(bw->ev_pm[i]); cudaEventDestroy(bw->ev_seg2[i]); cudaEventDestroy(bw->ev_pf[i]); cudaEventDestroy(bw->ev_comp[i]); cudaEventDestroy(bw->ev_hw[i]); }
    cudaEventDestroy(bw->ev_side);
  }
  delete bw;
  m->bf16_store = nullptr;
}

int dan_bf16_forward(dan_model* m, const DevInputs& in, int batch, float* heads_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  Bf16Weights* bw = static_cast<Bf16Weights*>(m->bf16_store);
  if (!bw) { dan_set_error("bf16 weights not packed"); return DAN_E_INVALID; }
  const Bf16Plan pl = make_plan(m, batch);
  if (ws_bytes < pl.total) { dan_set_error("workspace too small: %zu < %zu", ws_bytes, pl.total); return DAN_E_WORKSPACE; }
  char* base = static_cast<char*>(ws);
  const RowGeom g = m->geom;
  const int L = m->L, bott = m->bott, P = m->P, R = m->R;
  uint4* X0 = reinterpret_cast<uint4*>(base + pl.off_x0);
  uint4* H[4]; for (int i = 0; i < 4; ++i) H[i] = reinterpret_cast<uint4*>(base + pl.off_h[i]);
  uint4* Tset[2] = {reinterpret_cast<uint4*>(base + pl.off_t[0]), reinterpret_cast<uint4*>(base + pl.off_t[1])};
  float* POOLset[2] = {reinterpret_cast<float*>(base + pl.off_pool[0]), reinterpret_cast<float*>(base + pl.off_pool[1])};
  uint4* IM2COL = reinterpret_cast<uint4*>(base + pl.off_im2col);
  float* BMAPG = reinterpret_cast<float*>(base + pl.off_bmapg);
  uint4* BMAP = reinterpret_cast<uint4*>(base + pl.off_bmap);
  uint4* T = Tset[0];
  float* POOL = POOLset[0];
  uint8_t* AGREE = reinterpret_cast<uint8_t*>(base + pl.off_agree);
  const bool fast_encode = m->cfg.embed_dim == 20 && m->cfg.use_q_scores && m->cfg.use_strands && m->cfg.use_reads_ref_var_mask && m->CinPad == 48 && P <= 512 &&
                           encode_prod_smem_bytes(R) <= 48 * 1024;
  float* HWset[2] = {reinterpret_cast<float*>(base + pl.off_hw[0]), reinterpret_cast<float*>(base + pl.off_hw[1])};
  float* HW = HWset[0];
  uint4* FCIN = reinterpret_cast<uint4*>(base + pl.off_fcin);
  // highway compression (model.py:776) of every layer of a pass as ONE batched GEMM: HW[l][read][o] = T[l][read][:] . Wc[l][o][:]
  auto run_compression = [&](int l0, int nl, int reads, int set = 0) -> int {
    const long K = (long)P * bott;
    uint4* T = Tset[set]; float* HW = HWset[set];
    Gemm2Operand A{T + (long)l0 * pl.t_layer_pieces, reads, K * 2, pl.t_layer_pieces * 16};
    Gemm2Operand B{bw->wcomp[l0], bott, K * 2, K * 2 * bott};
    Gemm2Params gp{};
    gp.M = reads; gp.N = bott; gp.K = (int)K; gp.mode = kG2Raw;
    gp.out = HW + (long)l0 * pl.hw_layer_stride; gp.out_batch_stride = pl.hw_layer_stride; gp.ldo = bott;
    return run_gemm2(A, B, gp, nl, bw->num_sms, st);
  };
  int rc;

  static DanSmemAttr enc_attr, layer_attr, stack_attr[4];
  const size_t enc_smem = encode_smem_bytes(P, R, m->cfg.embed_dim);
  DAN_CUDA_TRY(enc_attr.ensure(encode_rows_bf16_kernel, enc_smem));
  DAN_CUDA_TRY(layer_attr.ensure(dan_layer_kernel, 227 * 1024));
  bool fused = m->P == 201 && m->geom.gap <= kStkLead && (!m->cfg.highway || bott == 32 || bott == 64) && !getenv("DAN_B200_LAYERWISE");
  for (int l = 1; l < L; ++l)      // a residual layer fed by a pool-add needs the un-pooled input as residual (model.py:732 vs :742)
    if (m->cfg.is_residual[l] && m->cfg.pool_after[l - 1]) fused = false;
  if (fused) {
    DAN_CUDA_TRY(stack_attr[0].ensure(dan_stack_kernel<0>, kStkSmemBytes));
    DAN_CUDA_TRY(stack_attr[1].ensure(dan_stack_kernel<1>, kStkSmemBytes));
    DAN_CUDA_TRY(stack_attr[2].ensure(dan_stack_kernel<2>, kStkSmemBytes));
    DAN_CUDA_TRY(stack_attr[3].ensure(dan_stack_kernel<3>, kStkSmemBytes));
  }
  // layer-wise path: halo rows (and the rows past the last tile) must read as zero: clear the row matrices once per call.
  // The fused path loads and stores exactly the P data rows of every read and keeps its zero rows in shared memory.
  if (!fused) DAN_CUDA_TRY(cudaMemsetAsync(base + pl.off_zero_begin, 0, pl.off_zero_end - pl.off_zero_begin, st));

  // one persistent launch of dan_stack_kernel over layers [l, l_end) of a pass (dan_stack.cuh)
  bool seg_bmap = false;     // the next launch_segment takes the pool term from the bias map instead of adding the pool table to every read
  auto launch_segment = [&](int l, int l_end, const uint4* seg_in, uint4* next, int set, bool with_pool, int ns) -> int {
          StackParams sp{};
          sp.in = seg_in; sp.in_kstride = pl.kstride; sp.out = next; sp.out_kstride = pl.kstride;
          sp.t_reads_stride = pl.readsPad; sp.num_reads = ns * R; sp.P = P; sp.pitch = g.pitch; sp.bott = bott > 0 ? bott : 32;
          sp.num_layers = l_end - l;
          sp.pool = with_pool && !seg_bmap ? POOLset[set] : nullptr; sp.reads_per_cand = R;
          sp.bmap = with_pool && seg_bmap ? BMAP : nullptr;
          for (int k = l; k < l_end; ++k) {
            StackLayer& SL = sp.layer[k - l];
            SL.wstream = bw->wstream[k]; SL.wreplica_stride = bw->wstream_bytes[k]; SL.chan = bw->chan_dev + (size_t)k * 4 * kC; SL.bbias = m->bottB[k];
            SL.tout = Tset[set] + (long)k * pl.t_layer_pieces;
            SL.kc_in = (k == 0 ? m->CinPad : kC) / 8; SL.conv_blocks = 3 * SL.kc_in / 2;
            SL.dil = m->cfg.dilation[k]; SL.residual = m->cfg.is_residual[k]; SL.highway = m->cfg.highway;
          }
          int grid = sp.num_reads / 2 < bw->num_sms ? (sp.num_reads + 1) / 2 : bw->num_sms;
          static const int stack_debug = getenv("DAN_B200_STACKDEBUG") ? atoi(getenv("DAN_B200_STACKDEBUG")) : 0;
          sp.debug = stack_debug;
          static const bool stack_prof = getenv("DAN_B200_STACKPROF") != nullptr;     // development aid: per-role cycle counters
          static const bool stack_trace = getenv("DAN_B200_STACKTRACE") != nullptr;
          static int trace_dumps = 0;
          if (stack_trace && trace_dumps < 2 && sp.num_layers > 2) { sp.trace_cap = 8000; DAN_CUDA_TRY(cudaMalloc(&sp.trace, sizeof(uint2) * sp.trace_cap)); DAN_CUDA_TRY(cudaMemsetAsync(sp.trace, 0, sizeof(uint2) * sp.trace_cap, st)); }
          if (stack_prof) { DAN_CUDA_TRY(cudaMalloc(&sp.prof, sizeof(unsigned long long) * 16 * grid)); DAN_CUDA_TRY(cudaMemsetAsync(sp.prof, 0, sizeof(unsigned long long) * 16 * grid, st)); }
          { DanProfScope ps(DAN_PROF_CONV_STACK, st);
            if (sp.trace) dan_stack_kernel<3><<<grid, kStkThreads, kStkSmemBytes, st>>>(sp);                  // development builds of the kernel
            else if (sp.prof) dan_stack_kernel<2><<<grid, kStkThreads, kStkSmemBytes, st>>>(sp);
            else if (sp.debug) dan_stack_kernel<1><<<grid, kStkThreads, kStkSmemBytes, st>>>(sp);
            else dan_stack_kernel<0><<<grid, kStkThreads, kStkSmemBytes, st>>>(sp); }
          dan_count_launch();
          DAN_CUDA_TRY(cudaGetLastError());
          static const bool sync_dbg = getenv("DAN_B200_SYNC") != nullptr;          // development: find the launch that hangs / faults
          if (sync_dbg) { fprintf(stderr, "[sync] stack layers %d-%d reads %d grid %d pool %d ... ", l + 1, l_end, sp.num_reads, grid, sp.pool != nullptr); fflush(stderr);
                          cudaError_t e = cudaStreamSynchronize(st); fprintf(stderr, "%s\n", cudaGetErrorString(e)); fflush(stderr); }
          if (sp.trace) {
            std::vector<uint2> h(sp.trace_cap);
            DAN_CUDA_TRY(cudaStreamSynchronize(st));
            DAN_CUDA_TRY(cudaMemcpy(h.data(), sp.trace, h.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
            cudaFree(sp.trace);
            char name[64]; snprintf(name, sizeof(name), "gpurun_out/stack_trace_%d.txt", trace_dumps++);
            if (FILE* f = fopen(name, "w")) { for (int k = 0; k < sp.trace_cap; ++k) if (h[k].y) fprintf(f, "%x %u\n", h[k].x, h[k].y); fclose(f); }
          }
          if (stack_prof) {
            std::vector<unsigned long long> h(16 * (size_t)grid);
            DAN_CUDA_TRY(cudaStreamSynchronize(st));
            DAN_CUDA_TRY(cudaMemcpy(h.data(), sp.prof, h.size() * 8, cudaMemcpyDeviceToHost));
            cudaFree(sp.prof);
            double a[16] = {0};
            for (int c = 0; c < grid; ++c) for (int k = 0; k < 16; ++k) a[k] += (double)h[c * 16 + k] / grid;
            fprintf(stderr, "[stackprof] layers %d-%d reads %d grid %d: issuer0 total %.0f dep-wait %.0f | epi0 wait %.0f main %.0f bott %.0f | epi1 wait %.0f main %.0f bott %.0f | store+load wait %.0f %.0f | issuer1 total %.0f dep-wait %.0f | weight-wait %.0f %.0f (cycles, mean over CTAs)\n",
                    l + 1, l_end, sp.num_reads, grid, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]);
          }
          return DAN_OK;
  };

  EncodeParams ep{};
  ep.in = in; ep.emb = m->emb; ep.pe = m->pe; ep.D = m->cfg.embed_dim; ep.Cin = m->Cin; ep.CinPad = m->CinPad;
  ep.use_q = m->cfg.use_q_scores; ep.use_s = m->cfg.use_strands; ep.use_m = m->cfg.use_reads_ref_var_mask; ep.g = g;

  auto encode_pass = [&](long cand0, int ns) -> int {
    if (fast_encode) {
      DanProfScope ps(DAN_PROF_ENCODE, st);
      agree_bits_kernel<<<ns, 128, 0, st>>>(in, cand0, P, R, AGREE);
      encode_prod_bf16_kernel<<<dim3((P + kEncPB - 1) / kEncPB, ns), 256, encode_prod_smem_bytes(R), st>>>(in, m->emb, m->pe, AGREE, cand0, g, X0, pl.kstride);
      dan_count_launch(2);
    } else {
      DanProfScope ps(DAN_PROF_ENCODE, st);
      encode_rows_bf16_kernel<<<ns, 256, enc_smem, st>>>(ep, cand0, X0, pl.kstride);
      dan_count_launch();
    }
    DAN_CUDA_TRY(cudaGetLastError());
    return DAN_OK;
  };

  // ---- software pipeline over the passes of one FC chunk (PROD structure: segment A = layers before the read-mean pool-add,
  // segment B = the rest). The HBM-bound read-axis reductions run on a side stream NEXT TO the persistent stack kernel of the
  // following pass (64-thread CTAs without shared memory fit beside its 640 threads / 226 KB on every SM):
  //   main : enc(0) A(0) | enc(1) A(1) B(0) comp(0) | enc(2) A(2) B(1) comp(1) | ...
  //   side :      mean(0)      mean(1)  final(0) hw(0)      mean(2)  final(1) hw(1)
  // Two sets of H / T / POOL / HW buffers (pass parity); events order producer -> consumer and consumer -> buffer reuse.
  int l_split = 0;
  for (int l = 1; l < L; ++l) if (m->cfg.pool_after[l - 1]) { if (!l_split) l_split = l; else { l_split = -1; break; } }
  // EXPERIMENTAL, opt-in (DAN_B200_PIPE=1): correct (GPU tests pass with it), but measured no gain — the side-stream kernels do not
  // become resident next to the persistent stack kernel on this driver (their class time stretches to the stack kernel's), so the
  // schedule degenerates to the sequential one. Kept for round 2.
  static const bool want_pipe = getenv("DAN_B200_PIPE") != nullptr;
  const bool pipelined = fused && l_split > 0 && l_split <= kStkMaxSeg && L - l_split <= kStkMaxSeg && m->cfg.highway && want_pipe &&
                         !getenv("DAN_B200_STACKPROF") && !getenv("DAN_B200_STACKTRACE") && !getenv("DAN_B200_SYNC");
  if (pipelined && !bw->side) {
    // same L1 / shared-memory split as the stack kernel, or the SM cannot hold both kernels' CTAs at once
    DAN_CUDA_TRY(cudaFuncSetAttribute(pool_mean_bf16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DAN_CUDA_TRY(cudaFuncSetAttribute(pool_final_bf16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DAN_CUDA_TRY(cudaFuncSetAttribute(highway_finish_bf16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DAN_CUDA_TRY(cudaStreamCreateWithFlags(&bw->side, cudaStreamNonBlocking));
    cudaEvent_t* evs[] = {bw->ev_seg1, bw->ev_pm, bw->ev_seg2, bw->ev_pf, bw->ev_comp, bw->ev_hw};
    for (auto e : evs) for (int i = 0; i < 2; ++i) DAN_CUDA_TRY(cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming));
    DAN_CUDA_TRY(cudaEventCreateWithFlags(&bw->ev_side, cudaEventDisableTiming));
  }
  auto run_chunk_pipelined = [&](int c0, int nb) -> int {
    cudaStream_t ss = bw->side;
    const int np = (nb + pl.S - 1) / pl.S;
    auto pass_ns = [&](int k) { const int s0 = k * pl.S; return nb - s0 < pl.S ? nb - s0 : pl.S; };
    auto seg_a = [&](int k) -> int {
      const int b = k & 1, ns = pass_ns(k);
      int rc2;
      if ((rc2 = encode_pass((long)c0 + (long)k * pl.S, ns))) return rc2;
      if ((rc2 = launch_segment(0, l_split, X0, H[b], b, false, ns))) return rc2;
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_seg1[b], st));
      DAN_CUDA_TRY(cudaStreamWaitEvent(ss, bw->ev_seg1[b], 0));
      { DanProfScope ps(DAN_PROF_POOL, ss); pool_mean_bf16_kernel<<<dim3((P + 63) / 64, ns, kKC), 64, 0, ss>>>(H[b], pl.kstride, POOLset[b], g); }
      dan_count_launch();
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_pm[b], ss));
      return DAN_OK;
    };
    int rc2;
    if ((rc2 = seg_a(0))) return rc2;
    for (int k = 0; k < np; ++k) {
      const int b = k & 1, ns = pass_ns(k), s0 = k * pl.S;
      if (k + 1 < np && (rc2 = seg_a(k + 1))) return rc2;
      DAN_CUDA_TRY(cudaStreamWaitEvent(st, bw->ev_pm[b], 0));
      if (k >= 2) DAN_CUDA_TRY(cudaStreamWaitEvent(st, bw->ev_pf[b], 0));                 // H[2 + b] has been reduced by pass k-2's pool kernel
      if ((rc2 = launch_segment(l_split, L, H[b], H[2 + b], b, true, ns))) return rc2;
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_seg2[b], st));
      DAN_CUDA_TRY(cudaStreamWaitEvent(ss, bw->ev_seg2[b], 0));
      { DanProfScope ps(DAN_PROF_POOL, ss); pool_final_bf16_kernel<<<dim3((P + 63) / 64, ns, kKC), 64, 0, ss>>>(H[2 + b], pl.kstride, FCIN, pl.fcKC, s0, g, m->cfg.skip_final_maxpool); }
      dan_count_launch();
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_pf[b], ss));
      if (k >= 2) DAN_CUDA_TRY(cudaStreamWaitEvent(st, bw->ev_hw[b], 0));                 // HW set b has been consumed by pass k-2
      if ((rc2 = run_compression(0, L, ns * R, b))) return rc2;
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_comp[b], st));
      DAN_CUDA_TRY(cudaStreamWaitEvent(ss, bw->ev_comp[b], 0));
      const int Lh = m->cfg.concat_hw_reads ? L : 1;
      { DanProfScope ps(DAN_PROF_POOL, ss); highway_finish_bf16_kernel<<<grid_for((long)ns * Lh * R * (bott / 8)), 256, 0, ss>>>(
          HWset[b], pl.hw_layer_stride, bw->comp_bias_ptrs, L, bott, R, m->cfg.concat_hw_reads, FCIN, pl.fcKC, m->pooled / 8, s0, ns); }
      dan_count_launch();
      DAN_CUDA_TRY(cudaEventRecord(bw->ev_hw[b], ss));
    }
    DAN_CUDA_TRY(cudaGetLastError());
    DAN_CUDA_TRY(cudaEventRecord(bw->ev_side, ss));
    DAN_CUDA_TRY(cudaStreamWaitEvent(st, bw->ev_side, 0));                               // FC reads every row of FCIN
    return DAN_OK;
  };

  for (int c0 = 0; c0 < batch; c0 += pl.Bc) {
    const int nb = batch - c0 < pl.Bc ? batch - c0 : pl.Bc;
    if (m->fcInPad != m->fcIn) DAN_CUDA_TRY(cudaMemsetAsync(FCIN, 0, (size_t)pl.fcKC * pl.BcPad * 16, st));
    if (pipelined) {
      if ((rc = run_chunk_pipelined(c0, nb))) return rc;
    } else
    for (int s0 = 0; s0 < nb; s0 += pl.S) {
      const int ns = nb - s0 < pl.S ? nb - s0 : pl.S;
      const long rows = g.rows_of(ns);
      const int num_tiles = (int)((rows + 127) / 128);
      if ((rc = encode_pass((long)c0 + s0, ns))) return rc;
      const uint4* cur = X0;
      int hsel = 0;
      if (fused) {
        // ---- fused path: one persistent launch per segment of layers without a pool-add in between (dan_stack.cuh) ----
        int l = 0;
        while (l < L) {
          int l_end = l + 1;
          while (l_end < L && !m->cfg.pool_after[l_end - 1] && l_end - l < kStkMaxSeg) ++l_end;
          const bool with_pool = l > 0 && m->cfg.pool_after[l - 1];      // the pool-add is fused into the segment's load
          uint4* next = H[(hsel + 1) % 3];
          if ((rc = launch_segment(l, l_end, cur, next, 0, with_pool, ns))) return rc;
          if (m->cfg.pool_after[l_end - 1] && l_end < L) {
            static const bool no_bmap = getenv("DAN_B200_NO_BMAP") != nullptr;       // development: per-read pool-add instead of the bias map
            seg_bmap = !no_bmap && bw->wbmap[l_end] != nullptr;
            { DanProfScope ps(DAN_PROF_POOL, st);
              pool_mean_bf16_kernel<<<dim3((P + 63) / 64, ns, kKC), 64, 0, st>>>(next, pl.kstride, POOL, g, seg_bmap ? IM2COL : nullptr, m->cfg.dilation[l_end]); }
            dan_count_launch();
            if (seg_bmap) {
              // bias map of layer l_end: conv(pool) on the tensor cores (rows = candidate positions, K = 3 taps x 128 channels), then + bias in fragment order
              Gemm2Operand A{IM2COL, (long)ns * P, 3L * kC * 2, 0};
              Gemm2Operand B{bw->wbmap[l_end], kC, 3L * kC * 2, 0};
              Gemm2Params gp{};
              gp.M = ns * P; gp.N = kC; gp.K = 3 * kC; gp.mode = kG2Raw; gp.out = BMAPG; gp.out_batch_stride = 0; gp.ldo = kC;
              if ((rc = run_gemm2(A, B, gp, 1, bw->num_sms, st))) return rc;
              { DanProfScope ps(DAN_PROF_POOL, st); bmap_pack_kernel<<<grid_for((long)ns * kStkBmapPerCand), 256, 0, st>>>(BMAPG, m->convB[l_end], BMAP, ns, P); }
              dan_count_launch();
              DAN_CUDA_TRY(cudaGetLastError());
            }
          }
          cur = next; hsel = (hsel + 1) % 3;
          l = l_end;
        }
        if (m->cfg.highway) {
          if ((rc = run_compression(0, L, ns * R))) return rc;
        }
      } else
      for (int l = 0; l < L; ++l) {
        const uint4* conv_in = cur;
        if (l > 0 && m->cfg.pool_after[l - 1]) {
          uint4* hp = H[(hsel + 2) % 3];
          { DanProfScope ps(DAN_PROF_POOL, st); add_pool_bf16_kernel<<<grid_for(rows * kKC), 256, 0, st>>>(cur, POOL, hp, pl.kstride, rows, g); }
          dan_count_launch();
          DAN_CUDA_TRY(cudaGetLastError());
          conv_in = hp;
        }
        uint4* next = H[(hsel + 1) % 3];
        LayerParams lp{};
        lp.in = conv_in; lp.in_kstride = pl.kstride; lp.out = next; lp.out_kstride = pl.kstride;
        lp.tout = T + (long)l * pl.t_layer_pieces; lp.t_reads_stride = pl.readsPad;
        lp.resid = (m->cfg.is_residual[l] && conv_in != cur) ? cur : nullptr;   // residual excludes the pool term (model.py:732 vs :742)
        lp.wconv = bw->wconv[l]; lp.wres = bw->wres[l]; lp.wbott = bw->wbott[l];
        lp.rows_total = rows; lp.num_tiles = num_tiles; lp.pitch = g.pitch; lp.P = P; lp.gap = g.gap; lp.dil = m->cfg.dilation[l];
        lp.kc_in = (l == 0 ? m->CinPad : kC) / 8; lp.residual = m->cfg.is_residual[l]; lp.highway = m->cfg.highway; lp.bott = bott;
        memcpy(lp.bias, bw->bias[l], sizeof(lp.bias)); memcpy(lp.scale, bw->scale[l], sizeof(lp.scale)); memcpy(lp.shift, bw->shift[l], sizeof(lp.shift));
        memcpy(lp.rbias, bw->rbias[l], sizeof(lp.rbias)); memcpy(lp.bbias, bw->bbias[l], sizeof(lp.bbias));
        const size_t smem = layer_smem_bytes(lp.kc_in, lp.residual, lp.highway, bott, g.gap);
        int grid = (num_tiles + kSlots - 1) / kSlots;
        if (grid > bw->num_sms) grid = bw->num_sms;
        { DanProfScope ps(DAN_PROF_CONV_STACK, st); dan_layer_kernel<<<grid, kLayerThreads, smem, st>>>(lp); }
        dan_count_launch();
        DAN_CUDA_TRY(cudaGetLastError());
        if (m->cfg.pool_after[l]) {
          { DanProfScope ps(DAN_PROF_POOL, st); pool_mean_bf16_kernel<<<dim3((P + 63) / 64, ns, kKC), 64, 0, st>>>(next, pl.kstride, POOL, g); }
          dan_count_launch();
          DAN_CUDA_TRY(cudaGetLastError());
        }
        if (m->cfg.highway) {
          if ((rc = run_compression(l, 1, ns * R))) return rc;
        }
        cur = next; hsel = (hsel + 1) % 3;
      }
      { DanProfScope ps(DAN_PROF_POOL, st); pool_final_bf16_kernel<<<dim3((P + 63) / 64, ns, kKC), 64, 0, st>>>(cur, pl.kstride, FCIN, pl.fcKC, s0, g, m->cfg.skip_final_maxpool); }
      dan_count_launch();
      DAN_CUDA_TRY(cudaGetLastError());
      if (m->cfg.highway) {
        const int Lh = m->cfg.concat_hw_reads ? L : 1;
        { DanProfScope ps(DAN_PROF_POOL, st); highway_finish_bf16_kernel<<<grid_for((long)ns * Lh * R * (bott / 8)), 256, 0, st>>>(
            HW, pl.hw_layer_stride, bw->comp_bias_ptrs, L, bott, R, m->cfg.concat_hw_reads, FCIN, pl.fcKC, m->pooled / 8, s0, ns); }
        dan_count_launch();
        DAN_CUDA_TRY(cudaGetLastError());
      }
    }
    // ---- FC trunk + heads (model.py:917-958) ----
    const uint4* x = FCIN; int K = m->fcInPad;
    for (int i = 0; i < m->cfg.num_fc; ++i) {
      const int N = m->cfg.fc_sizes[i];
      uint4* y = reinterpret_cast<uint4*>(base + pl.off_fcx[i]);
      Gemm2Operand A{x, nb, (long)K * 2, 0};
      Gemm2Operand B{bw->fcw[i], N, (long)K * 2, 0};
      Gemm2Params gp{};
      gp.M = nb; gp.N = N; gp.K = K; gp.mode = kG2BiasReluBf16; gp.bias = m->fcB[i];
      gp.out_bf16 = reinterpret_cast<__nv_bfloat16*>(y); gp.ld_bf16 = N;
      if ((rc = run_gemm2(A, B, gp, 1, bw->num_sms, st))) return rc;
      x = y; K = N;
    }
    {
      Gemm2Operand A{x, nb, (long)K * 2, 0};
      Gemm2Operand B{bw->headw, DAN_HEAD_PAD, (long)K * 2, 0};
      Gemm2Params gp{};
      gp.M = nb; gp.N = DAN_HEAD_PAD; gp.K = K; gp.mode = kG2Heads; gp.bias = m->headB;
      gp.out = heads_out + (long)c0 * DAN_NUM_HEAD_OUTPUTS;
      if ((rc = run_gemm2(A, B, gp, 1, bw->num_sms, st))) return rc;
    }
  }
  return DAN_OK;
}

// FC input in the REFERENCE feature order (fp32), undoing the bf16 path's feature permutation (test hook)
namespace {
__global__ void fcin_to_reference_order_kernel(const uint4* __restrict__ fcin, long kstride, int rows, float* __restrict__ out,
                                               int fcIn, int pooled, int P, int C, int R, int bott) {
  const long total = (long)rows * fcIn;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int f = (int)(i % fcIn); const long m = i / fcIn;
    long k;
    if (f < pooled) {
      const int blk = f / (C * P), rem = f % (C * P);
      const int c = rem / P, pp = rem % P;
      k = (long)blk * P * C + (long)pp * C + c;
    } else {
      const int h = f - pooled;
      const int l = h / (bott * R), rem = h % (bott * R);
      const int o = rem / R, r = rem % R;
      k = pooled + ((long)l * R + r) * bott + o;
    }
    out[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(fcin + m * kstride)[k]);      // kstride = pieces per row
  }
}
}  // namespace

int dan_bf16_debug_fc_input(dan_model* m, int batch, const void* ws, float* out, cudaStream_t st) {
  const Bf16Plan pl = make_plan(m, batch);
  const int nb = batch % pl.Bc == 0 ? pl.Bc : batch % pl.Bc;
  const uint4* FCIN = reinterpret_cast<const uint4*>(static_cast<const char*>(ws) + pl.off_fcin);
  fcin_to_reference_order_kernel<<<grid_for((long)nb * m->fcIn), 256, 0, st>>>(FCIN, pl.fcKC, nb, out, m->fcIn, m->pooled, m->P, kC, m->R, m->bott > 0 ? m->bott : 1);
  DAN_CUDA_TRY(cudaGetLastError());
  return nb;
}
