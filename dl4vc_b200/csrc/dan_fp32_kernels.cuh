// dan_fp32_kernels.cuh — fp32 (CUDA-core FFMA) device code shared by the inference path (dan_fp32.cu) and the training path
// (dan_train.cu): the pileup encoder on row matrices, the tap-gathering tiled SGEMM every contraction of the network maps to,
// read-axis pooling, and the weight re-layout kernels.
#pragma once
#include <cstdio>
#include <vector>
#include "dan_kernels_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ encoder
__global__ void __launch_bounds__(256) encode_rows_fp32_kernel(EncodeParams p, long cand0, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long cand = cand0 + blockIdx.x;
  EncodeSmem s = encode_stage(p, cand, smem_raw);
  const int q4 = p.CinPad / 4;
  const int rows = p.g.R * p.g.pitch;
  float4* dst = reinterpret_cast<float4*>(out + (long)blockIdx.x * rows * p.CinPad);
  for (int idx = threadIdx.x; idx < rows * q4; idx += blockDim.x) {
    const int row = idx / q4, c0 = (idx - row * q4) * 4;
    const int r = row / p.g.pitch, pp = row - r * p.g.pitch;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pp < p.g.P) {
      v.x = c0 + 0 < p.Cin ? encode_channel(p, s, c0 + 0, pp, r) : 0.f;
      v.y = c0 + 1 < p.Cin ? encode_channel(p, s, c0 + 1, pp, r) : 0.f;
      v.z = c0 + 2 < p.Cin ? encode_channel(p, s, c0 + 2, pp, r) : 0.f;
      v.w = c0 + 3 < p.Cin ? encode_channel(p, s, c0 + 3, pp, r) : 0.f;
    }
    dst[idx] = v;
  }
}

// rows [row][CinPad] -> reference order (B, Cin, R, P), for the bit-exactness test hook
__global__ void rows_to_reference_order_kernel(const float* __restrict__ rows, float* __restrict__ out, int cands,
                                               int Cin, int CinPad, RowGeom g) {
  const long total = (long)cands * Cin * g.R * g.P;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int pp = (int)(i % g.P); long t = i / g.P;
    int r = (int)(t % g.R); t /= g.R;
    int c = (int)(t % Cin); long b = t / Cin;
    out[i] = rows[((b * g.R + r) * g.pitch + pp) * CinPad + c];
  }
}

// ------------------------------------------------------------------------------------------------ SGEMM
struct GemmParams {
  const float* A; int lda; long a_rows;       // A[m][k]; rows m in [-gap, a_rows) are addressable, others read as 0
  int M;                                      // output rows (multiple of 128 not required)
  int ntaps; int tap_off[3]; int Kc;          // K = ntaps*Kc, Kc % 16 == 0; tap t reads row m + tap_off[t]
  const float* W; int N; int ldw;             // W[k][n]
  const float* bias; int relu;
  const float* scale; const float* shift;     // after relu (BatchNorm after ReLU, model.py:749-751)
  const float* resid; int ldr;                // + resid[m][n]
  int mask_pitch, mask_valid; long mask_total;  // rows with (m % pitch) >= valid or m >= total are written as 0
  float* out; int ldo;
  int head_act;                               // sigmoid on column 5, leaky_relu(0.01) on column 6 (model.py:954,956)
  int splits; long split_stride;              // split-K: raw partial sums to out + z*split_stride, no epilogue
};

template <int BN>
__global__ void __launch_bounds__(256) sgemm_taps_kernel(GemmParams p) {
  constexpr int BM = 128, BK = 16;
  constexpr int TN = BN / 16;                 // columns per thread: 8 (as 4+4) or 2
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int chunks_per_tap = p.Kc / BK;
  const int total_chunks = p.ntaps * chunks_per_tap;
  const int z = blockIdx.z;
  const int c_begin = (int)((long)total_chunks * z / p.splits), c_end = (int)((long)total_chunks * (z + 1) / p.splits);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // global -> register staging: A: thread = (row tid%128, half tid/128) 8 floats; W: BK*BN/256 floats
  const int a_row = tid & 127, a_half = tid >> 7;
  float4 ra[2];
  constexpr int WV = BK * BN / 4 / 256 > 0 ? BK * BN / 4 / 256 : 1;   // float4 per thread (2 for BN=128)
  float4 rb[WV];
  auto load_chunk = [&](int chunk) {
    const int t = chunk / chunks_per_tap, kc = (chunk - t * chunks_per_tap) * BK;
    const long row = m0 + a_row + p.tap_off[t];
    if (m0 + a_row < p.M && row < p.a_rows) {
      const float4* src = reinterpret_cast<const float4*>(p.A + row * p.lda + kc + a_half * 8);
      ra[0] = __ldg(src); ra[1] = __ldg(src + 1);
    } else {
      ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long kbase = (long)t * p.Kc + kc;
    if (BN == 128) {
#pragma unroll
      for (int v = 0; v < WV; ++v) {
        const int f = tid + v * 256, kk = f / (BN / 4), nn = (f % (BN / 4)) * 4;
        rb[v] = (n0 + nn < p.N) ? __ldg(reinterpret_cast<const float4*>(p.W + (kbase + kk) * p.ldw + n0 + nn))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {  // BN == 32: 128 float4 per chunk, threads 0..127
      if (tid < BK * BN / 4) {
        const int kk = tid / (BN / 4), nn = (tid % (BN / 4)) * 4;
        rb[0] = (n0 + nn < p.N) ? __ldg(reinterpret_cast<const float4*>(p.W + (kbase + kk) * p.ldw + n0 + nn))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_chunk = [&](int buf) {
    const float* f = reinterpret_cast<const float*>(ra);
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][a_half * 8 + j][a_row] = f[j];
    if (BN == 128) {
#pragma unroll
      for (int v = 0; v < WV; ++v) {
        const int fidx = tid + v * 256, kk = fidx / (BN / 4), nn = (fidx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = rb[v];
      }
    } else if (tid < BK * BN / 4) {
      const int kk = tid / (BN / 4), nn = (tid % (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = rb[0];
    }
  };

  if (c_begin < c_end) {
    load_chunk(c_begin);
    store_chunk(0);
    __syncthreads();
    for (int chunk = c_begin; chunk < c_end; ++chunk) {
      const int buf = (chunk - c_begin) & 1;
      if (chunk + 1 < c_end) load_chunk(chunk + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[8], b[TN];
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        if (BN == 128) {
          *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
          *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        } else {
          *reinterpret_cast<float2*>(b) = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (chunk + 1 < c_end) {
        store_chunk(buf ^ 1);
        __syncthreads();
      }
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    bool valid = true;
    if (p.mask_pitch > 0) valid = (m < p.mask_total) && ((int)(m % p.mask_pitch) < p.mask_valid);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (BN == 128 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx * 2 + j);
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.splits > 1) {
        p.out[(long)z * p.split_stride + m * p.ldo + n] = v;
        continue;
      }
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.scale) v = fmaf(v, p.scale[n], p.shift[n]);
      if (p.resid) v += p.resid[m * p.ldr + n];
      if (p.head_act) {
        if (n == 5) v = 1.f / (1.f + expf(-v));
        else if (n == 6) v = v >= 0.f ? v : 0.01f * v;
      }
      p.out[m * p.ldo + n] = valid ? v : 0.f;
    }
  }
}

// fixed-order reduction of split-K partials + bias / relu / head activations
__global__ void splitk_finish_kernel(const float* __restrict__ part, int splits, long split_stride, int M, int N,
                                     int ldo_part, const float* __restrict__ bias, int relu, int head_act,
                                     float* __restrict__ out, int ldo) {
  const long total = (long)M * N;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N); const long m = i / N;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[(long)s * split_stride + m * ldo_part + n];
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    if (head_act) {
      if (n == 5) v = 1.f / (1.f + expf(-v));
      else if (n == 6) v = v >= 0.f ? v : 0.01f * v;
    }
    out[m * ldo + n] = v;
  }
}

// ------------------------------------------------------------------------------------------------ pooling
// mean over reads -> pool[cand][p][c]   (model.py:772)
__global__ void pool_mean_kernel(const float* __restrict__ h, int ldh, float* __restrict__ pool, int C, RowGeom g) {
  const int cand = blockIdx.y, pp = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < g.R; ++r) s += h[(((long)cand * g.R + r) * g.pitch + pp) * ldh + c];
    pool[((long)cand * g.P + pp) * C + c] = s / (float)g.R;
  }
}

// out = h + pool broadcast over reads (model.py:742); gap rows stay zero
__global__ void add_pool_kernel(const float* __restrict__ h, const float* __restrict__ pool, float* __restrict__ out,
                                int C, RowGeom g, long rows) {
  const long total = rows * (C / 4);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / (C / 4); const int c4 = (int)(i % (C / 4));
    const int pp = (int)(row % g.pitch); const long cand = row / ((long)g.R * g.pitch);
    float4 v = reinterpret_cast<const float4*>(h)[i];
    if (pp < g.P) {
      const float4 a = reinterpret_cast<const float4*>(pool)[(cand * g.P + pp) * (C / 4) + c4];
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// final max ‖ mean over reads, flattened like model.py:833-839: feature (c*P + p) max block first, then mean
__global__ void pool_final_kernel(const float* __restrict__ h, int ldh, float* __restrict__ dst, int ld_dst, int C,
                                  RowGeom g, int skip_max) {
  const int cand = blockIdx.y, pp = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, mx = -INFINITY;
    for (int r = 0; r < g.R; ++r) {
      const float v = h[(((long)cand * g.R + r) * g.pitch + pp) * ldh + c];
      s += v; mx = fmaxf(mx, v);
    }
    float* row = dst + (long)cand * ld_dst;
    if (skip_max) {
      row[c * g.P + pp] = s / (float)g.R;
    } else {
      row[c * g.P + pp] = mx;
      row[(C + c) * g.P + pp] = s / (float)g.R;
    }
  }
}

// highway vectors hw[l][read][o] -> FC input section: relu(concat) at l*bott*R + o*R + r, or relu(mean over layers)
// (model.py:853-859)
__global__ void highway_assemble_kernel(const float* __restrict__ hw, long layer_stride, int L, int bott, int R,
                                        int concat, float* __restrict__ dst, int ld_dst, int cands) {
  const int per = bott * R;
  const long total = (long)cands * (concat ? L : 1) * per;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int feat = (int)(i % ((concat ? L : 1) * per)); const long cand = i / ((concat ? L : 1) * per);
    const int l = feat / per, o = (feat % per) / R, r = feat % R;
    float v;
    if (concat) {
      v = hw[l * layer_stride + (cand * R + r) * bott + o];
    } else {
      v = 0.f;
      for (int k = 0; k < L; ++k) v += hw[k * layer_stride + (cand * R + r) * bott + o];
      v /= (float)L;
    }
    dst[cand * ld_dst + feat] = fmaxf(v, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ packing
// out[k][n] = in[n][...] permutations, run once per load_state_dict
__global__ void pack_conv_w_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int CinPad, int taps) {
  // w (Cout, Cin, 1, taps) -> out[(t*CinPad + c)][n]
  const long total = (long)taps * CinPad * Cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout); const long k = i / Cout;
    const int c = (int)(k % CinPad), t = (int)(k / CinPad);
    out[i] = c < Cin ? w[((long)n * Cin + c) * taps + t] : 0.f;
  }
}
__global__ void pack_comp_w_kernel(const float* __restrict__ w, float* __restrict__ out, int bott, int P) {
  // w (O, Cb, 1, P) -> out[(p*Cb + c)][o]
  const long total = (long)P * bott * bott;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int o = (int)(i % bott); const long k = i / bott;
    const int c = (int)(k % bott), pp = (int)(k / bott);
    out[i] = w[((long)o * bott + c) * P + pp];
  }
}
__global__ void pack_linear_w_kernel(const float* __restrict__ w, float* __restrict__ out, int N, int K, int KPad, int NPad) {
  // w (N, K) -> out[k][n] with zero padding to (KPad, NPad)
  const long total = (long)KPad * NPad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i % NPad); const long k = i / NPad;
    out[i] = (n < N && k < K) ? w[(long)n * K + k] : 0.f;
  }
}
__global__ void pack_bn_kernel(const float* g, const float* b, const float* mean, const float* var, float* scale,
                               float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float inv = 1.0f / sqrtf(var[c] + 1e-5f);   // eps of nn.BatchNorm2d (model.py:217)
    const float sc = g[c] * inv;
    scale[c] = sc;
    shift[c] = b[c] - mean[c] * sc;
  }
}
__global__ void pad_copy_kernel(const float* src, float* dst, int n, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? src[i] : 0.f;
}

inline int grid_for(long total, int block = 256) {
  long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

int launch_gemm(const GemmParams& p, cudaStream_t st) {
  const int bn = p.N <= 32 ? 32 : 128;
  dim3 grid((unsigned)((p.M + 127) / 128), (unsigned)((p.N + bn - 1) / bn), (unsigned)p.splits);
  if (bn == 32) sgemm_taps_kernel<32><<<grid, 256, 0, st>>>(p);
  else sgemm_taps_kernel<128><<<grid, 256, 0, st>>>(p);
  dan_count_launch();
  DAN_CUDA_TRY(cudaGetLastError());
  return DAN_OK;
}

}  // namespace
