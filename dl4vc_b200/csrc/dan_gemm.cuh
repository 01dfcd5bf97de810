// dan_gemm.cuh — batched bf16 GEMM on tcgen05 with TMA-tiled operands (128-byte swizzle). Included by dan_bf16.cu.
//
//     C[b][m][n] (+)= sum_k A[b][m][k] * B[b][n][k]          A, B row-major bf16 (K contiguous), fp32 accumulation in TMEM
//
// Used for the highway compression conv (model.py:776: a (1 x 201) conv over 32 channels = one K = 6432 contraction per
// read, all layers of a pass batched in one launch), the FC trunk (model.py:917) and the six heads (model.py:919-958).
// Operands are fetched by cp.async.bulk.tensor (3-D maps: k, row, batch; out-of-range K is zero-filled by the TMA unit)
// into the canonical K-major SWIZZLE_128B shared-memory layout; one warp produces, one warp issues the MMAs
// (warp-uniform control flow, elected lane), four warps run the epilogue out of TMEM.
#pragma once
#include <cuda.h>

namespace {

constexpr int kG2Threads = 192;        // warp 0 TMA producer | warp 1 MMA issuer (+ TMEM alloc) | warps 2-5 epilogue
constexpr int kG2BK = 64;              // K elements per stage (= 128 bytes = one swizzle atom row)

enum { kG2Raw = 0, kG2BiasReluBf16 = 1, kG2Heads = 2 };

struct Gemm2Params {
  int M, N, K;                         // per batch entry
  int m_tiles, n_tiles, splits, batch;
  int mode;
  float* out; long out_batch_stride; int ldo;          // kG2Raw: fp32 [batch][M][ldo]; kG2Heads: fp32 [M][27]
  float* part;                                         // kG2Raw with splits > 1 (set by the caller): partial sums [split][batch][M][N], added in split order by splitk_reduce_kernel
  const float* bias;                                   // modes 1, 2
  __nv_bfloat16* out_bf16; int ld_bf16;                // kG2BiasReluBf16: bf16 [M][ld_bf16]
};

template <int BN>
__host__ __device__ constexpr int g2_stages() { return BN <= 32 ? 8 : 6; }
template <int BN>
__host__ __device__ constexpr size_t g2_smem_bytes() { return 1024 + (size_t)g2_stages<BN>() * (128 + BN) * 128 + 1024; }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: rows are 128 bytes apart, 8-row groups 1024 bytes apart (SBO), tile 1024-byte aligned
__device__ __forceinline__ uint64_t g2_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO: unused for swizzled K-major layouts
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(kG2Threads, 1) tma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ Gemm2Params p) {
  constexpr int NS = g2_stages<BN>();
  constexpr uint32_t kABytes = 128 * 128, kBBytes = BN * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + NS;
  uint64_t* done = empty + NS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* stage_base = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int w = blockIdx.x;
  const int m_tile = w % p.m_tiles; w /= p.m_tiles;
  const int n_tile = w % p.n_tiles; w /= p.n_tiles;
  const int split = w % p.splits; w /= p.splits;
  const int b = w;
  const int k_blocks = (p.K + kG2BK - 1) / kG2BK;
  const int per = (k_blocks + p.splits - 1) / p.splits;
  const int kb_begin = split * per, kb_end = min(k_blocks, kb_begin + per);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (kb_begin < kb_end) {
    if (warp == 0) {
      int i = 0; uint32_t par = 1;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty[i], par);
        if (elect_one()) {
          mbar_expect_tx(&full[i], kABytes + kBBytes);
          uint8_t* a_dst = stage_base + (size_t)i * (kABytes + kBBytes);
          tma_load_3d(a_dst, &mapA, kb * kG2BK, m_tile * 128, b, &full[i]);
          tma_load_3d(a_dst + kABytes, &mapB, kb * kG2BK, n_tile * BN, b, &full[i]);
        }
        __syncwarp();
        if (++i == NS) { i = 0; par ^= 1; }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16(128, BN);
      int i = 0; uint32_t par = 0, acc = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[i], par);
        tc_fence_after();
        const uint32_t a0 = smem_u32(stage_base + (size_t)i * (kABytes + kBBytes)), b0 = a0 + kABytes;
        const uint64_t ad = g2_desc(a0), bd = g2_desc(b0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kG2BK / 16; ++k) {          // +32 bytes per K = 16 step inside the 128-byte swizzle row
            umma_bf16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, acc);
            acc = 1;
          }
          umma_commit(&empty[i]);
        }
        __syncwarp();
        acc = 1;
        if (++i == NS) { i = 0; par ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int row = m_tile * 128 + q * 32 + lane;
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        if (row < p.M) {
          const int n0 = n_tile * BN + c * 32;
          if (p.mode == kG2Raw) {
            float* orow = p.splits > 1 ? p.part + (((long)split * p.batch + b) * p.M + row) * p.N + n0
                                       : p.out + (long)b * p.out_batch_stride + (long)row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else if (p.mode == kG2BiasReluBf16) {
            uint4* orow = reinterpret_cast<uint4*>(p.out_bf16 + (long)row * p.ld_bf16 + n0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(r[g * 8 + j]) + __ldg(p.bias + n0 + g * 8 + j), 0.f);
              orow[g] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
          } else {                                        // heads: [xbinary|xVT|sigmoid(xAF)|leaky_relu(xCov)|xVB|xVR], model.py:919-958
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              if (n < DAN_NUM_HEAD_OUTPUTS) {
                float v = __uint_as_float(r[j]) + __ldg(p.bias + n);
                if (n == 5) v = 1.f / (1.f + expf(-v));
                else if (n == 6) v = v >= 0.f ? v : 0.01f * v;
                p.out[(long)row * DAN_NUM_HEAD_OUTPUTS + n] = v;
              }
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
}

// fixed-order sum of the split-K partials (a candidate's numbers must not depend on the batch it arrives in)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int batch, int M, int N, float* __restrict__ out, long out_batch_stride, int ldo) {
  const long total = (long)batch * M * N;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[(long)s * total + i];
    const int n = (int)(i % N); const long t = i / N; const int row = (int)(t % M); const int b = (int)(t / M);
    out[(long)b * out_batch_stride + (long)row * ldo + n] = v;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*G2EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline G2EncodeTiledFn g2_encode_fn() {
  static G2EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<G2EncodeTiledFn>(ptr);
  }
  return fn;
}

// 3-D map (k, row, batch) over row-major bf16 matrices; strides in bytes (multiples of 16); box = 64 x box_rows x 1, 128-byte swizzle
inline int g2_make_map(CUtensorMap* map, const void* base, uint64_t K, uint64_t rows, uint64_t batch, uint64_t row_stride, uint64_t batch_stride, uint32_t box_rows) {
  G2EncodeTiledFn fn = g2_encode_fn();
  if (!fn) { dan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return DAN_E_CUDA; }
  const cuuint64_t dims[3] = {K, rows, batch};
  const cuuint64_t strides[2] = {row_stride, batch_stride ? batch_stride : row_stride * rows};
  const cuuint32_t box[3] = {(cuuint32_t)kG2BK, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dan_set_error("cuTensorMapEncodeTiled failed (%d): K %llu rows %llu batch %llu stride %llu", (int)r, (unsigned long long)K, (unsigned long long)rows, (unsigned long long)batch, (unsigned long long)row_stride); return DAN_E_CUDA; }
  return DAN_OK;
}

struct Gemm2Operand { const void* base; long rows; long row_stride_bytes; long batch_stride_bytes; };   // rows = allocated rows per batch entry

template <int BN>
int g2_launch(const CUtensorMap& ma, const CUtensorMap& mb, const Gemm2Params& g, cudaStream_t st) {
  static DanSmemAttr attr;
  DAN_CUDA_TRY(attr.ensure(tma_gemm_kernel<BN>, g2_smem_bytes<BN>()));
  { DanProfScope ps(DAN_PROF_GEMM, st); tma_gemm_kernel<BN><<<g.m_tiles * g.n_tiles * g.splits * g.batch, kG2Threads, g2_smem_bytes<BN>(), st>>>(ma, mb, g); }
  dan_count_launch();
  DAN_CUDA_TRY(cudaGetLastError());
  return DAN_OK;
}

// C = A * B^T for `batch` independent problems; `g` carries M, N, K, mode and the output pointers
int run_gemm2(const Gemm2Operand& A, const Gemm2Operand& B, Gemm2Params g, int batch, int num_sms, cudaStream_t st) {
  const int bn = (g.N % 64 == 0) ? 64 : 32;
  g.batch = batch;
  g.m_tiles = (g.M + 127) / 128; g.n_tiles = g.N / bn;
  const int k_blocks = (g.K + kG2BK - 1) / kG2BK;
  // split-K only with a partition that does not depend on the batch shape (`g.splits` fixed by the caller): a candidate's numbers
  // must not move with the batch it arrives in
  int splits = g.part && g.splits > 1 ? g.splits : 1;
  if (splits > k_blocks) splits = k_blocks;
  g.splits = splits;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = g2_make_map(&ma, A.base, (uint64_t)g.K, (uint64_t)A.rows, (uint64_t)batch, (uint64_t)A.row_stride_bytes, (uint64_t)A.batch_stride_bytes, 128))) return rc;
  if ((rc = g2_make_map(&mb, B.base, (uint64_t)g.K, (uint64_t)B.rows, (uint64_t)batch, (uint64_t)B.row_stride_bytes, (uint64_t)B.batch_stride_bytes, (uint32_t)bn))) return rc;
  if ((rc = bn == 64 ? g2_launch<64>(ma, mb, g, st) : g2_launch<32>(ma, mb, g, st))) return rc;
  if (splits > 1) {
    const long total = (long)batch * g.M * g.N;
    splitk_reduce_kernel<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, st>>>(g.part, splits, batch, g.M, g.N, g.out, g.out_batch_stride, g.ldo);
    dan_count_launch();
    DAN_CUDA_TRY(cudaGetLastError());
  }
  return DAN_OK;
}

}  // namespace
