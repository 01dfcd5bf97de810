// dan_stack.cuh — fused, persistent conv-stack kernel (bf16 tcgen05). Included by dan_bf16.cu.
//
// One launch runs a SEGMENT of consecutive conv layers (a maximal run without a read-axis pool-add in between: PROD = layers
// 1-2, then layers 3-7; dl4vc/model.py:728-778) for every read of a pass. A read (201 positions x 128 channels, bf16) is
// built (segment 1: embedding gather + positional / reference / q-score / strand / match-mask channels straight from the
// loader's uint8 tiles, model.py:450-627) or loaded once into shared memory, goes through all layers of the segment IN PLACE,
// and leaves the SM only as what the rest of the network consumes:
//   * the bottleneck outputs T of every layer (A operand of the highway compression GEMM, model.py:774-776),
//   * segment 1: the layer-2 output planes (re-read by segment 2) and their read-axis SUM (model.py:766-772),
//   * last segment: the read-axis MAX and SUM of the final layer (model.py:824-826); the per-read activations of the final
//     layer are never written.
// The read-axis reductions are done by the TMA engine on the read's output planes while they are still in shared memory
// (cp.reduce.async.bulk .max.bf16 / .add.noftz.bf16 into per-candidate accumulators in L2): max is exact; sums are kept per
// (block of 20 reads, slot) so that every accumulator sees at most 10 bf16 additions in a fixed order before the finishing
// kernel adds the groups in fp32 (deterministic, 7e-4 rms of the sum). The first read of a group STORES, so nothing is zeroed.
//
// Orientation ("D^T"): for one read and one layer the tensor core computes
//       D[cout 0..127][position 0..207] = sum_tap  W_tap[cout][cin] * X[position + (tap-1)*dil][cin]
// i.e. M = 128 output channels (A operand = weights, streamed from L2 through a shared-memory ring), N = 208 positions
// (B operand = the read's activation buffer; a dilated tap is a row offset of the descriptor start address, the zero
// rows around the read implement Conv2d's zero padding, model.py:214-229), accumulator = 208 TMEM columns.
// The epilogue reads the accumulator in the mma C-fragment layout (tcgen05.ld 16x256b), applies +bias -> ReLU -> BatchNorm
// (model.py:749-751) and writes bf16 back into the activation buffer with stmatrix.trans (free transpose).
// Residual layers (model.py:753-761): the epilogue pre-loads the accumulator with x + b_res (tcgen05.st), a second MMA pass
// accumulates W_res * y on top, and a second epilogue writes the layer output.
// Highway bottleneck (model.py:773-774): positions-as-M orientation (D3[position][32]); it reads the same buffer state as the
// NEXT layer's conv, so the two are issued as ONE op — bottleneck MMAs first, into a small accumulator of their own (TMEM
// columns 224..255 and 480..511, shared by the two slots under a small lock), then the conv MMAs: the bottleneck
// epilogue (relu, bf16, T store) runs under the conv MMAs instead of forming an op + epilogue round trip of its own.
//
// Per CTA: two independent read pipelines ("slots": own main accumulator, own epilogue warps, own issuer warp). ONE weight
// ring (7 x 16 KB stages) feeds both: slot 1 runs one op behind slot 0 and every stage is consumed twice before it is
// refilled. Work is handed out in blocks of 20 reads of one candidate; a block with an odd number of reads is completed by
// a phantom read (same op sequence, no global side effects), so both slots always run identical op sequences.
#pragma once

namespace {

constexpr int kStkThreads = 640;       // warps 0-7 epilogue slot 0 | 8-15 epilogue slot 1 | 16,17 weight producers | 18,19 MMA issuers
constexpr int kStkEpiThreads = 256;    // epilogue threads per slot: 4 TMEM lane quadrants x 2 position halves
constexpr int kStkLead = 2;            // zero rows in front of position 0 (>= largest dilation)
constexpr int kStkN = 208;             // MMA N = positions per read, padded to a multiple of 16
constexpr int kStkRB = kStkLead + kStkN + 2;   // rows per channel-chunk plane of a read buffer (212)
constexpr int kStkPlane = kStkRB * 16;         // bytes per plane (3392)
constexpr int kStkBuf = kKC * kStkPlane;       // bytes per read buffer (54272)
constexpr int kStkStageBytes = 16384;  // weight-ring stage = four k-step blocks of 4 KB
constexpr int kStkStages = 7;          // stages of the weight ring shared by the two slots (>= the largest op: bottleneck 1 + conv 6)
constexpr int kStkMaxSeg = 8;          // layers per segment
constexpr int kStkRegsIssue = 64, kStkRegsEpi = 104;   // setmaxnreg redistributes the launch allocation (640 x 96 = 61440): 128 x 64 + 512 x 104 = 61440
constexpr int kStkBmapChunk = 64;       // uint4 per (role, 16-position chunk) of the pool bias map: 32 lanes x 2 (bmap_pack_kernel, dan_bf16.cu)
constexpr int kStkBmapPerCand = 8 * 7 * kStkBmapChunk;   // uint4 per candidate: 8 roles (position half, lane quadrant) x 7 chunks
constexpr int kStkBott = 32;            // bottleneck width the fused kernel is built for (PROD; other widths take the layer-wise path)
constexpr int kStkAccStride = 256;      // TMEM columns between the two slots' main accumulators (208 used)
constexpr int kStkBottCol0 = 224, kStkBottCol1 = 480;   // shared bottleneck accumulator: position tile 0 / 1 (32 columns each) in the gaps behind the main accumulators
constexpr int kStkBlockReads = 20;      // work unit: consecutive reads of one candidate = one sum group per slot
constexpr int kStkSmemHeader = 3072;    // barriers, TMEM pointer, bottleneck biases of the segment
constexpr size_t kStkSmemBytes = kStkSmemHeader + 2 * (size_t)kStkBuf + kStkStages * kStkStageBytes;

enum { kStkInPlanes = 0, kStkInEncode = 1 };

struct StackLayer {
  const uint8_t* wstream;   // conv k-step blocks (tap-major) | residual blocks | bottleneck blocks, contiguous
  size_t wreplica_stride;   // byte distance between the kWeightReplicas copies of wstream
  const float* chan;        // [6][128] epilogue constants (dan_bf16.cu chan_table_kernel)
  const float* bbias;       // [bott]
  uint4* tout;              // T[read][c/8][p][c%8] (bf16) of this layer: row-major A operand (K = (c/8, p, c%8)) of the compression GEMM
  int conv_blocks;          // 3 * kc_in / 2
  int kc_in;                // 16-byte pieces per input row (CinPad/8 for layer 1, else 16)
  int dil, residual;
};

struct StackParams {
  int in_mode;                          // kStkInPlanes: chunk-major rows from `in`; kStkInEncode: built from the uint8 pileup tiles
  const uint4* in; long in_kstride;     // chunk-major input rows (dan_bf16.cu), kc_in planes
  DevInputs bytes; long cand0;          // encode mode: loader tensors of the batch and the first candidate of this pass
  const uint4* enc_tab;                 // encode mode: bf16(E[tok] + pe[p]) as [P][10 tokens][3 pieces] (enc_table_kernel)
  uint4* out; long out_kstride;         // optional: chunk-major output rows of the segment, 16 planes
  uint4* sums; int groups_per_cand;     // optional: read-axis sum groups [candidate][group][c/8][p] pieces (bf16)
  uint4* maxv; long max_stride;         // optional: read-axis max [candidate * max_stride + (c/8) * P + p] pieces (bf16), pre-set to -inf
  const uint4* bmap;                    // optional pool bias map conv(pool) + bias of the segment's first layer, [candidate][kStkBmapPerCand]
  int cands, R, P, pitch, highway, num_layers;
#ifdef DAN_STK_PROF
  unsigned long long* prof;             // development build: [grid][48] cycle counters
#endif
  StackLayer layer[kStkMaxSeg];
};

struct StackSmem {
  uint64_t w_full[kStkStages], w_empty[kStkStages];
  uint64_t acc_full[2], act_ready[2], in_full[2], bott_full[2];
  uint32_t bott_busy, bott_drained;    // the shared bottleneck accumulator: taken by an issuer (CAS 0 -> 1), given back by the last of the 256 epilogue threads that read it out
  uint32_t tmem_base;
#ifdef DAN_STK_PROF
  long long t0;
#endif
  uint32_t issued_ops;                 // ops fully issued by slot 0's issuer (slot 1 runs one op behind, see the issuer)
  float bbias[kStkMaxSeg][kStkBott];
};
static_assert(sizeof(StackSmem) <= kStkSmemHeader, "header too small");

// Walks the read pairs of a CTA's share of the pass: blocks of kStkBlockReads reads of one candidate, two reads (slot 0, slot 1) per
// step. Carried incrementally (no division on the per-read path): candidate, block within the candidate, pair within the block.
struct StkIter {
  int cand, bic, pr, left;      // left = blocks of this CTA still to do (including the current one)
  __device__ static int bpc(int R) { return (R + kStkBlockReads - 1) / kStkBlockReads; }
  __device__ void init(int cands, int R) {
    const int total = cands * bpc(R), per = total / (int)gridDim.x, rem = total % (int)gridDim.x;
    const int first = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
    left = per + ((int)blockIdx.x < rem ? 1 : 0);
    cand = first / bpc(R); bic = first - cand * bpc(R); pr = 0;
  }
  __device__ bool done() const { return left <= 0; }
  __device__ int nr(int R) const { return min(kStkBlockReads, R - bic * kStkBlockReads); }
  __device__ void next(int R) {
    if (++pr == (nr(R) + 1) >> 1) {
      pr = 0; --left;
      if (++bic == bpc(R)) { bic = 0; ++cand; }
    }
  }
  __device__ bool valid(int R, int s) const { return 2 * pr + s < nr(R); }       // slot 1 of the last pair of an odd block is a phantom
  __device__ int read_in_cand(int s) const { return bic * kStkBlockReads + 2 * pr + s; }
  __device__ int group(int s) const { return 2 * bic + s; }
};
__device__ inline int stk_total_pairs(int cands, int R) {
  StkIter it;
  it.init(cands, R);
  int n = 0;
  for (; it.left > 0; --it.left) {
    n += (it.nr(R) + 1) >> 1;
    if (++it.bic == StkIter::bpc(R)) it.bic = 0;
  }
  return n;
}

__device__ __forceinline__ void bulk_reduce_max_bf16(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.max.bf16 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(ptx::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_bf16(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(ptx::smem_u32(smem_src)), "r"(bytes) : "memory");
}
// barrier + AND-reduction of a predicate over the `threads` threads that use named barrier `id` (1 or 2)
__device__ __forceinline__ bool named_bar_and(uint32_t id, uint32_t threads, bool pred) {
  uint32_t r;
  if (id == 1) asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %2, 0;\n\tbarrier.cta.red.and.pred p, 1, %1, q;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(threads), "r"((uint32_t)pred) : "memory");
  else asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %2, 0;\n\tbarrier.cta.red.and.pred p, 2, %1, q;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(threads), "r"((uint32_t)pred) : "memory");
  return r != 0;
}

}  // namespace
#include "dan_stack_epi.cuh"
namespace {

#ifdef DAN_STK_PROF
// besides the per-category sums, CTA 0 logs (category, cycle since CTA start) of every mark of events [kStkTraceSkip, +kStkTraceCap) per role and slot
constexpr int kStkTraceRead0 = 10, kStkTraceReads = 4, kStkTraceCap = 1024;     // reads [10, 14) of each slot
#define STK_PROF_DECL long long prof_t0 = clock64(), prof_acc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_begin = prof_t0; \
  unsigned long long* prof_tr = nullptr; int prof_n = 0; bool prof_on = false
#define STK_TRACE_ROLE(cond, r) prof_tr = (blockIdx.x == 0 && (cond)) ? p.prof + (size_t)gridDim.x * 64 + (size_t)(r) * kStkTraceCap : nullptr
#define STK_PROF(i) do { const long long prof_t1 = clock64(); prof_acc[i] += prof_t1 - prof_t0; prof_t0 = prof_t1; \
    if (prof_tr && prof_on && prof_n < kStkTraceCap) prof_tr[prof_n++] = ((unsigned long long)(i) << 48) | (unsigned long long)(prof_t1 - sm->t0); } while (0)
#define STK_TRACE_GATE(rd) prof_on = (rd) >= kStkTraceRead0 && (rd) < kStkTraceRead0 + kStkTraceReads
#define STK_PROF_FLUSH(cond, base) do { if (cond) { for (int i_ = 0; i_ < 14; ++i_) p.prof[blockIdx.x * 64 + (base) + i_] = prof_acc[i_]; } } while (0)
#else
#define STK_PROF_DECL
#define STK_PROF(i)
#define STK_PROF_FLUSH(cond, base)
#define STK_TRACE_ROLE(cond, r)
#define STK_TRACE_GATE(rd)
#endif

// descriptor words: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14  (tcgen05_ptx.cuh make_smem_desc)
__device__ __forceinline__ uint64_t stk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// ---- pileup encoder of one read (model.py:450-627,719), run by the 256 epilogue threads of a slot: thread t owns position t.
// Bytes of the read (token, q-score, strand) are requested early (stk_enc_fetch) and turned into the six 8-channel planes of the
// conv-1 input later (stk_enc_rows / stk_enc_store):  0-19 E[tok]+pe | 20-39 E[ref]+pe | 40 q*0.01 | 41 strand*0.5 | 42 ref-match | 43 var-match |
// 44 var-length (from the REF mask, model.py:579,584) | 45-47 zero. Integer work (agreement of the read with the ref / var
// proposal over all 201 positions) is an AND-reduction over the slot's threads; the float channels come from a bf16 table of
// E[tok] + pe[p] built once per weight load, so every value is bf16_rn of the reference's fp32 value.
typedef uint32_t StkEncBytes;      // token | q-score << 8 | strand << 16

__device__ __forceinline__ StkEncBytes stk_enc_fetch(const StackParams& p, long cand, int r, int pos) {
  StkEncBytes b = 0u;
  if (pos < p.P) {
    const long off = (cand * p.P + pos) * p.R + r;
    b = (uint32_t)__ldg(p.bytes.reads + off) | (uint32_t)__ldg(p.bytes.q + off) << 8 | (uint32_t)__ldg(p.bytes.strands + off) << 16;
  }
  return b;
}

// per-candidate bytes of position `pos` (the same for every read of the candidate): ref token | ref-mask << 8 | var-mask << 16
__device__ __forceinline__ uint32_t stk_enc_fetch_cand(const StackParams& p, long cand, int pos) {
  if (pos >= p.P) return 0u;
  return (uint32_t)__ldg(p.bytes.ref + cand * p.P + pos) | (uint32_t)__ldg(p.bytes.ref_masks + cand * p.P + pos) << 8 | (uint32_t)__ldg(p.bytes.var_masks + cand * p.P + pos) << 16;
}
// The six 16-byte rows of position `pos` (one per input chunk plane). stk_enc_rows does the integer work and REQUESTS the table rows; it runs
// while the buffer is still being read out by the TMA engine, and stk_enc_store writes the rows once the buffer is free: the two block-wide
// reductions and the L2 latency of the table stay off the read-to-read chain.
struct StkEncRows { uint4 r0, r1, r2, a0, a1, a2; uint32_t misc; };
__device__ __forceinline__ StkEncRows stk_enc_rows(const StackParams& p, int pos, int s, StkEncBytes eb, uint32_t cb) {
  const uint32_t tok = eb & 0xFFu, refp = cb & 0xFFu, rmk = (cb >> 8) & 0xFFu, vmk = cb >> 16;
  const bool agreeR = named_bar_and(1 + s, kStkEpiThreads, rmk == 0 || tok == rmk);       // model.py:592-593
  const bool agreeV = named_bar_and(1 + s, kStkEpiThreads, vmk == 0 || tok == vmk);       // model.py:607-608
  StkEncRows e{};
  if (pos < p.P) {
    const uint4* tr = p.enc_tab + ((long)pos * DAN_VOCAB + min(tok, (uint32_t)DAN_VOCAB - 1)) * 3;
    const uint4* tf = p.enc_tab + ((long)pos * DAN_VOCAB + min(refp, (uint32_t)DAN_VOCAB - 1)) * 3;
    e.r0 = __ldg(tr); e.r1 = __ldg(tr + 1); e.r2 = __ldg(tr + 2); e.a0 = __ldg(tf); e.a1 = __ldg(tf + 1); e.a2 = __ldg(tf + 2);
    e.misc = (rmk != 0 && agreeR ? 1u : 0u) | (vmk != 0 && agreeV ? 2u : 0u) | (rmk != 0 ? 4u : 0u);
  }
  return e;
}
__device__ __forceinline__ void stk_enc_store(const StackParams& p, int pos, StkEncBytes eb, const StkEncRows& e, uint8_t* buf) {
  if (pos < p.P) {
    const uint32_t qv = (eb >> 8) & 0xFFu, sv = eb >> 16;
    const float m0 = (e.misc & 1u) ? 1.f : 0.f, m1 = (e.misc & 2u) ? 1.f : 0.f, m2 = (e.misc & 4u) ? 1.f : 0.f;
    uint4* row = reinterpret_cast<uint4*>(buf + (size_t)(kStkLead + pos) * 16);
    constexpr int kPl = kStkPlane / 16;
    row[0] = e.r0; row[kPl] = e.r1;
    row[2 * kPl] = make_uint4(e.r2.x, e.r2.y, e.a0.x, e.a0.y);
    row[3 * kPl] = make_uint4(e.a0.z, e.a0.w, e.a1.x, e.a1.y);
    row[4 * kPl] = make_uint4(e.a1.z, e.a1.w, e.a2.x, e.a2.y);
    row[5 * kPl] = make_uint4(pack_bf16x2((float)qv * 0.01f, (float)sv * 0.5f), pack_bf16x2(m0, m1), pack_bf16x2(m2, 0.f), 0u);   // model.py:24,16
  }
}

__global__ void __launch_bounds__(kStkThreads, 1) dan_stack_kernel(const __grid_constant__ StackParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  StackSmem* sm = reinterpret_cast<StackSmem*>(smem);
  uint8_t* bufs = smem + kStkSmemHeader;
  uint8_t* rings = bufs + 2 * (size_t)kStkBuf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t plane_bytes = (uint32_t)p.P * 16;
  const int in_kc = p.layer[0].kc_in;
  const int n_layers = p.num_layers;
  const bool highway = p.highway != 0;

  {  // zero rows / planes must read as 0 until an epilogue or a load writes them
    uint4* z = reinterpret_cast<uint4*>(bufs);
    for (int i = threadIdx.x; i < 2 * kStkBuf / 16; i += kStkThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
#ifdef DAN_STK_PROF
    sm->t0 = clock64();
#endif
    for (int i = 0; i < kStkStages; ++i) { mbar_init(&sm->w_full[i], 1); mbar_init(&sm->w_empty[i], 2); }   // both slots release a stage
    sm->issued_ops = 0;
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm->acc_full[s], 1); mbar_init(&sm->act_ready[s], kStkEpiThreads); mbar_init(&sm->in_full[s], 1); mbar_init(&sm->bott_full[s], 1);
    }
    sm->bott_busy = 0; sm->bott_drained = 0;
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < n_layers * kStkBott; i += kStkThreads) {
    const int l = i / kStkBott, c = i - l * kStkBott;
    sm->bbias[l][c] = highway ? p.layer[l].bbias[c] : 0.f;
  }
  fence_proxy_async_smem();
  if (warp == 18) tmem_alloc<512>(&sm->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;

  const int n_pairs = stk_total_pairs(p.cands, p.R);

  // register budget: the producer / issuer warpgroup (warps 16-19) hands registers to the four epilogue warpgroups
  if (warp >= 16) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kStkRegsIssue));
  if (warp == 16 || warp == 17) {
    // ===================== weight producer: ONE stream for both slots (every stage is consumed twice before it is refilled, which
    // halves the L2 -> SMEM weight traffic: measured ~100 W of board power against private per-slot rings at the same speed). Two
    // producer threads (lane 0 of warps 16 and 17) take alternate stages: one thread needs ~300 cycles per stage (empty-barrier
    // probe + expect_tx + copy issue), close to the rate at which the tensor pipe drains one.
    if (lane == 0) {
      const uint32_t mine = (uint32_t)(warp - 16);
      const uint64_t keep = l2_policy_evict_last();          // the weight images are re-read by every CTA for every pair of reads
      uint32_t idx = 0, par = 1, seq = 0;     // first pass over the ring: the "empty" phase counts as complete
      auto emit = [&](const uint8_t* src, uint32_t bytes) {
        for (uint32_t off = 0; off < bytes; off += kStkStageBytes) {
          if ((seq++ & 1u) == mine) {
            const uint32_t n = min((uint32_t)kStkStageBytes, bytes - off);
            mbar_wait(&sm->w_empty[idx], par);
            mbar_expect_tx(&sm->w_full[idx], n);
            bulk_g2s_hint(rings + (size_t)idx * kStkStageBytes, src + off, n, &sm->w_full[idx], keep);
          }
          if (++idx == kStkStages) { idx = 0; par ^= 1; }
        }
      };
      for (int pr = 0; pr < n_pairs; ++pr) {
        for (int l = 0; l < n_layers; ++l) {
          const StackLayer& L = p.layer[l];
          const uint32_t conv_bytes = (uint32_t)L.conv_blocks * 4096u;
          const uint8_t* w = L.wstream + (size_t)(blockIdx.x % kWeightReplicas) * L.wreplica_stride;
          emit(w, conv_bytes);
          if (L.residual) emit(w + conv_bytes, kKC / 2 * 4096u);
          if (highway) emit(w + conv_bytes + (L.residual ? kKC / 2 * 4096u : 0u), (uint32_t)kKC * kStkBott * 16u);
        }
      }
    }
  } else {
    // ===================== MMA issuer of slot s. The whole warp runs the (blocking, strictly sequential) op schedule of its slot —
    // warp-uniform control flow keeps descriptors in uniform registers — and one elected lane issues the tcgen05 instructions. The
    // two slots' issuers are independent warps: the tensor pipe interleaves their MMA streams, so one slot's epilogue runs under the
    // other slot's MMAs. Op sequence of a read:  conv 1 | [bott l-1 + conv l] (| residual l) ... | bott L.
    const int s = warp - 18;
    const uint32_t idesc_main = make_idesc_bf16(128, kStkN);
    const uint32_t idesc_bott = make_idesc_bf16(128, kStkBott);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo_w = (2048u >> 4) << 16, b_lbo_x = ((uint32_t)kStkPlane >> 4) << 16, b_lbo_bott = (((uint32_t)kStkBott * 16u) >> 4) << 16;
    constexpr int kBottPerStage = kKC / 2;                               // the bottleneck's 8 k-steps of weights (8 KB) sit in one ring stage
    static_assert(kKC / 2 * kStkBott * 32 <= kStkStageBytes, "bottleneck weights must fit one stage");
    const uint32_t ring_lo = smem_u32(rings) >> 4;
    uint64_t* const wfull = &sm->w_full[0];
    uint64_t* const wempty = &sm->w_empty[0];
    volatile uint32_t* const issued = &sm->issued_ops;
    uint32_t gops = 0;                                                                            // ops started by this issuer
    const uint32_t d_main = tmem_base + (uint32_t)s * kStkAccStride, d_bott0 = tmem_base + kStkBottCol0, d_bott1 = tmem_base + kStkBottCol1;
    const uint32_t x_lo = (smem_u32(bufs) >> 4) + (uint32_t)s * (kStkBuf >> 4) + kStkLead;       // centre row of chunk plane 0
    constexpr uint32_t kStep = 2 * (kStkPlane >> 4);                                              // one k-step = two chunk planes
    uint32_t wi = 0, wp = 0, opc = 0;      // ring position / parity, hand-overs awaited
    // The full-barrier probe of an op's first stage(s) is issued BEFORE the op's dependencies are awaited (prewait*): a probe costs
    // ~190 cycles even when the phase is complete, and there it would sit between the epilogue's hand-over and the op's first MMA.
    bool prewaited = false;
    STK_PROF_DECL;
    STK_TRACE_ROLE(lane == 0, s);
    auto wait_w = [&]() {
      if (prewaited) { prewaited = false; return; }
      STK_PROF(0);
      mbar_wait(&wfull[wi], wp);
      tc_fence_after();
      STK_PROF(1);
    };
    auto wait_w2 = [&](uint32_t wi1, uint32_t wp1) {
      if (prewaited) { prewaited = false; return; }
      STK_PROF(0);
      mbar_wait2(&wfull[wi], wp, &wfull[wi1], wp1);
      tc_fence_after();
      STK_PROF(1);
    };
    auto prewait1 = [&]() { wait_w(); prewaited = true; };
    auto prewait2 = [&]() {
      const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
      wait_w2(wi1, wp1);
      prewaited = true;
    };
    auto adv2 = [&](uint32_t wi1, uint32_t wp1) { wi = wi1 + 1; wp = wp1; if (wi == kStkStages) { wi = 0; wp ^= 1; } };
    auto wait_dep = [&](bool first_of_read, uint32_t k) {
      // slot 1 starts op n only after slot 0 has issued all of its op n: the tensor pipe then runs slot 1's MMAs under slot 0's
      // epilogue (and vice versa) instead of both slots computing and then both draining, and the lag between the two
      // consumers of the shared weight ring stays within one op (bottleneck + conv = all 7 stages).
      // (polled with a short sleep: a tight shared-memory spin would take issue slots from the epilogue warps of this warp's scheduler)
      STK_PROF(0);
      if (s == 1) { uint32_t spins = 0; while (*issued <= gops) { __nanosleep(32); if (++spins > (1u << 24)) __trap(); } }
      ++gops;
      STK_PROF(2);
      mbar_wait(&sm->act_ready[s], opc & 1);
      ++opc;
      STK_PROF(3);
      if (first_of_read && p.in_mode == kStkInPlanes) mbar_wait(&sm->in_full[s], k & 1);
      tc_fence_after();
      STK_PROF(4);
    };
    auto mark_issued = [&]() {
      if (s == 0 && elect_one()) { __threadfence_block(); *issued = gops; }
      __syncwarp();
    };
    auto commit_acc = [&]() {
      if (elect_one()) umma_commit(&sm->acc_full[s]);
      __syncwarp();
    };
    auto stage_done = [&]() {
      if (elect_one()) umma_commit(&wempty[wi]);
      __syncwarp();
      if (++wi == kStkStages) { wi = 0; wp ^= 1; }
    };
    // bottleneck 1x1 of the layer whose output sits in the buffer, positions-as-M orientation: A = activation rows (two 128-row tiles),
    // B = weights; accumulator = the shared columns, free once the other slot's epilogue has drained its previous use
    auto issue_bott = [&]() {
      STK_PROF(0);
      if (lane == 0) { uint32_t spins = 0; while (atomicCAS(&sm->bott_busy, 0u, 1u) != 0u) { __nanosleep(32); if (++spins > (1u << 24)) __trap(); } }
      __syncwarp();
      tc_fence_after();
      STK_PROF(5);
      uint32_t xa_lo = x_lo | b_lbo_x;
      for (int blk = 0; blk < kKC / 2; blk += kBottPerStage) {
        wait_w();
        const uint32_t w_lo = (ring_lo + wi * (kStkStageBytes >> 4)) | b_lbo_bott;
        if (elect_one()) {
          uint32_t xa = xa_lo, wl = w_lo;
#pragma unroll
          for (int u = 0; u < kBottPerStage; ++u) {
            umma_bf16(d_bott0, stk_desc(xa, desc_hi), stk_desc(wl, desc_hi), idesc_bott, (blk + u) > 0);
            umma_bf16(d_bott1, stk_desc(xa + 128u, desc_hi), stk_desc(wl, desc_hi), idesc_bott, (blk + u) > 0);
            xa += kStep;
            wl += (uint32_t)kStkBott * 2u;
          }
          umma_commit(&wempty[wi]);
        }
        __syncwarp();
        xa_lo += (uint32_t)kBottPerStage * kStep;
        if (++wi == kStkStages) { wi = 0; wp ^= 1; }
      }
      if (elect_one()) umma_commit(&sm->bott_full[s]);
      __syncwarp();
    };
    for (int pr = 0; pr < n_pairs; ++pr) {
      STK_TRACE_GATE(pr);
      for (int l = 0; l < n_layers; ++l) {
        const StackLayer& L = p.layer[l];
        const int ksteps = L.kc_in / 2, total = L.conv_blocks;
        const uint32_t dil = (uint32_t)L.dil;
        const bool with_bott = highway && l > 0;
        // ---- [bottleneck of layer l-1 +] conv of layer l: D[cout][pos] = sum over taps and input-channel k-steps ----
        if (with_bott || (ksteps & 7) != 0) prewait1(); else prewait2();
        wait_dep(l == 0, (uint32_t)pr);
        if (with_bott) issue_bott();
        if ((ksteps & 7) == 0) {
          // two ring stages (8 k-steps) per iteration: both full-barrier probes are in flight together and the eight MMAs and the two
          // stage releases go out from one elected region. The issuer shares its scheduler with four epilogue warps whose unrolled ALU
          // runs keep the issue slot while they are eligible (scripts/probes/issue_probe3.cu: one warp with four independent FFMA chains
          // doubles the cycles per MMA of this loop, four stretch them tenfold): what counts is the number of issuer instructions — and
          // dependency stalls, each of which hands the slot away — per MMA, so the bookkeeping is amortised over as many MMAs as a tap has.
          uint32_t acc = 0;
          for (int tap = 0; tap < 3; ++tap) {
            uint32_t bd_lo = (x_lo - dil + (uint32_t)tap * dil) | b_lbo_x;
            for (int j = 0; j < ksteps; j += 8) {
              const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
              wait_w2(wi1, wp1);
              const uint32_t a0 = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w, a1 = (ring_lo + wi1 * (kStkStageBytes >> 4)) | a_lbo_w;
              if (elect_one()) {
                umma_bf16(d_main, stk_desc(a0, desc_hi), stk_desc(bd_lo, desc_hi), idesc_main, acc);
#pragma unroll
                for (uint32_t u = 1; u < 4; ++u) umma_bf16(d_main, stk_desc(a0 + u * 256u, desc_hi), stk_desc(bd_lo + u * kStep, desc_hi), idesc_main, 1);
                umma_commit(&wempty[wi]);
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) umma_bf16(d_main, stk_desc(a1 + u * 256u, desc_hi), stk_desc(bd_lo + (4 + u) * kStep, desc_hi), idesc_main, 1);
                umma_commit(&wempty[wi1]);
              }
              __syncwarp();
              acc = 1;
              bd_lo += 8 * kStep;
              adv2(wi1, wp1);
            }
          }
        } else {
          uint32_t bd_lo = x_lo - dil;
          int jj = 0;
          for (int blk = 0; blk < total; blk += 4) {
            wait_w();
            const uint32_t a_lo = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (blk + u < total) {
                if (elect_one()) umma_bf16(d_main, stk_desc(a_lo + u * 256u, desc_hi), stk_desc(bd_lo | b_lbo_x, desc_hi), idesc_main, (blk + u) > 0);
                __syncwarp();
                bd_lo += kStep;
                if (++jj == ksteps) { jj = 0; bd_lo += dil - (uint32_t)ksteps * kStep; }
              }
            }
            stage_done();
          }
        }
        commit_acc();
        mark_issued();
        // ---- residual 1x1: accumulates on x + b_res stored by the epilogue ----
        if (L.residual) {
          prewait2();
          wait_dep(false, 0);
          const uint32_t bd_lo = x_lo | b_lbo_x;
          static_assert(kKC / 2 == 8 && kStkStageBytes == 4 * 4096, "the residual 1x1 (8 k-steps) is two ring stages");
          {
            const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
            wait_w2(wi1, wp1);
            const uint32_t a0 = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w, a1 = (ring_lo + wi1 * (kStkStageBytes >> 4)) | a_lbo_w;
            if (elect_one()) {
#pragma unroll
              for (uint32_t u = 0; u < 4; ++u) umma_bf16(d_main, stk_desc(a0 + u * 256u, desc_hi), stk_desc(bd_lo + u * kStep, desc_hi), idesc_main, 1);
              umma_commit(&wempty[wi]);
#pragma unroll
              for (uint32_t u = 0; u < 4; ++u) umma_bf16(d_main, stk_desc(a1 + u * 256u, desc_hi), stk_desc(bd_lo + (4 + u) * kStep, desc_hi), idesc_main, 1);
              umma_commit(&wempty[wi1]);
            }
            __syncwarp();
            adv2(wi1, wp1);
          }
          commit_acc();
          mark_issued();
        }
      }
      // ---- bottleneck of the segment's last layer: an op of its own (the next conv belongs to the next read) ----
      if (highway) {
        prewait1();
        wait_dep(false, 0);
        issue_bott();
        mark_issued();
      }
    }
    STK_PROF(0);
#ifdef DAN_STK_PROF
    prof_acc[9] = clock64() - prof_begin;
#endif
    STK_PROF_FLUSH(lane == 0, 14 * s);
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kStkRegsEpi));
    // ===================== epilogue warps of slot s: quadrant q = TMEM lanes / channels 32q.., half h = position range; the slot's
    // 8 warps also move the slot's reads in (bulk loads or the encoder) and out (bulk stores / reductions) ===============================
    const int s = warp >> 3, h = (warp >> 2) & 1, q = warp & 3;
    const int gtid = threadIdx.x & (kStkEpiThreads - 1);
    const int wl = warp & 7;            // this warp's index within the slot's epilogue group
    const uint32_t tbase = tmem_base + (uint32_t)s * kStkAccStride + ((uint32_t)(32 * q) << 16);
    const uint32_t tbott = tmem_base + (uint32_t)(h ? kStkBottCol1 : kStkBottCol0) + ((uint32_t)(32 * q) << 16);
    uint8_t* const buf = bufs + (size_t)s * kStkBuf;
    const uint32_t buf_addr = smem_u32(buf);
    // stmatrix / ldmatrix row address of this thread for position group 0: matrix lane>>3 = chunk plane 4q + (lane>>3), row lane&7
    const uint32_t saddr0 = buf_addr + (uint32_t)(4 * q + (lane >> 3)) * kStkPlane + (uint32_t)(kStkLead + (lane & 7)) * 16;
    const int g_begin = h == 0 ? 0 : 14, g_end = h == 0 ? 14 : 26;

    // brings the read `it` of this slot into the buffer (or arranges for it): encode mode = synchronous (bytes fetched earlier);
    // planes mode = arm the barrier, lane 0 of every warp issues the bulk loads of two chunk planes (a bulk-copy instruction costs its
    // issuing thread ~170 cycles, so the 2 x kc_in of them are spread over the slot's 8 warps)
    auto prepare_read = [&](const StkIter& it, StkEncBytes eb, const StkEncRows& rows) {
      const bool valid = it.valid(p.R, s);
      if (p.in_mode == kStkInEncode) {
        if (valid) stk_enc_store(p, gtid, eb, rows, buf);
        fence_proxy_async_smem();
      } else {
        if (gtid == 0) mbar_expect_tx(&sm->in_full[s], valid ? plane_bytes * in_kc : 0u);
        named_bar_sync(1 + s, kStkEpiThreads);                      // the barrier is armed before any copy can complete
        if (lane == 0 && valid) {
          const uint4* src = p.in + kLead + ((long)it.cand * p.R + it.read_in_cand(s)) * p.pitch;
          uint8_t* dst = buf + kStkLead * 16;
          const uint64_t once = l2_policy_evict_first();          // activations stream through: read once, written once
          for (int kc = 2 * wl; kc < 2 * wl + 2 && kc < in_kc; ++kc)
            bulk_g2s_hint(dst + (size_t)kc * kStkPlane, src + kc * p.in_kstride, plane_bytes, &sm->in_full[s], once);
        }
      }
    };
    // encode mode: integer work + table requests of the read `it` (its bytes in eb); every thread of the slot takes part in the two named-barrier
    // reductions, so the call is made slot-uniformly. cand_bytes caches the candidate-level bytes of this thread's position.
    uint32_t cand_bytes = 0u; int cand_cached = -1;
    auto encode_rows = [&](const StkIter& it, StkEncBytes eb) -> StkEncRows {
      if (p.in_mode != kStkInEncode || it.done() || !it.valid(p.R, s)) return StkEncRows{};
      if (it.cand != cand_cached) { cand_bytes = stk_enc_fetch_cand(p, p.cand0 + it.cand, gtid); cand_cached = it.cand; }
      return stk_enc_rows(p, gtid, s, eb, cand_bytes);
    };
    auto fetch_read = [&](const StkIter& it) -> StkEncBytes {
      if (it.done() || !it.valid(p.R, s)) return 0u;
      if (p.in_mode == kStkInEncode) return stk_enc_fetch(p, p.cand0 + it.cand, it.read_in_cand(s), gtid);
      if (lane == 0) {                                                                 // L2 prefetch: the load itself is then an L2 hit
        const uint4* src = p.in + kLead + ((long)it.cand * p.R + it.read_in_cand(s)) * p.pitch;
        for (int kc = 2 * wl; kc < 2 * wl + 2 && kc < in_kc; ++kc) bulk_prefetch_l2(src + kc * p.in_kstride, plane_bytes);
      }
      return 0u;
    };
    // bottleneck epilogue of layer lb: tile h, TMEM lane = position 128h + 32q + lane, columns = bottleneck channels; relu(. + bias) -> T
    uint32_t bfc = 0;
    STK_PROF_DECL;
    STK_TRACE_ROLE(gtid == 0, 2 + s);
#ifdef DAN_STK_PROF
    int prof_rd = 0;
#endif
    auto bott_epilogue = [&](int lb, int read_global, bool valid) {
      STK_PROF(0);
      mbar_wait(&sm->bott_full[s], bfc & 1);
      ++bfc;
      tc_fence_after();
      STK_PROF(1);
      uint32_t r[32];
      tmem_ld32(tbott, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && atomicAdd(&sm->bott_drained, 1u) == kStkEpiThreads / 32 - 1) {   // read out by all 8 warps: the next bottleneck MMAs (either slot) may overwrite it
        sm->bott_drained = 0;
        __threadfence_block();
        atomicExch(&sm->bott_busy, 0u);
      }
      const int pos = 128 * h + 32 * q + lane;
      if (valid && pos < p.P) {
        const float* bb = &sm->bbias[lb][0];
        uint4* dst = p.layer[lb].tout + ((long)read_global * (kStkBott / 8)) * p.P + pos;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;                                                                 // relu(bottleneck), model.py:774
          o.x = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 0]) + bb[g * 8 + 0], 0.f), fmaxf(__uint_as_float(r[g * 8 + 1]) + bb[g * 8 + 1], 0.f));
          o.y = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 2]) + bb[g * 8 + 2], 0.f), fmaxf(__uint_as_float(r[g * 8 + 3]) + bb[g * 8 + 3], 0.f));
          o.z = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 4]) + bb[g * 8 + 4], 0.f), fmaxf(__uint_as_float(r[g * 8 + 5]) + bb[g * 8 + 5], 0.f));
          o.w = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 6]) + bb[g * 8 + 6], 0.f), fmaxf(__uint_as_float(r[g * 8 + 7]) + bb[g * 8 + 7], 0.f));
          dst[(long)g * p.P] = o;                                                  // lanes = consecutive positions: 512 contiguous bytes per warp store
        }
      }
      STK_PROF(2);
    };

    StkIter it;
    it.init(p.cands, p.R);
    if (!it.done()) {
      const StkEncBytes eb = fetch_read(it);
      prepare_read(it, eb, encode_rows(it, eb));
    }
    mbar_arrive(&sm->act_ready[s]);       // initial credit: the issuer's first op waits for "phase 0"
    uint32_t opc = 0;
    while (!it.done()) {
#ifdef DAN_STK_PROF
      STK_TRACE_GATE(prof_rd); ++prof_rd;
#endif
      const bool valid = it.valid(p.R, s);
      const int cand = it.cand;
      const int read_global = cand * p.R + it.read_in_cand(s);
      StkEncBytes eb_next = 0u;
      for (int l = 0; l < n_layers; ++l) {
        const StackLayer& L = p.layer[l];
        const bool with_bmap = l == 0 && p.bmap != nullptr;     // the conv bias arrives inside the per-candidate pool bias map
        const bool last = l + 1 == n_layers;
        EpiConsts k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 32 * q + 8 * j + (lane >> 2);
          // y = scale * max(z, -b) + (scale * b + shift); with the pool bias map the conv bias is inside the map: max(z, -0) and shift alone
          k.scale[j] = __ldg(L.chan + kC + c); k.rbias[j] = __ldg(L.chan + 3 * kC + c);
          k.nb[j] = with_bmap ? -0.f : __ldg(L.chan + 4 * kC + c);
          k.c[j] = __ldg(L.chan + (with_bmap ? 2 : 5) * kC + c);
        }
        // pool bias map of this read's candidate, this thread's fragments: the first four chunks are requested before the
        // accumulator is awaited (their L2 latency hides under the conv MMAs), the rest as the chunks are consumed
        uint4 bpre[4][2];
        const uint4* bm = nullptr;
        if (with_bmap) {
          bm = p.bmap + ((long)cand * 8 + (h * 4 + q)) * (7 * kStkBmapChunk) + lane * 2;
#pragma unroll
          for (int c = 0; c < 4; ++c) { bpre[c][0] = __ldg(bm + c * kStkBmapChunk); bpre[c][1] = __ldg(bm + c * kStkBmapChunk + 1); }
        }
        if (highway && l > 0) bott_epilogue(l - 1, read_global, valid);      // runs under this layer's conv MMAs
        if (last) { StkIter nxt = it; nxt.next(p.R); eb_next = fetch_read(nxt); }     // next read of this slot: bytes / L2 prefetch requested early
        STK_PROF(0);
        mbar_wait(&sm->acc_full[s], opc & 1);
        ++opc;
        tc_fence_after();
        STK_PROF(3);
        if (with_bmap) stack_epi_bmap(tbase, saddr0, lane, p.P, g_begin, (g_end - g_begin) / 2, k, bm, bpre);
        else if (L.residual) stack_epi_main<kEpiPreRes>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
        else stack_epi_main<kEpiFinal>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
        fence_proxy_async_smem();
        tc_fence_before();
        // without a bottleneck op behind it, the read's last hand-over doubles as "next read is in place" and is made at the boundary
        if (L.residual || !last || highway) mbar_arrive(&sm->act_ready[s]);
        STK_PROF(4);
        if (L.residual) {
          mbar_wait(&sm->acc_full[s], opc & 1);
          ++opc;
          tc_fence_after();
          STK_PROF(3);
          stack_epi_main<kEpiPostRes>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
          fence_proxy_async_smem();
          tc_fence_before();
          if (!last || highway) mbar_arrive(&sm->act_ready[s]);
          STK_PROF(5);
        }
      }
      // ---- the read's segment output is final: hand its planes to the TMA engine (store / read-axis reductions), under the last
      // bottleneck MMA / epilogue. Two chunk planes per warp (bulk groups are per thread: each issuer commits and waits for its own).
      named_bar_sync(1 + s, kStkEpiThreads);
      // encode mode: the next read's integer work and table requests go out now, ahead of the output operations — their latency passes
      // while the engine reads the buffer out, and only the six stores per position are left for the moment the buffer is free
      StkIter nxt = it;
      nxt.next(p.R);
      const StkEncRows rows_next = encode_rows(nxt, eb_next);
      STK_PROF(10);
      if (lane == 0 && valid) {
        bulk_wait0();                                                     // earlier reductions into the same accumulators have landed (long ago)
        STK_PROF(11);
        const uint8_t* src = buf + kStkLead * 16;
        const bool first_of_group = it.pr == 0;
        const uint64_t once = l2_policy_evict_first();
        for (int kc = 2 * wl; kc < 2 * wl + 2; ++kc) {
          const uint8_t* sp = src + (size_t)kc * kStkPlane;
          if (p.out) bulk_s2g_hint(p.out + kLead + (long)read_global * p.pitch + kc * p.out_kstride, sp, plane_bytes, once);
          if (p.maxv) bulk_reduce_max_bf16(p.maxv + cand * p.max_stride + (long)kc * p.P, sp, plane_bytes);
          if (p.sums) {
            uint4* dst = p.sums + (((long)cand * p.groups_per_cand + it.group(s)) * kKC + kc) * p.P;
            if (first_of_group) bulk_s2g(dst, sp, plane_bytes); else bulk_reduce_add_bf16(dst, sp, plane_bytes);
          }
        }
        bulk_commit();
      }
      STK_PROF(6);
      if (highway) bott_epilogue(n_layers - 1, read_global, valid);
      // ---- read boundary: refill the slot as soon as the TMA engine has read the buffer ----
      STK_PROF(0);
      if (lane == 0) bulk_wait_read0();
      named_bar_sync(1 + s, kStkEpiThreads);
      STK_PROF(7);
      it = nxt;
      if (!it.done()) prepare_read(it, eb_next, rows_next);
      STK_PROF(8);
      mbar_arrive(&sm->act_ready[s]);
      STK_PROF(12);
    }
#ifdef DAN_STK_PROF
    prof_acc[9] = clock64() - prof_begin;
#endif
    STK_PROF_FLUSH(gtid == 0, 28 + 14 * s);
    if (lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 18) tmem_dealloc<512>(tmem_base);
}

}  // namespace
