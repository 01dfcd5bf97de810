// dan_stack.cuh — fused, persistent conv-stack kernel (bf16 tcgen05). Included by dan_bf16.cu.
//
// One launch runs a SEGMENT of consecutive conv layers (a maximal run without a read-axis pool-add in between:
// PROD = layers 1-2, then layers 3-7; dl4vc/model.py:728-778) for every read of a pass. A read (201 positions x 128
// channels, bf16) is loaded once into shared memory, goes through all layers of the segment IN PLACE and is written
// back once; per-read activations between layers never touch HBM.
//
// Orientation ("D^T"): for one read and one layer the tensor core computes
//       D[cout 0..127][position 0..207] = sum_tap  W_tap[cout][cin] * X[position + (tap-1)*dil][cin]
// i.e. M = 128 output channels (A operand = weights, streamed from L2 through a shared-memory ring), N = 208 positions
// (B operand = the read's activation buffer; a dilated tap is a row offset of the descriptor start address, the zero
// rows around the read implement Conv2d's zero padding, model.py:214-229), accumulator = 208 TMEM columns.
// 201 -> 208 padding costs 3.4 % (a positions-as-M tiling would cost 21 %). The epilogue reads the accumulator in the
// mma C-fragment layout (tcgen05.ld 16x256b), applies +bias -> ReLU -> BatchNorm with per-thread channel constants
// (model.py:749-751) and writes bf16 back into the activation buffer with stmatrix.trans, which performs the
// [channel][position] -> [position][channel] transpose for free.
// Residual layers (model.py:753-761): the same epilogue pre-loads the accumulator with x + b_res (tcgen05.st), a
// second MMA pass accumulates W_res * y on top, and a second epilogue writes the layer output.
// Highway bottleneck (model.py:773-774): positions-as-M orientation (D3[position][32]) so that N = 32 is legal; its
// epilogue writes relu(.)+bias to the T matrix consumed by the compression GEMM.
//
// Per CTA: two independent read pipelines ("slots": own accumulator, own epilogue warpgroup, own weight ring) that
// the single MMA-issuing thread multiplexes at weight-stage granularity, so one slot's epilogue runs under the other
// slot's MMAs; three activation buffers rotate so that the next read's load and the previous read's store overlap
// with compute.
#pragma once

namespace {

constexpr int kStkThreads = 640;       // warps 0-7 epilogue slot 0 | 8-15 epilogue slot 1 | 16,17 weight producers | 18,19 MMA issuers
constexpr int kStkEpiThreads = 256;    // epilogue threads per slot: 4 TMEM lane quadrants x 2 position halves
constexpr int kStkLead = 2;            // zero rows in front of position 0 (>= largest dilation)
constexpr int kStkN = 208;             // MMA N = positions per read, padded to a multiple of 16
constexpr int kStkRB = kStkLead + kStkN + 2;   // rows per channel-chunk plane of a read buffer (212)
constexpr int kStkPlane = kStkRB * 16;         // bytes per plane (3392)
constexpr int kStkBuf = kKC * kStkPlane;       // bytes per read buffer (54272)
constexpr int kStkStageBytes = 8192;   // weight-ring stage = two k-step blocks of 4 KB
constexpr int kStkStages = 14;         // stages of the weight ring shared by the two slots (>= largest op (12) + prefetch)
constexpr int kStkMaxSeg = 8;          // layers per segment
constexpr int kStkRegsIssue = 56, kStkRegsEpi = 104;   // setmaxnreg redistributes the launch allocation (640 x 96): 128 x 56 + 512 x 104 = 60416 <= 61440
constexpr int kStkBmapChunk = 64;       // uint4 per (role, 16-position chunk) of the pool bias map: 32 lanes x 2 (bmap_pack_kernel, dan_bf16.cu)
constexpr int kStkBmapPerCand = 8 * 7 * kStkBmapChunk;   // uint4 per candidate: 8 roles (position half, lane quadrant) x 7 chunks
constexpr int kStkSmemHeader = 3072;   // barriers, TMEM pointer, bottleneck biases of the segment
constexpr size_t kStkSmemBytes = kStkSmemHeader + 2 * (size_t)kStkBuf + kStkStages * kStkStageBytes;

struct StackLayer {
  const uint8_t* wstream;   // conv k-step blocks (tap-major) | residual blocks | bottleneck blocks, contiguous
  size_t wreplica_stride;   // byte distance between the kWeightReplicas copies of wstream
  const float* chan;        // [4][128]: conv bias, BN scale, BN shift, residual bias
  const float* bbias;       // [bott]
  uint4* tout;              // T[read][c/8][p][c%8] (bf16) of this layer: row-major A operand (K = (c/8, p, c%8)) of the compression GEMM
  int conv_blocks;          // 3 * kc_in / 2
  int kc_in;                // 16-byte pieces per input row (CinPad/8 for layer 1, else 16)
  int dil, residual, highway;
};

struct StackParams {
  const uint4* in; long in_kstride;     // chunk-major input rows (dan_bf16.cu), in_kc planes
  uint4* out; long out_kstride;         // chunk-major output rows, 16 planes
  long t_reads_stride;
  int num_reads, P, pitch, bott, num_layers;
  const float* pool; int reads_per_cand;   // optional read-mean of the previous segment, fp32 [candidate][c/8][p][8]: added to every read on load (model.py:742)
  const uint4* bmap;                   // optional pool bias map conv(pool) + bias of the segment's first layer (replaces `pool`), [candidate][kStkBmapPerCand]
  unsigned long long* prof;            // optional [grid][16] cycle counters (development aid), or null
  uint2* trace; int trace_cap;         // development: event trace of CTA 0 (id, clock), trace[0].x = count
  int debug;                           // development: bit 0 = skip the MMAs, bit 1 = skip epilogue math/stores
  StackLayer layer[kStkMaxSeg];
};

struct StackSmem {
  uint64_t w_full[kStkStages], w_empty[kStkStages];
  uint64_t acc_full[2], act_ready[2], in_full[2];
  uint32_t tmem_base;
  uint32_t issued_ops;                 // ops fully issued by slot 0's issuer (slot 1 runs one op behind, see the issuer)
  float bbias[kStkMaxSeg][64];
};
static_assert(sizeof(StackSmem) <= kStkSmemHeader, "header too small");

}  // namespace
#include "dan_stack_epi.cuh"
namespace {

// descriptor words: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14  (tcgen05_ptx.cuh make_smem_desc)
__device__ __forceinline__ uint64_t stk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// role = 0,1 issuer of slot 0,1; 2,3 epilogue of slot 0,1: each role logs into its own quarter of the buffer (no atomics)
__device__ __forceinline__ void stk_trace(const StackParams& p, int role, int& n, uint32_t id) {
  if (p.trace && blockIdx.x == 0) {
    const int cap = p.trace_cap / 4;
    if (n < cap) p.trace[role * cap + n++] = make_uint2(id, (uint32_t)clock64());
  }
}

// kDev: 0 = production, 1 = honours the debug skip flags only, 2 = + cycle counters, 3 = + event trace instead (development builds)
template <int kDev>
__global__ void __launch_bounds__(kStkThreads, 1) dan_stack_kernel(const __grid_constant__ StackParams p) {
  const bool prof_on = kDev == 2 && p.prof != nullptr, trace_on = kDev == 3 && p.trace != nullptr;
  extern __shared__ __align__(1024) uint8_t smem[];
  StackSmem* sm = reinterpret_cast<StackSmem*>(smem);
  uint8_t* bufs = smem + kStkSmemHeader;
  uint8_t* rings = bufs + 2 * (size_t)kStkBuf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int per = p.num_reads / (int)gridDim.x, rem = p.num_reads % (int)gridDim.x;
  const int r_begin = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
  const int n_reads = per + ((int)blockIdx.x < rem ? 1 : 0);
  const uint32_t plane_bytes_in = (uint32_t)p.P * 16;
  const int in_kc = p.layer[0].kc_in;

  {  // zero rows / planes must read as 0 until an epilogue or a load writes them
    uint4* z = reinterpret_cast<uint4*>(bufs);
    for (int i = threadIdx.x; i < 2 * kStkBuf / 16; i += kStkThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStkStages; ++i) { mbar_init(&sm->w_full[i], 1); mbar_init(&sm->w_empty[i], 2); }   // both slots release a stage
    sm->issued_ops = 0;
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm->acc_full[s], 1); mbar_init(&sm->act_ready[s], kStkEpiThreads); mbar_init(&sm->in_full[s], 1);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.num_layers * 64; i += kStkThreads) {
    const int l = i >> 6, c = i & 63;
    sm->bbias[l][c] = (p.layer[l].highway && c < p.bott) ? p.layer[l].bbias[c] : 0.f;
  }
  fence_proxy_async_smem();
  if (warp == 18) tmem_alloc<512>(&sm->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;

  // global -> shared load of local read i into its slot's buffer, spread over the slot's 8 epilogue warps: lane 0 of warp wl moves
  // chunk planes 2*wl and 2*wl + 1 and prefetches the same planes of the slot's next read into L2. A bulk-copy instruction costs its
  // issuing thread ~170 cycles (and divergent lanes of one warp are serialised), so one thread issuing 2 x in_kc of them kept the
  // slot waiting 5-6 k cycles at every read boundary. The barrier must be armed (arm_read) before any of the copies can complete.
  auto arm_read = [&](int i) { mbar_expect_tx(&sm->in_full[i & 1], plane_bytes_in * in_kc); };
  auto load_read_planes = [&](int i, int wl) {
    const int s = i & 1;
    const uint4* src = p.in + kLead + (long)(r_begin + i) * p.pitch;
    uint8_t* dst = bufs + (size_t)s * kStkBuf + kStkLead * 16;
    const uint64_t once = l2_policy_evict_first();          // activations stream through: read once, written once
    for (int kc = 2 * wl; kc < 2 * wl + 2 && kc < in_kc; ++kc)
      bulk_g2s_hint(dst + (size_t)kc * kStkPlane, src + kc * p.in_kstride, plane_bytes_in, &sm->in_full[s], once);
  };
  auto prefetch_read_planes = [&](int i, int wl) {           // read i into L2, so that its load (on the slot's critical path) is an L2 hit
    const uint4* src = p.in + kLead + (long)(r_begin + i) * p.pitch;
    for (int kc = 2 * wl; kc < 2 * wl + 2 && kc < in_kc; ++kc) bulk_prefetch_l2(src + kc * p.in_kstride, plane_bytes_in);
  };

  // register budget: the producer / issuer warpgroup (warps 16-19) hands registers to the four epilogue warpgroups
  if (warp >= 16) {
#ifndef DAN_STK_NO_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kStkRegsIssue));
#endif
  if (warp == 16 || warp == 17) {
    // ===================== weight producer: ONE stream for both slots. The two reads of a pair go through the same op
    // sequence one op apart (see the issuers), so every stage is consumed twice before it is refilled: the L2 -> SMEM
    // weight traffic (the binding resource of this kernel when each slot streamed its own copy) is halved. =========
    // Two producer threads (lane 0 of warps 16 and 17) take alternate stages of the stream: one thread needs ~300 cycles per
    // stage (empty-barrier probe + expect_tx + copy issue), which is close to the rate at which the tensor pipe drains a stage.
    if (lane == 0 && !(kDev != 0 && (p.debug & 16))) {
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
      for (int i = 0; i < n_reads; i += 2) {
        for (int l = 0; l < p.num_layers; ++l) {
          const StackLayer& L = p.layer[l];
          const uint32_t conv_bytes = (uint32_t)L.conv_blocks * 4096u;
          const uint8_t* w = L.wstream + (size_t)(blockIdx.x % kWeightReplicas) * L.wreplica_stride;
          emit(w, conv_bytes);
          if (L.residual) emit(w + conv_bytes, kKC / 2 * 4096u);
          if (L.highway) emit(w + conv_bytes + (L.residual ? kKC / 2 * 4096u : 0u), (uint32_t)kKC * p.bott * 16u);
        }
      }
    }
  } else {
    // ===================== MMA issuer of slot s. The whole warp runs the (blocking, strictly sequential) op schedule of
    // its slot — warp-uniform control flow keeps descriptors in uniform registers — and one elected lane issues the
    // tcgen05 instructions. The two slots' issuers are independent warps: the tensor pipe interleaves their MMA
    // streams, so one slot's epilogue runs under the other slot's MMAs without any software multiplexing. A single
    // warp retires one dependent instruction every ~4 cycles, so the loops below are kept to a few instructions
    // per MMA (descriptor words are advanced by constants). =========================================================
    const int s = warp - 18;
    const uint32_t idesc_main = make_idesc_bf16(128, kStkN);
    const uint32_t idesc_bott = make_idesc_bf16(128, p.bott);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo_w = (2048u >> 4) << 16, b_lbo_x = ((uint32_t)kStkPlane >> 4) << 16, b_lbo_bott = (((uint32_t)p.bott * 16u) >> 4) << 16;
    const int bott_per_stage = kStkStageBytes / (p.bott * 32);
    const uint32_t ring_lo = smem_u32(rings) >> 4;
    uint64_t* const wfull = &sm->w_full[0];
    uint64_t* const wempty = &sm->w_empty[0];
    volatile uint32_t* const issued = &sm->issued_ops;
    uint32_t gops = 0;                                                                            // ops started by this issuer
    int tr_n = 0;
    const uint32_t d_main = tmem_base + (uint32_t)s * 256u;
    const uint32_t x_lo = (smem_u32(bufs) >> 4) + (uint32_t)s * (kStkBuf >> 4) + kStkLead;       // centre row of chunk plane 0
    constexpr uint32_t kStep = 2 * (kStkPlane >> 4);                                              // one k-step = two chunk planes
    uint32_t wi = 0, wp = 0, opc = 0;
    const long long t_begin = clock64();
    long long t_dep = 0, t_wfull = 0;
    const bool do_mma = kDev == 0 || !(p.debug & 1);
    const bool no_w = kDev != 0 && (p.debug & 16);     // development: weights are not streamed (garbage operands, timing only)
    // The full-barrier probe of an op's first stage(s) is issued BEFORE the op's dependencies are awaited (prewait*): a probe costs
    // ~190 cycles even when the phase is complete, and there it would sit between the epilogue's hand-over and the op's first MMA.
    bool prewaited = false;
    auto wait_w = [&]() {
      if (prewaited) { prewaited = false; return; }
      if (no_w) return;
      if (prof_on) { const long long c0 = clock64(); mbar_wait(&wfull[wi], wp); t_wfull += clock64() - c0; }
      else mbar_wait(&wfull[wi], wp);
      tc_fence_after();
    };
    auto wait_w2 = [&](uint32_t wi1, uint32_t wp1) {
      if (prewaited) { prewaited = false; return; }
      if (no_w) return;
      if (prof_on) { const long long c0 = clock64(); mbar_wait2(&wfull[wi], wp, &wfull[wi1], wp1); t_wfull += clock64() - c0; }
      else mbar_wait2(&wfull[wi], wp, &wfull[wi1], wp1);
      tc_fence_after();
    };
    auto prewait1 = [&]() { wait_w(); prewaited = true; };
    auto prewait2 = [&]() {
      const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
      wait_w2(wi1, wp1);
      prewaited = true;
    };
    auto adv2 = [&](uint32_t wi1, uint32_t wp1) { wi = wi1 + 1; wp = wp1; if (wi == kStkStages) { wi = 0; wp ^= 1; } };
    auto wait_dep = [&](bool first_of_read, int k) {
      long long c0 = 0; if (prof_on) c0 = clock64();
      // slot 1 starts op n only after slot 0 has issued all of its op n: the tensor pipe then runs slot 1's MMAs under slot 0's
      // epilogue (and vice versa) instead of both slots computing and then both draining, and the lag between the two
      // consumers of the shared weight ring stays within one op (<= 12 of the 14 stages).
      // (polled with a short sleep: a tight shared-memory spin would take issue slots from the epilogue warps of this warp's scheduler)
      if (s == 1 && !(p.debug & 8)) { uint32_t spins = 0; while (*issued <= gops) { __nanosleep(32); if (++spins > (1u << 24)) __trap(); } }
      ++gops;
      mbar_wait(&sm->act_ready[s], opc & 1);
      if (first_of_read) mbar_wait(&sm->in_full[s], (uint32_t)k & 1);
      tc_fence_after();
      if (prof_on) t_dep += clock64() - c0;
      if (trace_on && lane == 0) stk_trace(p, s, tr_n, (uint32_t)s << 28 | 1u << 24 | (gops & 0xFFFFu));
    };
    auto op_done = [&]() {
      if (elect_one()) {
        umma_commit(&sm->acc_full[s]);
        if (s == 0) { __threadfence_block(); *issued = gops; }
      }
      __syncwarp();
      if (trace_on && lane == 0) stk_trace(p, s, tr_n, (uint32_t)s << 28 | 2u << 24 | (gops & 0xFFFFu));
      ++opc;
    };
    auto stage_done = [&]() {
      if (!no_w && elect_one()) umma_commit(&wempty[wi]);
      __syncwarp();
      if (++wi == kStkStages) { wi = 0; wp ^= 1; }
    };
    for (int i = s, k = 0; i < n_reads + (n_reads & 1); i += 2, ++k) {
      if (i >= n_reads) {
        // odd tail: slot 1 has no read in the last pair but must still release the stages streamed for slot 0's read
        for (int l = 0; l < p.num_layers; ++l) {
          const StackLayer& L = p.layer[l];
          const int stages = (L.conv_blocks + 1) / 2 + (L.residual ? kKC / 4 : 0) + (L.highway ? p.bott / 32 : 0);
          for (int st = 0; st < stages && !no_w; ++st) {
            mbar_wait(&wfull[wi], wp);
            if (elect_one()) mbar_arrive(&wempty[wi]);
            __syncwarp();
            if (++wi == kStkStages) { wi = 0; wp ^= 1; }
          }
        }
        break;
      }
      for (int l = 0; l < p.num_layers; ++l) {
        const StackLayer& L = p.layer[l];
        const int residual = L.residual, highway = L.highway, ksteps = L.kc_in / 2, total = L.conv_blocks;
        const uint32_t dil = (uint32_t)L.dil;
        // ---- conv: D[cout][pos] = sum over taps and input-channel k-steps ----
        if ((ksteps & 3) == 0) prewait2(); else prewait1();
        wait_dep(l == 0, k);
        if ((ksteps & 3) == 0) {
          // two ring stages (4 k-steps) per iteration: both full-barrier probes are in flight together and the four MMAs and
          // the two stage releases go out from one elected region — a single warp retires a dependent instruction only every
          // few cycles, so per-stage bookkeeping (not the tensor pipe) bounds the issue rate when it is done stage by stage
          uint32_t acc = 0;
          for (int tap = 0; tap < 3; ++tap) {
            uint32_t bd_lo = (x_lo - dil + (uint32_t)tap * dil) | b_lbo_x;
            for (int j = 0; j < ksteps; j += 4) {
              const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
              wait_w2(wi1, wp1);
              const uint32_t a0 = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w, a1 = (ring_lo + wi1 * (kStkStageBytes >> 4)) | a_lbo_w;
              if (elect_one()) {
                if (do_mma) {
                  umma_bf16(d_main, stk_desc(a0, desc_hi), stk_desc(bd_lo, desc_hi), idesc_main, acc);
                  umma_bf16(d_main, stk_desc(a0 + 256u, desc_hi), stk_desc(bd_lo + kStep, desc_hi), idesc_main, 1);
                  umma_bf16(d_main, stk_desc(a1, desc_hi), stk_desc(bd_lo + 2 * kStep, desc_hi), idesc_main, 1);
                  umma_bf16(d_main, stk_desc(a1 + 256u, desc_hi), stk_desc(bd_lo + 3 * kStep, desc_hi), idesc_main, 1);
                }
                if (!no_w) { umma_commit(&wempty[wi]); umma_commit(&wempty[wi1]); }
              }
              __syncwarp();
              acc = 1;
              bd_lo += 4 * kStep;
              adv2(wi1, wp1);
            }
          }
        } else {
          uint32_t bd_lo = x_lo - dil;
          int jj = 0;
          for (int blk = 0; blk < total; blk += 2) {
            wait_w();
            const uint32_t a_lo = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (blk + u < total) {
                if (do_mma && elect_one()) umma_bf16(d_main, stk_desc(a_lo + u * 256u, desc_hi), stk_desc(bd_lo | b_lbo_x, desc_hi), idesc_main, (blk + u) > 0);
                __syncwarp();
                bd_lo += kStep;
                if (++jj == ksteps) { jj = 0; bd_lo += dil - (uint32_t)ksteps * kStep; }
              }
            }
            stage_done();
          }
        }
        op_done();
        // ---- residual 1x1: accumulates on x + b_res stored by the epilogue ----
        if (residual) {
          prewait2();
          wait_dep(false, 0);
          uint32_t bd_lo = x_lo | b_lbo_x;
          for (int blk = 0; blk < kKC / 2; blk += 4) {
            const uint32_t wi1 = wi + 1 == kStkStages ? 0u : wi + 1, wp1 = wi1 == 0 ? wp ^ 1u : wp;
            wait_w2(wi1, wp1);
            const uint32_t a0 = (ring_lo + wi * (kStkStageBytes >> 4)) | a_lbo_w, a1 = (ring_lo + wi1 * (kStkStageBytes >> 4)) | a_lbo_w;
            if (elect_one()) {
              if (do_mma) {
                umma_bf16(d_main, stk_desc(a0, desc_hi), stk_desc(bd_lo, desc_hi), idesc_main, 1);
                umma_bf16(d_main, stk_desc(a0 + 256u, desc_hi), stk_desc(bd_lo + kStep, desc_hi), idesc_main, 1);
                umma_bf16(d_main, stk_desc(a1, desc_hi), stk_desc(bd_lo + 2 * kStep, desc_hi), idesc_main, 1);
                umma_bf16(d_main, stk_desc(a1 + 256u, desc_hi), stk_desc(bd_lo + 3 * kStep, desc_hi), idesc_main, 1);
              }
              if (!no_w) { umma_commit(&wempty[wi]); umma_commit(&wempty[wi1]); }
            }
            __syncwarp();
            bd_lo += 4 * kStep;
            adv2(wi1, wp1);
          }
          op_done();
        }
        // ---- bottleneck 1x1, positions-as-M orientation: A = activation rows (two 128-row tiles), B = weights ----
        if (highway) {
          prewait1();
          wait_dep(false, 0);
          uint32_t xa_lo = x_lo | b_lbo_x;
          for (int blk = 0; blk < kKC / 2; blk += bott_per_stage) {
            wait_w();
            const uint32_t w_lo = (ring_lo + wi * (kStkStageBytes >> 4)) | b_lbo_bott;
            if (elect_one()) {
              uint32_t xa = xa_lo, wl = w_lo;
              for (int u = 0; u < bott_per_stage; ++u) {
                if (do_mma) {
                  umma_bf16(d_main, stk_desc(xa, desc_hi), stk_desc(wl, desc_hi), idesc_bott, (blk + u) > 0);
                  umma_bf16(d_main + (uint32_t)p.bott, stk_desc(xa + 128u, desc_hi), stk_desc(wl, desc_hi), idesc_bott, (blk + u) > 0);
                }
                xa += kStep;
                wl += (uint32_t)p.bott * 2u;
              }
              if (!no_w) umma_commit(&wempty[wi]);
            }
            __syncwarp();
            xa_lo += (uint32_t)bott_per_stage * kStep;
            if (++wi == kStkStages) { wi = 0; wp ^= 1; }
          }
          op_done();
        }
      }
    }
    if (prof_on && lane == 0) { p.prof[blockIdx.x * 16 + 0 + 10 * s] = clock64() - t_begin; p.prof[blockIdx.x * 16 + 1 + 10 * s] = t_dep; p.prof[blockIdx.x * 16 + 12 + s] = t_wfull; }
  }
  } else {
#ifndef DAN_STK_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kStkRegsEpi));
#endif
    // ===================== epilogue warps of slot s: quadrant q = TMEM lanes / channels 32q.., half h = position range;
    // thread 0 of the slot's group also moves the slot's reads in and out ==========================================
    const int s = warp >> 3, h = (warp >> 2) & 1, q = warp & 3;
    const int gtid = threadIdx.x & (kStkEpiThreads - 1);
    const uint32_t tbase = tmem_base + (uint32_t)s * 256u + ((uint32_t)(32 * q) << 16);
    const uint32_t buf_addr = smem_u32(bufs + (size_t)s * kStkBuf);
    // stmatrix / ldmatrix row address of this thread for position group 0: matrix lane>>3 = chunk plane 4q + (lane>>3), row lane&7
    const uint32_t saddr0 = buf_addr + (uint32_t)(4 * q + (lane >> 3)) * kStkPlane + (uint32_t)(kStkLead + (lane & 7)) * 16;
    const int g_begin = h == 0 ? 0 : 14, g_end = h == 0 ? 14 : 26;
    // x += pool (model.py:734-742: the read-axis mean of the previous layer's output is added to the input of this segment's first
    // layer). Runs on the freshly loaded read before the slot is handed to the issuer; bf16(x + pool) like the stand-alone kernel.
    int tr_n = 0;
    auto add_pool = [&](int read_local, uint32_t parity) {
      mbar_wait(&sm->in_full[s], parity);
      if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 7u << 24);     // the read has landed
      const long cand = (long)(r_begin + read_local) / p.reads_per_cand;
      const float* pl = p.pool + cand * kKC * p.P * 8;
      uint8_t* b = bufs + (size_t)s * kStkBuf + kStkLead * 16;
      for (int pos = gtid; pos < p.P; pos += kStkEpiThreads) {
#pragma unroll 4
        for (int kc = 0; kc < kKC; ++kc) {
          uint4* px = reinterpret_cast<uint4*>(b + (size_t)kc * kStkPlane) + pos;
          const float4* pa = reinterpret_cast<const float4*>(pl + ((long)kc * p.P + pos) * 8);
          const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1);
          uint4 v = *px;
          v.x = pack_bf16x2(bf16_lo(v.x) + a0.x, bf16_hi(v.x) + a0.y); v.y = pack_bf16x2(bf16_lo(v.y) + a0.z, bf16_hi(v.y) + a0.w);
          v.z = pack_bf16x2(bf16_lo(v.z) + a1.x, bf16_hi(v.z) + a1.y); v.w = pack_bf16x2(bf16_lo(v.w) + a1.z, bf16_hi(v.w) + a1.w);
          *px = v;
        }
      }
      fence_proxy_async_smem();
    };
    const int wl = warp & 7;            // this warp's index within the slot's epilogue group
    if (s < n_reads) {
      if (gtid == 0) arm_read(s);
      named_bar_sync(1 + s, kStkEpiThreads);
      if (lane == 0) { load_read_planes(s, wl); if (s + 2 < n_reads) prefetch_read_planes(s + 2, wl); }
    }
    if (p.pool && s < n_reads) add_pool(s, 0);
    mbar_arrive(&sm->act_ready[s]);       // initial credit: the issuer's first op waits for "phase 0"
    uint32_t opc = 0, eops = 0;
    long long t_wait = 0, t_main = 0, t_bott = 0, t_io = 0, t0 = 0;
    const bool prof = prof_on && gtid == 0;
    for (int i = s; i < n_reads; i += 2) {
      for (int l = 0; l < p.num_layers; ++l) {
        const StackLayer& L = p.layer[l];
        const bool with_bmap = l == 0 && p.bmap != nullptr;     // the conv bias arrives inside the per-candidate pool bias map
        const bool last = l + 1 == p.num_layers;
        // without a bottleneck the layer's main epilogue is the read's last op: when the next read still gets the pool table added in
        // shared memory, the hand-over to the issuer has to wait for that (the issuer only waits for act_ready and the load)
        const bool defer_ready = last && !L.highway && p.pool != nullptr;
        EpiConsts k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 32 * q + 8 * j + (lane >> 2);
          const float b = with_bmap ? 0.f : __ldg(L.chan + c);
          k.scale[j] = __ldg(L.chan + kC + c); k.rbias[j] = __ldg(L.chan + 3 * kC + c);
          k.c[j] = fmaf(k.scale[j], b, __ldg(L.chan + 2 * kC + c)); k.nb[j] = -b;
        }
        // pool bias map of this read's candidate, this thread's fragments: the first four chunks are requested before the
        // accumulator is awaited (their L2 latency hides under the conv MMAs), the rest as the chunks are consumed
        uint4 bpre[4][2];
        const uint4* bm = nullptr;
        if (with_bmap) {
          const long cand = (long)(r_begin + i) / p.reads_per_cand;
          bm = p.bmap + (cand * 8 + (h * 4 + q)) * (7 * kStkBmapChunk) + lane * 2;
#pragma unroll
          for (int c = 0; c < 4; ++c) { bpre[c][0] = __ldg(bm + c * kStkBmapChunk); bpre[c][1] = __ldg(bm + c * kStkBmapChunk + 1); }
        }
        if (prof) t0 = clock64();
        mbar_wait(&sm->acc_full[s], opc & 1);
        tc_fence_after();
        if (prof) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
        if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 3u << 24 | (eops & 0xFFFFu));
        const bool do_epi = kDev == 0 || !(p.debug & 2);
        if (!do_epi) {}
        else if (with_bmap) stack_epi_bmap(tbase, saddr0, lane, p.P, g_begin, (g_end - g_begin) / 2, k, bm, bpre);
        else if (L.residual) stack_epi_main<kEpiPreRes>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
        else stack_epi_main<kEpiFinal>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
        fence_proxy_async_smem();
        tc_fence_before();
        if (L.residual || !defer_ready) mbar_arrive(&sm->act_ready[s]);
        ++opc;
        if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 4u << 24 | (eops++ & 0xFFFFu));
        if (prof) { const long long t1 = clock64(); t_main += t1 - t0; t0 = t1; }
        if (L.residual) {
          mbar_wait(&sm->acc_full[s], opc & 1);
          tc_fence_after();
          if (prof) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 3u << 24 | (eops & 0xFFFFu));
          if (do_epi) stack_epi_main<kEpiPostRes>(tbase, saddr0, lane, p.P, g_begin, g_end, k);
          fence_proxy_async_smem();
          tc_fence_before();
          if (!defer_ready) mbar_arrive(&sm->act_ready[s]);
          ++opc;
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 4u << 24 | (eops++ & 0xFFFFu));
          if (prof) { const long long t1 = clock64(); t_main += t1 - t0; t0 = t1; }
        }
        if (last) {
          // the segment output of this read is final: start writing it back now, under this layer's bottleneck MMA / epilogue
          named_bar_sync(1 + s, kStkEpiThreads);
          if (lane == 0) {           // two chunk planes per warp (bulk groups are per thread: each issuer commits and later waits for its own)
            uint4* dst = p.out + kLead + (long)(r_begin + i) * p.pitch;
            const uint8_t* src = bufs + (size_t)s * kStkBuf + kStkLead * 16;
            const uint64_t once = l2_policy_evict_first();
            for (int kc = 2 * wl; kc < 2 * wl + 2; ++kc) bulk_s2g_hint(dst + kc * p.out_kstride, src + (size_t)kc * kStkPlane, plane_bytes_in, once);
            bulk_commit();
          }
        }
        if (L.highway) {
          mbar_wait(&sm->acc_full[s], opc & 1);
          tc_fence_after();
          if (prof) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 3u << 24 | (eops & 0xFFFFu));
          // bottleneck tile h: TMEM lane = position 128h + 32q + lane, columns = bottleneck channels
          const int c8n = p.bott / 8;
          const int pos = 128 * h + 32 * q + lane;
          bool handed_over = false;
          for (int cc = 0; do_epi && cc < p.bott / 32; ++cc) {
            uint32_t r[32];
            tmem_ld32(tbase + (uint32_t)(h * p.bott + cc * 32), r);
            tmem_ld_wait();
            if (!last && cc + 1 == p.bott / 32) {
              // the accumulator has been read out: the next layer's conv MMAs (which overwrite these columns) may start while this
              // thread still converts and stores its T rows
              tc_fence_before();
              mbar_arrive(&sm->act_ready[s]);
              handed_over = true;
            }
            if (pos < p.P) {
              const float* bb = &sm->bbias[l][cc * 32];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 o;                                                                 // relu(bottleneck), model.py:774
                o.x = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 0]) + bb[g * 8 + 0], 0.f), fmaxf(__uint_as_float(r[g * 8 + 1]) + bb[g * 8 + 1], 0.f));
                o.y = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 2]) + bb[g * 8 + 2], 0.f), fmaxf(__uint_as_float(r[g * 8 + 3]) + bb[g * 8 + 3], 0.f));
                o.z = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 4]) + bb[g * 8 + 4], 0.f), fmaxf(__uint_as_float(r[g * 8 + 5]) + bb[g * 8 + 5], 0.f));
                o.w = pack_bf16x2(fmaxf(__uint_as_float(r[g * 8 + 6]) + bb[g * 8 + 6], 0.f), fmaxf(__uint_as_float(r[g * 8 + 7]) + bb[g * 8 + 7], 0.f));
                if (kDev == 0 || !(p.debug & 4)) L.tout[((long)(r_begin + i) * c8n + cc * 4 + g) * p.P + pos] = o;     // lanes = consecutive positions: 512 contiguous bytes per warp store
              }
            }
          }
          tc_fence_before();
          if (!last && !handed_over) mbar_arrive(&sm->act_ready[s]);
          ++opc;
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 4u << 24 | (eops++ & 0xFFFFu));
          if (prof) { const long long t1 = clock64(); t_bott += t1 - t0; t0 = t1; }
        }
        if (last) {
          // every MMA of this read has completed (the last accumulator was awaited above) and the write-back has been issued:
          // refill the slot as soon as the store has read the buffer
          if (lane == 0) bulk_wait_read0();                       // this warp's planes have been read out of the buffer
          if (gtid == 0 && i + 2 < n_reads) arm_read(i + 2);
          named_bar_sync(1 + s, kStkEpiThreads);                  // all planes drained, barrier armed
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 6u << 24 | (eops & 0xFFFFu));     // the store has drained the buffer
          if (lane == 0 && i + 2 < n_reads) load_read_planes(i + 2, wl);
          if (p.pool && i + 2 < n_reads) add_pool(i + 2, (uint32_t)((i + 2) >> 1) & 1u);
          if (L.highway || defer_ready) mbar_arrive(&sm->act_ready[s]);
          if (lane == 0 && i + 4 < n_reads) prefetch_read_planes(i + 4, wl);      // off the critical path: after the hand-over
          if (trace_on && gtid == 0) stk_trace(p, 2 + s, tr_n, (uint32_t)s << 28 | 5u << 24 | (eops & 0xFFFFu));
          if (prof) { const long long t1 = clock64(); t_io += t1 - t0; t0 = t1; }
        }
      }
    }
    if (lane == 0) bulk_wait0();
    if (prof) { unsigned long long* d = p.prof + blockIdx.x * 16 + 2 + 3 * s; d[0] = t_wait; d[1] = t_main; d[2] = t_bott; p.prof[blockIdx.x * 16 + 8 + s] = t_io; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 18) tmem_dealloc<512>(tmem_base);
}

}  // namespace
