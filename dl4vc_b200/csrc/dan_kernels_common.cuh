// dan_kernels_common.cuh — device code shared by the fp32 and bf16 paths: the pileup encoder and the
// read-axis pooling kernels.
#pragma once
#include "dan_internal.h"

// ---------------------------------------------------------------------------------------------------------
// Pileup encoder: uint8 tiles -> conv-1 input rows.  Restates dl4vc/model.py:450-451 (embedding gather),
// :463-470,506-507 (positional add), :501-503,517 (reference broadcast + concat), :534-561 (q-score / strand
// scaling), :576-625 (ref/var agreement masks) and :719 (transpose) in one pass; nothing is materialised in the
// reference's (B,201,100,45) order. One CTA per candidate: the three 201x100 byte tiles are staged in shared
// memory with coalesced 32-bit loads, the per-read agreement bits are an AND-reduction over the <=51 masked
// columns, and rows are written out fully coalesced.
// ---------------------------------------------------------------------------------------------------------
struct EncodeParams {
  DevInputs in;
  const float* emb;   // (10, D)
  const float* pe;    // (P, D)
  int D, Cin, CinPad;
  int use_q, use_s, use_m;
  RowGeom g;
  const uint8_t* removed;   // optional [candidate][R]: reads replaced by the empty-read encoding (training augmentation, model.py:633-716)
};

struct EncodeSmem {
  uint8_t* reads; uint8_t* q; uint8_t* st; uint8_t* ref; uint8_t* rm; uint8_t* vm; uint8_t* agreeR; uint8_t* agreeV;
  float* emb;
  const uint8_t* removed;   // this candidate's row of EncodeParams::removed, or null
};

__host__ __device__ inline size_t encode_smem_bytes(int P, int R, int D) {
  size_t tile = (size_t)((P * R + 15) / 16) * 16;
  size_t vec = (size_t)((P + 15) / 16) * 16;
  size_t ag = (size_t)((R + 15) / 16) * 16;
  return 3 * tile + 3 * vec + 2 * ag + (size_t)DAN_VOCAB * D * sizeof(float);
}

__device__ inline void stage_bytes(uint8_t* dst, const uint8_t* src, int n, bool present) {
  if (!present) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = 0;
    return;
  }
  if (((reinterpret_cast<uintptr_t>(src) & 3) == 0) && (n % 4 == 0)) {
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
  }
}

__device__ inline EncodeSmem encode_stage(const EncodeParams& p, long cand, unsigned char* smem_raw) {
  const int P = p.g.P, R = p.g.R;
  size_t tile = (size_t)((P * R + 15) / 16) * 16, vec = (size_t)((P + 15) / 16) * 16, ag = (size_t)((R + 15) / 16) * 16;
  EncodeSmem s;
  s.reads = smem_raw; s.q = s.reads + tile; s.st = s.q + tile; s.ref = s.st + tile; s.rm = s.ref + vec; s.vm = s.rm + vec;
  s.agreeR = s.vm + vec; s.agreeV = s.agreeR + ag; s.emb = reinterpret_cast<float*>(s.agreeV + ag);
  stage_bytes(s.reads, p.in.reads + cand * P * R, P * R, true);
  stage_bytes(s.q, p.in.q ? p.in.q + cand * P * R : nullptr, P * R, p.use_q && p.in.q);
  stage_bytes(s.st, p.in.strands ? p.in.strands + cand * P * R : nullptr, P * R, p.use_s && p.in.strands);
  stage_bytes(s.ref, p.in.ref + cand * P, P, true);
  stage_bytes(s.rm, p.in.ref_masks ? p.in.ref_masks + cand * P : nullptr, P, p.use_m && p.in.ref_masks);
  stage_bytes(s.vm, p.in.var_masks ? p.in.var_masks + cand * P : nullptr, P, p.use_m && p.in.var_masks);
  for (int i = threadIdx.x; i < DAN_VOCAB * p.D; i += blockDim.x) s.emb[i] = p.emb[i];
  s.removed = p.removed ? p.removed + cand * R : nullptr;
  __syncthreads();
  // agreement of read r with the ref / var proposal: every masked column must carry exactly the mask token
  // (integer compare; model.py:592-593 and :607-608 — unmasked columns compare 0 == 0 and always agree)
  for (int t = threadIdx.x; t < 2 * R; t += blockDim.x) {
    const int kind = t / R, r = t - kind * R;
    const uint8_t* mask = kind ? s.vm : s.rm;
    bool ok = true;
    for (int pp = 0; pp < P; ++pp) {
      const uint8_t mv = mask[pp];
      ok = ok && (mv == 0 || s.reads[pp * R + r] == mv);
    }
    (kind ? s.agreeV : s.agreeR)[r] = ok ? 1 : 0;
  }
  __syncthreads();
  return s;
}

// value of input channel c at (position pp, read r); channel order = torch.cat order of the reference
__device__ inline float encode_channel(const EncodeParams& p, const EncodeSmem& s, int c, int pp, int r) {
  const int D = p.D, R = p.g.R;
  if (s.removed && s.removed[r]) {      // the row of an empty read: pad embedding + reference, every other channel zero (model.py:519-568)
    if (c < D) return s.emb[c] + __ldg(p.pe + pp * D + c);
    if (c < 2 * D) return s.emb[min((int)s.ref[pp], DAN_VOCAB - 1) * D + (c - D)] + __ldg(p.pe + pp * D + (c - D));
    return 0.f;
  }
  if (c < D) return s.emb[min((int)s.reads[pp * R + r], DAN_VOCAB - 1) * D + c] + __ldg(p.pe + pp * D + c);
  if (c < 2 * D) return s.emb[min((int)s.ref[pp], DAN_VOCAB - 1) * D + (c - D)] + __ldg(p.pe + pp * D + (c - D));
  c -= 2 * D;
  if (p.use_q) { if (c == 0) return (float)s.q[pp * R + r] * 0.01f; --c; }       // Q_SCORE_SCALE_FACTOR, model.py:24
  if (p.use_s) { if (c == 0) return (float)s.st[pp * R + r] * 0.5f; --c; }       // STRAND_ENCODE_FACTOR, model.py:16
  if (p.use_m) {
    if (c == 0) return (s.rm[pp] != 0 && s.agreeR[r]) ? 1.f : 0.f;
    if (c == 1) return (s.vm[pp] != 0 && s.agreeV[r]) ? 1.f : 0.f;
    if (c == 2) return (s.rm[pp] != 0) ? 1.f : 0.f;                              // var_length uses the REF mask (model.py:579,584)
  }
  return 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// Read-axis pooling (model.py:766-772 mean for the pool-add; :824-833 max ‖ mean before the FC).
// Sums run over the 100 read slots in slot order (empty slots included, like AvgPool2d((100,1))) and are
// divided by R at the end, so the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------------------------
