// dan_records.h — host-side record decoding shared by dan_capi.cu (dan_make_mask_vectors) and dan_feeder.cu (dan_decode_records):
// the loader's proposal-mask rule (dl4vc/dataset.py:86-250) and the token tables of dl4vc/base_enum.py.
#pragma once
#include <cstring>
#include <vector>
#include "dan_internal.h"

// ---- proposal-mask decode (dl4vc/dataset.py:86-250), host code --------------------------------------------------------------
namespace {
// base_enum of dl4vc/base_enum.py:7-11 as the dict literal evaluates (later duplicate keys win: 's' -> 9); -1 = KeyError
int mask_base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 1;
    case 'T': case 't': case 'U': case 'u': return 2;
    case 'G': case 'g': return 3;
    case 'C': case 'c': return 4;
    case '-': case '*': case 'N': case 'n': case 'X': case 'x': case '.': return 5;
    case 'e': return 7;
    case '?': case 'M': case 'm': case 'K': case 'k': case 'R': case 'r': case 'Y': case 'y': case 'S': case 's': case 'W': case 'w':
    case 'B': case 'b': case 'V': case 'v': case 'H': case 'h': case 'D': case 'd': return 9;
    default: return -1;
  }
}
// real_bases_set as dataset.py sees it (the second definition, base_enum.py:25: 'g' is missing there)
bool mask_real_base(char c) { return c == 'A' || c == 'a' || c == 'T' || c == 't' || c == 'C' || c == 'c' || c == 'G'; }

int mask_one(const char* x, const char* y, const uint8_t* ref, uint8_t* ref_mask, uint8_t* var_mask) {
  constexpr int kLen = DAN_MASK_READ_LEN, kVarEncodeLen = 51;            // dataset.py:85
  const int lx = (int)strlen(x), ly_full = (int)strlen(y), ly = ly_full < kVarEncodeLen ? ly_full : kVarEncodeLen;
  std::vector<int> rv(lx), vv(ly);                                        // simple_variant_encoding_vectors(delete_limit=0, keep_pad=False)
  for (int i = 0; i < lx; ++i) { rv[i] = mask_base_code(x[i]); if (rv[i] < 0) return DAN_MASK_E_ALLELE_CHAR; }
  for (int i = 0; i < ly; ++i) { vv[i] = mask_base_code(y[i]); if (vv[i] < 0) return DAN_MASK_E_ALLELE_CHAR; }
  const bool snp = lx == 1 && ly_full == 1 && mask_real_base(x[0]) && mask_real_base(y[0]);
  if (!snp) {
    if (lx > ly_full) { if (ly_full != 1) return DAN_MASK_E_SHAPE; }      // delete: AT -> A
    else if (ly_full > lx) { if (lx != 1) return DAN_MASK_E_SHAPE; }      // insert: A -> ATT
    else return DAN_MASK_E_UNSUPPORTED;
  }
  int off = 100;                                                          // rewind past '-' columns (inserts of other alleles)
  while (off >= 0 && ref[off] == 5) --off;
  if (off < 0) return DAN_MASK_E_WINDOW;
  if (lx == 0 || ref[off] != rv[0]) return DAN_MASK_E_REF_MISMATCH;
  if ((int)rv.size() > 1) {                                               // delete: the variant is the first base followed by explicit deletes
    vv.resize(rv.size(), 5);
    bool same = off + (int)rv.size() <= kLen;
    for (size_t i = 0; same && i < rv.size(); ++i) same = ref[off + i] == rv[i];
    if (!same) {                                                          // gap columns inside the deleted stretch: walk the window
      std::vector<int> nr, nv;
      size_t bi = 0;
      for (int ri = off; ri < kLen; ++ri) {
        if (bi >= rv.size()) break;
        if (ref[ri] == rv[bi]) { nr.push_back(ref[ri]); nv.push_back(vv[bi]); ++bi; }
        else if (ref[ri] == 5) { nr.push_back(5); nv.push_back(8); }
        else return DAN_MASK_E_REF_MISMATCH;
      }
      if (bi < rv.size()) return DAN_MASK_E_REF_MISMATCH;
      for (auto& v : nr) if (v == 5) v = 0;                               // "match any reference as long as it does not delete"
      for (auto& v : nv) if (v == 8) v = 0;
      rv.swap(nr); vv.swap(nv);
    }
  }
  if (rv.size() == 1 && vv.size() > 1) rv.resize(vv.size(), 8);           // insert: 'noinsert' in the reference mask
  if (rv.size() != vv.size()) return DAN_MASK_E_SHAPE;
  if (off + (int)rv.size() > kLen) return DAN_MASK_E_WINDOW;
  for (size_t i = 0; i < rv.size(); ++i) { ref_mask[off + i] = (uint8_t)rv[i]; var_mask[off + i] = (uint8_t)vv[i]; }
  return DAN_MASK_OK;
}
}  // namespace

