// dan_feeder.cu — batched decode of pileup records into the model-facing uint8 tensors. Host code.
//
// Replaces the per-item Python path of the reference loader, ContextDatasetFromNumpy._get_generator (dl4vc/dataset.py:500-680) with its
// helpers sample_single_reads (:256-287), parse_vcf (dl4vc/utils.py:19-72), count_variants_from_single_reads (:340-362) and
// get_read_mask_vectors (:112-250), for a whole batch of record indices at once, writing straight into (pinned) host buffers in the
// layout dan_forward_host takes: [candidate][position][read]. Records are the compound type tools/convert_bam_single_reads.py writes
// (:694-698) stored back to back as raw bytes (np.memmap of that dtype; h5py / libhdf5 are not in this image).
#include <cstdlib>
#include <string>
#include "dan_records.h"

namespace {

// field offsets of the packed record (numpy compound dtype without alignment), --max-reads 200 --window-size 100
constexpr int kRecP = DAN_MASK_READ_LEN;      // 201 columns
constexpr int kRecRows = 200;                 // stored read rows (TOTAL_SINGLE_READS)
constexpr size_t kOffName = 0, kOffRef5 = 16, kOffReads5 = kOffRef5 + 5 * kRecP, kOffSingle = kOffReads5 + 2 * 5 * kRecP,
                 kOffRefBases = kOffSingle + (size_t)kRecRows * kRecP, kOffNumReads = kOffRefBases + kRecP, kOffLabel = kOffNumReads + 4,
                 kOffVcf = kOffLabel + 1, kOffQ = kOffVcf + 128, kOffStrand = kOffQ + (size_t)kRecRows * kRecP, kRecBytes = kOffStrand + (size_t)kRecRows * kRecP;
static_assert(kRecBytes == DAN_RECORD_BYTES, "record layout");

inline uint64_t splitmix(uint64_t& x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct VcfInfo { int is_snp, var_mode, ref_base, var_base, var_type, coverage; double allele_freq; };

// dl4vc/utils.py:19-72. Returns a DAN_REC_E_* status; every failure is one the reference raises on (KeyError / ValueError / IndexError).
int parse_vcf_record(const std::string& rec_in, VcfInfo* out, std::string* ref_allele, std::string* var_allele) {
  // str.strip() + split('\t')
  size_t b = 0, e = rec_in.size();
  while (b < e && isspace((unsigned char)rec_in[b])) ++b;
  while (e > b && isspace((unsigned char)rec_in[e - 1])) --e;
  std::vector<std::string> f;
  for (size_t i = b, start = b; i <= e; ++i)
    if (i == e || rec_in[i] == '\t') { f.emplace_back(rec_in.substr(start, i - start)); start = i + 1; }
  if (f.size() < 8) return DAN_REC_E_VCF_FIELDS;
  const std::string &x = f[3], &y = f[4];
  *ref_allele = x; *var_allele = y;
  VcfInfo r{};
  if (x.size() == 1 && y.size() == 1 && mask_real_base(x[0]) && mask_real_base(y[0])) {
    r.is_snp = 1; r.var_mode = 1;
    r.ref_base = mask_base_code(x[0]); r.var_base = mask_base_code(y[0]);
  } else if (x.size() > y.size()) {
    r.var_mode = 3;                                   // delete: GAAA -> G, reference base 'G', variant base '-'
    r.ref_base = mask_base_code(x[0]); r.var_base = 5;
  } else if (x.size() < y.size()) {
    if (x.empty()) return DAN_REC_E_ALLELE;           // ref_bases[0] raises IndexError
    r.var_mode = 2;                                   // insert: G -> GAAAA, variant base 'noinsert'
    r.ref_base = mask_base_code(x[0]); r.var_base = 8;
  } else {
    return DAN_REC_E_UNKNOWN_MUTATION;                // "Unknown mutation detected": no var_mode -> KeyError in the loader (dataset.py:596)
  }
  if (r.ref_base < 0 || r.var_base < 0) return DAN_REC_E_ALLELE;
  // INFO column: {n: v for (n, v) in [k.split('=') for k in rec[7].split(';')]}
  bool have_af = false, have_dp = false;
  const std::string& info = f[7];
  for (size_t i = 0, start = 0; i <= info.size(); ++i) {
    if (i != info.size() && info[i] != ';') continue;
    const std::string kv = info.substr(start, i - start);
    start = i + 1;
    const size_t eq = kv.find('=');
    if (eq == std::string::npos || kv.find('=', eq + 1) != std::string::npos) return DAN_REC_E_VCF_INFO;      // unpack of != 2 parts raises ValueError
    const std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
    if (k == "AF") {
      char* end = nullptr;
      r.allele_freq = strtod(v.c_str(), &end);
      if (end == v.c_str() || *end) return DAN_REC_E_VCF_INFO;
      have_af = true;
    } else if (k == "DP") {
      char* end = nullptr;
      const long dp = strtol(v.c_str(), &end, 10);
      if (end == v.c_str() || *end) return DAN_REC_E_VCF_INFO;
      r.coverage = (int)dp; have_dp = true;
    }
  }
  if (!have_af || !have_dp) return DAN_REC_E_VCF_INFO;
  r.var_type = 0;
  if (f.size() > 10) {
    const std::string& g = f[10];
    const size_t c = g.find(':');
    if (c == std::string::npos || g.find(':', c + 1) != std::string::npos) return DAN_REC_E_VCF_FIELDS;             // gt, var = rec[10].split(':')
    const std::string gt = g.substr(0, c), var = g.substr(c + 1);
    if (gt == "GT" && var.size() == 3 && (var[1] == '/' || var[1] == '|')) {
      if (var[0] == '1' && var[2] == '1') r.var_type = 2;
      else if ((var[0] == '0' && var[2] == '1') || (var[0] == '1' && var[2] == '0')) r.var_type = 1;
    }
  }
  *out = r;
  return DAN_REC_OK;
}

}  // namespace

extern "C" {

int dan_decode_records(const void* records, size_t record_stride, const int64_t* indices, int n, const dan_feeder_config* cfg, const dan_record_batch* out) {
  if (n < 0) { dan_set_error("negative record count"); return DAN_E_INVALID; }
  if (n == 0) return DAN_OK;
  if (!records || !indices || !cfg || !out || !out->reads || !out->ref || !out->ref_masks || !out->var_masks) { dan_set_error("null pointer"); return DAN_E_INVALID; }
  if (record_stride < DAN_RECORD_BYTES) { dan_set_error("record stride %zu below the %d-byte record", record_stride, DAN_RECORD_BYTES); return DAN_E_INVALID; }
  const int R = cfg->max_reads, S = cfg->store_max_reads < kRecRows ? cfg->store_max_reads : kRecRows;
  if (R < 1 || R > kRecRows || S < 1) { dan_set_error("max_reads %d / store_max_reads %d out of range", R, cfg->store_max_reads); return DAN_E_INVALID; }
  const size_t tile = (size_t)kRecP * R;
  int failed = 0;
  for (int i = 0; i < n; ++i) {
    const uint8_t* rec = static_cast<const uint8_t*>(records) + (size_t)indices[i] * record_stride;
    int32_t num_reads;
    memcpy(&num_reads, rec + kOffNumReads, 4);
    // A. window of stored rows (dataset.py:517-521), B. at most R of them (sample_single_reads, dataset.py:256-287)
    const int middle = (num_reads > S ? num_reads : S) / 2;
    int start = (int)((double)middle - (double)S / 2.0);
    if (start < 0) start = 0;
    const int avail = start < kRecRows ? (kRecRows - start < S ? kRecRows - start : S) : 0;      // columns of the transposed window
    const int max_reads = R < avail ? R : avail;
    int perm[kRecRows];
    if (max_reads >= num_reads) {
      for (int j = 0; j < max_reads; ++j) perm[j] = j;                                              // the first max_reads columns, deterministic
    } else {
      // a sorted random subset of the first min(avail, num_reads) columns: the reference draws it with an unseeded np.random.choice
      // (its inference is not reproducible for deep pileups); here the draw is a function of (seed, record index)
      const int total = avail < num_reads ? avail : num_reads;
      uint64_t st = cfg->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(indices[i] + 1));
      int pool[kRecRows];
      for (int j = 0; j < total; ++j) pool[j] = j;
      for (int j = 0; j < max_reads; ++j) { const int k = j + (int)(splitmix(st) % (uint64_t)(total - j)); const int t = pool[j]; pool[j] = pool[k]; pool[k] = t; }
      bool take[kRecRows] = {};
      for (int j = 0; j < max_reads; ++j) take[pool[j]] = true;
      for (int j = 0, k = 0; j < total; ++j) if (take[j]) perm[k++] = j;
    }
    uint8_t* reads = out->reads + (size_t)i * tile;
    uint8_t* q = out->q_scores ? out->q_scores + (size_t)i * tile : nullptr;
    uint8_t* sd = out->strands ? out->strands + (size_t)i * tile : nullptr;
    memset(reads, 0, tile);
    if (q) memset(q, 0, tile);
    if (sd) memset(sd, 0, tile);
    for (int j = 0; j < max_reads; ++j) {
      const uint8_t* row = rec + kOffSingle + (size_t)(start + perm[j]) * kRecP;
      const uint8_t* qrow = rec + kOffQ + (size_t)(start + perm[j]) * kRecP;
      const uint8_t* srow = rec + kOffStrand + (size_t)(start + perm[j]) * kRecP;
      for (int p = 0; p < kRecP; ++p) {
        reads[(size_t)p * R + j] = row[p];
        if (q && cfg->use_q_scores) q[(size_t)p * R + j] = qrow[p];
        if (sd && cfg->use_strands) sd[(size_t)p * R + j] = srow[p];
      }
    }
    uint8_t* ref = out->ref + (size_t)i * kRecP;
    memcpy(ref, rec + kOffRefBases, kRecP);
    if (out->label) out->label[i] = rec[kOffLabel];
    if (out->num_reads) out->num_reads[i] = num_reads;
    // vcfrec: S128, trailing NULs stripped like numpy does
    size_t len = 128;
    while (len > 0 && rec[kOffVcf + len - 1] == 0) --len;
    const std::string vcf(reinterpret_cast<const char*>(rec + kOffVcf), len);
    VcfInfo vi{};
    std::string xa, ya;
    int status = parse_vcf_record(vcf, &vi, &xa, &ya);
    uint8_t* rm = out->ref_masks + (size_t)i * kRecP;
    uint8_t* vm = out->var_masks + (size_t)i * kRecP;
    memset(rm, 0, kRecP); memset(vm, 0, kRecP);
    int blacklist = 0;
    if (status == DAN_REC_OK) {
      // coverage / allele frequency from the sampled reads at the proposal column (count_variants_from_single_reads, dataset.py:340-362)
      int ref_base; int col;
      if (vi.var_mode == 1) { ref_base = ref[100]; col = 100; }
      else if (vi.var_mode == 3) { ref_base = ref[101]; col = 101; }
      else { ref_base = 8; col = 101; }
      int agree = 0, disagree = 0;
      for (int j = 0; j < max_reads; ++j) {
        const int t = reads[(size_t)col * R + j];
        if (t == ref_base) ++agree;
        else if (t == 1 || t == 2 || t == 3 || t == 4 || t == 5 || t == 8 || t == 9) ++disagree;      // real_base_keys_set: A T C G - M noinsert
      }
      // columns beyond max_reads do not exist in the reference's slice; pad columns (token 0) count for neither side unless ref_base is 0
      if (ref_base == 0 && R > max_reads) { /* the reference's array has exactly max_reads columns */ }
      const int cover = agree + disagree;
      int coverage = vi.coverage; double af = vi.allele_freq;
      if (cover > 0) { coverage = cover; if (!cfg->keep_candidate_af) af = (double)disagree / (double)cover; }
      if (out->is_snp) out->is_snp[i] = (uint8_t)vi.is_snp;
      if (out->var_type) out->var_type[i] = vi.var_type;
      if (out->allele_freq) out->allele_freq[i] = (float)af;
      if (out->coverage) out->coverage[i] = coverage;
      if (out->var_base_enum) out->var_base_enum[i] = vi.var_base;
      if (out->var_ref_enum) out->var_ref_enum[i] = vi.ref_base;
      // proposal masks (get_read_mask_vectors); an AssertionError there blacklists the example and leaves all-pad masks (dataset.py:644-664)
      const int mcode = mask_one(xa.c_str(), ya.c_str(), ref, rm, vm);
      if (mcode != DAN_MASK_OK) {
        memset(rm, 0, kRecP); memset(vm, 0, kRecP);
        if (mcode == DAN_MASK_E_SHAPE || mcode == DAN_MASK_E_REF_MISMATCH) blacklist = 1;      // assert -> caught
        else status = DAN_REC_E_MASK;                                                            // KeyError / ValueError / UnboundLocalError propagate in the reference
      }
    }
    if (out->blacklist) out->blacklist[i] = (uint8_t)blacklist;
    if (out->status) out->status[i] = status;
    if (status != DAN_REC_OK) ++failed;
  }
  if (failed) { dan_set_error("%d of %d records raise in the reference loader (see status[])", failed, n); return DAN_E_INVALID; }
  return DAN_OK;
}

}  // extern "C"
