"""ORACLE tooling — the per-record thresholding of the REAL tools/format_vcf.py (filter_format_vcf, :52-215) run on a synthetic scored VCF
with one record per position (so that its multi-allele merge never triggers); the kept records, genotype strings and quality buckets go to
tests/golden/genotype_calls.npz.  TEST INFRASTRUCTURE ONLY.     python oracle/make_call_goldens.py"""
from __future__ import annotations

import argparse
import contextlib
import importlib.util
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# two threshold sets: the defaults of the script (indel / long-indel / delete fall back to the SNP values) and the README's tuned call
THRESHOLD_SETS = [
    dict(snp_threshold=0.3, indel_threshold=0.0, long_indel_threshold=0.0, delete_threshold=0.0, snp_zygo_threshold=0.5, indel_zygo_threshold=0.5,
         long_indel_zygo_threshold=0.5, delete_zygo_threshold=0.5),
    dict(snp_threshold=0.5, indel_threshold=0.1, long_indel_threshold=0.25, delete_threshold=0.15, snp_zygo_threshold=0.4, indel_zygo_threshold=0.45,
         long_indel_zygo_threshold=0.6, delete_zygo_threshold=0.35),
    dict(snp_threshold=0.42, indel_threshold=0.2, long_indel_threshold=0.0, delete_threshold=0.0, snp_zygo_threshold=0.55, indel_zygo_threshold=0.3,
         long_indel_zygo_threshold=0.9, delete_zygo_threshold=0.9),
]


def main():
    spec = importlib.util.spec_from_file_location("ref_format_vcf", "/root/reference/tools/format_vcf.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    n = 3000
    # scores as dan_scores writes them: [1 - p_bin0, P(no variant), P(het), P(hom)] fp32; many of them close to the thresholds in use
    vt = rng.dirichlet([0.6, 0.6, 0.6], n).astype(np.float32)
    near = rng.choice([0.3, 0.5, 0.1, 0.25, 0.15, 0.42, 0.2], n) + rng.normal(0, 2e-3, n)
    sel = rng.random(n) < 0.4
    vt[sel, 0] = np.clip(1.0 - near[sel], 0, 1).astype(np.float32)
    vt[sel, 2] = np.clip(rng.choice([0.5, 0.4, 0.45, 0.6, 0.35, 0.55, 0.3, 0.9], sel.sum()) + rng.normal(0, 2e-3, sel.sum()), 0, 1).astype(np.float32)
    vt[sel, 1] = np.clip(1.0 - vt[sel, 0] - vt[sel, 2], 0, 1)
    vt[:5] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.7, 0.3, 0.0], [0.5, 0.0, 0.5]]
    vt[-1] = [0.0, 0.5, 0.5]      # the script only flushes its last buffered position when the LAST record passes (format_vcf.py:210): make it pass
    scores = np.concatenate([rng.random((n, 1)).astype(np.float32), vt], axis=1).astype(np.float32)
    ref_len = rng.choice([1, 1, 1, 2, 3, 4, 7], n).astype(np.int32)
    var_len = rng.choice([1, 1, 1, 2, 3, 5], n).astype(np.int32)
    ref_len[-1] = var_len[-1] = 1
    bases = "ACGT"
    out = {"scores": scores, "ref_len": ref_len, "var_len": var_len, "num_sets": np.int32(len(THRESHOLD_SETS))}
    with tempfile.TemporaryDirectory() as tmp:
        # With the script's defaults (indel_threshold <= 0) `delete_threshold` is never assigned (format_vcf.py:57-80) and the first short
        # delete raises UnboundLocalError: threshold set 0 therefore runs on the records that are not short deletes (`defined0`).
        short_delete = (ref_len > 1) & (var_len == 1) & ~((ref_len >= 3) | (var_len >= 3))
        out["defined0"] = ~short_delete
        for k in range(len(THRESHOLD_SETS)):
          fin = os.path.join(tmp, "in%d.vcf" % k)
          with open(fin, "w") as f:
            f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")
            for i in range(n):
                if k == 0 and short_delete[i]:
                    continue
                info = "BP=%.8f;NV=%.8f;HV=%.8f;OV=%.8f" % tuple(float(x) for x in scores[i])           # utils.py:171-176
                ref = "".join(bases[(i + k) % 4] for k in range(ref_len[i])); alt = "".join(bases[(i + 2 * k + 1) % 4] for k in range(var_len[i]))
                f.write("chr1\t%d\t%s\t%s\t%s\t.\t.\t.\tGT\t0/1\n" % (1000 + 3 * i, info, ref, alt))
        for k, th in enumerate(THRESHOLD_SETS):
            fin = os.path.join(tmp, "in%d.vcf" % k)
            fout = os.path.join(tmp, "out%d.vcf" % k)
            args = argparse.Namespace(input_file=fin, output_file=fout, multiallele_second_threshold=0.7, multiallele_homozygous_second_threshold=0.9,
                                      debug=False, **th)
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                mod.filter_format_vcf(args)
            gt = np.zeros(n, np.int8); q = np.full(n, -1, np.int32)
            for line in open(fout):
                if line[0] == "#":
                    continue
                it = line.rstrip("\n").split("\t")
                i = (int(it[1]) - 1000) // 3
                g, qs = it[9].split(":")
                gt[i] = {"0/1": 1, "1/1": 2}[g]; q[i] = int(qs)
            out["gt%d" % k] = gt; out["q%d" % k] = q
            out["thr%d" % k] = np.array([th["snp_threshold"], th["snp_zygo_threshold"], th["indel_threshold"], th["indel_zygo_threshold"],
                                         th["long_indel_threshold"], th["long_indel_zygo_threshold"], th["delete_threshold"], th["delete_zygo_threshold"]], np.float64)
            print("set", k, "kept", int((gt > 0).sum()), "of", n, "hom", int((gt == 2).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "genotype_calls.npz"), **out)


if __name__ == "__main__":
    main()
