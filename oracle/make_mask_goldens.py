"""ORACLE helper (build container only, TEST INFRASTRUCTURE): golden vectors for the proposal-mask decode of the loader,
`get_read_mask_vectors` (reference dl4vc/dataset.py:112-250, with `simple_variant_encoding_vectors` :86-109), produced by running the
REAL reference function on seeded synthetic VCF records / reference windows. Writes tests/golden/mask_vectors.npz.

    python -m oracle.make_mask_goldens
"""
from __future__ import annotations

import os

import numpy as np

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mask_vectors.npz")
BASES = "ATGC"
CODE = {"A": 1, "T": 2, "G": 3, "C": 4}


def window(rng):
    """A 201-column reference window: bases with a few '-' (5) insert-gap columns, never at column 100 unless asked."""
    w = rng.integers(1, 5, size=201).astype(np.uint8)
    return w


def cases(rng, n):
    out = []
    for i in range(n):
        w = window(rng)
        kind = rng.choice(["snp", "del", "ins", "del_gap", "ins_long", "rewind", "bad_ref", "mnp", "lower", "snp_n"],
                          p=[0.2, 0.2, 0.2, 0.12, 0.05, 0.08, 0.05, 0.04, 0.03, 0.03])
        off = 100
        if kind == "rewind":                       # column 100 (and maybe 99) is an insert gap of another allele: rewind to the base
            g = int(rng.integers(1, 3))
            w[100 - g + 1:101] = 5
            off = 100 - g
            kind2 = rng.choice(["snp", "del", "ins"])
        else:
            kind2 = kind
        inv = {1: "A", 2: "T", 3: "G", 4: "C"}
        if kind2 in ("snp", "snp_n", "lower"):
            x = inv[int(w[off])]
            y = BASES[int(rng.integers(0, 4))]
            if kind2 == "snp_n":
                y = "N"
            if kind2 == "lower":
                x, y = x.lower(), y.lower()
        elif kind2 in ("del", "del_gap"):
            L = int(rng.integers(2, 14))
            if kind2 == "del_gap":                 # gap columns inside the deleted stretch (spurious insert of another allele)
                for _ in range(int(rng.integers(1, 4))):
                    w[off + int(rng.integers(1, L + 2))] = 5
            seq, j = [], off
            while len(seq) < L and j < 201:
                if w[j] != 5:
                    seq.append(inv[int(w[j])])
                j += 1
            x = "".join(seq)
            y = x[0]
        elif kind2 in ("ins", "ins_long"):
            L = int(rng.integers(2, 12)) if kind2 == "ins" else int(rng.integers(48, 70))
            x = inv[int(w[off])]
            y = x + "".join(BASES[int(k)] for k in rng.integers(0, 4, size=L - 1))
        elif kind2 == "bad_ref":                   # the record's reference base is not what the window holds
            x = inv[int(w[off]) % 4 + 1]
            y = BASES[int(rng.integers(0, 4))]
        else:                                      # mnp: equal lengths > 1 — no branch of the reference handles it
            x = "".join(inv[int(w[off + k])] for k in range(2))
            y = "".join(BASES[int(k)] for k in rng.integers(0, 4, size=2))
        rec = "\t".join(["chr1", "12345", ".", x, y, ".", ".", "."])
        out.append((rec, x, y, w))
    return out


def main():
    ds = ref_shim.import_reference_module("dl4vc.dataset")
    rng = np.random.default_rng(20261018)
    recs, xs, ys, wins, rms, vms, oks = [], [], [], [], [], [], []
    for rec, x, y, w in cases(rng, 600):
        try:
            rm, vm = ds.get_read_mask_vectors(rec, w.copy())
            ok = 1
        except Exception:                          # AssertionError / UnboundLocalError / KeyError / ValueError in the reference
            rm = vm = np.zeros(201, np.uint8)
            ok = 0
        recs.append(rec); xs.append(x); ys.append(y); wins.append(w)
        rms.append(np.asarray(rm, np.uint8)); vms.append(np.asarray(vm, np.uint8)); oks.append(ok)
    np.savez_compressed(OUT, ref_alleles=np.array(xs), var_alleles=np.array(ys), references=np.stack(wins),
                        ref_masks=np.stack(rms), var_masks=np.stack(vms), ok=np.array(oks, np.uint8))
    print("wrote", OUT, "cases", len(oks), "ok", int(sum(oks)), "reference raised", len(oks) - int(sum(oks)))


if __name__ == "__main__":
    main()
