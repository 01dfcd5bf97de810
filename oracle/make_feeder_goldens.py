"""ORACLE tooling — the REAL reference loader (dl4vc/dataset.py ContextDatasetFromNumpy.__getitem__, test-set settings of main.py:86-87) run on
synthetic records of the on-disk schema; its per-item dicts go to tests/golden/feeder.npz.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_feeder_goldens.py

h5py is not installed: a stand-in module whose File('...')['data'] is the numpy structured array lets the unmodified generator code run.
Records deeper than 100 reads are sampled by the reference with np.random.choice; np.random is seeded so the golden is reproducible, and
the test checks those records for the sampling PROPERTIES (sorted subset, same rows for q-scores / strands), everything else exactly."""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dl4vc_b200.synth import make_pileups      # noqa: E402
from oracle import ref_shim                    # noqa: E402

RECORD_DTYPE = np.dtype([("name", "S16"), ("ref", np.uint8, (5, 201)), ("reads", np.uint16, (5, 201)), ("single_reads", np.uint8, (200, 201)),
                         ("ref_bases", np.uint8, (201,)), ("num_reads", np.int32), ("label", np.uint8), ("vcfrec", "S128"),
                         ("q-scores", np.uint8, (200, 201)), ("strand", np.uint8, (200, 201))])      # tools/convert_bam_single_reads.py:694-698
assert RECORD_DTYPE.itemsize == 123965
LETTER = {1: "A", 2: "T", 3: "G", 4: "C", 5: "-"}


def synth_records(n=48, seed=3):
    rng = np.random.default_rng(seed)
    b = make_pileups(n, seed=seed, num_reads=200, coverage="ragged", max_depth=200)
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    for i in range(n):
        depth = int(b.num_reads[i])
        rec[i]["name"] = f"chr{1 + i % 22}:{1000 + i}".encode()
        rec[i]["single_reads"] = b.reads[i].T
        rec[i]["q-scores"] = b.q_scores[i].T
        rec[i]["strand"] = b.strands[i].T
        rec[i]["ref_bases"] = b.ref[i]
        rec[i]["num_reads"] = depth
        rec[i]["label"] = i % 3
        rm, vm = b.ref_masks[i], b.var_masks[i]
        nz = np.flatnonzero(rm)
        if b.kind[i] == 0:
            x, y = LETTER[int(rm[100])], LETTER[int(vm[100])]
        elif b.kind[i] == 1:      # insert: ref_mask = [b, noinsert...], var_mask = [b, inserted...]
            x, y = LETTER[int(rm[100])], "".join(LETTER[int(v)] for v in vm[nz])
        else:                      # delete: ref_mask = deleted reference bases, var_mask = [b0, '-', ...]
            x, y = "".join(LETTER[int(v)] for v in rm[nz]), LETTER[int(vm[100])]
        gt = ["", "\tGT:0/1", "\tGT:1/1", "\tGT:1|0", "\tGT:0/0"][i % 5]
        rec[i]["vcfrec"] = f"chr{1 + i % 22}\t{1000 + i}\t.\t{x}\t{y}\t50\t.\tAF={rng.random():.4f};DP={int(rng.integers(5, 90))}\tGT\t0/1{gt}".encode()
    # corner cases: lower-case SNP, mismatching reference (assert -> blacklist), equal-length non-SNP, broken INFO, unknown allele letter
    rec[0]["vcfrec"] = rec[0]["vcfrec"].replace(b"\t.\t", b"\t.\t", 1)
    f = rec[1]["vcfrec"].split(b"\t"); f[3] = b"AC" if b.ref[1][100] != 1 else b"TC"; f[4] = f[3][:1]; rec[1]["vcfrec"] = b"\t".join(f)
    f = rec[2]["vcfrec"].split(b"\t"); f[3], f[4] = b"AT", b"GC"; rec[2]["vcfrec"] = b"\t".join(f)
    f = rec[3]["vcfrec"].split(b"\t"); f[7] = b"AF=0.5;DPX"; rec[3]["vcfrec"] = b"\t".join(f)
    f = rec[4]["vcfrec"].split(b"\t"); f[3], f[4] = b"Z", b"ZQ"; rec[4]["vcfrec"] = b"\t".join(f)
    return rec


def main():
    rec = synth_records()
    h5 = types.ModuleType("h5py")

    class File:
        def __init__(self, path, mode="r"): pass
        def __enter__(self): return {"data": rec}
        def __exit__(self, *a): return False
    h5.File = File
    sys.modules["h5py"] = h5
    ds_mod = ref_shim.import_reference_module("dl4vc.dataset")
    ds_mod.h5py = h5
    args = types.SimpleNamespace(model_use_q_scores=True, model_use_strands=True, aux_keep_candidate_af=False)
    ds = ds_mod.ContextDatasetFromNumpy("synthetic.hdf", args, augment_single_reads=False, augment_refernce=False)      # main.py:86-87
    np.random.seed(11)
    keys = ("reads", "q-scores", "strands", "ref", "ref_mask", "var_mask")
    out = {k: [] for k in keys}
    scal = {k: [] for k in ("label", "num_reads", "is_snp", "var_type", "allele_freq", "coverage", "var_base_enum", "var_ref_enum", "blacklist")}
    raised = []
    for i in range(len(rec)):
        try:
            ds._h5_gen = None                       # a generator that raised cannot be resumed
            item = ds[i]
        except Exception as e:
            raised.append(type(e).__name__)
            for k in keys: out[k].append(np.zeros((201, 100) if k in ("reads", "q-scores", "strands") else (201,), np.uint8))
            for k in scal: scal[k].append(0)
            continue
        raised.append("")
        for k in keys: out[k].append(np.asarray(item[k], np.uint8))
        for k in scal: scal[k].append(item[k])
    path = os.path.join(ROOT, "tests", "golden", "feeder.npz")
    np.savez_compressed(path, records=rec.view(np.uint8).reshape(len(rec), -1), raised=np.array(raised),
                        **{"out_" + k.replace("-", "_"): np.stack(v) for k, v in out.items()},
                        **{"s_" + k: np.asarray(v, dtype=np.float64) for k, v in scal.items()})
    print(f"{len(rec)} records, raised: {[r for r in raised if r]} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
