"""Host-side feeder and score post-processing around the DAN forward (north_star item 4).

PinnedBatchFeeder  collates the loader's per-candidate dicts (the schema `dl4vc/dataset.py:672-680` yields: uint8 `reads`,
                   `q-scores`, `strands` of shape (201, 100), `ref`, `ref_mask`, `var_mask` of shape (201,)) straight into
                   rotating PINNED uint8 batch buffers — no int64 inflation (`trainer.py:520-528` moves 8x the bytes) — and
                   hands them to `Basic2DNet.forward_heads_host`, whose library side overlaps the H2D staging of a chunk
                   with the kernels of the previous one. Collating batch k+1 on the host overlaps the GPU work of batch k.
scores_from_heads  the caller-side post-ops of `trainer.py:611-623`: softmax over xbinary / xVT and the variant score 1 - p0.
make_mask_vectors  batch proposal-mask decode of the loader (`dataset.py:112-250`) in the C-ABI library, pinned to the reference's outputs.
scores_on_device   the same post-ops as one CUDA kernel behind the C-ABI (`dan_scores`): 4 floats per candidate leave the GPU.
RecordFile / decode_records  raw fixed-record pileup file (the reference's HDF5 compound type as a flat np.memmap: h5py is not in the image) and
                   the batched C decode of `ContextDatasetFromNumpy._get_generator` (`dataset.py:500-680`: row window, sample_single_reads,
                   parse_vcf, count_variants_from_single_reads, proposal masks) straight into a pinned HostBatch — no per-item Python.
format_vcf_info    the `BP=..;NV=..;HV=..;OV=..` field `utils.append_vcf_records` splices into VCF column 3 (`utils.py:162-178`).
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

ITEM_KEYS = (("reads", "reads"), ("q-scores", "q"), ("strands", "strands"), ("ref", "ref"), ("ref_mask", "ref_masks"),
             ("var_mask", "var_masks"))


class HostBatch:
    """One pinned (when CUDA is available) uint8 batch in the layout the C-ABI takes: [batch][position][read]."""

    def __init__(self, capacity: int, read_len: int = 201, num_reads: int = 100, pin: bool | None = None):
        pin = torch.cuda.is_available() if pin is None else pin
        mk = lambda *shape: (torch.empty(shape, dtype=torch.uint8).pin_memory() if pin else torch.empty(shape, dtype=torch.uint8))
        self.reads, self.q, self.strands = (mk(capacity, read_len, num_reads) for _ in range(3))
        self.ref, self.ref_masks, self.var_masks = (mk(capacity, read_len) for _ in range(3))
        self.capacity, self.size = capacity, 0
        self.meta = []

    def fill(self, items) -> "HostBatch":
        """Copy a list of dataset items into the buffers (values are 0..93, any integer dtype narrows losslessly)."""
        n = len(items)
        if n > self.capacity:
            raise ValueError(f"{n} items > capacity {self.capacity}")
        views = {"reads": self.reads.numpy(), "q": self.q.numpy(), "strands": self.strands.numpy(), "ref": self.ref.numpy(),
                 "ref_masks": self.ref_masks.numpy(), "var_masks": self.var_masks.numpy()}
        for i, it in enumerate(items):
            for key, dst in ITEM_KEYS:
                a = np.asarray(it[key])
                if a.shape != views[dst].shape[1:]:
                    raise ValueError(f"item {i}: '{key}' has shape {a.shape}, expected {views[dst].shape[1:]} (dataset.py:672-680)")
                views[dst][i] = a          # numpy casts to uint8
        self.size = n
        self.meta = [(it.get("name"), it.get("vcfrec")) for it in items]
        return self

    def tensors(self):
        n = self.size
        return self.reads[:n], self.ref[:n], self.q[:n], self.strands[:n], self.ref_masks[:n], self.var_masks[:n]


class PinnedBatchFeeder:
    """submit(items) -> pending result; results() yields (meta, heads[n, 27]) in submission order.

    `depth` host batches rotate: while the GPU works on batch k (asynchronously, on `stream`), batch k+1 is collated into the
    next pinned buffer. A buffer is reused only after its result has been consumed, so at most `depth` batches are in flight."""

    def __init__(self, model, batch_size: int, depth: int = 2, stream: torch.cuda.Stream | None = None):
        self.model, self.batch_size, self.depth = model, batch_size, depth
        self.stream = stream
        self.free = deque(HostBatch(batch_size, model.single_read_len, model.num_single_reads) for _ in range(depth))
        self.pending = deque()

    def submit(self, items):
        if not self.free:
            raise RuntimeError("all host batches are in flight: consume results() first")
        hb = self.free.popleft().fill(items)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            out = self.model.forward_heads_host(*hb.tensors())
            ev = torch.cuda.Event()
            ev.record()
        self.pending.append((hb, out, ev))

    def results(self, drain: bool = True):
        """Yield finished (meta, heads) pairs; with drain=True waits for everything submitted so far."""
        while self.pending and (drain or self.pending[0][2].query()):
            hb, out, ev = self.pending.popleft()
            ev.synchronize()
            heads = out[: hb.size].clone()
            meta = hb.meta
            self.free.append(hb)
            yield meta, heads


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def scores_from_heads(heads: torch.Tensor):
    """(B,27) head matrix -> (bin_score (B,), vt_probs (B,3)) exactly as trainer.test computes them (trainer.py:611-623,
    use_var_type_threshold off): bin_score = 1 - softmax(xbinary)[:,0]; vt_probs = softmax(xVT) = P{no variant, het, hom}."""
    xbinary, xvt = heads[:, 0:2], heads[:, 2:5]
    bin_score = 1.0 - torch.softmax(xbinary, dim=1)[:, 0]
    return bin_score, torch.softmax(xvt, dim=1)


def scores_on_device(heads: torch.Tensor) -> torch.Tensor:
    """Device version of `scores_from_heads` through the C-ABI (`dan_scores`, include/dan_b200.h): (B,27) CUDA head matrix ->
    (B,4) CUDA tensor [bin_score, P(no variant), P(het), P(hom)], so only 16 bytes per candidate cross PCIe on the way to
    the VCF writer. No CPU path: raises on a CPU tensor."""
    from . import _lib
    if not heads.is_cuda:
        raise RuntimeError("scores_on_device needs the CUDA head matrix written by forward_heads (no CPU path)")
    if heads.dim() != 2 or heads.shape[1] != 27 or heads.dtype != torch.float32:
        raise RuntimeError("heads must be (B, 27) float32")
    heads = heads.contiguous()
    out = torch.empty((heads.shape[0], 4), dtype=torch.float32, device=heads.device)
    with torch.cuda.device(heads.device):
        _lib.check(_lib.load_library().dan_scores(heads.data_ptr(), heads.shape[0], out.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "dan_scores")
    return out


CALL_THRESHOLD_FIELDS = ("snp", "snp_zygo", "indel", "indel_zygo", "long_indel", "long_indel_zygo", "delete", "delete_zygo")


def resolve_call_thresholds(snp=0.3, snp_zygo=0.5, indel=0.0, indel_zygo=0.5, long_indel=0.0, long_indel_zygo=0.5, delete=0.0, delete_zygo=0.5):
    """The threshold fall-backs of tools/format_vcf.py:57-80 (arguments = its command-line options, same defaults). The script leaves the
    delete pair unassigned when --indel_threshold is not given (and raises on the first short delete); here it follows the indel pair."""
    if not indel > 0.0:
        indel, indel_zygo = snp, snp_zygo
        long_indel, long_indel_zygo = indel, indel_zygo
        delete, delete_zygo = indel, indel_zygo
    else:
        if not long_indel > 0.0:
            long_indel, long_indel_zygo = indel, indel_zygo
        if not delete > 0.0:
            delete, delete_zygo = indel, indel_zygo
    return (snp, snp_zygo, indel, indel_zygo, long_indel, long_indel_zygo, delete, delete_zygo)


def genotype_calls(scores, ref_len, var_len, **thresholds):
    """Host restatement of the per-record part of tools/format_vcf.py:107-138, vectorised: (gt, q) with gt 0 = dropped, 1 = "0/1", 2 = "1/1"
    and q the script's quality bucket (-1 when dropped). `scores` (n, 4) = [BP, NV, HV, OV] as written into the VCF ("%.8f")."""
    t = resolve_call_thresholds(**thresholds)
    s = np.asarray(scores, dtype=np.float32)
    nv = np.rint(s[:, 1].astype(np.float64) * 1e8) / 1e8
    ov = np.rint(s[:, 3].astype(np.float64) * 1e8) / 1e8
    rl, vl = np.asarray(ref_len), np.asarray(var_len)
    snp, lng, dele = (rl == 1) & (vl == 1), (rl >= 3) | (vl >= 3), (rl > 1) & (vl == 1)
    thr = np.where(snp, t[0], np.where(lng, t[4], np.where(dele, t[6], t[2])))
    hz = np.where(snp, t[1], np.where(lng, t[5], np.where(dele, t[7], t[3])))
    margin = (1.0 - nv) - thr
    keep = margin >= 0.0
    gt = np.where(keep, np.where(ov >= hz, 2, 1), 0).astype(np.int8)
    q = np.where(keep, (margin / (1.0 - thr) * 50.0).astype(np.int64), -1).astype(np.int32)
    return gt, q


def genotype_calls_on_device(scores: torch.Tensor, ref_len: torch.Tensor, var_len: torch.Tensor, **thresholds):
    """Device version through the C-ABI (`dan_genotype_calls`): scores (B,4) CUDA float32 from `scores_on_device`, allele lengths (B,) int32
    CUDA tensors -> (gt int8, q int32) CUDA tensors. No CPU path."""
    import ctypes as C
    from . import _lib
    if not scores.is_cuda:
        raise RuntimeError("genotype_calls_on_device needs CUDA tensors (no CPU path; genotype_calls is the host restatement)")
    B = scores.shape[0]
    scores = scores.contiguous().float()
    rl = ref_len.to(device=scores.device, dtype=torch.int32).contiguous(); vl = var_len.to(device=scores.device, dtype=torch.int32).contiguous()
    gt = torch.empty(B, dtype=torch.int8, device=scores.device); q = torch.empty(B, dtype=torch.int32, device=scores.device)
    d = dict(zip(CALL_THRESHOLD_FIELDS, (0.3, 0.5, 0.0, 0.5, 0.0, 0.5, 0.0, 0.5)))
    d.update(thresholds)
    thr = (C.c_double * 8)(*[d[k] for k in CALL_THRESHOLD_FIELDS])
    with torch.cuda.device(scores.device):
        _lib.check(_lib.load_library().dan_genotype_calls(scores.data_ptr(), rl.data_ptr(), vl.data_ptr(), B, C.cast(thr, C.c_void_p), gt.data_ptr(), q.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream), "dan_genotype_calls")
    return gt, q


def format_vcf_info(bin_score, vt_probs):
    """The strings utils.append_vcf_records writes into VCF column 3 (utils.py:171-176), vectorised over the batch."""
    b = np.asarray(bin_score, dtype=np.float64)
    v = np.asarray(vt_probs, dtype=np.float64)
    return ["BP=%.8f;NV=%.8f;HV=%.8f;OV=%.8f" % (b[i], v[i, 0], v[i, 1], v[i, 2]) for i in range(len(b))]


def make_mask_vectors(ref_alleles, var_alleles, references, strict: bool = True):
    """Batch version of the loader's `get_read_mask_vectors` (dl4vc/dataset.py:112-250) through the C-ABI (`dan_make_mask_vectors`):
    REF / ALT strings of n VCF records + their encoded reference windows (n, 201) uint8 -> (ref_masks, var_masks, status), each mask
    (n, 201) uint8. `status[i]` != 0 marks a record the reference itself would have raised on (masks all zero); with `strict` those
    raise a RuntimeError like the reference's assert would."""
    import ctypes as C
    from . import _lib
    refs = np.ascontiguousarray(np.asarray(references, dtype=np.uint8))
    n = len(ref_alleles)
    if refs.shape != (n, 201) or len(var_alleles) != n:
        raise RuntimeError("references must be (n, 201) uint8 with one REF / ALT string per row")
    xs = (C.c_char_p * max(n, 1))(*[str(a).encode("ascii") for a in ref_alleles])
    ys = (C.c_char_p * max(n, 1))(*[str(a).encode("ascii") for a in var_alleles])
    rm = np.zeros((n, 201), np.uint8); vm = np.zeros((n, 201), np.uint8); st = np.zeros(n, np.int32)
    lib = _lib.load_library()
    rc = lib.dan_make_mask_vectors(xs, ys, refs.ctypes.data, n, rm.ctypes.data, vm.ctypes.data, st.ctypes.data)
    if rc != 0 and (strict or not st.any()):
        _lib.check(rc, "dan_make_mask_vectors")
    return rm, vm, st


def format_vcf_info_native(scores) -> list:
    """Same strings as `format_vcf_info` from a (B,4) float32 score matrix [bin_score, P(no variant), P(het), P(hom)] (the output of
    `scores_on_device`, copied to the host), formatted by the C-ABI's `dan_format_vcf_info` instead of a Python loop."""
    import ctypes as C
    from . import _lib
    arr = np.ascontiguousarray(np.asarray(scores, dtype=np.float32))
    if arr.ndim != 2 or arr.shape[1] != 4:
        raise RuntimeError("scores must be (B, 4)")
    n, stride = arr.shape[0], 56
    buf = C.create_string_buffer(max(n, 1) * stride)
    _lib.check(_lib.load_library().dan_format_vcf_info(arr.ctypes.data, n, C.addressof(buf), len(buf)), "dan_format_vcf_info")
    raw = buf.raw
    return [raw[i * stride:i * stride + 55].decode("ascii") for i in range(n)]


def splice_vcf_records(vcf_records, bin_score, vt_probs):
    """Records with the score field spliced in, one per line, as append_vcf_records would append them (utils.py:166-178)."""
    info = format_vcf_info(bin_score, vt_probs)
    if len(info) != len(vcf_records):
        raise AssertionError("mis-match between results and VCF to save")          # utils.py:165
    out = []
    for rec, txt in zip(vcf_records, info):
        items = rec.strip().split("\t")
        assert items[2] == ".", "DANGER -- would replace non-empty INFO -- check the hack"   # utils.py:172
        items[2] = txt
        out.append("\t".join(items))
    return out


# ---- raw record files + batched decode (dan_decode_records) --------------------------------------------------------------------------
RECORD_DTYPE = np.dtype([("name", "S16"), ("ref", np.uint8, (5, 201)), ("reads", np.uint16, (5, 201)), ("single_reads", np.uint8, (200, 201)),
                         ("ref_bases", np.uint8, (201,)), ("num_reads", np.int32), ("label", np.uint8), ("vcfrec", "S128"),
                         ("q-scores", np.uint8, (200, 201)), ("strand", np.uint8, (200, 201))])      # tools/convert_bam_single_reads.py:694-698
RECORD_BYTES = 123965
assert RECORD_DTYPE.itemsize == RECORD_BYTES
REC_STATUS = {0: "ok", 1: "vcf record has too few columns", 2: "allele letter outside base_enum", 3: "unknown mutation (equal-length non-SNP)",
              4: "INFO column without AF= / DP=", 5: "proposal masks: the reference raises (not an assertion)"}


class RecordFile:
    """np.memmap over back-to-back records of RECORD_DTYPE (what `hdfile['data']` holds in the reference, dataset.py:501-503)."""

    def __init__(self, path_or_array):
        if isinstance(path_or_array, np.ndarray):
            a = path_or_array
            self.raw = np.ascontiguousarray(a.view(np.uint8).reshape(len(a), -1)) if a.dtype == RECORD_DTYPE else np.ascontiguousarray(a, dtype=np.uint8)
        else:
            self.raw = np.memmap(path_or_array, dtype=np.uint8, mode="r").reshape(-1, RECORD_BYTES)
        assert self.raw.ndim == 2 and self.raw.shape[1] == RECORD_BYTES

    def __len__(self):
        return self.raw.shape[0]

    def field(self, idx, name):
        return self.raw[idx].view(RECORD_DTYPE)[0][name]


def decode_records(records: RecordFile, indices, batch: HostBatch | None = None, use_q_scores=True, use_strands=True, keep_candidate_af=False, seed=0,
                   max_reads=100, store_max_reads=200, strict=True):
    """Decode `indices` of a RecordFile into a (pinned) HostBatch + the per-example scalars the trainer reads (dataset.py:672-680).
    Returns (batch, scalars dict of numpy arrays incl. `status` / `blacklist`). strict: raise if the reference's loader would."""
    import ctypes as C

    from . import _lib
    lib = _lib.load_library()

    class Cfg(C.Structure):
        _fields_ = [("max_reads", C.c_int32), ("store_max_reads", C.c_int32), ("use_q_scores", C.c_int32), ("use_strands", C.c_int32),
                    ("keep_candidate_af", C.c_int32), ("seed", C.c_uint64)]

    class Out(C.Structure):
        _fields_ = [(n, C.c_void_p) for n in ("reads", "q_scores", "strands", "ref", "ref_masks", "var_masks", "label", "num_reads", "is_snp", "var_type",
                                              "allele_freq", "coverage", "var_base_enum", "var_ref_enum", "blacklist", "status")]

    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
    n = len(idx)
    if n and (idx.min() < 0 or idx.max() >= len(records)):
        raise IndexError("record index out of range")
    if batch is None:
        batch = HostBatch(max(n, 1), num_reads=max_reads)
    if n > batch.capacity:
        raise ValueError(f"{n} records > capacity {batch.capacity}")
    sc = {"label": np.zeros(n, np.uint8), "num_reads": np.zeros(n, np.int32), "is_snp": np.zeros(n, np.uint8), "var_type": np.zeros(n, np.int32),
          "allele_freq": np.zeros(n, np.float32), "coverage": np.zeros(n, np.int32), "var_base_enum": np.zeros(n, np.int32),
          "var_ref_enum": np.zeros(n, np.int32), "blacklist": np.zeros(n, np.uint8), "status": np.zeros(n, np.int32)}
    out = Out(batch.reads.data_ptr(), batch.q.data_ptr(), batch.strands.data_ptr(), batch.ref.data_ptr(), batch.ref_masks.data_ptr(),
              batch.var_masks.data_ptr(), *[sc[k].ctypes.data for k in ("label", "num_reads", "is_snp", "var_type", "allele_freq", "coverage",
                                                                         "var_base_enum", "var_ref_enum", "blacklist", "status")])
    cfg = Cfg(max_reads, store_max_reads, int(use_q_scores), int(use_strands), int(keep_candidate_af), int(seed))
    lib.dan_decode_records.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(Cfg), C.POINTER(Out)]
    rc = lib.dan_decode_records(records.raw.ctypes.data, RECORD_BYTES, idx.ctypes.data, n, C.byref(cfg), C.byref(out))
    batch.size = n
    batch.meta = [(bytes(records.field(int(i), "name")).decode(), bytes(records.field(int(i), "vcfrec")).decode()) for i in idx]
    if rc < 0 and strict:
        bad = np.flatnonzero(sc["status"])
        raise ValueError(f"{len(bad)} records cannot be decoded (the reference loader raises on them): " +
                         ", ".join(f"#{int(idx[b])}: {REC_STATUS.get(int(sc['status'][b]), '?')}" for b in bad[:5]))
    return batch, sc
