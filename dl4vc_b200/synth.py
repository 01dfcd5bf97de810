"""Synthetic pileup candidates in the model-facing layout of the reference loader.

Shapes/dtypes follow what `ContextDatasetFromNumpy.__getitem__` returns (reference: dl4vc/dataset.py:672-680):
reads / q-scores / strands uint8 [201 positions, 100 reads] per candidate, ref / ref_mask / var_mask uint8 [201].
The proposal masks follow the rule of `get_read_mask_vectors` (reference: dl4vc/dataset.py:112-250) for the three
proposal kinds (SNP, insert, delete) — see SURVEY App. C for probed examples. Recipe: SURVEY §8d.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

P = 201
CENTER = 100
TOK_GAP, TOK_START, TOK_END, TOK_NOINS = 5, 6, 7, 8


@dataclass
class PileupBatch:
    reads: np.ndarray      # (B, 201, R) uint8
    q_scores: np.ndarray   # (B, 201, R) uint8
    strands: np.ndarray    # (B, 201, R) uint8
    ref: np.ndarray        # (B, 201) uint8
    ref_masks: np.ndarray  # (B, 201) uint8
    var_masks: np.ndarray  # (B, 201) uint8
    num_reads: np.ndarray  # (B,) int32 populated rows
    kind: np.ndarray       # (B,) uint8  0 SNP, 1 insert, 2 delete
    var_fraction: np.ndarray = None   # (B,) float32 planted share of reads carrying the proposed allele: 0 / 0.5 / 1 (a genotype label)

    def __len__(self):
        return self.reads.shape[0]

    def arrays(self):
        return (self.reads, self.q_scores, self.strands, self.ref, self.ref_masks, self.var_masks)

    def slice(self, lo, hi):
        return PileupBatch(*(a[lo:hi] for a in (self.reads, self.q_scores, self.strands, self.ref, self.ref_masks,
                                                 self.var_masks, self.num_reads, self.kind)),
                           var_fraction=None if self.var_fraction is None else self.var_fraction[lo:hi])

    @property
    def nbytes(self):
        return sum(a.nbytes for a in self.arrays())


def make_pileups(n: int, seed: int = 20261018, num_reads: int = 100, coverage: str = "full",
                 max_depth: int = 100) -> PileupBatch:
    """n synthetic candidates.

    coverage: "full" (all rows populated, the dense worst case), "poisson" (n ~ Poisson(55) clipped to
    [1, num_reads]), "ragged" (on-disk depth uniform in [1, max_depth]; depths above num_reads are reduced to a
    sorted random num_reads-subset exactly like the reference's sample_single_reads, dl4vc/dataset.py:256-287 —
    after which all rows are populated), "empty" (no reads at all).
    """
    rng = np.random.default_rng(seed)
    R = num_reads
    if coverage == "full":
        depth = np.full(n, R, dtype=np.int64)
    elif coverage == "poisson":
        depth = np.clip(rng.poisson(55, n), 1, R)
    elif coverage == "ragged":
        depth = np.minimum(rng.integers(1, max_depth + 1, n), R)  # sampling a subset of iid rows == fewer iid rows
    elif coverage == "empty":
        depth = np.zeros(n, dtype=np.int64)
    else:
        raise ValueError(coverage)

    ref = rng.integers(1, 5, (n, P)).astype(np.uint8)
    kind = rng.choice(3, n, p=[0.8, 0.1, 0.1]).astype(np.uint8)
    vlen = rng.integers(2, 12, n)                      # total mask length for indels
    ref_masks = np.zeros((n, P), np.uint8)
    var_masks = np.zeros((n, P), np.uint8)
    cols = np.arange(P)
    span = (cols[None, :] > CENTER) & (cols[None, :] < CENTER + vlen[:, None])   # columns after the anchor base
    is_ins, is_del, is_snp = kind == 1, kind == 2, kind == 0
    # insert: reference shows '-' in the inserted columns; ref_mask = [b, noinsert...], var_mask = [b, ins...]
    ref = np.where(is_ins[:, None] & span, TOK_GAP, ref).astype(np.uint8)
    ins_bases = rng.integers(1, 5, (n, P)).astype(np.uint8)
    anchor = ref[:, CENTER]
    ref_masks[:, CENTER] = anchor
    var_masks[:, CENTER] = anchor
    alt = (anchor - 1 + rng.integers(1, 4, n)) % 4 + 1           # SNP: a different base
    var_masks[is_snp, CENTER] = alt[is_snp].astype(np.uint8)
    ref_masks = np.where(is_ins[:, None] & span, TOK_NOINS, ref_masks).astype(np.uint8)
    var_masks = np.where(is_ins[:, None] & span, ins_bases, var_masks).astype(np.uint8)
    # delete: ref_mask = deleted reference bases, var_mask = [b0, '-', '-', ...]
    ref_masks = np.where(is_del[:, None] & span, ref, ref_masks).astype(np.uint8)
    var_masks = np.where(is_del[:, None] & span, TOK_GAP, var_masks).astype(np.uint8)

    # ---- reads, built as (n, R, P) then transposed to the loader's [position, read] order --------------
    row = np.arange(R)
    populated = row[None, :] < depth[:, None]                               # (n, R)
    start = rng.integers(-50, 151, (n, R))
    length = np.clip(np.rint(rng.normal(150, 10, (n, R))), 30, 400).astype(np.int64)
    end = start + length - 1
    c = cols[None, None, :]
    inside = (c >= start[..., None]) & (c <= end[..., None]) & populated[..., None]
    base = np.broadcast_to(ref[:, None, :], (n, R, P))
    u = rng.random((n, R, P), dtype=np.float32)
    noise = rng.integers(1, 5, (n, R, P)).astype(np.uint8)
    reads = np.where(u < 0.01, noise, base)
    reads = np.where(u > 0.998, TOK_GAP, reads)
    reads = np.where((base == TOK_GAP), TOK_NOINS, reads)                    # no insertion in this read
    # proposal support: fraction of reads carrying the variant allele is 0 / 0.5 / 1 per candidate
    frac = rng.choice([0.0, 0.5, 1.0], n)
    carries = rng.random((n, R)) < frac[:, None]
    vm = np.broadcast_to(var_masks[:, None, :], (n, R, P))
    rm = np.broadcast_to(ref_masks[:, None, :], (n, R, P))
    keep_noise = rng.random((n, R)) < 0.05                                   # a few reads keep sequencing noise
    plant_v = carries[..., None] & (vm != 0)
    plant_r = (~carries & ~keep_noise)[..., None] & (rm != 0)
    reads = np.where(plant_v, vm, reads)
    reads = np.where(plant_r, rm, reads)
    reads = np.where(inside, reads, 0)
    reads = np.where((c == start[..., None] - 1) & populated[..., None], TOK_START, reads)
    reads = np.where((c == end[..., None] + 1) & populated[..., None], TOK_END, reads)
    reads = reads.astype(np.uint8)
    q = np.clip(np.rint(rng.normal(33, 6, (n, R, P))), 2, 41).astype(np.uint8)
    q = np.where(inside, q, 0).astype(np.uint8)
    strand_of_read = rng.integers(1, 3, (n, R)).astype(np.uint8)
    strands = np.where(inside, strand_of_read[..., None], 0).astype(np.uint8)

    tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
    return PileupBatch(tr(reads), tr(q), tr(strands), ref, ref_masks, var_masks, depth.astype(np.int32), kind, frac.astype(np.float32))


def edge_case_pileups(num_reads: int = 100) -> PileupBatch:
    """A handful of hand-made corner cases: no reads, all-pad masks (blacklisted example, dataset.py:644-664),
    extreme q-scores/strands, every token value, a fully agreeing and a fully disagreeing pileup."""
    b = make_pileups(8, seed=7, num_reads=num_reads, coverage="full")
    R = num_reads
    # 0: completely empty pileup
    b.reads[0] = 0; b.q_scores[0] = 0; b.strands[0] = 0
    # 1: masks all zero (blacklisted example) -> every read "agrees" with an all-pad mask
    b.ref_masks[1] = 0; b.var_masks[1] = 0
    # 2: q-scores at the extremes 0..93 and 255, strands cycling 0/1/2
    b.q_scores[2] = (np.arange(P * R).reshape(P, R) % 95).astype(np.uint8)
    b.q_scores[2, 0, :] = 255
    b.strands[2] = (np.arange(P * R).reshape(P, R) % 3).astype(np.uint8)
    # 3: every token value 0..9 appears in reads and ref
    b.reads[3] = (np.arange(P * R).reshape(P, R) % 10).astype(np.uint8)
    b.ref[3] = (np.arange(P) % 10).astype(np.uint8)
    # 4: all reads carry the variant exactly; 5: all reads carry the reference exactly
    for i, m in ((4, b.var_masks), (5, b.ref_masks)):
        nz = m[i] != 0
        b.reads[i][nz, :] = m[i][nz][:, None]
    # 6: mask touching the right edge of the window
    b.ref_masks[6] = 0; b.var_masks[6] = 0
    b.ref_masks[6, 198:201] = [1, 2, 3]; b.var_masks[6, 198:201] = [1, 5, 5]
    b.reads[6, 198:201, ::2] = np.array([1, 5, 5], np.uint8)[:, None]
    # 7: single populated read
    b.reads[7, :, 1:] = 0; b.q_scores[7, :, 1:] = 0; b.strands[7, :, 1:] = 0
    b.num_reads[:] = [0, R, R, R, R, R, R, 1]
    return b
