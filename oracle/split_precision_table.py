"""ORACLE tooling — error of reduced-precision tensor-core schemes against the fp32 goldens.  TEST INFRASTRUCTURE ONLY.

Emulates, on the CPU, what a tensor-core implementation of the DAN forward would compute when every conv / linear
operand pair is fed to the MMA as

  bf16        one pass, operands rounded to bf16 (the shipped DAN_PRECISION_BF16 path)
  bf16x3      split-bf16: a = a_hi + a_lo (both bf16), three passes  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
  tf32        one pass, operands rounded to TF32 (10 explicit mantissa bits)
  tf32x3      split-TF32, three passes as above

with fp32 accumulation (torch CPU fp32 conv/linear of the rounded operands: products of the rounded values are exact in
fp32 for bf16 and accurate to 2^-24 for tf32, so only the operand rounding differs from fp32). Activations between
layers are kept in the precision the scheme can store (bf16: bf16; bf16x3: hi + lo pair = 16 mantissa bits; tf32*: fp32).

    python -m oracle.split_precision_table          # prints the table committed as profiles/r02_split_precision_error.md

It answers VERDICT r01 item 4 ("measure 3-pass split-bf16 and 3xTF32 error against the goldens first"). The forward is the
op-for-op restatement in oracle/dan_torch_cpu.py with the two matmul-like ops swapped out.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dl4vc_b200.config import DanConfig  # noqa: E402
from dl4vc_b200.weights import synth_state_dict  # noqa: E402
from oracle import dan_torch_cpu  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rn_bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def rn_tf32(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to 10 explicit mantissa bits"""
    i = x.contiguous().view(torch.int32)
    bias = ((i >> 13) & 1) + 0x0FFF
    return ((i + bias) & ~0x1FFF).view(torch.float32)


class Scheme:
    def __init__(self, name):
        self.name = name
        self.rn = {"bf16": rn_bf16, "bf16x3": rn_bf16, "tf32": rn_tf32, "tf32x3": rn_tf32}.get(name)
        self.split = name.endswith("x3")

    def parts(self, x):
        if self.rn is None:
            return [x]
        hi = self.rn(x)
        return [hi, self.rn(x - hi)] if self.split else [hi]

    def store(self, x):
        """what the scheme can keep between layers"""
        if self.name == "bf16":
            return rn_bf16(x)
        if self.name == "bf16x3":
            hi = rn_bf16(x)
            return hi + rn_bf16(x - hi)
        return x

    def op(self, fn, x, w, b, **kw):
        xs, ws = self.parts(x), self.parts(w)
        out = fn(xs[0], ws[0], b, **kw)
        if self.split:
            out = out + fn(xs[1], ws[0], None, **kw) + fn(xs[0], ws[1], None, **kw)
        return out


def forward_with(scheme: Scheme, cfg, sd, arrays):
    """dan_torch_cpu.forward with conv2d / linear routed through the scheme (same op order)."""
    real_conv, real_lin = F.conv2d, F.linear

    def conv(x, w, b=None, **kw):
        return scheme.store(scheme.op(real_conv, scheme.store(x), w, b, **kw))

    def lin(x, w, b=None):
        return scheme.op(real_lin, x, w, b)

    F.conv2d, F.linear = conv, lin
    try:
        reads, q, st, ref, rm, vm = arrays
        return dan_torch_cpu.forward(cfg, sd, reads, ref, q, st, rm, vm).numpy()
    finally:
        F.conv2d, F.linear = real_conv, real_lin


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    d = json.loads(str(g["config"]))
    d["conv_1d_pool_layers"] = tuple(d["conv_1d_pool_layers"])
    d["layer_sizes"] = tuple(d["layer_sizes"])
    cfg = DanConfig(**d)
    arrays = tuple(g[k] for k in ("reads", "q_scores", "strands", "ref", "ref_masks", "var_masks"))
    return cfg, int(g["seed"]), arrays, g["heads"]


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.abs(b).max())))


def main():
    torch.set_num_threads(8)
    cases = ["prod_full", "prod_smallfc_mixed", "prod_smallfc_edge", "variant_a", "min_smallfc", "reads300_ragged"]
    schemes = ["fp32", "bf16", "tf32", "bf16x3", "tf32x3"]
    print("| golden | " + " | ".join(schemes) + " |")
    print("|---|" + "---|" * len(schemes))
    worst = {s: 0.0 for s in schemes}
    for name in cases:
        cfg, seed, arrays, heads = load(name)
        sd = synth_state_dict(cfg, seed=seed)
        row = []
        for s in schemes:
            e = rel_err(forward_with(Scheme(s), cfg, sd, arrays), heads)
            worst[s] = max(worst[s], e)
            row.append("%.2e" % e)
        print("| %s | %s |" % (name, " | ".join(row)), flush=True)
    print("| **worst** | " + " | ".join("%.2e" % worst[s] for s in schemes) + " |")


if __name__ == "__main__":
    main()
