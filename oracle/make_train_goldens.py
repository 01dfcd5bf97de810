"""ORACLE tooling — gradients of the REAL reference in training mode (build container only).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_train_goldens.py        # needs /root/reference, CPU, ~1 min

tests/golden/train_smallfc.npz: PROD topology with a small FC trunk and hidden_dropout 0 (dropout masks are the one thing that cannot
be shared with torch's generator), 3 candidates, reference Basic2DNet in .train() mode (batch-statistics BatchNorm, running
statistics updated) -> loss = sum(heads * cw) for a fixed random cw (so d loss / d heads = cw) -> loss.backward(). Kept: the inputs,
cw, the training-mode head outputs, the gradient of EVERY parameter (tensors above 200 k elements as a strided sample) from the fp32 run
and from a float64 run of the same reference code (grad64:*, rounded to fp32), and the
BatchNorm running statistics after the step. Weights are regenerated from the seed (dl4vc_b200.weights.synth_state_dict).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dl4vc_b200.config import small_config            # noqa: E402
from dl4vc_b200.synth import make_pileups             # noqa: E402
from dl4vc_b200.weights import synth_state_dict       # noqa: E402
from oracle import ref_shim                           # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SAMPLE_ABOVE, SAMPLE_STRIDE = 200_000, 97


def run_case(name, cfg, batch, seed):
    import torch
    sd = synth_state_dict(cfg, seed=seed)
    model, _ = ref_shim.build_reference_model(cfg, sd)
    model.train()
    reads, q, s, ref, rm, vm = (torch.from_numpy(a).long() for a in batch.arrays())
    out = model(reads, ref, q_scores=q, strands=s, binary_trust_vector=None, af_scores=None, ref_bases=None, var_bases=None, ref_masks=rm, var_masks=vm)
    heads = torch.cat([o.reshape(o.shape[0], -1) for o in out[:6]], dim=1)
    cw = np.random.default_rng(seed + 1000).standard_normal(tuple(heads.shape)).astype(np.float32)
    (heads * torch.from_numpy(cw)).sum().backward()
    g = dict(config=json.dumps(cfg.to_dict()), seed=np.int64(seed), cw=cw, heads=heads.detach().numpy().astype(np.float32),
             reads=batch.reads, q_scores=batch.q_scores, strands=batch.strands, ref=batch.ref, ref_masks=batch.ref_masks, var_masks=batch.var_masks)
    for pname, p in model.named_parameters():
        if p.grad is None:
            continue
        a = p.grad.detach().numpy().astype(np.float32)
        g["grad:" + pname] = a.reshape(-1)[::SAMPLE_STRIDE].copy() if a.size > SAMPLE_ABOVE else a
    # the same step of the same reference code in float64: tells rounding noise of the fp32 reference apart from real differences
    model64, _ = ref_shim.build_reference_model(cfg, sd)
    model64 = model64.double().train()
    out64 = model64(reads, ref, q_scores=q, strands=s, binary_trust_vector=None, af_scores=None, ref_bases=None, var_bases=None, ref_masks=rm, var_masks=vm)
    heads64 = torch.cat([o.reshape(o.shape[0], -1) for o in out64[:6]], dim=1)
    (heads64 * torch.from_numpy(cw).double()).sum().backward()
    for pname, p in model64.named_parameters():
        if p.grad is None:
            continue
        a = p.grad.detach().numpy()
        g["grad64:" + pname] = (a.reshape(-1)[::SAMPLE_STRIDE].copy() if a.size > SAMPLE_ABOVE else a).astype(np.float32)
    for bname, b in model.named_buffers():
        if "running_" in bname:
            g["buf:" + bname] = b.detach().numpy().astype(np.float32)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **g)
    print(f"{name}: {len([k for k in g if k.startswith('grad:')])} gradient tensors -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    run_case("train_smallfc", small_config(hidden_dropout=0.0), make_pileups(3, seed=61, coverage="poisson"), seed=7)
