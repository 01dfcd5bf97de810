"""Development / evidence: genotype-call agreement of the bf16 path against the fp32 path (the 1e-4 stand-in for the reference) on N
PROD candidates (mixed SNP / insert / delete proposals, Poisson depth). Prints overall and margin-conditioned agreement.
Usage (GPU box): python scripts/gpu_agree.py [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dl4vc_b200.config import prod_config
from dl4vc_b200.factory import build_model
from dl4vc_b200.synth import make_pileups
from dl4vc_b200.weights import synth_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10360
cfg = prod_config()
model = build_model(cfg, synth_state_dict(cfg, seed=1), precision="fp32")
ref_all, got_all = [], []
for k in range(0, n, 1036):
    m = min(1036, n - k)
    batch = make_pileups(m, seed=20261018 + k, coverage="poisson")
    t = [torch.from_numpy(np.ascontiguousarray(a)) for a in batch.arrays()]
    r, q, s, ref, rm, vm = t
    ref_all.append(model.set_precision("fp32").forward_heads(r, ref, q, s, rm, vm).cpu().numpy())
    got_all.append(model.set_precision("bf16").forward_heads(r, ref, q, s, rm, vm).cpu().numpy())
ref32, got = np.concatenate(ref_all), np.concatenate(got_all)
scale = np.abs(ref32).max()
err = np.abs(got - ref32).max() / scale
vt_ref, vt_got = ref32[:, 2:5], got[:, 2:5]
srt = np.sort(vt_ref, axis=1)
margin = srt[:, -1] - srt[:, -2]
agree = vt_ref.argmax(1) == vt_got.argmax(1)
print(f"candidates {n}  max |bf16-fp32| / max|logit| = {err:.3e}  (max|logit| {scale:.3f})")
print(f"genotype argmax agreement: all {agree.mean():.5f} ({(~agree).sum()} differ)")
for f in (0.5, 1, 2, 4):
    thr = f * err * scale
    c = margin > thr
    print(f"  fp32 margin > {f} x max error ({thr:.4f}): {c.sum()} candidates, agreement {agree[c].mean():.5f}")
print("margin percentiles (1, 5, 25, 50):", np.percentile(margin, [1, 5, 25, 50]).round(4), " margins of disagreeing calls (max):", margin[~agree].max() if (~agree).any() else None)
bagree = ref32[:, 0:2].argmax(1) == got[:, 0:2].argmax(1)
print(f"variant / no-variant argmax agreement: {bagree.mean():.5f}")
