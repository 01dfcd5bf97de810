"""ORACLE tooling — the trainer's loss block run with the REAL reference code (dl4vc/objectives.py + the F.* calls of dl4vc/trainer.py:252-255,
309-313,426-427) on random model outputs; losses, d(total)/d(outputs) from autograd and the close flags go to tests/golden/losses.npz.
TEST INFRASTRUCTURE ONLY.     python oracle/make_loss_goldens.py"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim      # noqa: E402

CFG = dict(label_smoothing=0.001, close_match_window=2.0, focal_gamma=0.2, focal_alpha=1.0, fp_train_weight=0.2, binary_weight=1.0,
           aux_weight=1.0, aux_allele_weight=0.001, aux_bases_weight=0.01)      # train_variant_caller.sh:104-131


def main():
    import torch
    import torch.nn.functional as F
    obj = ref_shim.import_reference_module("dl4vc.objectives")
    rng = np.random.default_rng(5)
    B = 83
    raw = rng.standard_normal((B, 27)).astype(np.float32) * 3.0
    raw[:20, 0:5] *= 6.0                                   # some confidently classified examples (close flags, focal weights near 0)
    raw[:, 5] = 1.0 / (1.0 + np.exp(-raw[:, 5]))           # xAF is a sigmoid output (model.py:954)
    tb = rng.integers(0, 2, B); tvt = rng.integers(0, 3, B)
    raw[:10, 0:2] = 0.0; raw[np.arange(10), tb[:10]] = 14.0                 # right and saturated
    taf = rng.random(B).astype(np.float32); tcov = (rng.integers(1, 100, B) * 0.01).astype(np.float32)
    tvb = rng.integers(0, 10, B); tvr = rng.integers(0, 10, B)
    wex = rng.choice([0.5, 1.0, 2.0], B).astype(np.float32)
    heads = torch.tensor(raw, requires_grad=True)
    w = torch.from_numpy(wex).unsqueeze(1)
    crit = dict(label_smoothing=CFG["label_smoothing"], close_match_window=CFG["close_match_window"], alpha=CFG["focal_alpha"], gamma=CFG["focal_gamma"])
    bin_crit = obj.SoftBCEWithLogitsFocalLoss(num_classes=2, pos_weight=torch.Tensor([CFG["fp_train_weight"], 1.]), **crit)        # trainer.py:90-97
    vt_crit = obj.SoftBCEWithLogitsFocalLoss(num_classes=3, pos_weight=torch.Tensor([CFG["fp_train_weight"], 1., 1.]), **crit)
    binary_loss, bin_close = bin_crit(heads[:, 0:2], torch.from_numpy(tb).long().unsqueeze(1), weight=w)                              # trainer.py:252
    vt_loss, vt_close = vt_crit(heads[:, 2:5], torch.from_numpy(tvt).long().unsqueeze(1), weight=w)                                   # trainer.py:255
    af_loss = F.binary_cross_entropy(heads[:, 5:6], torch.from_numpy(taf).unsqueeze(1), weight=w)                                     # trainer.py:309
    cov_loss = F.mse_loss(heads[:, 6:7], torch.from_numpy(tcov).unsqueeze(1))
    bw = torch.Tensor([0.001, 1., 1., 1., 1., 1., 0.001, 0.001, 1., 0.001])
    vb_loss = F.cross_entropy(heads[:, 7:17], torch.from_numpy(tvb).long(), weight=bw)
    vr_loss = F.cross_entropy(heads[:, 17:27], torch.from_numpy(tvr).long(), weight=bw)
    loss = binary_loss * CFG["binary_weight"]                                                                                          # trainer.py:426-427
    loss = loss + (vt_loss + af_loss * CFG["aux_allele_weight"] + cov_loss + (vb_loss + vr_loss) * CFG["aux_bases_weight"]) * CFG["aux_weight"]
    loss.backward()
    comps = np.array([binary_loss.item(), vt_loss.item(), af_loss.item(), cov_loss.item(), vb_loss.item(), vr_loss.item(), loss.item(), float(vt_close.sum())], np.float32)
    path = os.path.join(ROOT, "tests", "golden", "losses.npz")
    np.savez_compressed(path, heads=raw, target_binary=tb, target_var_type=tvt, target_allele_freq=taf, target_coverage=tcov, target_var_base=tvb,
                        target_ref_base=tvr, example_weight=wex, losses=comps, dheads=heads.grad.numpy(), close_vt=vt_close.numpy().astype(np.uint8),
                        close_bin=bin_close.numpy().astype(np.uint8), **{"cfg_" + k: np.float32(v) for k, v in CFG.items()})
    print(f"losses {comps} close {int(vt_close.sum())}/{B} -> {path}")


if __name__ == "__main__":
    main()
