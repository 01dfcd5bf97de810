"""ORACLE tooling — the 10 360-candidate PROD golden (BASELINE.json configs[0] shape) and the checkpoint-shaped weights it is
quoted on, both produced by the REAL reference in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_scale_golden.py features     # ~4 min : reference forward on the 2048-candidate training set, FC1 outputs cached in /tmp
    python oracle/make_scale_golden.py tune         # ~1 min : fits conv2hidden.4 + the genotype / binary heads on those features
    python oracle/make_scale_golden.py golden       # ~20 min: reference forward on the 10 360 test candidates with the tuned weights

Why tuned weights: the bf16 acceptance bar is "identical genotype argmax on >= 99.99 % of candidates". With
random-init heads 1 % of the candidates sit within 0.3 logit units of a tie, which no 8-bit-mantissa arithmetic resolves; a trained
checkpoint separates the classes. The shipped checkpoint is unreachable (S3), so the last FC layer and the two classification heads
(263 k parameters, committed as tests/golden/prod_tuned_head.npz) are fitted to the generator's genotype label (share of reads that
carry the proposed allele: 0 -> no variant, 0.5 -> het, 1 -> hom) on top of the seeded random-init stack, FC1 included. Everything
else stays synth_state_dict(cfg, seed=1). The golden holds only what the reference printed: heads (10 360 x 27 fp32); the inputs are
regenerated from the seeds below by dl4vc_b200.synth.make_pileups.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dl4vc_b200.config import prod_config            # noqa: E402
from dl4vc_b200.synth import make_pileups            # noqa: E402
from dl4vc_b200.weights import synth_state_dict      # noqa: E402
from oracle import ref_shim                          # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CACHE = os.environ.get("DL4VC_SCALE_CACHE", "/tmp/dl4vc_scale_cache")
WEIGHT_SEED = 1
TRAIN_SEED, TRAIN_N, TRAIN_CHUNK = 777001, 2048, 128
TEST_SEED, TEST_CHUNK, TEST_CHUNKS = 20261018, 1036, 10          # chunk k: make_pileups(1036, seed=TEST_SEED + 1036 * k, "poisson")
LORA_RANK = 8
TUNED_KEYS = ("conv2hidden.4.weight", "conv2hidden.4.bias", "fcHidden2VT.weight", "fcHidden2VT.bias",
              "fcHidden2BinTarget.weight", "fcHidden2BinTarget.bias")


def tuned_state_dict(cfg, path=None):
    """synth_state_dict(cfg, seed=1) with the committed tuned tensors laid over it."""
    import torch
    sd = synth_state_dict(cfg, seed=WEIGHT_SEED, tracer=True)
    t = np.load(path or os.path.join(GOLDEN_DIR, "prod_tuned_head.npz"))
    for k in TUNED_KEYS:
        assert tuple(sd[k].shape) == t[k].shape, k
        sd[k] = torch.from_numpy(t[k].astype(np.float32))
    sd["conv2hidden.1.weight"] = torch.from_numpy(apply_low_rank(sd["conv2hidden.1.weight"].numpy(), t["fc1_u"], t["fc1_v"]))
    return sd


def labels_of(batch):
    vt = np.rint(batch.var_fraction * 2).astype(np.int64)          # 0 none, 1 het, 2 hom  (dl4vc/utils.py:18)
    return vt, (vt > 0).astype(np.int64)


def stage_features():
    import torch
    torch.set_num_threads(int(os.environ.get("DL4VC_THREADS", "6")))
    cfg = prod_config()
    model, _ = ref_shim.build_reference_model(cfg, synth_state_dict(cfg, seed=WEIGHT_SEED, tracer=True))
    grabbed = {}
    model.conv2hidden.register_forward_pre_hook(lambda m, inp: grabbed.__setitem__("fc_in", inp[0].detach().clone()))
    os.makedirs(CACHE, exist_ok=True)
    feats, vts = [], []
    t0 = time.time()
    for k in range(0, TRAIN_N, TRAIN_CHUNK):
        b = make_pileups(TRAIN_CHUNK, seed=TRAIN_SEED + k, coverage="poisson")
        ref_shim.reference_forward(model, b.arrays())
        feats.append(grabbed["fc_in"].numpy().astype(np.float16)); vts.append(labels_of(b)[0])
        print(f"features {k + TRAIN_CHUNK}/{TRAIN_N}  {time.time() - t0:.0f} s", flush=True)
    np.savez(os.path.join(CACHE, "train_features.npz"), fc_in=np.concatenate(feats), vt=np.concatenate(vts))


def apply_low_rank(w1, u, v):
    """conv2hidden.1.weight + u v^T, evaluated in float64 in a fixed order (rank-1 terms one after the other) and rounded to fp32
    once: elementwise IEEE arithmetic, hence bit-identical on every host (a BLAS matmul would not be)."""
    out = np.empty_like(w1, dtype=np.float32)
    for lo in range(0, w1.shape[0], 128):
        acc = w1[lo:lo + 128].astype(np.float64)
        for k in range(u.shape[1]):
            acc += u[lo:lo + 128, k:k + 1].astype(np.float64) * v[None, :, k].astype(np.float64)
        out[lo:lo + 128] = acc.astype(np.float32)
    return out


def stage_tune():
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    torch.set_num_threads(int(os.environ.get("DL4VC_THREADS", "8")))
    cfg = prod_config()
    sd = synth_state_dict(cfg, seed=WEIGHT_SEED, tracer=True)
    d = np.load(os.path.join(CACHE, "train_features.npz"))
    x = torch.from_numpy(d["fc_in"].astype(np.float32)); vt = torch.from_numpy(d["vt"]); bn = (vt > 0).long()
    n_fit = int(0.85 * len(x))
    t0 = time.time()
    z0 = F.linear(x, sd["conv2hidden.1.weight"], sd["conv2hidden.1.bias"])          # frozen part of FC1
    print(f"frozen FC1 product: {time.time() - t0:.0f} s", flush=True)
    xs = x / x.abs().mean()                                                            # conditioning of the low-rank factor only
    scale = float(x.abs().mean())
    u = torch.zeros((z0.shape[1], LORA_RANK), requires_grad=True)
    v = (1e-3 * torch.randn((x.shape[1], LORA_RANK))).requires_grad_(True)
    params = {k: sd[k].clone().requires_grad_(True) for k in TUNED_KEYS}
    opt = torch.optim.Adam([{"params": list(params.values()), "lr": 1e-3}, {"params": [u, v], "lr": 3e-3}], weight_decay=1e-4)

    def heads(idx):
        h1 = F.relu(z0[idx] + (xs[idx] @ v) @ u.t())
        h = F.relu(F.linear(h1, params["conv2hidden.4.weight"], params["conv2hidden.4.bias"]))
        return F.linear(h, params["fcHidden2VT.weight"], params["fcHidden2VT.bias"]), F.linear(h, params["fcHidden2BinTarget.weight"], params["fcHidden2BinTarget.bias"])

    everything = torch.arange(len(x))
    for ep in range(int(os.environ.get("DL4VC_TUNE_EPOCHS", "60"))):
        perm = torch.randperm(n_fit)
        for i in range(0, n_fit, 128):
            idx = perm[i:i + 128]
            lv, lb = heads(idx)
            loss = F.cross_entropy(lv, vt[idx]) + F.cross_entropy(lb, bn[idx])
            opt.zero_grad(); loss.backward(); opt.step()
        if ep % 5 == 4 or ep == 0:
            with torch.no_grad():
                lv, lb = heads(everything)
                acc = (lv.argmax(1) == vt).float()
                srt = lv.sort(1).values
                margin = (srt[:, -1] - srt[:, -2])[n_fit:]
                print(f"epoch {ep + 1}: loss {loss.item():.4f}  fit acc {acc[:n_fit].mean():.4f}  held-out acc {acc[n_fit:].mean():.4f}  "
                      f"held-out margin percentiles 0.1/1/5/50: {np.percentile(margin.numpy(), [0.1, 1, 5, 50]).round(3)}  max|logit| {lv.abs().max():.2f}  {time.time() - t0:.0f} s", flush=True)
    out = {k: p.detach().numpy().astype(np.float32) for k, p in params.items()}
    out["fc1_u"] = u.detach().numpy().astype(np.float32)
    out["fc1_v"] = (v.detach().numpy() / scale).astype(np.float32)                      # folded back to the raw feature scale
    path = os.path.join(GOLDEN_DIR, "prod_tuned_head.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def stage_golden():
    import torch
    torch.set_num_threads(int(os.environ.get("DL4VC_THREADS", "6")))
    cfg = prod_config()
    model, _ = ref_shim.build_reference_model(cfg, tuned_state_dict(cfg))
    heads, vts = [], []
    t0 = time.time()
    for k in range(TEST_CHUNKS):
        b = make_pileups(TEST_CHUNK, seed=TEST_SEED + TEST_CHUNK * k, coverage="poisson")
        for lo in range(0, TEST_CHUNK, 148):
            h, _, _ = ref_shim.reference_forward(model, b.slice(lo, min(lo + 148, TEST_CHUNK)).arrays())
            heads.append(h.astype(np.float32))
        vts.append(labels_of(b)[0])
        print(f"golden chunk {k + 1}/{TEST_CHUNKS}  {time.time() - t0:.0f} s", flush=True)
        np.savez(os.path.join(CACHE, "golden_partial.npz"), heads=np.concatenate(heads))
    heads = np.concatenate(heads); vt = np.concatenate(vts)
    path = os.path.join(GOLDEN_DIR, "prod_scale10k.npz")
    np.savez_compressed(path, heads=heads, vt_label=vt.astype(np.uint8), weight_seed=np.int64(WEIGHT_SEED), test_seed=np.int64(TEST_SEED),
                        chunk=np.int64(TEST_CHUNK), chunks=np.int64(TEST_CHUNKS), coverage="poisson")
    srt = np.sort(heads[:, 2:5], axis=1)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)  reference genotype accuracy vs label {(heads[:, 2:5].argmax(1) == vt).mean():.4f}, "
          f"margin percentiles 0.01/0.1/1/50: {np.percentile(srt[:, -1] - srt[:, -2], [0.01, 0.1, 1, 50]).round(4)}")


if __name__ == "__main__":
    {"features": stage_features, "tune": stage_tune, "golden": stage_golden}[sys.argv[1]]()
