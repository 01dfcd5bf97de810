"""ORACLE tooling — generate tests/golden/*.npz by running the REAL reference in the build container.
TEST INFRASTRUCTURE ONLY.

    python oracle/make_goldens.py            # needs /root/reference (read-only mount), CPU only, ~1 min

Each golden holds: the uint8 inputs, the config, the weight seed (weights are regenerated from the seed by
dl4vc_b200.weights.synth_state_dict — a 311 MB state_dict cannot be committed), and what the unmodified reference
(dl4vc/model.py, executed through oracle/ref_shim.py) produced for them: the six head outputs (B,27), a SHA-256
and slices of the conv-1 input tensor (bit-exact target for the integer/encoding work), per-layer activation
statistics and the highway vectors of candidate 0.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dl4vc_b200.config import DanConfig, prod_config, small_config, min_config  # noqa: E402
from dl4vc_b200.synth import make_pileups, edge_case_pileups, PileupBatch       # noqa: E402
from dl4vc_b200.weights import synth_state_dict                                  # noqa: E402
from oracle import ref_shim                                                       # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def cat_batches(bs):
    return PileupBatch(*(np.concatenate([getattr(b, f) for b in bs], axis=0)
                         for f in ("reads", "q_scores", "strands", "ref", "ref_masks", "var_masks", "num_reads", "kind")))


def run_case(name, cfg: DanConfig, batch: PileupBatch, seed: int):
    import torch

    sd = synth_state_dict(cfg, seed=seed)
    model, _ = ref_shim.build_reference_model(cfg, sd)
    acts = {}
    hooks = []
    L = cfg.total_conv_layers

    # layer outputs as the reference defines them (after BN / residual): input of the next conv, or of the pools
    def grab(key):
        return lambda m, inp: acts.__setitem__(key, inp[0].detach())
    # With the highway on, the bottleneck conv sees exactly the layer output (model.py:774); otherwise the next
    # conv's input is used, which for pool layers already includes the "+ pool" term (model.py:742).
    includes_pool = np.zeros(L, dtype=np.uint8)
    for l in range(1, L):
        if cfg.highway:
            hooks.append(model.conv1D_bottleneck_layers[l - 1].register_forward_pre_hook(grab(f"layer{l}")))
        else:
            hooks.append(model.conv1D_layers[l].register_forward_pre_hook(grab(f"layer{l}")))
            includes_pool[l - 1] = int(l in cfg.conv_1d_pool_layers)
    hooks.append(model.avgPool1D.register_forward_pre_hook(grab(f"layer{L}")))
    if cfg.highway:
        for l in range(L):
            hooks.append(model.conv1D_compression_layers[l].register_forward_hook(
                lambda m, i, o, l=l: acts.__setitem__(f"hw{l}", o.detach())))
    hooks.append(model.conv2hidden.register_forward_pre_hook(grab("fc_in")))
    heads, out, cap = ref_shim.reference_forward(model, batch.arrays(), capture_conv_input=True)
    for h in hooks:
        h.remove()
    x0 = cap["x0"].contiguous().numpy()            # (B,Cin,R,P) logical order
    g = dict(
        config=json.dumps(cfg.to_dict()), seed=np.int64(seed),
        reads=batch.reads, q_scores=batch.q_scores, strands=batch.strands, ref=batch.ref,
        ref_masks=batch.ref_masks, var_masks=batch.var_masks,
        heads=heads.astype(np.float32),
        x0_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(x0).tobytes()).digest(), dtype=np.uint8),
        x0_mask_channels=x0[:, 2 * cfg.embed_dim:, :, :].astype(np.float16) if cfg.in_channels > 2 * cfg.embed_dim else np.zeros(0, np.float16),
        x0_cand0_reads=x0[0][:, [0, 1, cfg.num_reads // 2, cfg.num_reads - 1], :],
        fc_in_stats=np.stack([acts["fc_in"].numpy().mean(1), np.abs(acts["fc_in"].numpy()).mean(1)]),
        fc_in_cand0_sample=acts["fc_in"][0].numpy()[::37].copy(),
    )
    # per-layer statistics of candidate 0 (mean and mean-abs per channel) and a strided sample
    lay_mean, lay_abs, lay_samp = [], [], []
    for l in range(1, L + 1):
        a = acts[f"layer{l}"][0].numpy()              # (C,R,P)
        lay_mean.append(a.mean(axis=(1, 2))); lay_abs.append(np.abs(a).mean(axis=(1, 2)))
        lay_samp.append(a[::16, ::33, ::20].copy())
    g.update(layer_includes_pool=includes_pool, layer_mean=np.stack(lay_mean), layer_absmean=np.stack(lay_abs), layer_sample=np.stack(lay_samp))
    if cfg.highway:
        g["highway_cand0"] = np.stack([acts[f"hw{l}"][0, :, :, 0].numpy() for l in range(L)])   # (L, bott, R)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **g)
    print(f"{name}: B={len(batch)} heads range [{heads.min():.3f},{heads.max():.3f}] -> {path} "
          f"({os.path.getsize(path)/1024:.0f} KiB)")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]          # optional: names of the cases to (re)generate
    global run_case
    _run = run_case
    run_case = lambda name, *a, **k: _run(name, *a, **k) if (not only or name in only) else None
    mixed = cat_batches([make_pileups(3, seed=11, coverage="poisson"), make_pileups(2, seed=12, coverage="full"),
                         make_pileups(1, seed=13, coverage="ragged")])
    # PROD topology, small FC trunk: mixed coverage + the hand-made corner cases
    run_case("prod_smallfc_mixed", small_config(), mixed, seed=1)
    run_case("prod_smallfc_edge", small_config(), edge_case_pileups(), seed=2)
    # the full shipped shape (77.7 M parameters)
    run_case("prod_full", prod_config(), cat_batches([make_pileups(2, seed=21, coverage="poisson"),
                                                       make_pileups(1, seed=22, coverage="full")]), seed=3)
    # argparse-default topology ("MIN": no BN / highway / masks, pool_combine Linear), FC shrunk via layer_sizes
    run_case("min_smallfc", min_config(layer_sizes=(64, 32), pool_combine_dimension=96), make_pileups(3, seed=31, coverage="poisson"), seed=4)
    # variations the constructor allows: dilation 1, no residual, two pool layers, averaged highway, mean-pool only
    run_case("variant_a", small_config(total_conv_layers=4, residual_layer_start=3, middle_layer_dilation=1,
                                       final_layer_dilation=2, conv_1d_pool_layers=(1, 3), concat_hw_reads=False,
                                       skip_final_maxpool=True, use_strands=False, hidden_dropout=0.0),
             make_pileups(3, seed=41, coverage="poisson"), seed=5)
    # BASELINE configs[3]: 300 read rows, ragged depth 1..300 (reference built with MAX_READS = num_single_reads = 300)
    run_case("reads300_ragged", small_config(num_reads=300), make_pileups(3, seed=51, num_reads=300, coverage="ragged", max_depth=300), seed=6)


if __name__ == "__main__":
    main()
