"""GPU (-m gpu): training-mode forward + backward through dan_train_forward / dan_backward against the REAL reference's autograd
(tests/golden/train_smallfc.npz, oracle/make_train_goldens.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from dl4vc_b200.factory import build_model
from dl4vc_b200.synth import make_pileups
from dl4vc_b200.weights import synth_state_dict

pytestmark = pytest.mark.gpu

GRAD_TOL = 5e-3          # per tensor, relative to its largest |gradient|. Measured against a float64 run of the reference: <= 8e-4 in layers
                         # 1-3 (where the reference's own fp32 run is off by up to 3.7e-3: conv / BatchNorm gradients are sums over 60 900 rows that
                         # cancel 10^4-fold), 2e-3 in the conv weights of layers 4-5, <= 1e-4 everywhere else (FC / heads / highway: 1e-6)
SAMPLE_ABOVE, SAMPLE_STRIDE = 200_000, 97


def _train_step(model, g):
    r, q, s, ref, rm, vm = (torch.from_numpy(np.ascontiguousarray(a)).long() for a in g["arrays"])      # trainer.py:119-127 hands CPU int64
    dummy = torch.zeros(r.shape[0], dtype=torch.long)
    out = model(r, ref, q, s, dummy, dummy, dummy, dummy, rm, vm)
    heads = torch.cat([o.reshape(o.shape[0], -1) for o in out[:6]], dim=1)
    (heads * torch.from_numpy(g["cw"]).to(heads.device)).sum().backward()
    return heads


def test_training_forward_and_gradients_match_reference_autograd():
    g = load_golden("train_smallfc")
    cfg = g["cfg"]
    model = build_model(cfg, synth_state_dict(cfg, seed=g["seed"]), precision="fp32").train()
    heads = _train_step(model, g)
    assert rel_err(heads.detach().cpu().numpy(), g["heads"]) < 1e-4, "training-mode forward (batch-statistics BatchNorm)"
    checked, errs, e64 = 0, {}, {}
    for name, p in model.named_parameters():
        key = "grad:" + name
        if key not in g:
            assert p.grad is None or name in ("bin_output_weights", "vt_output_weights"), name
            continue
        assert p.grad is not None, f"no gradient for {name}"
        got = p.grad.detach().cpu().numpy().astype(np.float32)
        want = g[key]
        if got.size > SAMPLE_ABOVE:
            got = got.reshape(-1)[::SAMPLE_STRIDE]
        assert got.shape == want.shape, name
        scale = np.abs(want).max()
        err = np.abs(got - want).max() / max(scale, 1e-20)
        errs[name] = float(err)
        truth = g["grad64:" + name]                                   # the same reference step in float64
        e64[name] = (float(np.abs(got - truth).max() / max(np.abs(truth).max(), 1e-20)), float(np.abs(want - truth).max() / max(np.abs(truth).max(), 1e-20)))
        checked += 1
    print("gradient error vs float64 reference, ours | reference fp32:", {k: f"{a:.1e} | {b:.1e}" for k, (a, b) in e64.items() if "conv1D_layers" in k or "bn1D" in k or "emb" in k})
    bad = {k: v for k, v in errs.items() if not v < GRAD_TOL}
    assert not bad, f"gradients off: {bad}"
    assert checked == sum(k.startswith("grad:") for k in g)
    # running statistics after one training step (momentum 0.1, unbiased variance); num_batches_tracked counts like nn.BatchNorm2d
    for name, b in model.named_buffers():
        if "running_" in name:
            np.testing.assert_allclose(b.detach().cpu().numpy(), g["buf:" + name], rtol=2e-5, atol=1e-6, err_msg=name)
    assert int(model.bn1D_layers[0].num_batches_tracked) == 101
    # the eval path picks up the moved running statistics (packed-weight cache keyed on the buffers too)
    model.eval()
    with torch.no_grad():
        r, q, s, ref, rm, vm = (torch.from_numpy(np.ascontiguousarray(a)) for a in g["arrays"])
        after = model.forward_heads(r, ref, q, s, rm, vm).cpu().numpy()
    fresh = build_model(cfg, {k: v.detach().cpu() for k, v in model.state_dict().items()}, precision="fp32")
    assert np.array_equal(after, fresh.forward_heads(r, ref, q, s, rm, vm).cpu().numpy())


def test_training_step_with_dropout_and_read_removal_runs_and_is_reproducible():
    """hidden_dropout 0.1 (PROD) + the trainer's read-removal augmentation (trainer.py:175-192): same torch seed -> same masks, same
    removed reads -> bitwise equal outputs and gradients; a different seed changes them; removed reads change the outputs."""
    from dl4vc_b200.config import small_config
    cfg = small_config()
    sd = synth_state_dict(cfg, seed=9)
    batch = make_pileups(4, seed=17, coverage="poisson")
    r, q, s, ref, rm, vm = (torch.from_numpy(np.ascontiguousarray(a)).long() for a in batch.arrays())
    dummy = torch.zeros(4, dtype=torch.long)

    def run(seed, rm_var):
        model = build_model(cfg, sd, precision="fp32").train()
        torch.manual_seed(seed)
        out = model(r, ref, q, s, dummy, dummy, dummy, dummy, rm, vm, rm_non_var_reads=0, rm_var_reads=rm_var)
        loss = out[0].sum() + out[1].pow(2).sum()
        loss.backward()
        return torch.cat(out[:2], dim=1).detach().cpu().numpy(), model.conv1D_layers[3].weight.grad.cpu().numpy(), model.embeddings.weight.grad.cpu().numpy()

    a, b, c, d = run(1, 0), run(1, 0), run(2, 0), run(1, 2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert not np.array_equal(a[0], c[0])
    assert not np.array_equal(a[0], d[0])
    assert np.isfinite(a[1]).all() and np.abs(a[1]).max() > 0 and np.all(a[2][0] == 0)      # padding_idx row gets no gradient


def test_device_losses_match_reference_loss_block():
    """dan_losses (one kernel: focal soft-BCE x2, allele-frequency BCE, coverage MSE, two weighted cross entropies, weighted total, gradient
    w.r.t. the model outputs, close flags) against the REAL reference's objectives.py + trainer.py loss lines and torch autograd
    (tests/golden/losses.npz, oracle/make_loss_goldens.py)."""
    import os
    from conftest import GOLDEN_DIR
    from dl4vc_b200.losses import LossConfig, fused_losses, update_close_table
    g = np.load(os.path.join(GOLDEN_DIR, "losses.npz"))
    cfg = LossConfig(**{k[4:]: float(g[k]) for k in g.files if k.startswith("cfg_")})
    heads = torch.tensor(g["heads"], device="cuda", requires_grad=True)
    t = lambda k: torch.from_numpy(g[k]).cuda()
    total, comps, close_vt, close_bin = fused_losses(heads, t("target_binary"), t("target_var_type"), t("target_allele_freq"), t("target_coverage"),
                                                     t("target_var_base"), t("target_ref_base"), t("example_weight"), cfg)
    (total * 1.0).backward()
    np.testing.assert_allclose(comps.cpu().numpy(), g["losses"], rtol=2e-5, atol=1e-6)
    assert abs(total.item() - g["losses"][6]) < 2e-5 * abs(g["losses"][6])
    want = g["dheads"]
    assert np.abs(heads.grad.cpu().numpy() - want).max() < 2e-5 * np.abs(want).max()
    assert np.array_equal(close_vt.cpu().numpy(), g["close_vt"]) and np.array_equal(close_bin.cpu().numpy(), g["close_bin"])
    table = torch.zeros(200, dtype=torch.uint8, device="cuda")
    idx = torch.arange(len(want), device="cuda") * 2
    update_close_table(table, idx, close_vt)
    assert int(table.sum()) == int(g["close_vt"].sum()) and bool(table[::2][: len(want)].cpu().eq(torch.from_numpy(g["close_vt"])).all())
