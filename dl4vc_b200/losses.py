"""The trainer's loss block on the device: one kernel (dan_losses, include/dan_b200.h) computes the focal soft-BCE of the
binary and genotype heads, the auxiliary losses, their weighted total, d(total)/d(heads) and the "close example" flags that drive the
easy-example down-sampling — what dl4vc/trainer.py:252-255,309-313,426-427 and dl4vc/objectives.py:49-112 compute with ~60 small torch ops and
three host round trips per step (trainer.py:258,263,267)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib

LOSS_NAMES = ("binary", "genotype", "allele_freq", "coverage", "var_base", "ref_base", "total", "num_close")


@dataclass(frozen=True)
class LossConfig:
    """Defaults = the flag set of train_variant_caller.sh:104-131."""
    label_smoothing: float = 0.001
    close_match_window: float = 2.0
    focal_gamma: float = 0.2
    focal_alpha: float = 1.0
    fp_train_weight: float = 0.2
    binary_weight: float = 1.0
    aux_weight: float = 1.0
    aux_allele_weight: float = 0.001
    aux_bases_weight: float = 0.01


class DanLossConfigC(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("label_smoothing", "close_match_window", "focal_gamma", "focal_alpha", "fp_train_weight", "binary_weight",
                                         "aux_weight", "aux_allele_weight", "aux_bases_weight")]


def _bind(lib):
    if getattr(lib, "_losses_bound", False):
        return lib
    vp, i32 = C.c_void_p, C.c_int
    lib.dan_losses.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(DanLossConfigC), vp, vp, vp, vp, vp]
    lib.dan_close_table_update.argtypes = [vp, C.c_int64, vp, vp, i32, vp]
    lib._losses_bound = True
    return lib


class _FusedLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heads, target_binary, target_var_type, target_allele_freq, target_coverage, target_var_base, target_ref_base, example_weight, cfg):
        if heads.device.type != "cuda":
            raise RuntimeError("dan_losses runs on the device the heads live on (no CPU path)")
        lib = _bind(_lib.load_library())
        dev = heads.device
        B = int(heads.shape[0])
        h = heads.detach().contiguous().float()
        i32 = lambda t: t.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
        f32 = lambda t: t.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        tb, tv, vb, vr = i32(target_binary), i32(target_var_type), i32(target_var_base), i32(target_ref_base)
        af, cov = f32(target_allele_freq), f32(target_coverage)
        w = None if example_weight is None else f32(example_weight)
        losses = torch.empty(8, dtype=torch.float32, device=dev)
        dheads = torch.empty_like(h)
        close_vt = torch.empty(B, dtype=torch.uint8, device=dev)
        close_bin = torch.empty(B, dtype=torch.uint8, device=dev)
        c = DanLossConfigC(*[float(getattr(cfg, n)) for n, _ in DanLossConfigC._fields_])
        with torch.cuda.device(dev):
            _lib.check(lib.dan_losses(h.data_ptr(), B, tb.data_ptr(), tv.data_ptr(), af.data_ptr(), cov.data_ptr(), vb.data_ptr(), vr.data_ptr(),
                                      None if w is None else w.data_ptr(), C.byref(c), losses.data_ptr(), dheads.data_ptr(), close_vt.data_ptr(),
                                      close_bin.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "dan_losses")
        ctx.save_for_backward(dheads)
        ctx.mark_non_differentiable(losses, close_vt, close_bin)
        return losses[6].clone(), losses, close_vt, close_bin

    @staticmethod
    def backward(ctx, g_total, g_losses, g_cv, g_cb):
        (dheads,) = ctx.saved_tensors
        return (dheads * g_total, None, None, None, None, None, None, None, None)


def fused_losses(heads, target_binary, target_var_type, target_allele_freq, target_coverage, target_var_base, target_ref_base,
                 example_weight=None, cfg: LossConfig = LossConfig()):
    """heads: (B, 27) as returned by the training forward ([xbinary|xVT|xAF|xCov|xVB|xVR]); targets as the trainer builds them
    (trainer.py:130-146; coverage already scaled by 0.01). Returns (total — autograd-connected to `heads` —, components (8,) tensor in
    LOSS_NAMES order, close_vt (B,) uint8, close_bin (B,) uint8); nothing leaves the device."""
    return _FusedLosses.apply(heads, target_binary, target_var_type, target_allele_freq, target_coverage, target_var_base, target_ref_base, example_weight, cfg)


def update_close_table(table: torch.Tensor, idx: torch.Tensor, flags: torch.Tensor):
    """table[idx[b]] = flags[b] on the device (trainer.py:263 -> dataset.py:480, without the .cpu() round trip)."""
    lib = _bind(_lib.load_library())
    assert table.dtype == torch.uint8 and table.is_cuda and table.is_contiguous()
    idx = idx.to(device=table.device, dtype=torch.int64).contiguous()
    flags = flags.to(device=table.device, dtype=torch.uint8).contiguous()
    with torch.cuda.device(table.device):
        _lib.check(lib.dan_close_table_update(table.data_ptr(), table.numel(), idx.data_ptr(), flags.data_ptr(), int(idx.numel()),
                                              torch.cuda.current_stream(table.device).cuda_stream), "dan_close_table_update")
    return table
