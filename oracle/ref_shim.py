"""ORACLE helper — import and run the REAL reference model (build container only).  TEST INFRASTRUCTURE ONLY.

/root/reference is mounted read-only in the build container and does not exist on the GPU box, so this module is
used solely by oracle/make_goldens.py and by container-side tests that skip when the path is absent.
Recipe: SURVEY App. F — stub the two absent imports (h5py, pysam), neutralise the unconditional `.cuda()` calls
(dl4vc/model.py:459,481,538,553,565,936,941) and build Basic2DNet exactly like main.py:99-112 does.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DL4VC_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dl4vc", "model.py"))


def cfg_to_flags(cfg):
    """DanConfig -> main.py command-line flags (arguments.py:106-124)."""
    f = ["--test_file", "x.hdf",
         "--model-conv-layers", str(cfg.total_conv_layers),
         "--model-residual-layer-start", str(cfg.residual_layer_start),
         "--model-init-conv-channels", str(cfg.channels), "--model-final-conv-channels", str(cfg.channels),
         "--model-bottleneck-size", str(cfg.bottleneck),
         "--model_middle_layer_dilation", str(cfg.middle_layer_dilation),
         "--model_final_layer_dilation", str(cfg.final_layer_dilation),
         "--model_pool_combine_dimension", str(cfg.pool_combine_dimension),
         "--model-hidden-dropout", str(cfg.hidden_dropout)]
    f += ["--model-ave-pool-layers"] + [str(x) for x in cfg.conv_1d_pool_layers]
    for flag, on in (("--model-batchnorm", cfg.use_batchnorm), ("--model-use-q-scores", cfg.use_q_scores),
                     ("--model-use-strands", cfg.use_strands), ("--model-use-reads-ref-var-mask", cfg.use_reads_ref_var_mask),
                     ("--model-highway-single-reads", cfg.highway), ("--model_concat_hw_reads", cfg.concat_hw_reads),
                     ("--model_skip_final_maxpool", cfg.skip_final_maxpool)):
        if on:
            f.append(flag)
    return f


def import_reference():
    import torch

    for n in ("h5py", "pysam"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # our repo also has a package called dl4vc_b200, never `dl4vc`, so the reference's `dl4vc` resolves cleanly
    from arguments import create_arg_parser          # reference: arguments.py:5
    from dl4vc.model import Basic2DNet               # reference: dl4vc/model.py:31
    return create_arg_parser, Basic2DNet


def import_reference_module(name: str):
    """Import another module of the reference tree (e.g. 'dl4vc.utils') under the same stubs."""
    import importlib

    import_reference()
    return importlib.import_module(name)


def build_reference_model(cfg, state_dict=None, quiet=True):
    """Construct the reference Basic2DNet the way main.py:99-112 does and optionally load weights."""
    create_arg_parser, Basic2DNet = import_reference()
    args = create_arg_parser().parse_args(cfg_to_flags(cfg))
    sink = io.StringIO()
    # A read axis other than 100 rows (BASELINE configs[3], up to 300): the reference sizes its pooling kernels from the module
    # constant MAX_READS and its FC input from num_single_reads (model.py:194,303-304,336) — SURVEY App. F recipe.
    import dl4vc.model as ref_model_module
    extra = {}
    if cfg.num_reads != 100:
        ref_model_module.MAX_READS = cfg.num_reads
        extra["num_single_reads"] = cfg.num_reads
    else:
        ref_model_module.MAX_READS = 100
    with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
        model = Basic2DNet(
            target_size=3, init_conv_channels=args.model_init_conv_channels,
            final_conv_channels=args.model_final_conv_channels, hidden_dropout=args.model_hidden_dropout,
            use_batchnorm=args.model_batchnorm, skip_final_maxpool=args.model_skip_final_maxpool,
            pool_combine_dimension=args.model_pool_combine_dimension, early_loss_layers=args.early_loss_layers,
            use_q_scores=args.model_use_q_scores, use_strands=args.model_use_strands,
            total_conv_layers=args.model_conv_layers, residual_layer_start=args.model_residual_layer_start,
            conv_1d_pool_layers=args.model_ave_pool_layers, final_layer_dilation=args.model_final_layer_dilation,
            middle_layer_dilation=args.model_middle_layer_dilation,
            append_bottleneck_highway_reads=args.model_highway_single_reads,
            bottleneck_channels=args.model_bottleneck_size, bottleneck_linear_outputs=args.model_bottleneck_size,
            concat_hw_reads=args.model_concat_hw_reads, use_naive_variant_encoding=args.model_use_naive_var_vector,
            use_reads_ref_var_mask=args.model_use_reads_ref_var_mask, append_allele_frequency=args.model_use_AF,
            layer_sizes=list(cfg.layer_sizes), args=args, **extra)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model.eval(), args


def reference_forward(model, batch, capture_conv_input=False):
    """Run the unmodified reference forward on a PileupBatch-like tuple of uint8 arrays.
    Inputs are cast to int64 like the trainer does (dl4vc/trainer.py:520-528)."""
    import torch

    reads, q, s, ref, rm, vm = (torch.from_numpy(a).long() for a in batch)
    captured = {}
    hook = None
    if capture_conv_input:
        hook = model.conv1D_layers[0].register_forward_pre_hook(lambda m, inp: captured.__setitem__("x0", inp[0].detach().clone()))
    with torch.no_grad():
        out = model(reads, ref, q_scores=q, strands=s, binary_trust_vector=None, af_scores=None,
                    ref_bases=None, var_bases=None, ref_masks=rm, var_masks=vm)
    if hook is not None:
        hook.remove()
    heads = torch.cat([o.reshape(o.shape[0], -1) for o in out[:6]], dim=1)
    return heads.numpy(), out, captured
