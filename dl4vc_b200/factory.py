"""Build a dl4vc_b200.Basic2DNet from a DanConfig with the keyword set main.py uses (reference main.py:99-112)."""
from __future__ import annotations

from .config import DanConfig


def ctor_kwargs(cfg: DanConfig) -> dict:
    return dict(
        target_size=3, layer_sizes=list(cfg.layer_sizes), init_conv_channels=cfg.channels, final_conv_channels=cfg.channels,
        hidden_dropout=cfg.hidden_dropout, use_batchnorm=cfg.use_batchnorm, skip_final_maxpool=cfg.skip_final_maxpool,
        pool_combine_dimension=cfg.pool_combine_dimension, early_loss_layers=[], use_q_scores=cfg.use_q_scores,
        use_strands=cfg.use_strands, total_conv_layers=cfg.total_conv_layers, residual_layer_start=cfg.residual_layer_start,
        conv_1d_pool_layers=list(cfg.conv_1d_pool_layers), final_layer_dilation=cfg.final_layer_dilation,
        middle_layer_dilation=cfg.middle_layer_dilation, append_bottleneck_highway_reads=cfg.highway,
        bottleneck_channels=cfg.bottleneck, bottleneck_linear_outputs=cfg.bottleneck, concat_hw_reads=cfg.concat_hw_reads,
        use_naive_variant_encoding=False, use_reads_ref_var_mask=cfg.use_reads_ref_var_mask, append_allele_frequency=False,
        embed_dim=cfg.embed_dim, num_single_reads=cfg.num_reads, single_read_len=cfg.read_len)


def build_model(cfg: DanConfig, state_dict=None, device="cuda", precision="bf16"):
    from .model import Basic2DNet

    model = Basic2DNet(**ctor_kwargs(cfg))
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model = model.eval()
    if device is not None:
        model = model.to(device)
    return model.set_precision(precision)
