"""Shape/config description of the DAN ("Basic2DNet") forward path.

One frozen dataclass carries every knob of the reference constructor that changes the arithmetic of the
hot path (reference: dl4vc/model.py:35-53, flag sets: call_variants.sh:101-147 and arguments.py:106-124).
It is shared by the drop-in module (dl4vc_b200/model.py), the C-ABI packer, the oracle and the benchmarks,
so that all of them agree on names, shapes and flattening orders.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import List, Tuple

# Token vocabulary (reference: dl4vc/base_enum.py:7-13): 0 pad, 1 A, 2 T, 3 G, 4 C, 5 '-', 6 start, 7 end,
# 8 'noinsert', 9 unknown.
VOCAB_SIZE = 10
# reference: dl4vc/model.py:16,24,25 and dl4vc/dataset.py:398
STRAND_ENCODE_FACTOR = 0.5
Q_SCORE_SCALE_FACTOR = 1.0 / 100.0
SINGLE_READ_LENGTH = 201
MAX_READS = 100
BN_EPS = 1e-5
# Head order in the fused 27-wide head matrix: xbinary(2) xVT(3) xAF(1) xCov(1) xVB(10) xVR(10)
HEAD_NAMES = ("fcHidden2BinTarget", "fcHidden2VT", "fcHidden2AF", "fcHidden2Coverage", "fcHidden2VB", "fcHidden2VR")
HEAD_SIZES = (2, 3, 1, 1, VOCAB_SIZE, VOCAB_SIZE)
NUM_HEAD_OUTPUTS = sum(HEAD_SIZES)  # 27


@dataclass(frozen=True)
class DanConfig:
    total_conv_layers: int = 7
    channels: int = 128                      # init_conv_channels == final_conv_channels
    embed_dim: int = 20
    use_q_scores: bool = True
    use_strands: bool = True
    use_reads_ref_var_mask: bool = True
    middle_layer_dilation: int = 2
    final_layer_dilation: int = 2
    use_batchnorm: bool = True
    residual_layer_start: int = 5            # 0 = no residual layers
    conv_1d_pool_layers: Tuple[int, ...] = (2,)
    highway: bool = True                     # append_bottleneck_highway_reads
    bottleneck: int = 32                     # bottleneck_channels == bottleneck_linear_outputs
    concat_hw_reads: bool = True
    pool_combine_dimension: int = 0
    skip_final_maxpool: bool = False
    layer_sizes: Tuple[int, ...] = (1024, 256)
    hidden_dropout: float = 0.1
    num_reads: int = MAX_READS
    read_len: int = SINGLE_READ_LENGTH

    # ---- derived shapes ------------------------------------------------------------------------
    @property
    def in_channels(self) -> int:
        """reference: dl4vc/model.py:169-178 (ref_concat_at_reads=True, no reads_sum)"""
        return 2 * self.embed_dim + int(self.use_q_scores) + int(self.use_strands) + (3 if self.use_reads_ref_var_mask else 0)

    def dilation(self, layer: int) -> int:
        """layer is 1-based. reference: dl4vc/model.py:213-229"""
        if layer == 1:
            return 1
        return self.middle_layer_dilation if layer < self.total_conv_layers else self.final_layer_dilation

    def is_residual(self, layer: int) -> bool:
        """reference: dl4vc/model.py:246"""
        return self.residual_layer_start > 0 and layer >= self.residual_layer_start

    @property
    def num_residual_layers(self) -> int:
        return sum(self.is_residual(l) for l in range(1, self.total_conv_layers + 1))

    @property
    def pooled_features(self) -> int:
        """reference: dl4vc/model.py:296"""
        return (1 if self.skip_final_maxpool else 2) * self.channels * self.read_len

    @property
    def highway_features(self) -> int:
        """reference: dl4vc/model.py:336"""
        if not self.highway:
            return 0
        return (self.total_conv_layers if self.concat_hw_reads else 1) * self.bottleneck * self.num_reads

    @property
    def fc_in_features(self) -> int:
        """reference: dl4vc/model.py:327,338"""
        base = self.pool_combine_dimension if self.pool_combine_dimension > 0 else self.pooled_features
        return base + self.highway_features

    @property
    def fc_indices(self) -> List[int]:
        """Indices of the nn.Linear modules inside the `conv2hidden` Sequential (reference: dl4vc/model.py:369-377)."""
        first = 1 if self.hidden_dropout else 0
        return [first + 3 * i for i in range(len(self.layer_sizes))]

    def macs_per_candidate(self) -> int:
        """Dense multiply-accumulates of one forward (matches SURVEY App. E for PROD: 8 059 296 000)."""
        R, P, C = self.num_reads, self.read_len, self.channels
        total = 0
        cin = self.in_channels
        for l in range(1, self.total_conv_layers + 1):
            total += R * P * C * cin * 3
            cin = C
            if self.is_residual(l):
                total += R * P * C * C
            if self.highway:
                total += R * P * C * self.bottleneck
                total += R * self.bottleneck * self.bottleneck * P
        feat = self.pooled_features
        if self.pool_combine_dimension > 0:
            total += feat * self.pool_combine_dimension
        sizes = [self.fc_in_features] + list(self.layer_sizes)
        for a, b in zip(sizes[:-1], sizes[1:]):
            total += a * b
        total += sizes[-1] * NUM_HEAD_OUTPUTS
        return total

    def to_dict(self):
        d = asdict(self)
        d["conv_1d_pool_layers"] = list(self.conv_1d_pool_layers)
        d["layer_sizes"] = list(self.layer_sizes)
        return d


def prod_config(**over) -> DanConfig:
    """The shipped model: flag set of call_variants.sh:101-147 (column PROD)."""
    return DanConfig(**over)


def min_config(**over) -> DanConfig:
    """argparse defaults (arguments.py:106-124; SURVEY App. D, column MIN)."""
    base = dict(total_conv_layers=5, use_q_scores=False, use_strands=False, use_reads_ref_var_mask=False,
                middle_layer_dilation=1, final_layer_dilation=1, use_batchnorm=False, residual_layer_start=0,
                highway=False, bottleneck=32, concat_hw_reads=False, pool_combine_dimension=2048,
                hidden_dropout=0.0)
    base.update(over)
    return DanConfig(**base)


def small_config(**over) -> DanConfig:
    """PROD topology with a small FC trunk, for cheap parity cases (layer_sizes is a reference ctor argument)."""
    base = dict(layer_sizes=(64, 32))
    base.update(over)
    return DanConfig(**base)


# ---- state_dict layout ---------------------------------------------------------------
def state_dict_spec(cfg: DanConfig):
    """Ordered list of (name, shape, kind) for every tensor of the reference state_dict.

    kind in {"param", "buffer", "counter"}. Order follows the registration order of the reference
    constructor (dl4vc/model.py:143-431) so that `list(model.state_dict())` compares equal.
    """
    C, L = cfg.channels, cfg.total_conv_layers
    spec = []
    spec.append(("bin_output_weights", (1,), "param"))
    spec.append(("vt_output_weights", (1,), "param"))
    spec.append(("pe", (1, cfg.read_len, cfg.embed_dim), "buffer"))
    spec.append(("embeddings.weight", (VOCAB_SIZE, cfg.embed_dim), "param"))
    for l in range(L):
        cin = cfg.in_channels if l == 0 else C
        spec.append((f"conv1D_layers.{l}.weight", (C, cin, 1, 3), "param"))
        spec.append((f"conv1D_layers.{l}.bias", (C,), "param"))
    for l in range(L):
        spec.append((f"bn1D_layers.{l}.weight", (C,), "param"))
        spec.append((f"bn1D_layers.{l}.bias", (C,), "param"))
        spec.append((f"bn1D_layers.{l}.running_mean", (C,), "buffer"))
        spec.append((f"bn1D_layers.{l}.running_var", (C,), "buffer"))
        spec.append((f"bn1D_layers.{l}.num_batches_tracked", (), "counter"))
    if cfg.highway:
        for l in range(L):
            spec.append((f"conv1D_bottleneck_layers.{l}.weight", (cfg.bottleneck, C, 1, 1), "param"))
            spec.append((f"conv1D_bottleneck_layers.{l}.bias", (cfg.bottleneck,), "param"))
        for l in range(L):
            spec.append((f"conv1D_compression_layers.{l}.weight", (cfg.bottleneck, cfg.bottleneck, 1, cfg.read_len), "param"))
            spec.append((f"conv1D_compression_layers.{l}.bias", (cfg.bottleneck,), "param"))
    for i in range(cfg.num_residual_layers):
        spec.append((f"residual_conv_layers.{i}.weight", (C, C, 1, 1), "param"))
        spec.append((f"residual_conv_layers.{i}.bias", (C,), "param"))
    if cfg.pool_combine_dimension > 0:
        spec.append(("post_pool_conv1D.weight", (cfg.pool_combine_dimension, cfg.pooled_features), "param"))
        spec.append(("post_pool_conv1D.bias", (cfg.pool_combine_dimension,), "param"))
    sizes = [cfg.fc_in_features] + list(cfg.layer_sizes)
    for idx, (a, b) in zip(cfg.fc_indices, zip(sizes[:-1], sizes[1:])):
        spec.append((f"conv2hidden.{idx}.weight", (b, a), "param"))
        spec.append((f"conv2hidden.{idx}.bias", (b,), "param"))
    for name, n in zip(HEAD_NAMES, HEAD_SIZES):
        spec.append((f"{name}.weight", (n, sizes[-1]), "param"))
        spec.append((f"{name}.bias", (n,), "param"))
    return spec
