"""Deterministic "checkpoint-shaped" weights for parity tests and benchmarks.

The shipped checkpoint is 311 MB of fp32 and only reachable over the network, so goldens cannot carry a
state_dict. Instead every tensor is regenerated from (seed, tensor name) with numpy's PCG64 streams, which makes
the same weights available to the real reference (via load_state_dict, in the build container), to the oracle
and to the CUDA path on the GPU box. Scales follow torch's default initialisers (kaiming-uniform bound
1/sqrt(fan_in)) so activations have trained-network magnitudes; BatchNorm statistics are made non-trivial and
the heads are scaled up so that logit gaps are large against bf16 noise.
"""
from __future__ import annotations

import math
import zlib

import numpy as np

from .config import DanConfig, state_dict_spec


def positional_encoding(read_len: int, embed_dim: int) -> np.ndarray:
    """Sinusoidal buffer 'pe' (reference: dl4vc/model.py:154-162), computed like torch does (fp32 ops)."""
    import torch

    pe = torch.zeros(read_len, embed_dim)
    position = torch.arange(0.0, read_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0.0, embed_dim, 2) * -(math.log(10000.0) / embed_dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).numpy().copy()


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def plant_tracer_channels(cfg: DanConfig, sd: dict, gain: float = 2.0):
    """Make the stack sensitive to the genotype evidence, the way a trained checkpoint is: output channels 0 / 1 of every non-residual
    conv layer carry the var-match / ref-match input channels (model.py:576-625) straight through (centre tap, zero bias; ReLU is the
    identity on these non-negative values, BatchNorm stays an affine map with the layer's own statistics), and the residual layers add
    their usual terms on top. The read-mean of those channels at the proposal columns is then an affine function of the share of reads
    that support the proposed / reference allele — the signal the classification heads of the scale golden are fitted on
    (oracle/make_scale_golden.py). numpy arrays in, modified in place."""
    assert cfg.use_reads_ref_var_mask
    first = 2 * cfg.embed_dim + int(cfg.use_q_scores) + int(cfg.use_strands)      # ref-match channel; var-match is first + 1
    for l in range(cfg.total_conv_layers):
        if cfg.is_residual(l + 1):
            break
        w, b = sd[f"conv1D_layers.{l}.weight"], sd[f"conv1D_layers.{l}.bias"]
        for out_ch, src in ((0, first + 1 if l == 0 else 0), (1, first if l == 0 else 1)):
            w[out_ch] = 0.0
            w[out_ch, src, 0, 1] = gain
            b[out_ch] = 0.0
    return sd


def synth_state_dict(cfg: DanConfig, seed: int = 1, head_gain: float = 8.0, as_torch: bool = True, prefix: str = "", tracer: bool = False):
    """name -> tensor for every entry of state_dict_spec(cfg). tracer: see plant_tracer_channels."""
    out = {}
    spec = state_dict_spec(cfg)
    shapes = {n: s for n, s, _ in spec}
    for name, shape, kind in spec:
        g = _rng(seed, name)
        if name == "pe":
            a = positional_encoding(cfg.read_len, cfg.embed_dim)
        elif kind == "counter":
            a = np.array(100, dtype=np.int64)
        elif name in ("bin_output_weights", "vt_output_weights"):
            a = np.full(shape, 0.1, dtype=np.float32)
        elif name == "embeddings.weight":
            a = g.standard_normal(shape, dtype=np.float32)
            a[0] = 0.0  # padding_idx row (reference: dl4vc/model.py:143-145)
        elif name.startswith("bn1D_layers"):
            leaf = name.rsplit(".", 1)[1]
            if leaf == "weight":
                a = g.uniform(0.6, 1.4, shape).astype(np.float32)
            elif leaf == "bias":
                a = (0.1 * g.standard_normal(shape)).astype(np.float32)
            elif leaf == "running_mean":
                a = g.uniform(0.0, 0.3, shape).astype(np.float32)
            else:  # running_var
                a = g.uniform(0.05, 0.6, shape).astype(np.float32)
        else:
            if name.endswith(".weight"):
                fan_in = int(np.prod(shape[1:]))
            else:  # bias: fan_in of the matching weight
                wshape = shapes[name[:-4] + "weight"]
                fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            gain = 1.0
            if name.startswith("fcHidden2"):
                gain = head_gain
            elif name.startswith("conv1D_layers") and name.endswith(".weight"):
                gain = math.sqrt(3.0)  # keep post-ReLU activations O(1) through 7 layers
            a = g.uniform(-bound * gain, bound * gain, shape).astype(np.float32)
        out[prefix + name] = a
    if tracer:
        assert not prefix
        plant_tracer_channels(cfg, out)
    if as_torch:
        import torch

        out = {k: (torch.from_numpy(np.ascontiguousarray(v)) if v.ndim else torch.tensor(int(v), dtype=torch.int64))
               for k, v in out.items()}
    return out
