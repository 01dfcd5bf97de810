"""ORACLE — torch-CPU restatement of the DL4VC DAN forward, op for op.  TEST INFRASTRUCTURE ONLY.

Purpose: the CPU *baseline* of bench.py (`cpu_baseline`, `--impl reference`). The reference itself is pure PyTorch
(dl4vc/model.py) and cannot travel to the GPU box (/root/reference is not mounted there), so this file issues the
same sequence of torch library calls the reference issues on its own CPU path — nn.functional.embedding, conv2d
(MKLDNN), batch_norm (eval), avg/max_pool2d, linear — for the configurations of dl4vc_b200.config.DanConfig.
Because the library kernels are the same, its CPU throughput equals the reference's (checked in the build container,
BASELINE.md §2: 11.7 candidates/s PROD on 8 threads) and its results equal the reference's goldens
(tests/test_oracle_golden.py::test_torch_cpu_port_matches_reference_goldens).

Only tests/ and bench.py's CPU-baseline legs may import this file; the product path never does.

Line references: dl4vc/model.py:450-627 (encode), :719-778 (conv stack), :824-861 (pool / highway),
:911-958 (FC / heads).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from dl4vc_b200.config import BN_EPS, DanConfig, HEAD_NAMES, Q_SCORE_SCALE_FACTOR, STRAND_ENCODE_FACTOR


def _t(x, dtype=None):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
    return t if dtype is None else t.to(dtype)


@torch.no_grad()
def forward(cfg: DanConfig, sd, reads, ref, q_scores, strands, ref_masks, var_masks) -> torch.Tensor:
    """(B, 27) fp32 head matrix [xbinary|xVT|sigmoid(xAF)|leaky_relu(xCov)|xVB|xVR]."""
    reads = _t(reads, torch.int64); ref = _t(ref, torch.int64)                       # trainer.py:520-528 (.long())
    E = _t(sd["embeddings.weight"]).float(); pe = _t(sd["pe"]).float()
    B, P, R = reads.shape
    # ---- encode (model.py:450-627) ----
    reads_emb = F.embedding(reads, E, padding_idx=0) + pe.unsqueeze(2)               # (B,P,R,D)
    ref_emb = (F.embedding(ref, E, padding_idx=0) + pe).unsqueeze(2).expand(-1, -1, R, -1)
    x = torch.cat((reads_emb, ref_emb), dim=3)
    if cfg.use_q_scores:
        x = torch.cat((x, (_t(q_scores).float() * Q_SCORE_SCALE_FACTOR).unsqueeze(3)), dim=3)
    if cfg.use_strands:
        x = torch.cat((x, (_t(strands).float() * STRAND_ENCODE_FACTOR).unsqueeze(3)), dim=3)
    if cfg.use_reads_ref_var_mask:
        chans = []
        nzs = []
        for m in (ref_masks, var_masks):
            m = _t(m, torch.int64).unsqueeze(2)                                      # (B,P,1)
            nz = (m != 0).long()
            agree = ((reads * nz) == m).long().sum(dim=1) == P                       # (B,R)  model.py:592-593
            chans.append((nz * agree.unsqueeze(1).long()).float())
            nzs.append(nz)
        chans.append(nzs[0].expand(-1, -1, R).float())                               # var_len from the REF mask (sic)
        x = torch.cat([x] + [c.unsqueeze(3) for c in chans], dim=3)
    h = x.permute(0, 3, 2, 1)                                                        # (B,Cin,R,P) view, model.py:719
    # ---- conv stack (model.py:728-778) ----
    pool = None
    hw = []
    res_i = 0
    for l in range(1, cfg.total_conv_layers + 1):
        resid = h
        if pool is not None:
            h = h + pool
            pool = None
        d = cfg.dilation(l)
        h = F.relu(F.conv2d(h, _t(sd[f"conv1D_layers.{l-1}.weight"]), _t(sd[f"conv1D_layers.{l-1}.bias"]),
                            padding=(0, d), dilation=d))
        if cfg.use_batchnorm:
            p = f"bn1D_layers.{l-1}."
            h = F.batch_norm(h, _t(sd[p + "running_mean"]), _t(sd[p + "running_var"]), _t(sd[p + "weight"]),
                             _t(sd[p + "bias"]), training=False, eps=BN_EPS)
        if cfg.is_residual(l):
            h = F.conv2d(h, _t(sd[f"residual_conv_layers.{res_i}.weight"]), _t(sd[f"residual_conv_layers.{res_i}.bias"])) + resid
            res_i += 1
        if l in cfg.conv_1d_pool_layers:
            pool = F.avg_pool2d(h, kernel_size=(R, 1), ceil_mode=True)
        if cfg.highway:
            t = F.relu(F.conv2d(h, _t(sd[f"conv1D_bottleneck_layers.{l-1}.weight"]), _t(sd[f"conv1D_bottleneck_layers.{l-1}.bias"])))
            o = F.conv2d(t, _t(sd[f"conv1D_compression_layers.{l-1}.weight"]), _t(sd[f"conv1D_compression_layers.{l-1}.bias"]))
            hw.append(o.reshape(B, -1))
    # ---- read-axis pooling (model.py:824-844) ----
    mean = F.avg_pool2d(h, kernel_size=(R, 1), ceil_mode=True)
    if cfg.skip_final_maxpool:
        pooled = mean.reshape(B, -1)
    else:
        pooled = torch.cat((F.max_pool2d(h, kernel_size=(R, 1), ceil_mode=True), mean), dim=1).reshape(B, -1)
    if cfg.pool_combine_dimension > 0:
        pooled = F.relu(F.linear(pooled, _t(sd["post_pool_conv1D.weight"]), _t(sd["post_pool_conv1D.bias"])))
    if cfg.highway:
        hwv = torch.cat(hw, dim=1) if cfg.concat_hw_reads else torch.stack(hw, 0).mean(0)
        pooled = torch.cat((pooled, F.relu(hwv)), dim=1)
    # ---- FC trunk + heads (model.py:911-958); Dropout is identity in eval ----
    xh = pooled
    for idx in cfg.fc_indices:
        xh = F.relu(F.linear(xh, _t(sd[f"conv2hidden.{idx}.weight"]), _t(sd[f"conv2hidden.{idx}.bias"])))
    outs = [F.linear(xh, _t(sd[n + ".weight"]), _t(sd[n + ".bias"])) for n in HEAD_NAMES]
    outs[2] = torch.sigmoid(outs[2])
    outs[3] = F.leaky_relu(outs[3], 0.01)
    return torch.cat(outs, dim=1)
