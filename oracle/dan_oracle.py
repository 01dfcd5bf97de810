"""ORACLE — CPU restatement (numpy, fp32) of the DL4VC DAN forward pass.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product path (dl4vc_b200/) never does and fails loudly without its CUDA library.

Restates, function by function, what /root/reference/dl4vc/model.py computes for the configurations described by
dl4vc_b200.config.DanConfig (PROD = call_variants.sh:101-147 flag set):
    encode()        dl4vc/model.py:450-451, 463-470, 501-519, 534-561, 576-627, 719
    conv_stack()    dl4vc/model.py:728-778
    pool_and_fc()   dl4vc/model.py:824-861, 911-921, 953-958
The arithmetic lives in a third-party dependency of the reference (torch==1.2.0, requirements.txt:9: Embedding,
Conv2d with zero padding / dilation, BatchNorm2d eval eps=1e-5, Avg/MaxPool2d, Linear); this file restates their
published semantics. Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the REAL reference executed in the build container (oracle/make_goldens.py →
tests/golden/*.npz; tests/test_oracle_golden.py) — encoder bit-exact, heads <= 1e-6 abs.
"""
from __future__ import annotations

import numpy as np

from dl4vc_b200.config import (DanConfig, HEAD_NAMES, HEAD_SIZES, Q_SCORE_SCALE_FACTOR, STRAND_ENCODE_FACTOR, BN_EPS)

f32 = np.float32


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def match_masks(reads, ref_masks, var_masks):
    """Integer part of model.py:576-623. reads (B,P,R) ; masks (B,P).
    Returns agreeR, agreeV (B,R) bool and nzR, nzV (B,P) bool."""
    reads = _np(reads).astype(np.int64)
    out = []
    for m in (ref_masks, var_masks):
        m = _np(m).astype(np.int64)[:, :, None]                 # (B,P,1)
        nz = (m != 0)
        agree = ((reads * nz) == m).sum(axis=1) == reads.shape[1]   # model.py:592-593
        out.append((agree, nz[:, :, 0]))
    (aR, nzR), (aV, nzV) = out
    return aR, aV, nzR, nzV


def encode(cfg: DanConfig, sd, reads, ref, q_scores, strands, ref_masks, var_masks):
    """45-channel conv-1 input, returned as (B, Cin, R, P) fp32 (the view produced by model.py:719)."""
    reads = _np(reads).astype(np.int64); ref = _np(ref).astype(np.int64)
    E = _np(sd["embeddings.weight"]).astype(f32)
    pe = _np(sd["pe"]).astype(f32)[0]                           # (P, D)
    B, P, R = reads.shape
    chans = []
    reads_emb = E[reads] + pe[None, :, None, :]                 # (B,P,R,D)  model.py:450,506
    ref_emb = (E[ref] + pe[None, :, :])[:, :, None, :]          # (B,P,1,D)  model.py:451,502-507
    chans.append(reads_emb)
    chans.append(np.broadcast_to(ref_emb, reads_emb.shape))
    if cfg.use_q_scores:                                        # model.py:536
        chans.append((_np(q_scores).astype(f32) * f32(Q_SCORE_SCALE_FACTOR))[..., None])
    if cfg.use_strands:                                         # model.py:551
        chans.append((_np(strands).astype(f32) * f32(STRAND_ENCODE_FACTOR))[..., None])
    if cfg.use_reads_ref_var_mask:                              # model.py:576-625
        aR, aV, nzR, nzV = match_masks(reads, ref_masks, var_masks)
        ref_match = (nzR[:, :, None] & aR[:, None, :]).astype(f32)
        var_match = (nzV[:, :, None] & aV[:, None, :]).astype(f32)
        var_len = np.broadcast_to(nzR[:, :, None], ref_match.shape).astype(f32)   # ref mask, sic (model.py:579,584)
        chans += [ref_match[..., None], var_match[..., None], var_len[..., None]]
    x = np.concatenate(chans, axis=3)                           # (B,P,R,Cin)
    return np.ascontiguousarray(x.transpose(0, 3, 2, 1))        # (B,Cin,R,P)


def conv1xk(x, w, b, dilation):
    """Per-read 1-D convolution over positions: Conv2d kernel (1,k), zero padding (0,dilation) (model.py:214-229)."""
    B, Ci, R, P = x.shape
    Co, _, _, k = w.shape
    pad = dilation * (k // 2)
    xp = np.zeros((B, Ci, R, P + 2 * pad), f32)
    xp[..., pad:pad + P] = x
    out = np.zeros((B, Co, R * P), f32)
    for t in range(k):
        xs = np.ascontiguousarray(xp[..., t * dilation:t * dilation + P]).reshape(B, Ci, R * P)
        out += np.matmul(w[:, :, 0, t].astype(f32), xs)
    out = out.reshape(B, Co, R, P)
    if b is not None:
        out += b.astype(f32)[None, :, None, None]
    return out


def conv1x1(x, w, b):
    B, Ci, R, P = x.shape
    out = np.matmul(w[:, :, 0, 0].astype(f32), x.reshape(B, Ci, R * P)).reshape(B, w.shape[0], R, P)
    return out + b.astype(f32)[None, :, None, None]


def batchnorm_eval(x, sd, prefix):
    """BatchNorm2d in eval mode (model.py:750-751); same operation order as ATen's CPU kernel:
    (x - mean) * (1/sqrt(var+eps)) * gamma + beta, folded to scale/shift like ATen does."""
    g = _np(sd[prefix + ".weight"]).astype(f32); b = _np(sd[prefix + ".bias"]).astype(f32)
    m = _np(sd[prefix + ".running_mean"]).astype(f32); v = _np(sd[prefix + ".running_var"]).astype(f32)
    inv = f32(1.0) / np.sqrt(v + f32(BN_EPS))
    scale = g * inv
    shift = b - m * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def conv_stack(cfg: DanConfig, sd, x0, keep=False):
    """model.py:728-778. Returns final activations (B,C,R,P), list of highway outputs (B, bott*R) and, if keep,
    every layer's output."""
    h = x0
    pool = None
    highway, layers = [], []
    res_idx = 0
    for l in range(1, cfg.total_conv_layers + 1):
        resid = h                                                        # model.py:732 (before the pool add)
        if (l - 1) in cfg.conv_1d_pool_layers:                           # model.py:734-742
            h = h + pool
        W = _np(sd[f"conv1D_layers.{l-1}.weight"]); b = _np(sd[f"conv1D_layers.{l-1}.bias"])
        h = np.maximum(conv1xk(h, W, b, cfg.dilation(l)), 0)             # model.py:749
        if cfg.use_batchnorm:
            h = batchnorm_eval(h, sd, f"bn1D_layers.{l-1}")              # model.py:751
        if cfg.is_residual(l):                                           # model.py:753-761
            h = conv1x1(h, _np(sd[f"residual_conv_layers.{res_idx}.weight"]), _np(sd[f"residual_conv_layers.{res_idx}.bias"]))
            h = h + resid
            res_idx += 1
        if l in cfg.conv_1d_pool_layers:                                 # model.py:766-772, unmasked mean over reads
            pool = h.mean(axis=2, keepdims=True, dtype=f32)
        if cfg.highway:                                                  # model.py:773-778
            t = np.maximum(conv1x1(h, _np(sd[f"conv1D_bottleneck_layers.{l-1}.weight"]),
                                   _np(sd[f"conv1D_bottleneck_layers.{l-1}.bias"])), 0)
            Wc = _np(sd[f"conv1D_compression_layers.{l-1}.weight"]).astype(f32)[:, :, 0, :]   # (O, Cb, P)
            hw = np.einsum("bcrp,ocp->bor", t, Wc, optimize=True).astype(f32)
            hw = hw + _np(sd[f"conv1D_compression_layers.{l-1}.bias"]).astype(f32)[None, :, None]
            highway.append(hw.reshape(hw.shape[0], -1))                  # index o*R + r
        if keep:
            layers.append(h)
    return h, highway, layers


def pool_and_fc(cfg: DanConfig, sd, h, highway):
    """model.py:824-861, 911-921, 953-958. Returns the six head outputs and the hidden vector."""
    B = h.shape[0]
    avg = h.mean(axis=2, dtype=f32)                                      # (B,C,P)
    if cfg.skip_final_maxpool:
        xc = avg.reshape(B, -1)
    else:
        xc = np.concatenate([h.max(axis=2), avg], axis=1).reshape(B, -1)  # max block first, index c*P+p
    if cfg.pool_combine_dimension > 0:                                   # model.py:841-843
        xc = np.maximum(xc @ _np(sd["post_pool_conv1D.weight"]).T + _np(sd["post_pool_conv1D.bias"]), 0)
    if cfg.highway:                                                      # model.py:853-859, 912
        hw = np.concatenate(highway, axis=1) if cfg.concat_hw_reads else (sum(highway) / f32(len(highway)))
        xc = np.concatenate([xc, np.maximum(hw, 0)], axis=1)
    x = xc.astype(f32)
    for idx in cfg.fc_indices:                                           # model.py:917 (dropout = identity in eval)
        x = np.maximum(x @ _np(sd[f"conv2hidden.{idx}.weight"]).T.astype(f32) + _np(sd[f"conv2hidden.{idx}.bias"]), 0).astype(f32)
    outs = []
    for name in HEAD_NAMES:
        outs.append((x @ _np(sd[name + ".weight"]).T + _np(sd[name + ".bias"])).astype(f32))
    outs[2] = (f32(1) / (f32(1) + np.exp(-outs[2]))).astype(f32)         # sigmoid, model.py:954
    outs[3] = np.where(outs[3] >= 0, outs[3], f32(0.01) * outs[3]).astype(f32)   # leaky_relu, model.py:956
    return outs, x, xc


def forward(cfg: DanConfig, sd, reads, ref, q_scores, strands, ref_masks, var_masks, keep=False):
    """Full forward. Returns dict(heads=(B,27) fp32 [xbinary|xVT|xAF|xCov|xVB|xVR], plus intermediates if keep)."""
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
    x0 = encode(cfg, sd, reads, ref, q_scores, strands, ref_masks, var_masks)
    h, highway, layers = conv_stack(cfg, sd, x0, keep=keep)
    outs, hidden, fc_in = pool_and_fc(cfg, sd, h, highway)
    res = {"heads": np.concatenate(outs, axis=1).astype(f32)}
    if keep:
        res.update(x0=x0, layers=layers, highway=highway, hidden=hidden, fc_in=fc_in)
    return res


def forward_chunked(cfg, sd, arrays, chunk=4, **kw):
    reads = arrays[0]
    outs = []
    for lo in range(0, reads.shape[0], chunk):
        r, q, s, ref, rm, vm = (a[lo:lo + chunk] for a in arrays)
        outs.append(forward(cfg, sd, r, ref, q, s, rm, vm, **kw)["heads"])
    return np.concatenate(outs, axis=0) if outs else np.zeros((0, sum(HEAD_SIZES)), f32)
