"""Synthetic weights and inputs for tests, smoke and bench (no datasets or checkpoints exist here).

`make_state_dict` emits a checkpoint in the REFERENCE's key layout (the weight ABI,
FlashVTG/inference.py:471 `load_state_dict(ckpt["model"], strict=True)`) with each tensor drawn
from the distribution the reference's own initialisers use (model.py:15-23,116-117,
transformer.py:76-80, torch defaults elsewhere).  Values come from numpy's PCG64 so that the GPU
box regenerates bit-identical tensors from the seed; tests/golden pins that with checksums.

`make_inputs` follows SURVEY.md §8(d): per-group L2-normalised Gaussian clip features with TEF
appended (start_end_dataset.py:174-180,524-530), L2-normalised Gaussian query tokens.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .config import ModelConfig


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _normal(rng, shape, std=1.0):
    return torch.from_numpy((rng.standard_normal(size=shape) * std).astype(np.float32))


def _trunc_normal(rng, shape, std):
    # nn.init.trunc_normal_(p, std=.02) truncates at +-2 absolute == +-100 sigma: a plain normal.
    return _normal(rng, shape, std)


def _linear(rng, sd, prefix, n_out, n_in, trunc=False):
    if trunc:
        sd[prefix + ".weight"] = _trunc_normal(rng, (n_out, n_in), 0.02)
    else:
        sd[prefix + ".weight"] = _uniform(rng, (n_out, n_in), 1.0 / math.sqrt(n_in))
    sd[prefix + ".bias"] = _uniform(rng, (n_out,), 1.0 / math.sqrt(n_in))


def _ln(sd, prefix, n, rng=None, jitter=0.0):
    w = torch.ones(n)
    b = torch.zeros(n)
    if rng is not None and jitter > 0:
        w = w + _normal(rng, (n,), jitter)
        b = b + _normal(rng, (n,), jitter)
    sd[prefix + ".weight"] = w
    sd[prefix + ".bias"] = b


def _enc_layer(rng, sd, prefix, with_in_proj, trunc, jitter):
    d, ff = 256, 1024
    if with_in_proj:
        if trunc:
            sd[prefix + ".self_attn.in_proj_weight"] = _trunc_normal(rng, (3 * d, d), 0.02)
        else:  # xavier_uniform_ on (768, 256)
            sd[prefix + ".self_attn.in_proj_weight"] = _uniform(rng, (3 * d, d),
                                                                math.sqrt(6.0 / (3 * d + d)))
        sd[prefix + ".self_attn.in_proj_bias"] = (
            _normal(rng, (3 * d,), jitter) if jitter > 0 else torch.zeros(3 * d))
    _linear(rng, sd, prefix + ".self_attn.out_proj", d, d, trunc)
    if jitter <= 0:
        sd[prefix + ".self_attn.out_proj.bias"] = torch.zeros(d)
    _linear(rng, sd, prefix + ".linear1", ff, d, trunc)
    _linear(rng, sd, prefix + ".linear2", d, ff, trunc)
    _ln(sd, prefix + ".norm1", d, rng, jitter)
    _ln(sd, prefix + ".norm2", d, rng, jitter)
    a = 0.25 if jitter <= 0 else float(0.25 + 0.1 * rng.standard_normal())
    sd[prefix + ".activation.weight"] = torch.tensor([a], dtype=torch.float32)


def make_state_dict(cfg: ModelConfig, seed: int = 2024, spread: bool = False) -> dict:
    """Random-init checkpoint in the reference key layout.

    spread=False: the reference's init distributions ("plain init").
    spread=True : the SURVEY §7 "spread-init" set - LayerNorm affine / biases / PReLU slopes
                  jittered, the last MLP layer of both score heads scaled x64, coef ~ U(0.5,1.5),
                  x = 0.35 - so that scores span (0,1) and every parameter is exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    jit = 0.1 if spread else 0.0
    d = 256
    sd: dict = {}
    nd = cfg.num_dummies
    sd["dummy_rep_token"] = _normal(rng, (nd, d))
    sd["dummy_rep_pos"] = _normal(rng, (nd, d))
    nlev = cfg.num_levels
    sd["coef"] = (_uniform(rng, (nlev,), 0.5) + 1.0) if spread else torch.ones(nlev)
    sd["x"] = torch.tensor(0.35 if spread else 0.5)
    for i in range(cfg.t2v_layers):
        _enc_layer(rng, sd, f"transformer.t2v_encoder.layers.{i}", False, True, jit)
    for i in range(cfg.enc_layers):
        _enc_layer(rng, sd, f"transformer.encoder.layers.{i}", True, True, jit)
    sd["txt_position_embed.position_embeddings.weight"] = _normal(rng, (cfg.max_q_l, d))
    _ln(sd, "txt_position_embed.LayerNorm", d)
    _linear(rng, sd, "saliency_proj1", d, d)
    _linear(rng, sd, "saliency_proj2", d, d)
    for name, dim in (("input_txt_proj", cfg.t_feat_dim), ("input_vid_proj", cfg.v_feat_dim)):
        _ln(sd, f"{name}.0.LayerNorm", dim, rng, jit)
        _linear(rng, sd, f"{name}.0.net.1", d, dim)
        _ln(sd, f"{name}.1.LayerNorm", d, rng, jit)
        _linear(rng, sd, f"{name}.1.net.1", d, d)
    sd["token_type_embeddings.weight"] = _normal(rng, (2, d), 0.02)
    for i in range(cfg.dummy_layers):
        _enc_layer(rng, sd, f"txtproj_encoder.layers.{i}", True, False, jit)
    for l in range(1, nlev):
        for j in range(l):
            sd[f"pyramid.blocks.{l}.{1 + 5 * j}.weight"] = _uniform(rng, (d, d, 2),
                                                                   1.0 / math.sqrt(2 * d))
            sd[f"pyramid.blocks.{l}.{1 + 5 * j}.bias"] = _uniform(rng, (d,), 1.0 / math.sqrt(2 * d))
            _ln(sd, f"pyramid.blocks.{l}.{3 + 5 * j}", d, rng, jit)
    sd["pooling.att.weight"] = _uniform(rng, (1, d), 1.0 / math.sqrt(d))
    k = cfg.kernel_size
    for head in ("conf_head", "class_head"):
        for c in range(cfg.num_conv_layers):
            bound = 1.0 / math.sqrt(d * k)
            sd[f"{head}.convs.{c}.weight"] = _uniform(rng, (d, d, 1, k), bound)
            sd[f"{head}.convs.{c}.bias"] = _uniform(rng, (d,), bound)
        dims = [d] + [128] * (cfg.num_mlp_layers - 1) + [1]
        for m in range(cfg.num_mlp_layers):
            _linear(rng, sd, f"{head}.fc.layers.{m}", dims[m + 1], dims[m])
        if spread:
            sd[f"{head}.fc.layers.{cfg.num_mlp_layers - 1}.weight"] *= 64.0
    ck = cfg.coord_kernel
    bound = 1.0 / math.sqrt(d * ck)
    sd["coord_head.module.1.weight"] = _uniform(rng, (d, d, ck), bound)
    sd["coord_head.module.1.bias"] = _uniform(rng, (d,), bound)
    sd["coord_head.module.3.weight"] = _uniform(rng, (2, d, ck), bound)
    sd["coord_head.module.3.bias"] = _uniform(rng, (2,), bound)
    return sd


def state_dict_checksum(sd: dict) -> float:
    """Order-independent fp64 checksum used to pin regenerated weights to the golden fixtures."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double().reshape(-1)
        w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 97 + 1.0
        tot += float((v * w).sum()) * (1 + (len(k) % 7))
    return tot


def _video_groups(v_feat_dim: int):
    base = v_feat_dim - 2  # TEF is the last 2 columns
    if base == 2816:
        return [2304, 512]  # SlowFast ‖ CLIP, normalised per feature dir
    return [base]


def make_inputs(cfg: ModelConfig, B: int, Lv: int, Lt: int, seed: int = 1234,
                ragged: bool = False, min_lv: int | None = None, min_lt: int = 4):
    """Returns dict(src_vid (B,Lv,Dv), src_vid_mask (B,Lv), src_txt (B,Lt,Dt), src_txt_mask (B,Lt),
    vid_len (B,), txt_len (B,), duration (B,)) as CPU fp32 / int32 tensors.  Padded rows are zero
    (start_end_collate -> pad_sequences_1d, utils/tensor_utils.py:5-53)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if ragged:
        lo = min_lv if min_lv is not None else max(1, Lv // 2)
        vlen = rng.integers(lo, Lv + 1, size=B)
        tlen = rng.integers(min(min_lt, Lt), Lt + 1, size=B)
        vlen[0] = Lv  # batch max defines the padded length
        tlen[0] = Lt
    else:
        vlen = np.full(B, Lv)
        tlen = np.full(B, Lt)
    Dv, Dt = cfg.v_feat_dim, cfg.t_feat_dim
    vid = np.zeros((B, Lv, Dv), np.float32)
    txt = np.zeros((B, Lt, Dt), np.float32)
    for b in range(B):
        lv, lt = int(vlen[b]), int(tlen[b])
        col = 0
        for g in _video_groups(Dv):
            x = rng.standard_normal((lv, g)).astype(np.float32)
            x /= (np.linalg.norm(x, axis=-1, keepdims=True) + 1e-5)
            vid[b, :lv, col:col + g] = x
            col += g
        tef_st = np.arange(lv, dtype=np.float32) / lv
        vid[b, :lv, col] = tef_st
        vid[b, :lv, col + 1] = tef_st + 1.0 / lv
        q = rng.standard_normal((lt, Dt)).astype(np.float32)
        q /= (np.linalg.norm(q, axis=-1, keepdims=True) + 1e-5)
        txt[b, :lt] = q
    ar_v = np.arange(Lv)[None, :]
    ar_t = np.arange(Lt)[None, :]
    return dict(
        src_vid=torch.from_numpy(vid),
        src_vid_mask=torch.from_numpy((ar_v < vlen[:, None]).astype(np.float32)),
        src_txt=torch.from_numpy(txt),
        src_txt_mask=torch.from_numpy((ar_t < tlen[:, None]).astype(np.float32)),
        vid_len=torch.from_numpy(vlen.astype(np.int32)),
        txt_len=torch.from_numpy(tlen.astype(np.int32)),
        duration=torch.from_numpy((vlen * cfg.clip_length).astype(np.float32)),
    )
