"""Reference checkpoint layout -> the packed device weights the C-ABI consumes (FvtgWeights).

The weight ABI of the drop-in is the reference's `state_dict` (FlashVTG/inference.py:471,
`model.load_state_dict(checkpoint["model"], strict=True)`); `expected_shapes` is that key set,
derived from the module constructors (FlashVTG/model.py:81-135, transformer.py:311-330,387-405,
blocks/blocks.py:23-50,93-101).  `pack` converts it once per device into bf16 K-major GEMM
operands (conv taps folded into K) and fp32 vectors, as include/flashvtg_b200.h documents.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import ModelConfig

D, FF, MLP_H = 256, 1024, 128


def expected_shapes(cfg: ModelConfig) -> dict:
    """name -> shape of every entry of the reference FlashVTG.state_dict() for `cfg`."""
    s: dict = {}
    nd, k, ck = cfg.num_dummies, cfg.kernel_size, cfg.coord_kernel
    s["dummy_rep_token"] = (nd, D)
    s["dummy_rep_pos"] = (nd, D)
    s["coef"] = (cfg.num_levels,)
    s["x"] = ()

    def layer(prefix, in_proj):
        if in_proj:
            s[prefix + ".self_attn.in_proj_weight"] = (3 * D, D)
            s[prefix + ".self_attn.in_proj_bias"] = (3 * D,)
        s[prefix + ".self_attn.out_proj.weight"] = (D, D)
        s[prefix + ".self_attn.out_proj.bias"] = (D,)
        s[prefix + ".linear1.weight"] = (FF, D)
        s[prefix + ".linear1.bias"] = (FF,)
        s[prefix + ".linear2.weight"] = (D, FF)
        s[prefix + ".linear2.bias"] = (D,)
        for n in ("norm1", "norm2"):
            s[f"{prefix}.{n}.weight"] = (D,)
            s[f"{prefix}.{n}.bias"] = (D,)
        s[prefix + ".activation.weight"] = (1,)

    for i in range(cfg.t2v_layers):
        layer(f"transformer.t2v_encoder.layers.{i}", False)
    for i in range(cfg.enc_layers):
        layer(f"transformer.encoder.layers.{i}", True)
    for i in range(cfg.dummy_layers):
        layer(f"txtproj_encoder.layers.{i}", True)
    s["txt_position_embed.position_embeddings.weight"] = (cfg.max_q_l, D)
    s["txt_position_embed.LayerNorm.weight"] = (D,)
    s["txt_position_embed.LayerNorm.bias"] = (D,)
    for n in ("saliency_proj1", "saliency_proj2"):
        s[n + ".weight"] = (D, D)
        s[n + ".bias"] = (D,)
    for name, dim in (("input_txt_proj", cfg.t_feat_dim), ("input_vid_proj", cfg.v_feat_dim)):
        s[f"{name}.0.LayerNorm.weight"] = (dim,)
        s[f"{name}.0.LayerNorm.bias"] = (dim,)
        s[f"{name}.0.net.1.weight"] = (D, dim)
        s[f"{name}.0.net.1.bias"] = (D,)
        s[f"{name}.1.LayerNorm.weight"] = (D,)
        s[f"{name}.1.LayerNorm.bias"] = (D,)
        s[f"{name}.1.net.1.weight"] = (D, D)
        s[f"{name}.1.net.1.bias"] = (D,)
    s["token_type_embeddings.weight"] = (2, D)
    for l in range(1, cfg.num_levels):
        for j in range(l):
            s[f"pyramid.blocks.{l}.{1 + 5 * j}.weight"] = (D, D, 2)
            s[f"pyramid.blocks.{l}.{1 + 5 * j}.bias"] = (D,)
            s[f"pyramid.blocks.{l}.{3 + 5 * j}.weight"] = (D,)
            s[f"pyramid.blocks.{l}.{3 + 5 * j}.bias"] = (D,)
    s["pooling.att.weight"] = (1, D)
    dims = [D] + [MLP_H] * (cfg.num_mlp_layers - 1) + [1]
    for head in ("conf_head", "class_head"):
        for c in range(cfg.num_conv_layers):
            s[f"{head}.convs.{c}.weight"] = (D, D, 1, k)
            s[f"{head}.convs.{c}.bias"] = (D,)
        for m in range(cfg.num_mlp_layers):
            s[f"{head}.fc.layers.{m}.weight"] = (dims[m + 1], dims[m])
            s[f"{head}.fc.layers.{m}.bias"] = (dims[m + 1],)
    s["coord_head.module.1.weight"] = (D, D, ck)
    s["coord_head.module.1.bias"] = (D,)
    s["coord_head.module.3.weight"] = (2, D, ck)
    s["coord_head.module.3.bias"] = (2,)
    return s


def check_state_dict(cfg: ModelConfig, sd: dict, strict: bool = True):
    """torch's strict load semantics: returns (missing, unexpected); raises like nn.Module does."""
    exp = expected_shapes(cfg)
    missing = [k for k in exp if k not in sd]
    unexpected = [k for k in sd if k not in exp]
    errors = []
    for k, shp in exp.items():
        if k in sd and tuple(sd[k].shape) != tuple(shp):
            errors.append(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {shp}")
    if strict and (missing or unexpected):
        if missing:
            errors.insert(0, "Missing key(s) in state_dict: " + ", ".join(map(repr, missing)))
        if unexpected:
            errors.insert(0, "Unexpected key(s) in state_dict: " + ", ".join(map(repr, unexpected)))
    if errors or (not strict and missing):
        if not errors:
            errors = ["Missing key(s) in state_dict: " + ", ".join(map(repr, missing))]
        raise RuntimeError("Error(s) in loading state_dict for FlashVTG:\n\t" + "\n\t".join(errors))
    return missing, unexpected


def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


class PackedWeights:
    """Device-resident packed weights + the FvtgWeights struct pointing into them."""

    def __init__(self, cfg: ModelConfig, sd: dict, device: torch.device):
        self.device = device
        self._keep: list = []
        self.struct = _lib.FvtgWeights()
        self.nbytes = 0
        w = self.struct

        def f32(t):
            t = t.detach().to(torch.float32).contiguous().to(device)
            self._keep.append(t)
            self.nbytes += t.numel() * 4
            return t.data_ptr()

        def b16(t, k_pad=None, n_pad=None):
            t = t.detach().to(torch.float32)
            n, k = t.shape
            k_pad = k_pad or pad64(k)
            n_pad = n_pad or n
            out = torch.zeros(n_pad, k_pad, dtype=torch.float32)
            out[:n, :k] = t
            out = out.to(torch.bfloat16).contiguous().to(device)
            self._keep.append(out)
            self.nbytes += out.numel() * 2
            return out.data_ptr()

        def lin(dst, weight, bias, k_pad=None, n_pad=None):
            dst.w = b16(weight, k_pad, n_pad)
            if n_pad and n_pad != bias.numel():
                bb = torch.zeros(n_pad)
                bb[: bias.numel()] = bias
                bias = bb
            dst.b = f32(bias)

        def ln(dst, prefix):
            dst.g = f32(sd[prefix + ".weight"])
            dst.b = f32(sd[prefix + ".bias"])

        te = sd["token_type_embeddings.weight"].float()
        for name, dst, dim, row in (("input_vid_proj", w.vid, cfg.v_feat_dim, 1),
                                    ("input_txt_proj", w.txt, cfg.t_feat_dim, 0)):
            ln(dst.ln0, f"{name}.0.LayerNorm")
            # LayerNorm(raw dim) folded into the first projection (csrc/inproj.cu):
            #   LN(x) W^T + b = rstd * ((x - m0) Wg^T - mean(x - m0) * rowsum(Wg)) + (W beta + b)
            w0 = sd[f"{name}.0.net.1.weight"].float()
            g0 = sd[f"{name}.0.LayerNorm.weight"].float()
            b0 = sd[f"{name}.0.LayerNorm.bias"].float()
            wg = w0 * g0[None, :]
            lin(dst.fc0, wg, w0 @ b0 + sd[f"{name}.0.net.1.bias"].float(), pad64(dim))
            dst.fc0_wsum = f32(wg.to(torch.bfloat16).float().sum(1))
            ln(dst.ln1, f"{name}.1.LayerNorm")
            # token_type_embeddings row folded into the bias (model.py:151-152)
            lin(dst.fc1, sd[f"{name}.1.net.1.weight"], sd[f"{name}.1.net.1.bias"].float() + te[row])
        w.dummy_tok = f32(sd["dummy_rep_token"])
        w.dummy_pos = f32(sd["dummy_rep_pos"])

        def layer(dst, prefix, in_proj):
            if in_proj:
                lin(dst.in_proj, sd[prefix + ".self_attn.in_proj_weight"],
                    sd[prefix + ".self_attn.in_proj_bias"])
            lin(dst.out_proj, sd[prefix + ".self_attn.out_proj.weight"],
                sd[prefix + ".self_attn.out_proj.bias"])
            ln(dst.norm1, prefix + ".norm1")
            lin(dst.ff1, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"])
            lin(dst.ff2, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])
            ln(dst.norm2, prefix + ".norm2")
            dst.prelu = float(sd[prefix + ".activation.weight"].reshape(-1)[0])

        for i in range(cfg.dummy_layers):
            layer(w.dummy[i], f"txtproj_encoder.layers.{i}", True)
        for i in range(cfg.t2v_layers):
            layer(w.t2v[i], f"transformer.t2v_encoder.layers.{i}", False)
        for i in range(cfg.enc_layers):
            layer(w.enc[i], f"transformer.encoder.layers.{i}", True)
        w.sal_w1 = f32(sd["saliency_proj1.weight"])
        w.sal_b1 = f32(sd["saliency_proj1.bias"])
        w.sal_w2t = f32(sd["saliency_proj2.weight"].t())
        w.sal_b2 = f32(sd["saliency_proj2.bias"])
        for l in range(1, cfg.num_levels):
            for j in range(l):
                cw = sd[f"pyramid.blocks.{l}.{1 + 5 * j}.weight"]      # (out, in, tap)
                lin(w.pyr[l][j].conv, cw.permute(0, 2, 1).reshape(D, 2 * D),
                    sd[f"pyramid.blocks.{l}.{1 + 5 * j}.bias"])
                ln(w.pyr[l][j].ln, f"pyramid.blocks.{l}.{3 + 5 * j}")
        k = cfg.kernel_size
        for head, dst in (("class_head", w.cls), ("conf_head", w.conf)):
            for c in range(cfg.num_conv_layers):
                cw = sd[f"{head}.convs.{c}.weight"][:, :, 0, :]      # (out, in, tap)
                lin(dst.conv[c], cw.permute(0, 2, 1).reshape(D, k * D), sd[f"{head}.convs.{c}.bias"])
            for m in range(cfg.num_mlp_layers - 1):
                lin(dst.mlp[m], sd[f"{head}.fc.layers.{m}.weight"], sd[f"{head}.fc.layers.{m}.bias"])
            last = cfg.num_mlp_layers - 1
            dst.last_w = f32(sd[f"{head}.fc.layers.{last}.weight"].reshape(-1))
            dst.last_b = float(sd[f"{head}.fc.layers.{last}.bias"].reshape(-1)[0])
        ck = cfg.coord_kernel
        lin(w.coord1, sd["coord_head.module.1.weight"].permute(0, 2, 1).reshape(D, ck * D),
            sd["coord_head.module.1.bias"])
        lin(w.coord2, sd["coord_head.module.3.weight"].permute(0, 2, 1).reshape(2, ck * D),
            sd["coord_head.module.3.bias"], n_pad=16)
        coef = sd["coef"].float().reshape(-1).tolist()
        for i in range(_lib.MAX_LEVELS):
            w.coef[i] = coef[i] if i < len(coef) else 1.0
        w.x = float(sd["x"])

    def ref(self):
        return C.byref(self.struct)


def make_cfg_struct(cfg: ModelConfig) -> "_lib.FvtgCfg":
    c = _lib.FvtgCfg()
    c.abi_version = _lib.ABI_VERSION
    c.v_dim, c.t_dim = cfg.v_feat_dim, cfg.t_feat_dim
    c.v_dim_pad, c.t_dim_pad = pad64(cfg.v_feat_dim), pad64(cfg.t_feat_dim)
    c.num_dummies = cfg.num_dummies
    c.dummy_layers, c.t2v_layers, c.enc_layers = cfg.dummy_layers, cfg.t2v_layers, cfg.enc_layers
    c.num_levels = cfg.num_levels
    c.head_k = cfg.kernel_size
    c.num_conv_layers = cfg.num_conv_layers
    c.num_mlp_layers = cfg.num_mlp_layers
    c.coord_k = cfg.coord_kernel
    c.max_num_moment = cfg.max_num_moment
    c.clip_len = cfg.clip_length
    return c
