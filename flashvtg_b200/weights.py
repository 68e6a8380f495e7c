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
    """Device-resident packed weights + the FvtgWeights struct pointing into them.

    The packing itself (bf16 K-major operands, conv taps folded into K, LayerNorm / token-type folding) is
    fvtg_pack_weights in the C-ABI library (csrc/pack.cu): this class only hands it the state_dict as
    (name, host pointer, numel) triples and owns the one device buffer the result lives in, so a non-Python
    host loads a checkpoint through exactly the same code."""

    def __init__(self, cfg: ModelConfig, sd: dict, device: torch.device):
        self.device = device
        self.struct = _lib.FvtgWeights()
        lib = _lib.load()
        cs = make_cfg_struct(cfg)
        self.nbytes = int(lib.fvtg_packed_weights_bytes(C.byref(cs)))
        if self.nbytes == 0:
            raise RuntimeError("fvtg_packed_weights_bytes rejected the configuration: "
                               + lib.fvtg_last_error().decode(errors="replace"))
        host = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in sd.items()}
        params = (_lib.FvtgParam * len(host))()
        self._names = [k.encode() for k in host]            # keep the C strings alive for the call
        for i, (k, t) in enumerate(host.items()):
            params[i].name = self._names[i]
            params[i].data = t.data_ptr()
            params[i].numel = t.numel()
        if torch.device(device).type == "cuda":
            self.buffer = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                rc = lib.fvtg_pack_weights(C.byref(cs), params, len(host), self.buffer.data_ptr(), self.nbytes,
                                           C.byref(self.struct), _lib.stream_ptr())
        else:   # layout inspection on the host (tests): same packer, no CUDA call
            raw = torch.empty(self.nbytes + 256, dtype=torch.uint8)
            skip = (-raw.data_ptr()) % 256
            self.buffer = raw[skip:skip + self.nbytes]
            rc = lib.fvtg_pack_weights_host(C.byref(cs), params, len(host), self.buffer.data_ptr(), self.nbytes,
                                            None, C.byref(self.struct))
        _lib.check(rc, "fvtg_pack_weights")

    def view(self, ptr: int, shape, dtype=torch.float32) -> torch.Tensor:
        """Tensor view of the packed buffer at address `ptr` (a pointer field of .struct)."""
        n = 1
        for d in shape:
            n *= d
        off = ptr - self.buffer.data_ptr()
        esz = torch.empty(0, dtype=dtype).element_size()
        return self.buffer[off:off + n * esz].view(dtype).view(*shape)

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
