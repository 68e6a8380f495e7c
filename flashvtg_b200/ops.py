"""torch.library registration of the hot path (namespace `fvtg`), CUDA only.

`BASELINE.json:north_star` asks for the kernels to be reachable as torch operators: one operator per
kernel group of SURVEY.md section 8(b) plus the whole path and the standalone NMS, each a thin dispatcher
entry over the C-ABI call of the same name (include/flashvtg_b200.h):

    torch.ops.fvtg.fusion_fwd        -> fvtg_fusion_fwd         (group A: projections .. encoder, saliency)
    torch.ops.fvtg.pyramid_heads_fwd -> fvtg_pyramid_heads_fwd  (group B: pyramid + class / conf / coord heads)
    torch.ops.fvtg.decode_nms        -> fvtg_decode_nms         (group C: ASR, decode, top-k, post-process, NMS)
    torch.ops.fvtg.forward           -> fvtg_forward            (A + B + C in one call)
    torch.ops.fvtg.temporal_nms      -> fvtg_temporal_nms / fvtg_temporal_nms_hull_f64

There is deliberately no CPU / Meta implementation: dispatching them on a CPU tensor raises
NotImplementedError from the dispatcher.  Weights stay owned by a FlashVTGB200 module; the operators take
the integer handle register_model() returns.
"""
from __future__ import annotations

from typing import List

import torch

_MODELS: dict = {}
_NMS_NAMES = {-1: None, 0: "normal", 1: "linear", 2: "hull"}


def register_model(model) -> int:
    """Returns the integer handle the fvtg:: operators take (weights stay owned by the module)."""
    h = id(model)
    _MODELS[h] = model
    return h


@torch.library.custom_op("fvtg::fusion_fwd", mutates_args=(), device_types="cuda")
def fvtg_fusion_fwd(src_vid: torch.Tensor, vid_len: torch.Tensor, src_txt: torch.Tensor,
                    txt_len: torch.Tensor, model: int) -> List[torch.Tensor]:
    """-> [video_emb (B,Lv,256), saliency (B,Lv), t2vattn (B,Lv), dummy_tokens (B,nd,256)]"""
    return list(_MODELS[model].fusion(src_vid, vid_len, src_txt, txt_len))


@torch.library.custom_op("fvtg::pyramid_heads_fwd", mutates_args=(), device_types="cuda")
def fvtg_pyramid_heads_fwd(video_emb: torch.Tensor, vid_len: torch.Tensor, model: int) -> List[torch.Tensor]:
    """-> [cls_logit (B,N), conf_logit (B,N), coord (B,N,2)]"""
    return list(_MODELS[model].pyramid_heads(video_emb, vid_len))


@torch.library.custom_op("fvtg::decode_nms", mutates_args=(), device_types="cuda")
def fvtg_decode_nms(cls_logit: torch.Tensor, conf_logit: torch.Tensor, coord: torch.Tensor,
                    vid_len: torch.Tensor, duration: torch.Tensor, Lv: int, model: int, nms_mode: int,
                    nms_thd: float) -> List[torch.Tensor]:
    """-> [boundary (B,K,3), windows (B,K,3), nms_windows (B,K,3), nms_order (B,K), count (B), nms_count (B)]"""
    return list(_MODELS[model].decode(cls_logit, conf_logit, coord, vid_len, Lv, duration=duration,
                                      nms=_NMS_NAMES[nms_mode], nms_thd=nms_thd))


@torch.library.custom_op("fvtg::forward", mutates_args=(), device_types="cuda")
def fvtg_forward(src_vid: torch.Tensor, vid_len: torch.Tensor, src_txt: torch.Tensor,
                 txt_len: torch.Tensor, duration: torch.Tensor, model: int, nms_mode: int,
                 nms_thd: float) -> List[torch.Tensor]:
    """-> [saliency (B,Lv), t2vattn (B,Lv), boundary (B,K,3), windows (B,K,3), count (B),
           nms_windows (B,K,3), nms_order (B,K)]  (the last two are empty when nms_mode == -1)."""
    m = _MODELS[model]
    r = m.infer(src_vid, vid_len, src_txt, txt_len, duration=duration, nms=_NMS_NAMES[nms_mode],
                nms_thd=nms_thd)
    empty = torch.empty(0, device=src_vid.device)
    # saliency / count / nms_windows are views of ONE packed record buffer (FvtgResult.packed); an operator's
    # returns may not alias each other, so those three leave as copies
    return [r.saliency.clone(), r.t2vattn, r.boundary, r.windows, r.count.clone(),
            r.nms_windows.clone() if r.nms_windows is not None else empty,
            r.nms_order if r.nms_order is not None else empty.to(torch.int32)]


@torch.library.custom_op("fvtg::temporal_nms", mutates_args=(), device_types="cuda")
def fvtg_temporal_nms(windows: torch.Tensor, count: torch.Tensor, thd: float, mode: int,
                      max_after_nms: int) -> List[torch.Tensor]:
    """-> [out_windows (B,M,3), order (B,M) int32, out_count (B) int32]"""
    from .postprocessing import temporal_nms
    out, order, cnt = temporal_nms(windows, count, thd, _NMS_NAMES[mode], max_after_nms)
    return [out, order, cnt]
