"""torch.library registration of the hot path (namespace `fvtg`), CUDA only.

`BASELINE.json:north_star` asks for the kernels to be reachable as torch operators; these are thin
dispatcher entries over the same C-ABI calls `FlashVTGB200.infer` / `postprocessing.temporal_nms`
make.  There is deliberately no CPU / Meta implementation: dispatching them on a CPU tensor
raises NotImplementedError from the dispatcher.
"""
from __future__ import annotations

from typing import List

import torch

_MODELS: dict = {}
_NMS_NAMES = {-1: None, 0: "normal", 1: "linear", 2: "hull"}


def register_model(model) -> int:
    """Returns the integer handle `fvtg::forward` takes (weights stay owned by the module)."""
    h = id(model)
    _MODELS[h] = model
    return h


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
    return [r.saliency, r.t2vattn, r.boundary, r.windows, r.count,
            r.nms_windows if r.nms_windows is not None else empty,
            r.nms_order if r.nms_order is not None else empty.to(torch.int32)]


@torch.library.custom_op("fvtg::temporal_nms", mutates_args=(), device_types="cuda")
def fvtg_temporal_nms(windows: torch.Tensor, count: torch.Tensor, thd: float, mode: int,
                      max_after_nms: int) -> List[torch.Tensor]:
    """-> [out_windows (B,M,3), order (B,M) int32, out_count (B) int32]"""
    from .postprocessing import temporal_nms
    out, order, cnt = temporal_nms(windows, count, thd, _NMS_NAMES[mode], max_after_nms)
    return [out, order, cnt]
