"""Device-resident input pipeline: the reference's per-item host work between the
feature files and `model.forward` - `l2_normalize_np_array` per feature directory, concatenation,
TEF, zero padding and masks (FlashVTG/start_end_dataset.py:160-182,462-531,534-570;
utils/basic_utils.py:84-86; utils/tensor_utils.py:5-53) - as one CUDA pass (`fvtg_prepare_inputs`).

Raw feature arrays may be fp32, fp16 or bf16: the loader casts to fp32 before normalising, so a
feature store kept in half precision crosses PCIe at half the bytes and is widened on the device.
No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

_DT = {torch.float32: _lib.RAW_F32, torch.float16: _lib.RAW_F16, torch.bfloat16: _lib.RAW_BF16}


def prepare_inputs(raw_vid: Sequence[torch.Tensor], vid_len: torch.Tensor, raw_txt: torch.Tensor,
                   txt_len: torch.Tensor, normalize_v: bool = True, normalize_t: bool = True,
                   use_tef: bool = True, want_masks: bool = True):
    """raw_vid: per feature directory a CUDA tensor (B, Lv, D_g) (rows >= vid_len[b] ignored);
    raw_txt (B, Lt, Dt); vid_len / txt_len int32 (B,) on the same device.
    Returns (src_vid (B, Lv, sum D_g + 2*use_tef) fp32, src_vid_mask, src_txt fp32, src_txt_mask),
    i.e. the tensors `prepare_batch_inputs` hands to `model.forward`."""
    if not raw_vid or len(raw_vid) > _lib.RAW_MAX_GROUPS:
        raise ValueError(f"1..{_lib.RAW_MAX_GROUPS} video feature groups expected")
    dev = raw_txt.device
    if dev.type != "cuda":
        raise RuntimeError("prepare_inputs runs on a CUDA (sm_100a) device only; there is no CPU path")
    dt = raw_txt.dtype
    if dt not in _DT:
        raise ValueError(f"raw features must be fp32 / fp16 / bf16, got {dt}")
    B, Lt, Dt = raw_txt.shape
    Lv = raw_vid[0].shape[1]
    for g in raw_vid:
        if g.dtype != dt or g.device != dev or g.dim() != 3 or g.shape[0] != B or g.shape[1] != Lv:
            raise ValueError("all raw feature arrays must share dtype, device, batch size and length")
    raw_vid = [g.contiguous() for g in raw_vid]
    raw_txt = raw_txt.contiguous()
    vid_len = vid_len.to(device=dev, dtype=torch.int32).contiguous()
    txt_len = txt_len.to(device=dev, dtype=torch.int32).contiguous()
    Dv = sum(g.shape[2] for g in raw_vid) + (2 if use_tef else 0)
    src_vid = torch.empty(B, Lv, Dv, dtype=torch.float32, device=dev)
    src_txt = torch.empty(B, Lt, Dt, dtype=torch.float32, device=dev)
    vmask = torch.empty(B, Lv, dtype=torch.float32, device=dev) if want_masks else None
    tmask = torch.empty(B, Lt, dtype=torch.float32, device=dev) if want_masks else None
    rb = _lib.FvtgRawBatch()
    rb.B, rb.Lv, rb.Lt, rb.n_groups = B, Lv, Lt, len(raw_vid)
    for i, g in enumerate(raw_vid):
        rb.group_dim[i] = g.shape[2]
        rb.vid[i] = g.data_ptr()
    rb.t_dim, rb.dtype = Dt, _DT[dt]
    rb.normalize_v, rb.normalize_t, rb.use_tef = int(normalize_v), int(normalize_t), int(use_tef)
    rb.txt, rb.vid_len, rb.txt_len = raw_txt.data_ptr(), vid_len.data_ptr(), txt_len.data_ptr()
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.fvtg_prepare_inputs(C.byref(rb), src_vid.data_ptr(), _lib.ptr(vmask), src_txt.data_ptr(),
                                     _lib.ptr(tmask), _lib.stream_ptr())
    _lib.check(rc, "fvtg_prepare_inputs")
    return src_vid, vmask, src_txt, tmask
