"""Host mirror of the reference's list-based post-processing API, backed by the CUDA kernels.

  temporal_nms            : batched device entry (fvtg_temporal_nms)
  post_processing_mr_nms  : same signature / in-place semantics as FlashVTG/inference.py:36-57
  temporal_nms_list       : same signature as utils/temporal_nms.py:25 (the hull variant)
  compute_mr_results      : counterpart of FlashVTG/inference.py:232-355 for a FlashVTGB200 model -
                            forward + compose (clamp, 4-dp) + PostProcessorDETR in one device pass

No CPU fallback: everything below needs the sm_100a library and a CUDA device.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib

_MODES = {"normal": _lib.NMS_NORMAL, "linear": _lib.NMS_LINEAR, "hull": _lib.NMS_HULL}


def temporal_nms(windows: torch.Tensor, count: Optional[torch.Tensor], thd: float,
                 mode: str = "normal", max_after_nms: int = 100):
    """windows fp32 (B, M, 3) CUDA rows (st, ed, score), count int32 (B,) or None (= M each).
    Returns (out_windows (B,M,3), order (B,M) int32 source row or -1, out_count (B,) int32)."""
    if mode not in _MODES:
        raise ValueError(f"Unknown nms_type: {mode}")  # inference.py:50
    if not windows.is_cuda:
        raise RuntimeError("temporal_nms runs on a CUDA (sm_100a) device only; there is no CPU path")
    if windows.dtype != torch.float32 or windows.dim() != 3 or windows.shape[2] != 3:
        raise ValueError("windows must be fp32 (B, M, 3)")
    windows = windows.contiguous()
    B, M, _ = windows.shape
    lib = _lib.load()
    out = torch.empty_like(windows)
    order = torch.empty(B, M, dtype=torch.int32, device=windows.device)
    ocount = torch.empty(B, dtype=torch.int32, device=windows.device)
    if B == 0:
        return out, order, ocount
    if count is not None:
        count = count.to(device=windows.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(windows.device):
        rc = lib.fvtg_temporal_nms(windows.data_ptr(), _lib.ptr(count), B, M, float(thd), _MODES[mode],
                                   int(max_after_nms), out.data_ptr(), order.data_ptr(),
                                   ocount.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "fvtg_temporal_nms")
    return out, order, ocount


def _batch_lists(lists, device):
    n = [len(w) for w in lists]
    M = max(max(n), 1) if n else 1
    if M > _lib.MAX_TOPK:
        raise ValueError(f"at most {_lib.MAX_TOPK} windows per query are supported (got {M}); the "
                         "reference never produces more than max_num_moment = 50")
    buf = np.zeros((len(lists), M, 3), np.float32)
    for i, w in enumerate(lists):
        if n[i]:
            buf[i, : n[i]] = np.asarray(w, dtype=np.float64).astype(np.float32).reshape(-1, 3)
    return (torch.from_numpy(buf).to(device, non_blocking=False),
            torch.tensor(n, dtype=torch.int32, device=device), n)


def post_processing_mr_nms(mr_res, nms_thd, max_before_nms, max_after_nms, nms_type,  # noqa: ARG001
                           device="cuda"):
    """Drop-in for FlashVTG/inference.py:36-57: every entry's "pred_relevant_windows" is replaced
    by the NMS result (all rows kept, suppressed scores zeroed, sorted by score descending).
    `max_before_nms` / `max_after_nms` are accepted and ignored, exactly like the reference."""
    if nms_type not in ("normal", "linear"):
        raise ValueError(f"Unknown nms_type: {nms_type}")
    if not mr_res:
        return []
    win, cnt, n = _batch_lists([e["pred_relevant_windows"] for e in mr_res], torch.device(device))
    out, _, _ = temporal_nms(win, cnt, nms_thd, nms_type)
    out = out.cpu()
    res = []
    for i, e in enumerate(mr_res):
        e["pred_relevant_windows"] = out[i, : n[i]].tolist()
        res.append(e)
    return res


def temporal_nms_list(predictions, nms_thd, max_after_nms=100, device="cuda"):
    """Drop-in for utils/temporal_nms.py:25-74 (hull 'union', strict >, removal) on one list.
    Rows come back as the original python floats (the reference only re-orders / drops them)."""
    if len(predictions) == 1:
        return predictions
    if not predictions:
        return []
    win64 = np.asarray([list(map(float, p)) for p in predictions], np.float64).reshape(1, -1, 3)
    M = win64.shape[1]
    if M > _lib.MAX_TOPK:
        raise ValueError(f"at most {_lib.MAX_TOPK} windows are supported")
    dev = torch.device(device)
    win = torch.from_numpy(win64).to(dev)
    order = torch.empty(1, M, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.fvtg_temporal_nms_hull_f64(win.data_ptr(), None, 1, M, float(nms_thd),
                                            int(max_after_nms), order.data_ptr(), cnt.data_ptr(),
                                            _lib.stream_ptr())
    _lib.check(rc, "fvtg_temporal_nms_hull_f64")
    k = int(cnt[0].item())
    return [predictions[j] for j in order[0, :k].tolist()]


def round4_host(x: np.ndarray) -> np.ndarray:
    """float(f"{e:.4f}") vectorised: fp32 -> fp64, x*1e4 exact, rint half-even, /1e4."""
    return np.rint(np.asarray(x, np.float64) * 1e4) / 1e4


@torch.no_grad()
def compute_mr_results(model, eval_loader, opt=None, nms: Optional[str] = None):
    """Counterpart of compute_mr_results (inference.py:232-355) for a FlashVTGB200 model.

    eval_loader yields the reference's (query_meta, batched_inputs) pairs where batched_inputs is
    the dict `prepare_batch_inputs` returns (src_txt, src_txt_mask, src_vid, src_vid_mask); any
    batch size is accepted.  Returns the same list of dicts (qid, query, vid,
    pred_relevant_windows, pred_saliency_scores) AFTER PostProcessorDETR; with `nms` set the
    windows additionally went through post_processing_mr_nms on the device."""
    mr_res = []
    for query_meta, inp in eval_loader:
        dev = inp["src_vid"].device
        vid_len = inp["src_vid_mask"].sum(1).to(torch.int32)
        txt_len = inp["src_txt_mask"].sum(1).to(torch.int32)
        dur = torch.tensor([float(m["duration"]) for m in query_meta], dtype=torch.float32, device=dev)
        r = model.infer(inp["src_vid"].contiguous().float(), vid_len,
                        inp["src_txt"].contiguous().float(), txt_len, duration=dur, nms=nms,
                        nms_thd=getattr(opt, "nms_thd", None))
        win = (r.nms_windows if nms else r.windows).cpu().numpy()
        cnt = r.count.cpu().tolist()
        sal = r.saliency.cpu().numpy()
        vl = vid_len.cpu().tolist()
        for i, meta in enumerate(query_meta):
            w = win[i, : cnt[i]].astype(np.float64)
            if not nms:
                w[:, 2] = round4_host(w[:, 2])   # the 4-dp score is a python float (postprocessing.py:33)
            mr_res.append(dict(qid=meta["qid"], query=meta["query"], vid=meta["vid"],
                               pred_relevant_windows=w.tolist(),
                               pred_saliency_scores=round4_host(sal[i, : vl[i]]).tolist()))
    return mr_res


def highlight_ap(pred, label, topk: Optional[int] = None) -> float:
    """Average precision of one ranked binary relevance list as FlashVTG/inference.py:169-186,191-213
    computes it (trapezoidal precision-recall area, UMT convention): `pred` scores (n,), `label` 0/1 (n,);
    the list is ranked by score (descending) and cut to `topk` (5 for TVSum, None for YouTube-HL)."""
    pred = np.asarray(pred, dtype=np.float64)
    label = np.asarray(label, dtype=np.float64)
    order = np.argsort(-pred, kind="stable")
    rel = label[order][:topk] if topk is not None else label[order]
    num_gt = rel.sum()
    if num_gt == 0:
        return 0.0
    hits = np.cumsum(rel)
    rec = hits / num_gt
    prc = hits / np.arange(1, len(rel) + 1)
    rec_prev = np.concatenate([[0.0], rec[:-1]])
    prc_prev = np.concatenate([[1.0], prc[:-1]])
    return float(np.sum((rec - rec_prev) * (prc_prev + prc) / 2))


def compute_hl_results(model, eval_loader, opt=None):
    """Counterpart of compute_hl_results (FlashVTG/inference.py:118-229) for a FlashVTGB200 model: forward
    (saliency_scores on the device, any batch size), then the per-video top-5 mAP (TVSum: 20 annotators,
    label binarised at its median) or full-list AP (YouTube-HL).  eval_loader yields (query_meta,
    batched_inputs) with batched_inputs as prepare_batch_inputs returns them and meta["label"] the raw
    labels.  Returns dict(mAP=...) like the reference's `submmission`."""
    dset = getattr(opt, "dset_name", model.cfg.dset_name)
    if dset not in ("tvsum", "youtube_uni"):
        raise ValueError("No such dataset")   # inference.py:214-216
    aps = []
    for query_meta, inp in eval_loader:
        vid_len = inp["src_vid_mask"].sum(1).to(torch.int32)
        txt_len = inp["src_txt_mask"].sum(1).to(torch.int32)
        r = model.infer(inp["src_vid"].contiguous().float(), vid_len, inp["src_txt"].contiguous().float(),
                        txt_len, nms=None)
        sal = r.saliency.cpu().numpy()
        for i, meta in enumerate(query_meta):
            label = np.asarray(meta["label"], dtype=np.float64)
            pred = sal[i, : len(label)]
            if dset == "tvsum":
                video_ap = []
                for a in range(label.shape[1]):
                    cur = label[:, a]
                    # torch.median returns the LOWER of the two middle values for even counts
                    med = np.sort(cur)[(len(cur) - 1) // 2]
                    video_ap.append(highlight_ap(pred, (cur > med).astype(np.float64), topk=5))
                aps.append(video_ap)
            else:
                # quirk kept (inference.py:197-200): for a video without a positive label the reference's
                # `continue` leaves the per-video loop BEFORE video_ap_collected.append, so the video is left
                # out of the mean instead of counting as AP 0 (the TVSum branch's `continue` only skips one
                # annotator and keeps the video)
                if label.reshape(-1).sum() == 0:
                    continue
                aps.append([highlight_ap(pred, label.reshape(-1))])
    return dict(mAP=round(float(np.mean(aps)), 5)) if aps else dict(mAP=0.0)
