"""QVHighlights evaluation with the per-query work on the device - the drop-in for the reference's
`standalone_eval.eval.eval_submission` (standalone_eval/eval.py:271-345; same name, arguments, returned
dict and 2-decimal formatting).

The reference spreads Python loops over queries on an 8-process pool (eval.py:24-69 detection AP,
:72-102 R1 / mIoU, :173-236 highlight Hit1 / AP).  Here `fvtg_eval_submission` (csrc/evalmetrics.cu)
computes every per-query number in fp64 with numpy's operation order, so they are bit-identical to the
CPU restatement; this module only packs the rows, averages over queries and formats.

`eval_arrays` is the array-level entry for predictions that already live on the device (the ranked
windows and saliency scores `FlashVTGB200.infer` returns): nothing but a few KB of per-query results
crosses PCIe.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import _lib

AP_THDS = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]      # eval.py:26
R1_THDS = [float(f"{e:.2f}") for e in np.linspace(0.3, 0.95, 14)]      # eval.py:74
RANGE_NAMES = ("short", "middle", "long", "full")                        # eval.py:141-143
HL_NAMES = ("Fair", "Good", "VeryGood")                                  # eval.py:251-252


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def eval_arrays(pred_win: Optional[torch.Tensor], pred_cnt: Optional[torch.Tensor], gt_win: torch.Tensor,
                gt_cnt: torch.Tensor, pred_sal: Optional[torch.Tensor] = None,
                pred_sal_len: Optional[torch.Tensor] = None, gt_sal: Optional[torch.Tensor] = None,
                gt_clips: Optional[torch.Tensor] = None, max_pred_windows: int = 10):
    """Per-query metrics on the device.  pred_win [Q][P][3] (start, end, score; any float dtype, listed
    order), gt_win [Q][G][2], pred_sal [Q][L], gt_sal u8 [Q][C][3] (annotator scores of every clip),
    counts int32 [Q].  Returns (mr, hl): mr = (ap [4][Q][10], iou [4][Q], valid [4][Q]) or None,
    hl = (ap [3][Q][3], hit [3][Q]) or None - device tensors."""
    lib = _lib.load()
    dev = gt_win.device
    if dev.type != "cuda":
        raise RuntimeError("flashvtg_b200.evaluation runs on a CUDA device (no CPU fallback)")
    Q = int(gt_win.shape[0])
    f64 = dict(device=dev, dtype=torch.float64)
    i32 = dict(device=dev, dtype=torch.int32)
    keep = []

    def prep(t, dtype):
        t = t.to(device=dev, dtype=dtype).contiguous()
        keep.append(t)
        return t

    b = _lib.FvtgEvalBatch()
    b.n_queries = Q
    gt_win = prep(gt_win, torch.float64)
    gt_cnt = prep(gt_cnt, torch.int32)
    b.max_gt = int(gt_win.shape[1])
    b.gt_win, b.gt_cnt = _ptr(gt_win), _ptr(gt_cnt)
    mr = hl = None
    mr_ap = mr_iou = mr_valid = hl_ap = hl_hit = None
    b.max_pred = b.max_sal = b.max_clips = 1
    if pred_win is not None:
        pred_win = prep(pred_win, torch.float64)
        pred_cnt = prep(pred_cnt, torch.int32)
        b.max_pred = int(pred_win.shape[1])
        b.pred_win, b.pred_cnt = _ptr(pred_win), _ptr(pred_cnt)
        mr_ap = torch.empty(4, Q, 10, **f64)
        mr_iou = torch.empty(4, Q, **f64)
        mr_valid = torch.empty(4, Q, device=dev, dtype=torch.uint8)
    if pred_sal is not None:
        pred_sal = prep(pred_sal, torch.float64)
        pred_sal_len = prep(pred_sal_len, torch.int32)
        gt_sal = prep(gt_sal, torch.uint8)
        gt_clips = prep(gt_clips, torch.int32)
        b.max_sal = int(pred_sal.shape[1])
        b.max_clips = int(gt_sal.shape[1])
        b.pred_sal, b.pred_sal_len = _ptr(pred_sal), _ptr(pred_sal_len)
        b.gt_sal, b.gt_clips = _ptr(gt_sal), _ptr(gt_clips)
        hl_ap = torch.empty(3, Q, 3, **f64)
        hl_hit = torch.empty(3, Q, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.fvtg_eval_submission(C.byref(b), int(max_pred_windows), _ptr(mr_ap), _ptr(mr_iou), _ptr(mr_valid),
                                  _ptr(hl_ap), _ptr(hl_hit), C.c_void_p(st))
    _lib.check(rc, "fvtg_eval_submission")
    if mr_ap is not None:
        mr = (mr_ap, mr_iou, mr_valid)
    if hl_ap is not None:
        hl = (hl_ap, hl_hit)
    return mr, hl


def _pct(v) -> float:
    return float(f"{100 * v:.2f}")


def format_metrics(mr=None, hl=None) -> OrderedDict:
    """Means over queries + the reference's dict layout (eval.py:56-69, 95-101, 155-169, 255-262, 300-345)."""
    metrics, brief = {}, OrderedDict()
    if mr is not None:
        ap, iou, valid = (np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x) for x in mr)
        valid = valid.astype(bool)
        for r, name in enumerate(RANGE_NAMES):
            v = valid[r]
            if not v.any():      # eval.py:153-158
                dummy = {k: 0. for k in np.linspace(0.5, 0.95, 19)}
                dummy["average"] = 0.
                metrics[name] = {"MR-mAP": dummy, "MR-R1": dummy}
                continue
            ap_thds = ap[r][v].mean(0)
            d_ap = dict(zip([str(e) for e in AP_THDS], ap_thds))
            d_ap["average"] = np.mean(ap_thds)
            d_ap = {k: _pct(x) for k, x in d_ap.items()}
            ious = iou[r][v]
            metrics[name] = {"MR-mIoU": _pct(np.mean(ious)), "MR-mAP": d_ap,
                             "MR-R1": {str(t): _pct(np.mean(ious >= t)) for t in R1_THDS}}
        full = metrics["full"]
        b = {
            "MR-full-mAP": full["MR-mAP"]["average"], "MR-full-mAP@0.5": full["MR-mAP"]["0.5"],
            "MR-full-mAP@0.75": full["MR-mAP"]["0.75"],
            "MR-short-mAP": metrics["short"]["MR-mAP"]["average"],
            "MR-middle-mAP": metrics["middle"]["MR-mAP"]["average"],
            "MR-long-mAP": metrics["long"]["MR-mAP"]["average"],
            "MR-full-mIoU": full["MR-mIoU"], "MR-full-R1@0.3": full["MR-R1"]["0.3"],
            "MR-full-R1@0.5": full["MR-R1"]["0.5"], "MR-full-R1@0.7": full["MR-R1"]["0.7"],
        }
        brief.update(sorted(b.items(), key=lambda x: x[0]))
    if hl is not None:
        ap, hit = (np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x) for x in hl)
        hl_metrics = {}
        for m, name in enumerate(HL_NAMES):
            hl_metrics[f"HL-min-{name}"] = {"HL-mAP": _pct(np.mean(ap[m])),
                                            "HL-Hit1": _pct(np.mean(hit[m].astype(np.float64)))}
        metrics.update(hl_metrics)
        brief.update({f"{k}-{sk.split('-')[1]}": v[sk] for k, v in hl_metrics.items() for sk in v})
    out = OrderedDict()
    out["brief"] = brief
    out.update(sorted(metrics.items(), key=lambda x: x[0]))
    return out


def pack_submission(submission, ground_truth, clip_length: int = 2):
    """jsonl rows -> pinned host arrays, one row per submission entry (GT matched by qid)."""
    gt_by_qid = {d["qid"]: d for d in ground_truth}
    Q = len(submission)
    has_mr = "pred_relevant_windows" in submission[0]
    has_hl = "pred_saliency_scores" in submission[0]
    gts = [gt_by_qid[d["qid"]] for d in submission]
    G = max(max(len(g["relevant_windows"]) for g in gts), 1)
    out = {"gt_win": np.zeros((Q, G, 2)), "gt_cnt": np.zeros(Q, np.int32)}
    for i, g in enumerate(gts):
        w = np.asarray(g["relevant_windows"], dtype=np.float64).reshape(-1, 2)
        out["gt_win"][i, :len(w)] = w
        out["gt_cnt"][i] = len(w)
    if has_mr:
        P = max(max(len(d["pred_relevant_windows"]) for d in submission), 1)
        out["pred_win"] = np.zeros((Q, P, 3))
        out["pred_cnt"] = np.zeros(Q, np.int32)
        for i, d in enumerate(submission):
            w = np.asarray(d["pred_relevant_windows"], dtype=np.float64).reshape(-1, 3)
            if len(w) == 0:
                raise ValueError(f"qid {d['qid']}: no predicted windows (the reference indexes the first one)")
            out["pred_win"][i, :len(w)] = w
            out["pred_cnt"][i] = len(w)
    if has_hl:
        L = max(max(len(d["pred_saliency_scores"]) for d in submission), 1)
        C_ = max(max(int(g["duration"] / clip_length) for g in gts), 1)
        out["pred_sal"] = np.zeros((Q, L))
        out["pred_sal_len"] = np.zeros(Q, np.int32)
        out["gt_sal"] = np.zeros((Q, C_, 3), np.uint8)
        out["gt_clips"] = np.zeros(Q, np.int32)
        for i, (d, g) in enumerate(zip(submission, gts)):
            s = np.asarray(d["pred_saliency_scores"], dtype=np.float64)
            out["pred_sal"][i, :len(s)] = s
            out["pred_sal_len"][i] = len(s)
            out["gt_clips"][i] = int(g["duration"] / clip_length)
            if len(g.get("relevant_clip_ids", [])):          # mk_gt_scores, eval.py:239-246
                out["gt_sal"][i, np.asarray(g["relevant_clip_ids"])] = np.asarray(g["saliency_scores"], np.uint8)
    return out


def pack_ground_truth(ground_truth, qids, clip_length: int = 2, device="cuda"):
    """Ground-truth rows (jsonl dicts) of the queries `qids`, in that order -> device tensors for
    `eval_predictions`.  Done once per dataset split; the per-epoch evaluation then needs no host work."""
    gt_by_qid = {d["qid"]: d for d in ground_truth}
    rows = [gt_by_qid[q] for q in qids]
    Q = len(rows)
    G = max(max(len(g["relevant_windows"]) for g in rows), 1)
    C_ = max(max(int(g["duration"] / clip_length) for g in rows), 1)
    gt_win = np.zeros((Q, G, 2))
    gt_cnt = np.zeros(Q, np.int32)
    gt_sal = np.zeros((Q, C_, 3), np.uint8)
    gt_clips = np.zeros(Q, np.int32)
    for i, g in enumerate(rows):
        w = np.asarray(g["relevant_windows"], dtype=np.float64).reshape(-1, 2)
        gt_win[i, :len(w)] = w
        gt_cnt[i] = len(w)
        gt_clips[i] = int(g["duration"] / clip_length)
        if len(g.get("relevant_clip_ids", [])):
            gt_sal[i, np.asarray(g["relevant_clip_ids"])] = np.asarray(g["saliency_scores"], np.uint8)
    return {k: torch.from_numpy(v).to(device) for k, v in
            dict(gt_win=gt_win, gt_cnt=gt_cnt, gt_sal=gt_sal, gt_clips=gt_clips).items()}


def eval_predictions(windows: torch.Tensor, counts: torch.Tensor, saliency: Optional[torch.Tensor],
                     sal_len: Optional[torch.Tensor], gt: dict, round_4dp: bool = True) -> OrderedDict:
    """Metrics straight from the device-resident outputs of `FlashVTGB200.infer` - ranked windows
    [Q][P][3] (start, end, score) with their counts, saliency [Q][L] with the clip counts - against
    `pack_ground_truth(...)`: what `eval_epoch` gets from writing the submission rows and calling
    `eval_submission` (inference.py:355-385), without the predictions leaving the GPU.  `windows` are the
    POST-PROCESSED rows (FvtgResult.windows / nms_windows): their start / end already went through the compose
    rounding and round_to_multiple_clip_lengths and are used as they are (postprocessing.py:31-33 re-rounds the
    score only - re-rounding start / end would move them when clip_length is not a 4-decimal number, e.g.
    0.166666 for Charades-VGG); round_4dp applies the submission's 4-decimal rounding to the scores and the
    saliency (inference.py:286-290, 318-320)."""
    w = windows.to(torch.float64)
    sal = None if saliency is None else saliency.to(torch.float64)
    if round_4dp:
        w = torch.cat([w[..., :2], torch.round(w[..., 2:3] * 1e4) / 1e4], -1)
        if sal is not None:
            sal = torch.round(sal * 1e4) / 1e4
    mr, hl = eval_arrays(w, counts, gt["gt_win"], gt["gt_cnt"], sal, sal_len,
                         gt["gt_sal"] if sal is not None else None, gt["gt_clips"] if sal is not None else None)
    return format_metrics(mr, hl)


def eval_submission(submission, ground_truth, verbose: bool = True, match_number: bool = True,
                    device: str = "cuda"):
    """Same contract as standalone_eval.eval.eval_submission (eval.py:271-345)."""
    pred_qids = {e["qid"] for e in submission}
    gt_qids = {e["qid"] for e in ground_truth}
    if match_number:
        assert pred_qids == gt_qids, \
            "qids in ground_truth and submission must match. " \
            "use `match_number=False` if you wish to disable this check"
    else:
        shared = pred_qids & gt_qids
        submission = [e for e in submission if e["qid"] in shared]
        ground_truth = [e for e in ground_truth if e["qid"] in shared]
    a = pack_submission(submission, ground_truth)
    t = {k: torch.from_numpy(v).to(device) for k, v in a.items()}
    mr, hl = eval_arrays(t.get("pred_win"), t.get("pred_cnt"), t["gt_win"], t["gt_cnt"], t.get("pred_sal"),
                         t.get("pred_sal_len"), t.get("gt_sal"), t.get("gt_clips"))
    out = format_metrics(mr, hl)
    if verbose:
        print({k: v for k, v in out["brief"].items()})
    return out
