"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's QVHighlights evaluation
(standalone_eval/eval.py + standalone_eval/utils.py), the oracle of flashvtg_b200/evaluation.py and
csrc/evalmetrics.cu.  Works on the array form both sides share (see `pack`), in float64 like numpy.

  cross / paired IoU            utils.py:16-63 (paired: "union" = max end - min start)
  detection AP per query        utils.py:83-159 (score-sorted predictions, greedy GT locking per IoU
                                threshold) + interpolated_precision_recall utils.py:65-80
  MR mAP / R1 / mIoU            eval.py:24-102 (first 10 predicted windows; R1 uses the FIRST listed
                                window against the GT window with the highest IoU)
  length ranges                 eval.py:109-170 (short (0,10], middle (10,30], long (30,150], full)
  HL Hit1 / mAP                 eval.py:173-268 (3 annotators, minimum score 2/3/4) + get_ap
                                utils.py:162-209 on sklearn's precision_recall_curve (restated here:
                                distinct-score thresholds, no truncation at full recall - the points a
                                pre-1.1 scikit-learn dropped never enter the average)
  formatting                    float(f"{100 * v:.2f}") everywhere, eval.py:67-69,98-101,186,214

Pinned: tests/golden/make_eval_golden.py runs the UNMODIFIED reference eval_submission on the
reference's own sample submission and checks this restatement against it and against
sample_val_preds_metrics_raw.json; the packed sample and the expected metrics are committed as
tests/golden/eval_sample.npz / eval_sample_metrics.json.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

AP_THDS = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]
R1_THDS = [float(f"{e:.2f}") for e in np.linspace(0.3, 0.95, 14)]
LENGTH_RANGES = (("short", 0, 10), ("middle", 10, 30), ("long", 30, 150), ("full", 0, 150))
HL_MINS = (("Fair", 2), ("Good", 3), ("VeryGood", 4))


def pack(submission, ground_truth, clip_length=2):
    """Lists of dicts (the reference's jsonl rows) -> aligned arrays, one row per submission entry."""
    gt_by_qid = {d["qid"]: d for d in ground_truth}
    Q = len(submission)
    has_mr = "pred_relevant_windows" in submission[0]
    has_hl = "pred_saliency_scores" in submission[0]
    P = max([len(d["pred_relevant_windows"]) for d in submission] + [1]) if has_mr else 1
    G = max([len(gt_by_qid[d["qid"]]["relevant_windows"]) for d in submission] + [1])
    L = max([len(d["pred_saliency_scores"]) for d in submission] + [1]) if has_hl else 1
    C = max([int(gt_by_qid[d["qid"]]["duration"] / clip_length) for d in submission] + [1])
    out = dict(
        qid=np.zeros(Q, np.int64),
        pred_win=np.zeros((Q, P, 3)), pred_cnt=np.zeros(Q, np.int32),
        gt_win=np.zeros((Q, G, 2)), gt_cnt=np.zeros(Q, np.int32),
        pred_sal=np.zeros((Q, L)), pred_sal_len=np.zeros(Q, np.int32),
        gt_sal=np.zeros((Q, C, 3), np.uint8), gt_clips=np.zeros(Q, np.int32),
    )
    for i, d in enumerate(submission):
        g = gt_by_qid[d["qid"]]
        out["qid"][i] = d["qid"]
        if has_mr:
            w = np.asarray(d["pred_relevant_windows"], dtype=np.float64).reshape(-1, 3)
            out["pred_win"][i, :len(w)] = w
            out["pred_cnt"][i] = len(w)
        gw = np.asarray(g["relevant_windows"], dtype=np.float64).reshape(-1, 2)
        out["gt_win"][i, :len(gw)] = gw
        out["gt_cnt"][i] = len(gw)
        if has_hl:
            s = np.asarray(d["pred_saliency_scores"], dtype=np.float64)
            out["pred_sal"][i, :len(s)] = s
            out["pred_sal_len"][i] = len(s)
        n_clips = int(g["duration"] / clip_length)
        out["gt_clips"][i] = n_clips
        if "relevant_clip_ids" in g and len(g["relevant_clip_ids"]):
            out["gt_sal"][i, np.asarray(g["relevant_clip_ids"])] = np.asarray(g["saliency_scores"], np.uint8)
    out["has_mr"], out["has_hl"] = has_mr, has_hl
    return out


def iou_cross(span, gts):
    left = np.maximum(span[0], gts[:, 0])
    right = np.minimum(span[1], gts[:, 1])
    inter = np.clip(right - left, 0, None)
    union = (span[1] - span[0]) + (gts[:, 1] - gts[:, 0]) - inter
    return inter / union


def iou_paired(p, g):
    inter = max(0.0, min(p[1], g[1]) - max(p[0], g[0]))
    union = max(p[1], g[1]) - min(p[0], g[0])
    return inter / union if union != 0 else 0.0


def interpolated_ap(precision, recall):
    mp = np.hstack([[0], precision, [0]])
    mr = np.hstack([[0], recall, [1]])
    for i in range(len(mp) - 1)[::-1]:
        mp[i] = max(mp[i], mp[i + 1])
    idx = np.where(mr[1:] != mr[:-1])[0] + 1
    return np.sum((mr[idx] - mr[idx - 1]) * mp[idx])


def detection_ap(gt, pred, thds=AP_THDS):
    """gt (G, 2), pred (P, 3) in listed order -> AP per IoU threshold."""
    n_thd, G, P = len(thds), len(gt), len(pred)
    ap = np.zeros(n_thd)
    if P == 0:
        return ap
    order = sorted(range(P), key=lambda i: -pred[i][2])           # stable, like list.sort
    lock = -np.ones((n_thd, G))
    tp = np.zeros((n_thd, P))
    fp = np.zeros((n_thd, P))
    for idx, pi in enumerate(order):
        if G == 0:
            fp[:, idx] = 1
            continue
        tiou = iou_cross(pred[pi, :2], gt)
        by_iou = tiou.argsort()[::-1]
        for t, thd in enumerate(thds):
            for j in by_iou:
                if tiou[j] < thd:
                    fp[t, idx] = 1
                    break
                if lock[t, j] >= 0:
                    continue
                tp[t, idx] = 1
                lock[t, j] = idx
                break
            if fp[t, idx] == 0 and tp[t, idx] == 0:
                fp[t, idx] = 1
    tpc = np.cumsum(tp, axis=1).astype(float)
    fpc = np.cumsum(fp, axis=1).astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = tpc / float(G)
        prec = tpc / (tpc + fpc)
    for t in range(n_thd):
        ap[t] = interpolated_ap(prec[t], rec[t])
    return ap


def range_windows(gt_win, lo, hi):
    if lo == 0 and hi == 150:
        return gt_win
    ln = gt_win[:, 1] - gt_win[:, 0]
    return gt_win[(lo < ln) & (ln <= hi)]


def mr_per_query(a, max_pred_windows=10):
    """-> ap [4][Q][10], iou [4][Q], valid [4][Q] (query has a GT window in the range)."""
    Q = len(a["pred_cnt"])
    ap = np.zeros((4, Q, 10))
    iou = np.zeros((4, Q))
    valid = np.zeros((4, Q), bool)
    for r, (_, lo, hi) in enumerate(LENGTH_RANGES):
        for i in range(Q):
            gt = range_windows(a["gt_win"][i, :a["gt_cnt"][i]], lo, hi)
            if len(gt) == 0:
                continue
            valid[r, i] = True
            n = int(a["pred_cnt"][i])
            pred = a["pred_win"][i, :n]
            ap[r, i] = detection_ap(gt, pred[:max_pred_windows])
            top = pred[0, :2]
            best = int(np.argmax(iou_cross(top, gt)))
            iou[r, i] = iou_paired(top, gt[best])
    return ap, iou, valid


def pr_curve(y_true, y_score):
    """sklearn.metrics.precision_recall_curve for binary 0/1 labels, no sample weights."""
    order = np.argsort(y_score, kind="mergesort")[::-1]
    ys, yt = y_score[order], y_true[order]
    distinct = np.where(np.diff(ys))[0]
    idxs = np.r_[distinct, len(yt) - 1]
    tps = np.cumsum(yt.astype(np.float64))[idxs]
    fps = 1 + idxs.astype(np.float64) - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return np.r_[precision[::-1], 1.0], np.r_[recall[::-1], 0.0]


def highlight_ap(y_true, y_predict):
    if len(set(y_true)) == 1:
        return 0 if y_true[0] == 0 else 1
    precision, recall = pr_curve(y_true, y_predict)
    recall = recall.astype(np.float32)
    for i in range(1, len(precision)):
        precision[i] = max(precision[i - 1], precision[i])
    return np.mean(precision[np.where(np.diff(recall))])


def hl_per_query(a):
    """-> ap [3][Q][3] (min score, query, annotator), hit [3][Q]."""
    Q = len(a["gt_clips"])
    ap = np.zeros((3, Q, 3))
    hit = np.zeros((3, Q))
    for m, (_, smin) in enumerate(HL_MINS):
        for i in range(Q):
            n = int(a["gt_clips"][i])
            binary = (a["gt_sal"][i, :n] >= smin).astype(float)
            pred = a["pred_sal"][i, :a["pred_sal_len"][i]]
            top = int(np.argmax(pred))
            if top < n:
                hit[m, i] = binary[top].max() if n else 0.0
            y = np.zeros(n)
            k = min(n, len(pred))
            y[:k] = pred[:k]
            for w in range(3):
                ap[m, i, w] = highlight_ap(binary[:, w], y)
    return ap, hit


def _pct(v):
    return float(f"{100 * v:.2f}")


def assemble(mr=None, hl=None, n_total=None):
    """Per-query results -> the reference's metrics dict (eval.py:271-345).  Shared with the product's
    host wrapper ONLY as a specification: flashvtg_b200/evaluation.py carries its own copy."""
    metrics, brief = {}, OrderedDict()
    if mr is not None:
        ap, iou, valid = mr
        for r, (name, _, _) in enumerate(LENGTH_RANGES):
            v = valid[r]
            if not v.any():
                dummy = {k: 0. for k in np.linspace(0.5, 0.95, 19)}
                dummy["average"] = 0.
                metrics[name] = {"MR-mAP": dummy, "MR-R1": dummy}
                continue
            ap_thds = ap[r][v].mean(0)
            d_ap = dict(zip([str(e) for e in AP_THDS], ap_thds))
            d_ap["average"] = np.mean(ap_thds)
            d_ap = {k: _pct(x) for k, x in d_ap.items()}
            ious = iou[r][v]
            d_r1 = {str(t): _pct(np.mean(ious >= t)) for t in R1_THDS}
            metrics[name] = {"MR-mIoU": _pct(np.mean(ious)), "MR-mAP": d_ap, "MR-R1": d_r1}
        f = metrics["full"]
        b = {"MR-full-mAP": f["MR-mAP"]["average"], "MR-full-mAP@0.5": f["MR-mAP"]["0.5"],
             "MR-full-mAP@0.75": f["MR-mAP"]["0.75"], "MR-short-mAP": metrics["short"]["MR-mAP"]["average"],
             "MR-middle-mAP": metrics["middle"]["MR-mAP"]["average"],
             "MR-long-mAP": metrics["long"]["MR-mAP"]["average"], "MR-full-mIoU": f["MR-mIoU"],
             "MR-full-R1@0.3": f["MR-R1"]["0.3"], "MR-full-R1@0.5": f["MR-R1"]["0.5"],
             "MR-full-R1@0.7": f["MR-R1"]["0.7"]}
        brief.update(sorted(b.items(), key=lambda x: x[0]))
    if hl is not None:
        ap, hit = hl
        hl_metrics = {}
        for m, (name, _) in enumerate(HL_MINS):
            hl_metrics[f"HL-min-{name}"] = {"HL-mAP": _pct(np.mean(ap[m])), "HL-Hit1": _pct(np.mean(hit[m]))}
        metrics.update(hl_metrics)
        brief.update({f"{k}-{sk.split('-')[1]}": v[sk] for k, v in hl_metrics.items() for sk in v})
    out = OrderedDict()
    out["brief"] = brief
    out.update(sorted(metrics.items(), key=lambda x: x[0]))
    return out


def eval_submission(submission, ground_truth, match_number=True):
    pred_qids = {e["qid"] for e in submission}
    gt_qids = {e["qid"] for e in ground_truth}
    if match_number:
        assert pred_qids == gt_qids, "qids in ground_truth and submission must match. " \
                                     "use `match_number=False` if you wish to disable this check"
    else:
        shared = pred_qids & gt_qids
        submission = [e for e in submission if e["qid"] in shared]
        ground_truth = [e for e in ground_truth if e["qid"] in shared]
    a = pack(submission, ground_truth)
    return assemble(mr_per_query(a) if a["has_mr"] else None, hl_per_query(a) if a["has_hl"] else None)
