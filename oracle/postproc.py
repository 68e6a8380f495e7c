"""TEST INFRASTRUCTURE ONLY - CPU restatement of the list-based post-processing after the model.

numpy, fp32 where the reference computes in fp32 tensors, fp64 where it uses Python floats.
Pinned against the unmodified reference functions by tests/golden/make_golden.py (fixtures) and
tests/test_oracle_vs_reference.py; temporal IoU also against the docstring known answers of
FlashVTG/span_utils.py:53-59.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def round4(x):
    """float(f"{e:.4f}") element-wise (inference.py:268,288-290): fp32 -> exact double ->
    correctly rounded 4-decimal string -> double.  x * 1e4 is exact in fp64 for fp32 x, so
    rint (half-to-even) / 1e4 is the same correctly rounded value."""
    x = np.asarray(x, dtype=np.float64)
    return np.rint(x * 1e4) / 1e4


def compose_windows(boundary, duration):
    """inference.py:286-290: clamp ALL three columns to [0, duration], then 4-dp every number.
    boundary fp32 (n,3) -> fp64 (n,3) python-float values."""
    b = np.clip(np.asarray(boundary, dtype=F32), F32(0), F32(duration))
    return round4(b)


def post_process(windows, clip_len, clip_ts=True, min_ts=0.0, max_ts=150.0, round_multiple=True):
    """PostProcessorDETR.__call__ (postprocessing.py:25-50) on one query.
    windows: (n,3) python-float values -> torch.tensor(...) is fp32.  Returns fp64 (n,3) holding
    the .tolist() values (fp32-exact windows, 4-dp score)."""
    w = np.asarray(windows, dtype=np.float64).astype(F32)
    se = w[:, :2]
    if clip_ts:
        se = np.clip(se, F32(min_ts), F32(max_ts))
    if round_multiple:
        se = (np.rint(se / F32(clip_len)) * F32(clip_len)).astype(F32)
    out = np.concatenate([se.astype(np.float64), round4(w[:, 2:3])], axis=1)
    return out


def temporal_iou_f32(a, b):
    """nncore.ops.temporal_iou / span_utils.temporal_iou (span_utils.py:61-70): a (2,), b (m,2)."""
    a = np.asarray(a, F32)
    b = np.asarray(b, F32)
    area1 = a[1] - a[0]
    area2 = b[:, 1] - b[:, 0]
    inter = np.maximum(np.minimum(a[1], b[:, 1]) - np.maximum(a[0], b[:, 0]), F32(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / ((area1 + area2) - inter)).astype(F32)


def nms_reference_order(windows, thd, mode="normal"):
    """post_processing_mr_nms for one query (inference.py:36-57), fp32.

    Returns (out (n,3) fp32 in final order, order (n,) index into the input rows,
    sel (n,) the selection sequence = source index swapped into position i at step i).
    Final sort: descending score, ties broken by position after the selection loop (the reference
    uses an unstable sort; this is the documented contract)."""
    bnd = np.array(windows, dtype=np.float64).astype(F32).reshape(-1, 3).copy()
    n = bnd.shape[0]
    src = np.arange(n)
    for i in range(n):
        m = int(np.argmax(bnd[i:, 2])) + i      # first index among ties, NaN counts as max
        if np.isnan(bnd[i:, 2]).any():           # torch.argmax returns the first NaN
            m = int(np.flatnonzero(np.isnan(bnd[i:, 2]))[0]) + i
        if m != i:
            bnd[[i, m]] = bnd[[m, i]]
            src[[i, m]] = src[[m, i]]
        if i + 1 < n:
            iou = temporal_iou_f32(bnd[i, :2], bnd[i + 1:, :2])
            if mode == "normal":
                sup = iou >= F32(thd)            # NaN compares false -> kept
                bnd[i + 1:, 2][sup] = 0
            elif mode == "linear":
                bnd[i + 1:, 2] = (bnd[i + 1:, 2] * (F32(1) - iou)).astype(F32)
            else:
                raise ValueError(f"Unknown nms_type: {mode}")
    sel = src.copy()
    key = bnd[:, 2].copy()
    # descending stable sort; NaN sorts first (torch.sort descending puts NaN first)
    nanmask = np.isnan(key)
    keyc = np.where(nanmask, np.inf, key)
    order = np.argsort(-keyc, kind="stable")
    return bnd[order], src[order], sel


def temporal_nms_hull(preds, thd, max_after_nms=100):
    """utils/temporal_nms.py:25-74 in fp64 (python floats): hull 'union', strict >, removal.
    Returns (kept rows, kept source indices)."""
    preds = [list(map(float, p)) for p in preds]
    if len(preds) == 1:
        return preds, [0]
    idx = sorted(range(len(preds)), key=lambda i: preds[i][2], reverse=True)  # stable, like sorted()
    ts = [preds[i][0] for i in idx]
    te = [preds[i][1] for i in idx]
    sc = [preds[i][2] for i in idx]
    src = list(idx)
    out, out_src = [], []

    def iou(i, j):
        inter = max(0, min(te[i], te[j]) - max(ts[i], ts[j]))
        union = max(te[i], te[j]) - min(ts[i], ts[j])
        return 0 if union == 0 else 1.0 * inter / union

    while len(ts) > 1 and len(out) < max_after_nms:
        j = 1
        while j < len(ts):
            if iou(0, j) > thd:
                ts.pop(j), te.pop(j), sc.pop(j), src.pop(j)
            else:
                j += 1
        out.append([ts.pop(0), te.pop(0), sc.pop(0)])
        out_src.append(src.pop(0))
    if len(out) < max_after_nms and len(ts) >= 1:
        out.append([ts.pop(0), te.pop(0), sc.pop(0)])
        out_src.append(src.pop(0))
    return out, out_src


def full_postproc(boundary, duration, clip_len, clip_ts, min_ts, max_ts, round_multiple,
                  nms_thd=0.7, nms_mode="normal"):
    """boundary (n,3) fp32 -> (windows fp64 (n,3), nms_windows fp32 (n,3), nms_order (n,))."""
    w = post_process(compose_windows(boundary, duration), clip_len, clip_ts, min_ts, max_ts,
                     round_multiple)
    out, order, _ = nms_reference_order(w, nms_thd, nms_mode)
    return w, out, order
