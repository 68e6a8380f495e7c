"""TEST INFRASTRUCTURE ONLY - nncore.ops.temporal_iou restated (see ../__init__.py)."""


def temporal_iou(windows1, windows2):
    """IoU matrix (N, M) of 1-D windows [st, ed]: inter / (len1 + len2 - inter)."""
    area1 = windows1[:, 1] - windows1[:, 0]
    area2 = windows2[:, 1] - windows2[:, 0]
    start = windows1[:, None, 0].maximum(windows2[:, 0])
    end = windows1[:, None, 1].minimum(windows2[:, 1])
    inter = (end - start).clamp(min=0)
    return inter / (area1[:, None] + area2 - inter)
