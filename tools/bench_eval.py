"""Throughput of the device evaluation (fvtg_eval_submission) on the reference's sample submission,
replicated to a large query count, beside the CPU restatement on the box's host cores.  Prints one JSON line.

    python tools/bench_eval.py [--replicate 64] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flashvtg_b200 import evaluation as ev  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicate", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cpu-queries", type=int, default=400)
    args = ap.parse_args()
    from test_eval_metrics import load_sample
    a, _, _ = load_sample()
    dev = "cuda:0"
    keys = ("pred_win", "pred_cnt", "gt_win", "gt_cnt", "pred_sal", "pred_sal_len", "gt_sal", "gt_clips")
    host = {k: torch.from_numpy(np.ascontiguousarray(np.concatenate([a[k]] * args.replicate))).pin_memory()
            for k in keys}
    Q = host["gt_cnt"].shape[0]
    t = {k: v.to(dev) for k, v in host.items()}

    def run(tt):
        return ev.eval_arrays(tt["pred_win"], tt["pred_cnt"], tt["gt_win"], tt["gt_cnt"], tt["pred_sal"],
                              tt["pred_sal_len"], tt["gt_sal"], tt["gt_clips"])

    for _ in range(3):
        run(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mr, hl = run(t)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / args.steps
    # end to end: pinned host arrays -> device -> per-query results -> formatted metrics on the host
    t0 = time.perf_counter()
    for _ in range(3):
        tt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        mr, hl = run(tt)
        out = ev.format_metrics(mr, hl)
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    # CPU restatement (the checker) on a bounded sample, single core like the reference's per-worker loop
    from oracle import eval_metrics as om
    n = args.cpu_queries
    sub = {k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    t0 = time.perf_counter()
    om.assemble(om.mr_per_query(sub), om.hl_per_query(sub))
    cpu_s = time.perf_counter() - t0
    in_bytes = sum(v.numel() * v.element_size() for v in host.values())
    print(json.dumps({
        "metric": "queries/sec QVHighlights eval_submission (MR mAP/R1/mIoU x 4 length ranges + HL mAP/Hit1 x 3)",
        "value": Q / dev_ms * 1e3, "unit": "queries/s", "queries": Q, "ms_per_step": dev_ms,
        "e2e": {"value": Q / e2e_ms * 1e3, "unit": "queries/s", "ms": e2e_ms, "h2d_bytes": in_bytes,
                "d2h_bytes": int(Q * (4 * 10 * 8 + 4 * 8 + 4 + 9 * 8 + 3))},
        "cpu_baseline": {"value": n / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port",
                         "sample": f"{n} queries of the sample submission, oracle/eval_metrics.py, {cpu_s:.1f} s"},
        "input_GBps": in_bytes / dev_ms / 1e6,
        "brief": out["brief"],
    }))


if __name__ == "__main__":
    main()
