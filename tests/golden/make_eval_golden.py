"""Golden vectors of the evaluation row.  Run in the build container, where
/root/reference exists:

    python tests/golden/make_eval_golden.py

1. runs the UNMODIFIED reference `standalone_eval.eval.eval_submission` on the reference's own sample
   submission (standalone_eval/sample_val_preds.jsonl vs data/highlight_val_release.jsonl) and checks the
   result against the reference's published standalone_eval/sample_val_preds_metrics_raw.json;
2. checks the oracle restatement (oracle/eval_metrics.py) against both, per metric;
3. writes the packed sample (arrays, see oracle.eval_metrics.pack) as tests/golden/eval_sample.npz and the
   expected metrics as tests/golden/eval_sample_metrics.json, plus the oracle's per-query results for a
   subset so that the GPU tests can compare element by element without the oracle's run time.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_jsonl(p):
    with open(p) as f:
        return [json.loads(l) for l in f if l.strip()]


def main():
    sys.path.insert(0, REF)
    from standalone_eval.eval import eval_submission as ref_eval  # noqa: E402
    from oracle import eval_metrics as om

    sub = load_jsonl(os.path.join(REF, "standalone_eval/sample_val_preds.jsonl"))
    gt = load_jsonl(os.path.join(REF, "data/highlight_val_release.jsonl"))
    published = json.load(open(os.path.join(REF, "standalone_eval/sample_val_preds_metrics_raw.json")))
    ref = json.loads(json.dumps(ref_eval(sub, gt, verbose=False)))
    ours = json.loads(json.dumps(om.eval_submission(sub, gt)))

    def flat(d, pre=""):
        out = {}
        for k, v in d.items():
            if isinstance(v, dict):
                out.update(flat(v, pre + str(k) + "/"))
            else:
                out[pre + str(k)] = v
        return out

    fr, fo, fp = flat(ref), flat(ours), flat(published)
    assert fr.keys() == fo.keys(), (sorted(set(fr) ^ set(fo)))
    bad = {k: (fr[k], fo[k]) for k in fr if fr[k] != fo[k]}
    assert not bad, f"oracle differs from the reference run: {bad}"
    missing = {k: (fp[k], fr.get(k)) for k in fp if fr.get(k) != fp[k]}
    print(f"reference run vs published json: {len(fp) - len(missing)}/{len(fp)} equal; differing: {missing}")
    print(f"oracle vs reference run: {len(fr)} metrics identical")

    a = om.pack(sub, gt)
    # saliency predictions of the sample are fp16 values, window scores 4-decimal: store compactly, exactly
    sal16 = a["pred_sal"].astype(np.float16)
    assert np.array_equal(sal16.astype(np.float64), a["pred_sal"])
    win_i = np.rint(a["pred_win"] * 1e4).astype(np.int32)
    assert np.array_equal(win_i / 1e4, a["pred_win"])
    gt_i = a["gt_win"].astype(np.int16)
    assert np.array_equal(gt_i.astype(np.float64), a["gt_win"])
    np.savez_compressed(
        os.path.join(ROOT, "tests/golden/eval_sample.npz"),
        qid=a["qid"].astype(np.int32), pred_win_1e4=win_i, pred_cnt=a["pred_cnt"], gt_win=gt_i, gt_cnt=a["gt_cnt"],
        pred_sal_f16=sal16, pred_sal_len=a["pred_sal_len"], gt_sal=a["gt_sal"], gt_clips=a["gt_clips"])
    mr_ap, mr_iou, mr_valid = om.mr_per_query(a)
    hl_ap, hl_hit = om.hl_per_query(a)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/eval_sample_per_query.npz"),
                        mr_ap=mr_ap, mr_iou=mr_iou, mr_valid=mr_valid, hl_ap=hl_ap, hl_hit=hl_hit.astype(np.uint8))
    with open(os.path.join(ROOT, "tests/golden/eval_sample_metrics.json"), "w") as f:
        json.dump({"reference_run": ref, "published": published,
                   "published_keys_not_reproduced": sorted(missing)}, f, indent=1, sort_keys=True)
    for n in ("eval_sample.npz", "eval_sample_per_query.npz", "eval_sample_metrics.json"):
        print(n, os.path.getsize(os.path.join(ROOT, "tests/golden", n)), "bytes")


if __name__ == "__main__":
    main()
