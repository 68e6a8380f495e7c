"""Evaluation row: the oracle restatement of standalone_eval against the golden
sample (CPU), and the device kernels against the oracle (GPU, bit-exact per query)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eval_metrics as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_sample():
    z = np.load(os.path.join(GOLD, "eval_sample.npz"))
    a = dict(qid=z["qid"].astype(np.int64), pred_win=z["pred_win_1e4"] / 1e4, pred_cnt=z["pred_cnt"],
             gt_win=z["gt_win"].astype(np.float64), gt_cnt=z["gt_cnt"],
             pred_sal=z["pred_sal_f16"].astype(np.float64), pred_sal_len=z["pred_sal_len"],
             gt_sal=z["gt_sal"], gt_clips=z["gt_clips"], has_mr=True, has_hl=True)
    exp = json.load(open(os.path.join(GOLD, "eval_sample_metrics.json")))
    pq = np.load(os.path.join(GOLD, "eval_sample_per_query.npz"))
    return a, exp, pq


def flat(d, pre=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(flat(v, pre + str(k) + "/"))
        else:
            out[pre + str(k)] = v
    return out


def random_case(seed, Q=96, P=12, G=5, L=90, C=80):
    """Ragged synthetic submission with ties, short lists, saliency / GT length mismatches."""
    rng = np.random.default_rng(seed)
    a = dict(pred_win=np.zeros((Q, P, 3)), pred_cnt=np.zeros(Q, np.int32), gt_win=np.zeros((Q, G, 2)),
             gt_cnt=np.zeros(Q, np.int32), pred_sal=np.zeros((Q, L)), pred_sal_len=np.zeros(Q, np.int32),
             gt_sal=np.zeros((Q, C, 3), np.uint8), gt_clips=np.zeros(Q, np.int32), has_mr=True, has_hl=True)
    for i in range(Q):
        n = int(rng.integers(1, P + 1))
        st = rng.integers(0, 70, n) * 2.0
        ed = st + rng.integers(1, 40, n) * 2.0
        sc = np.round(rng.random(n), 1 if i % 3 == 0 else 4)      # coarse scores -> ties
        a["pred_win"][i, :n] = np.stack([st, np.minimum(ed, 150.0), sc], 1)
        a["pred_cnt"][i] = n
        g = int(rng.integers(1, G + 1))
        gs = rng.integers(0, 70, g) * 2.0
        ge = gs + rng.choice([2, 4, 8, 10, 12, 20, 30, 32, 60, 100], g)
        if i % 5 == 0:                                               # identical GT windows -> IoU ties
            gs[:] = gs[0]
            ge[:] = ge[0]
        a["gt_win"][i, :g] = np.stack([gs, ge], 1)
        a["gt_cnt"][i] = g
        nc = int(rng.integers(1, C + 1))
        a["gt_clips"][i] = nc
        a["gt_sal"][i, :nc] = rng.integers(0, 5, (nc, 3))
        if i % 7 == 0:
            a["gt_sal"][i, :nc] = 4 if i % 14 == 0 else 0               # all positive / all negative
        lp = int(rng.integers(1, L + 1)) if i % 4 else nc
        sal = rng.standard_normal(lp)
        if i % 3 == 1:
            sal = np.round(sal, 1)                                    # tied saliency scores
        a["pred_sal"][i, :lp] = sal
        a["pred_sal_len"][i] = lp
    return a


def test_oracle_reproduces_reference_sample_metrics():
    a, exp, pq = load_sample()
    mr = om.mr_per_query(a)
    hl = om.hl_per_query(a)
    got = flat(json.loads(json.dumps(om.assemble(mr, hl))))
    ref = flat(exp["reference_run"])
    assert got == ref
    pub = flat(exp["published"])
    assert exp["published_keys_not_reproduced"] == []
    assert all(got[k] == v for k, v in pub.items())
    assert np.array_equal(mr[0], pq["mr_ap"]) and np.array_equal(mr[1], pq["mr_iou"])
    assert np.array_equal(hl[0], pq["hl_ap"]) and np.array_equal(hl[1].astype(np.uint8), pq["hl_hit"])


def test_oracle_against_unmodified_reference_on_random_cases():
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference checkout not present (GPU box)")
    import sys
    sys.path.insert(0, ref_root)
    from standalone_eval.eval import eval_submission as ref_eval
    a = random_case(11, Q=40)
    sub, gt = [], []
    for i in range(len(a["pred_cnt"])):
        nc = int(a["gt_clips"][i])
        sub.append({"qid": i, "pred_relevant_windows": a["pred_win"][i, :a["pred_cnt"][i]].tolist(),
                    "pred_saliency_scores": a["pred_sal"][i, :a["pred_sal_len"][i]].tolist()})
        gt.append({"qid": i, "duration": 2 * nc, "relevant_windows": a["gt_win"][i, :a["gt_cnt"][i]].tolist(),
                   "relevant_clip_ids": list(range(nc)), "saliency_scores": a["gt_sal"][i, :nc].tolist()})
    ref = flat(json.loads(json.dumps(ref_eval(sub, gt, verbose=False))))
    got = flat(json.loads(json.dumps(om.eval_submission(sub, gt))))
    assert got == ref


def _device_eval(a):
    from flashvtg_b200 import evaluation as ev
    dev = "cuda:0"
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in a.items() if isinstance(v, np.ndarray)}
    mr, hl = ev.eval_arrays(t["pred_win"], t["pred_cnt"], t["gt_win"], t["gt_cnt"], t["pred_sal"],
                            t["pred_sal_len"], t["gt_sal"], t["gt_clips"])
    torch.cuda.synchronize()
    return mr, hl


@pytest.mark.gpu
def test_device_metrics_bit_exact_on_reference_sample():
    from flashvtg_b200 import evaluation as ev
    a, exp, pq = load_sample()
    mr, hl = _device_eval(a)
    assert np.array_equal(mr[2].cpu().numpy().astype(bool), pq["mr_valid"])
    assert np.array_equal(mr[0].cpu().numpy(), pq["mr_ap"])          # fp64, bit for bit
    assert np.array_equal(mr[1].cpu().numpy(), pq["mr_iou"])
    assert np.array_equal(hl[0].cpu().numpy(), pq["hl_ap"])
    assert np.array_equal(hl[1].cpu().numpy(), pq["hl_hit"])
    got = flat(json.loads(json.dumps(ev.format_metrics(mr, hl))))
    assert got == flat(exp["reference_run"])
    assert all(got[k] == v for k, v in flat(exp["published"]).items())


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_device_metrics_bit_exact_on_ragged_cases(seed):
    a = random_case(seed)
    mr, hl = _device_eval(a)
    o_ap, o_iou, o_valid = om.mr_per_query(a)
    h_ap, h_hit = om.hl_per_query(a)
    assert np.array_equal(mr[2].cpu().numpy().astype(bool), o_valid)
    assert np.array_equal(mr[0].cpu().numpy(), o_ap)
    assert np.array_equal(mr[1].cpu().numpy(), o_iou)
    assert np.array_equal(hl[0].cpu().numpy(), h_ap)
    assert np.array_equal(hl[1].cpu().numpy().astype(np.float64), h_hit)


@pytest.mark.gpu
def test_eval_submission_dict_api_and_long_videos():
    """jsonl-row API (the reference's call) + more than 128 clips per video (numpy's recursive pairwise sum)."""
    from flashvtg_b200 import evaluation as ev
    a = random_case(5, Q=24, L=300, C=300)
    sub, gt = [], []
    for i in range(len(a["pred_cnt"])):
        nc = int(a["gt_clips"][i])
        sub.append({"qid": 100 + i, "pred_relevant_windows": a["pred_win"][i, :a["pred_cnt"][i]].tolist(),
                    "pred_saliency_scores": a["pred_sal"][i, :a["pred_sal_len"][i]].tolist()})
        gt.append({"qid": 100 + i, "duration": 2 * nc, "relevant_windows": a["gt_win"][i, :a["gt_cnt"][i]].tolist(),
                   "relevant_clip_ids": list(range(nc)), "saliency_scores": a["gt_sal"][i, :nc].tolist()})
    got = flat(json.loads(json.dumps(ev.eval_submission(sub, gt[::-1], verbose=False))))
    exp = flat(json.loads(json.dumps(om.eval_submission(sub, gt))))
    assert got == exp
    with pytest.raises(AssertionError):
        ev.eval_submission(sub[:-1], gt, verbose=False)


@pytest.mark.gpu
def test_eval_predictions_from_device_tensors_equals_the_row_api():
    """Device-resident predictions + packed ground truth give the same dict as the jsonl-row API."""
    from flashvtg_b200 import evaluation as ev
    a = random_case(9, Q=64)
    sub, gt = [], []
    for i in range(len(a["pred_cnt"])):
        nc = int(a["gt_clips"][i])
        sub.append({"qid": 7000 + i, "pred_relevant_windows": a["pred_win"][i, :a["pred_cnt"][i]].tolist(),
                    "pred_saliency_scores": a["pred_sal"][i, :a["pred_sal_len"][i]].tolist()})
        gt.append({"qid": 7000 + i, "duration": 2 * nc, "relevant_windows": a["gt_win"][i, :a["gt_cnt"][i]].tolist(),
                   "relevant_clip_ids": list(range(nc)), "saliency_scores": a["gt_sal"][i, :nc].tolist()})
    exp = flat(json.loads(json.dumps(ev.eval_submission(sub, gt, verbose=False))))
    dev = "cuda:0"
    g = ev.pack_ground_truth(gt[::-1], [d["qid"] for d in sub], device=dev)
    got = ev.eval_predictions(torch.from_numpy(a["pred_win"]).to(dev), torch.from_numpy(a["pred_cnt"]).to(dev),
                              torch.from_numpy(a["pred_sal"]).to(dev), torch.from_numpy(a["pred_sal_len"]).to(dev),
                              g, round_4dp=False)
    assert flat(json.loads(json.dumps(got))) == exp
