"""CPU: the oracle (oracle/) against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known-answer docstrings."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, c_nms_f32, c_nms_hull, denan, load_forward_index, max_rel, regen_case
from oracle import forward as O
from oracle import postproc as P


@pytest.mark.parametrize("entry", load_forward_index(), ids=lambda e: e["file"][:-4])
def test_forward_oracle_matches_reference_golden(entry):
    cfg, sd, batch, gold = regen_case(entry)
    outs = O.forward_batch(sd, cfg, batch)
    for b, o in enumerate(outs):
        assert max_rel(o["saliency"], gold[f"saliency_{b}"]) < 2e-5
        assert max_rel(o["t2vattn"], gold[f"t2vattn_{b}"]) < 2e-5
        assert max_rel(o["logit"], gold[f"logit_{b}"]) < 2e-5
        assert max_rel(o["coord"], gold[f"coord_{b}"]) < 2e-5
        assert o["boundary"].shape == gold[f"boundary_{b}"].shape
        assert max_rel(o["boundary"], gold[f"boundary_{b}"]) < 2e-5
        pt = gold[f"point_{b}"]
        np.testing.assert_array_equal(o["point_t"].numpy(), pt[:, 0])
        np.testing.assert_array_equal(o["point_s"].numpy(), pt[:, 3])
        if f"dummy_tokens_{b}" in gold:
            assert max_rel(o["dummy_tokens"], gold[f"dummy_tokens_{b}"]) < 2e-5
            assert max_rel(torch.relu(o["video_emb"]), gold[f"video_emb_relu_{b}"]) < 2e-5


def test_temporal_iou_known_answers():
    # FlashVTG/span_utils.py:53-59 docstring
    s1 = [[0, 0.2], [0.5, 1.0]]
    s2 = np.array([[0, 0.3], [0.0, 1.0]], np.float32)
    exp = np.array([[0.6667, 0.2], [0.0, 0.5]])
    for i, a in enumerate(s1):
        got = P.temporal_iou_f32(a, s2)
        np.testing.assert_allclose(got, exp[i], atol=5e-5)


def _nms_cases():
    with open(os.path.join(GOLDEN, "nms_cases.json")) as f:
        return json.load(f)


def _same_rows(a, b):
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_nms_oracles_match_reference_golden():
    cases = _nms_cases()
    assert len(cases) >= 60
    for rec in cases:
        w = denan(rec["windows"])
        for mode in ("normal", "linear"):
            for thd in (0.7, 0.5):
                gold = np.array(denan(rec[f"{mode}_{thd}"]), np.float64).reshape(-1, 3)
                out, order, _ = P.nms_reference_order(w, thd, mode)
                cout, corder, _ = c_nms_f32(w, thd, mode)
                np.testing.assert_array_equal(order, corder)
                assert _same_rows(out, cout)
                # reference sort is unstable: rows must agree wherever scores are distinct
                gs = gold[:, 2]
                assert np.array_equal(np.sort(gs)[::-1], np.sort(out[:, 2].astype(np.float64))[::-1],
                                      equal_nan=True)
                distinct = np.array([np.sum(gs == s) == 1 for s in gs])
                assert _same_rows(gold[distinct], out.astype(np.float64)[distinct])
        for thd, mx in ((0.7, 100), (0.5, 5)):
            gold = rec[f"hull_{thd}_{mx}"]
            out, _ = P.temporal_nms_hull(w, thd, mx)
            cout, _ = c_nms_hull(w, thd, mx)
            assert _same_rows(gold, out)
            assert _same_rows(gold, cout)


def test_postproc_oracle_matches_reference_golden():
    from flashvtg_b200.config import PRESETS, postprocessor_preset
    with open(os.path.join(GOLDEN, "postproc_cases.json")) as f:
        cases = json.load(f)
    for rec in cases:
        cfg = PRESETS[rec["preset"]]
        clip_ts, mn, mx, rnd = postprocessor_preset(cfg)
        comp = P.compose_windows(np.array(rec["boundary"], np.float32), rec["duration"])
        np.testing.assert_array_equal(comp, np.array(rec["composed"], np.float64))
        proc = P.post_process(comp, cfg.clip_length, clip_ts, mn, mx, rnd)
        np.testing.assert_array_equal(proc, np.array(rec["processed"], np.float64))


def test_round4_ties_half_even():
    x = np.array([0.03125, 0.09375, 1.03125, 2.5e-5], np.float32)
    exp = [float(f"{float(v):.4f}") for v in x]
    np.testing.assert_array_equal(P.round4(x), np.array(exp))
