"""Highlight-detection path (data/HD.py: one pyramid level; compute_hl_results, inference.py:118-229)."""
import numpy as np
import pytest
import torch


def _ref_ap(pred, label, topk):
    """The loop of FlashVTG/inference.py:166-186 restated literally (torch argsort, running sums)."""
    inds = torch.argsort(torch.as_tensor(pred), descending=True, dim=-1)
    cur_label = torch.as_tensor(label, dtype=torch.float32)[inds].tolist()
    if topk is not None:
        cur_label = cur_label[:topk]
    num_gt = sum(cur_label)
    if num_gt == 0:
        return 0
    hits = ap = rec = 0
    prc = 1
    for j, gt in enumerate(cur_label):
        hits += gt
        _rec = hits / num_gt
        _prc = hits / (j + 1)
        ap += (_rec - rec) * (prc + _prc) / 2
        rec, prc = _rec, _prc
    return ap


def test_highlight_ap_matches_reference_loop():
    from flashvtg_b200.postprocessing import highlight_ap
    rng = np.random.Generator(np.random.PCG64(1))
    for n in (1, 4, 5, 37, 200):
        for _ in range(20):
            pred = rng.standard_normal(n).astype(np.float32)
            label = (rng.random(n) > 0.6).astype(np.float32)
            for topk in (5, None):
                assert abs(highlight_ap(pred, label, topk) - _ref_ap(pred, label, topk)) < 1e-12
    assert highlight_ap([0.3, 0.1], [0, 0], 5) == 0.0


def test_hd_presets():
    from flashvtg_b200.config import PRESETS
    for name in ("tvsum", "youtube_uni"):
        cfg = PRESETS[name]
        assert cfg.strides == (1,) and cfg.num_points(120) == 120 and cfg.buffer_size == 2048


@pytest.mark.gpu
@pytest.mark.parametrize("preset,B,Lv,Lt", [("tvsum", 3, 180, 5), ("youtube_uni", 2, 97, 3)])
def test_highlight_forward_matches_oracle(preset, B, Lv, Lt):
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.model import FlashVTGB200
    from helpers import max_rel
    from oracle import forward as O
    cfg = PRESETS[preset]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    batch = synth.make_inputs(cfg, B, Lv, Lt, seed=77, ragged=True, min_lv=Lv // 2, min_lt=1)
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    r = m.infer(batch["src_vid"].to(dev), batch["vid_len"].to(dev), batch["src_txt"].to(dev),
                batch["txt_len"].to(dev), nms=None, want_heads=True)
    torch.cuda.synchronize()
    outs = O.forward_batch(sd, cfg, batch)
    x = float(sd["x"])
    for b, o in enumerate(outs):
        lv = int(batch["vid_len"][b])
        assert max_rel(r.saliency[b, :lv].cpu().numpy(), o["saliency"].numpy()) < 1e-2
        n = o["logit"].shape[0]
        logit = x * r.cls_logit[b, :n].cpu() + (1 - x) * r.conf_logit[b, :n].cpu()
        assert max_rel(torch.sigmoid(logit).numpy(), o["score"].numpy()) < 1e-2


@pytest.mark.gpu
def test_compute_hl_results_runs_and_ranks_like_the_oracle():
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.model import FlashVTGB200
    from flashvtg_b200.postprocessing import compute_hl_results, highlight_ap
    from oracle import forward as O
    cfg = PRESETS["tvsum"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    batch = synth.make_inputs(cfg, 2, 60, 4, seed=5, ragged=False)
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    rng = np.random.Generator(np.random.PCG64(2))
    metas = [{"label": rng.integers(1, 6, size=(60, 20)).tolist()} for _ in range(2)]
    inp = {k: batch[k].to(dev) for k in ("src_vid", "src_vid_mask", "src_txt", "src_txt_mask")}
    got = compute_hl_results(m, [(metas, inp)])
    outs = O.forward_batch(sd, cfg, batch)
    aps = []
    for meta, o in zip(metas, outs):
        lab = np.asarray(meta["label"], dtype=np.float64)
        va = []
        for a in range(20):
            cur = torch.tensor(lab[:, a])
            va.append(highlight_ap(o["saliency"].numpy(), (cur > cur.median()).double().numpy(), 5))
        aps.append(va)
    assert abs(got["mAP"] - round(float(np.mean(aps)), 5)) < 0.05   # bf16 saliency may swap near-ties


@pytest.mark.gpu
def test_youtube_uni_video_without_positive_label_is_left_out_of_the_mean():
    """inference.py:197-200: the youtube_uni branch `continue`s before video_ap_collected.append when a video has no
    positive label, so that video does not count as AP 0 - it is dropped from the mean."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.model import FlashVTGB200
    from flashvtg_b200.postprocessing import compute_hl_results, highlight_ap
    cfg = PRESETS["youtube_uni"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    batch = synth.make_inputs(cfg, 3, 40, 4, seed=6, ragged=False)
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    rng = np.random.Generator(np.random.PCG64(3))
    metas = [{"label": (rng.random((40, 1)) > 0.5).astype(np.float64).tolist()},
             {"label": np.zeros((40, 1)).tolist()},                                  # no positive clip
             {"label": (rng.random((40, 1)) > 0.5).astype(np.float64).tolist()}]
    inp = {k: batch[k].to(dev) for k in ("src_vid", "src_vid_mask", "src_txt", "src_txt_mask")}
    got = compute_hl_results(m, [(metas, inp)])
    r = m.infer(inp["src_vid"], inp["src_vid_mask"].sum(1).int(), inp["src_txt"], inp["src_txt_mask"].sum(1).int(),
                nms=None)
    sal = r.saliency.cpu().numpy()
    want = np.mean([[highlight_ap(sal[i], np.asarray(metas[i]["label"]).reshape(-1))] for i in (0, 2)])
    assert got["mAP"] == round(float(want), 5)
    assert compute_hl_results(m, [([metas[1]], {k: v[1:2] for k, v in inp.items()})]) == dict(mAP=0.0)
