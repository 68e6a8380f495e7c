"""Generates the committed golden fixtures by running the UNMODIFIED reference (imported from
/root/reference through oracle/shim) on seeded synthetic weights and inputs.

Run in the build container only:  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read the .npz / .json files written here and
regenerate weights / inputs from the recorded seeds (checksums in the fixtures pin the RNG).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flashvtg_b200 import synth  # noqa: E402
from flashvtg_b200.config import PRESETS, postprocessor_preset  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (preset, B, Lv, Lt, ragged, spread, keep_big)
FORWARD_CASES = [
    ("qvh_iv2", 3, 75, 32, True, True, True),
    ("qvh_iv2", 1, 75, 29, False, False, False),
    ("qvh_sfclip", 2, 75, 32, True, True, False),
    ("charades_vgg", 2, 61, 8, True, True, False),
    ("charades_iv2", 2, 33, 11, True, True, False),
    ("tacos", 2, 130, 12, True, True, False),
    ("tacos_deep", 1, 201, 9, False, True, False),
]


def run_reference_case(cfg, sd, batch):
    model = ref_loader.build_reference_model(cfg, sd)
    outs = []
    for b in range(batch["src_vid"].shape[0]):
        lv, lt = int(batch["vid_len"][b]), int(batch["txt_len"][b])
        v = batch["src_vid"][b:b + 1, :lv]
        t = batch["src_txt"][b:b + 1, :lt]
        vm = torch.ones(1, lv)
        tm = torch.ones(1, lt)
        # eval path: the ranked top-k boundary exactly as inference.py consumes it
        ev = ref_loader.reference_forward_bs1(model, t, tm, v, vm)
        # same modules, top-level training flag only: exposes the pre-sort tensors
        # (out_class / out_coord / video_emb, model.py:218-226) with dropout still off.
        model.training = True
        with torch.no_grad():
            tr = model(src_txt=t, src_txt_mask=tm, src_vid=v, src_vid_mask=vm, vid=None, qid=None,
                       targets={})
        model.training = False
        outs.append(dict(
            saliency=ev["saliency_scores"][0].numpy(), t2vattn=ev["t2vattnvalues"][0].numpy(),
            dummy_tokens=ev["dummy_tokens"][0].numpy(), boundary=ev["_out"]["boundary"].numpy(),
            logit=tr["out_class"][0, :, 0].numpy(), coord=tr["out_coord"][0].numpy(),
            video_emb_relu=tr["video_emb"][0].numpy(), point=tr["point"].numpy()))
    return outs


def make_forward():
    index = []
    for (name, B, Lv, Lt, ragged, spread, big) in FORWARD_CASES:
        cfg = PRESETS[name]
        wseed, iseed = (2025 if spread else 2024), 1234
        sd = synth.make_state_dict(cfg, wseed, spread=spread)
        batch = synth.make_inputs(cfg, B, Lv, Lt, seed=iseed, ragged=ragged)
        outs = run_reference_case(cfg, sd, batch)
        tag = f"fwd_{name}_B{B}_Lv{Lv}_Lt{Lt}_{'spread' if spread else 'plain'}"
        arrays = {}
        for b, o in enumerate(outs):
            for k, v in o.items():
                if k in ("dummy_tokens", "video_emb_relu") and not big:
                    continue
                arrays[f"{k}_{b}"] = v.astype(np.float32)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **arrays)
        index.append(dict(
            file=tag + ".npz", preset=name, B=B, Lv=Lv, Lt=Lt, ragged=ragged, spread=spread,
            weight_seed=wseed, input_seed=iseed, vid_len=batch["vid_len"].tolist(),
            txt_len=batch["txt_len"].tolist(),
            weights_checksum=synth.state_dict_checksum(sd),
            inputs_checksum=float(batch["src_vid"].double().sum() + batch["src_txt"].double().sum())))
        print("wrote", tag)
    with open(os.path.join(OUT, "forward_index.json"), "w") as f:
        json.dump(index, f, indent=1)


def _edge_cases():
    rng = np.random.Generator(np.random.PCG64(7))
    cases = []
    # zero-length pairs (NaN IoU), exact ties, IoU exactly at the threshold on the 2-s grid
    cases.append([[10.0, 10.0, 0.9], [10.0, 10.0, 0.8], [0.0, 20.0, 0.7], [10.0, 10.0, 0.6]])
    cases.append([[0.0, 20.0, 0.5], [0.0, 20.0, 0.5], [2.0, 20.0, 0.5], [0.0, 14.0, 0.5]])
    cases.append([[0.0, 20.0, 0.9], [0.0, 14.0, 0.8], [6.0, 20.0, 0.7], [0.0, 28.0, 0.6],
                  [0.0, 28.58, 0.55]])
    cases.append([[4.0, 12.0, 0.3]])
    cases.append([[0.0, 150.0, 0.0], [0.0, 150.0, 0.0], [10.0, 20.0, 0.0]])
    for n in (2, 7, 50, 50, 50):
        st = np.round(rng.uniform(0, 140, size=n) / 2) * 2
        ln = np.round(rng.uniform(0, 40, size=n) / 2) * 2
        sc = np.round(rng.uniform(0, 1, size=n), 4)
        if n == 50:
            sc[rng.integers(0, n, size=10)] = sc[0]  # score ties
        cases.append(np.stack([st, np.minimum(st + ln, 150.0), sc], 1).tolist())
    # un-rounded windows (Charades-style clip_len) incl. negative / reversed spans
    for n in (13, 50):
        st = rng.uniform(-2, 30, size=n)
        ed = st + rng.uniform(-1, 12, size=n)
        cases.append(np.stack([st, ed, rng.uniform(0, 1, size=n)], 1).astype(np.float32)
                     .astype(np.float64).tolist())
    return cases


def make_nms():
    nms = ref_loader.reference_nms()
    hull = ref_loader.reference_temporal_nms()
    cases = _edge_cases()
    sample = os.path.join(str(ref_loader.REF_ROOT), "standalone_eval", "sample_val_preds.jsonl")
    with open(sample) as f:
        lines = [json.loads(x) for x in f]
    for e in lines[:60]:
        cases.append(e["pred_relevant_windows"])
    out = []
    for w in cases:
        rec = dict(windows=w)
        for mode in ("normal", "linear"):
            for thd in (0.7, 0.5):
                res = nms([dict(pred_relevant_windows=[list(r) for r in w])], nms_thd=thd,
                          max_before_nms=1000, max_after_nms=10, nms_type=mode)
                rec[f"{mode}_{thd}"] = res[0]["pred_relevant_windows"]
        for thd, mx in ((0.7, 100), (0.5, 5)):
            rec[f"hull_{thd}_{mx}"] = hull([list(r) for r in w], thd, max_after_nms=mx)
        out.append(rec)

    def nan_safe(o):
        if isinstance(o, float) and (o != o):
            return "nan"
        if isinstance(o, list):
            return [nan_safe(x) for x in o]
        if isinstance(o, dict):
            return {k: nan_safe(v) for k, v in o.items()}
        return o
    with open(os.path.join(OUT, "nms_cases.json"), "w") as f:
        json.dump(nan_safe(out), f)
    print("wrote nms_cases.json", len(out))


def make_postproc():
    """compose (inference.py:286-290, the three reference lines quoted verbatim below) +
    PostProcessorDETR (postprocessing.py:25-50) on seeded raw boundaries."""
    PP = ref_loader.reference_postprocessor()
    rng = np.random.Generator(np.random.PCG64(11))
    out = []
    for name in ("qvh_iv2", "charades_vgg", "charades_sfclip", "tacos"):
        cfg = PRESETS[name]
        clip_ts, mn, mx, rnd = postprocessor_preset(cfg)
        names = (("clip_ts",) if clip_ts else ()) + (("round_multiple",) if rnd else ())
        pp = PP(clip_length=cfg.clip_length, min_ts_val=mn, max_ts_val=mx, min_w_l=0, max_w_l=mx,
                move_window_method="left", process_func_names=names)
        for _ in range(6):
            n = int(rng.integers(1, 51))
            duration = float(rng.uniform(10, 200))
            st = rng.uniform(-5, duration, size=n)
            b = np.stack([st, st + rng.uniform(0, 60, size=n), rng.uniform(0, 1, size=n)], 1)
            b[rng.integers(0, n)] = [0.03125, 1.03125, 0.03125]  # exact 4-dp rounding ties
            boundary = torch.from_numpy(b.astype(np.float32))
            # --- reference lines, FlashVTG/inference.py:286-290 ---
            spans = torch.clamp(boundary, 0, duration)
            cur_ranked_preds = spans.tolist()
            cur_ranked_preds = [[float(f"{e:.4f}") for e in row] for row in cur_ranked_preds]
            # -------------------------------------------------------
            composed = [list(r) for r in cur_ranked_preds]
            res = pp([dict(pred_relevant_windows=[list(r) for r in cur_ranked_preds])])
            out.append(dict(preset=name, duration=duration, boundary=boundary.numpy().tolist(),
                            composed=composed, processed=res[0]["pred_relevant_windows"]))
    with open(os.path.join(OUT, "postproc_cases.json"), "w") as f:
        json.dump(out, f)
    print("wrote postproc_cases.json", len(out))


def make_inputs_golden():
    """tests/golden/inputs_case.npz: the reference loader's own functions (l2_normalize_np_array,
    the TEF expression of start_end_dataset.py:175-177, pad_sequences_1d) on tests/test_inputs._raw_case."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_inputs import _raw_case
    sys.path.insert(0, "/root/reference")
    from utils.basic_utils import l2_normalize_np_array
    from utils.tensor_utils import pad_sequences_1d
    groups, txt, vlen, tlen = _raw_case()
    vs, qs = [], []
    for b in range(len(vlen)):
        fl = [l2_normalize_np_array(g[b, :vlen[b]].astype(np.float32)) for g in groups]
        v = torch.from_numpy(np.concatenate(fl, axis=1))
        L = len(v)
        tef_st = torch.arange(0, L, 1.0) / L
        tef_ed = tef_st + 1.0 / L
        vs.append(torch.cat([v, torch.stack([tef_st, tef_ed], dim=1)], dim=1))
        qs.append(torch.from_numpy(l2_normalize_np_array(txt[b, :tlen[b]].astype(np.float32))))
    pv, mv = pad_sequences_1d(vs, dtype=torch.float32, fixed_length=None)
    pq, mq = pad_sequences_1d(qs, dtype=torch.float32, fixed_length=None)
    np.savez_compressed(os.path.join(OUT, "inputs_case.npz"), src_vid=pv.numpy(), vid_mask=mv.numpy(),
                        src_txt=pq.numpy(), txt_mask=mq.numpy())


if __name__ == "__main__":
    torch.set_num_threads(8)
    assert ref_loader.available(), "needs /root/reference"
    make_nms()
    make_postproc()
    make_forward()
    make_inputs_golden()
