"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle on the same
seeded inputs, against the committed golden fixtures generated from the unmodified reference,
and - at the bench's full batch size - through size-independent properties.

Tolerances (BASELINE.json north_star): floating-point tensors of the bf16 forward within 1e-2
per-tensor max-norm relative error  max|a-b| / max|b|  (SURVEY §7 explains why element-wise
relative error is meaningless for logits / saliency near 0); everything that is an index or a
keep decision (top-k order on identical scores, NMS order / zero mask, window rounding) bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, c_nms_f32, c_nms_hull, denan, load_forward_index, max_rel, regen_case

pytestmark = pytest.mark.gpu

TOL = 1e-2
TOL_LOGIT = 2.5e-2   # pre-sigmoid logits: profiles/r02_logit_error.md (worst measured 1.9e-2; the bf16-emulating oracle alone is at 2.2e-2)
TOL_EMU = 4e-3
TOL_EMU_LOGIT = 2e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(cfg, sd):
    from flashvtg_b200.model import FlashVTGB200
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    return m


def _run(cfg, sd, batch, **kw):
    dev = torch.device("cuda:0")
    m = _model(cfg, sd)
    r = m.infer(batch["src_vid"].to(dev), batch["vid_len"].to(dev), batch["src_txt"].to(dev),
                batch["txt_len"].to(dev), duration=batch["duration"].to(dev), want_heads=True,
                want_emb=True, want_dummy=True, **kw)
    torch.cuda.synchronize()
    return m, r


def test_library_loaded_and_native(lib):
    import flashvtg_b200._lib as L0
    assert lib.fvtg_abi_version() == L0.ABI_VERSION
    import flashvtg_b200._lib as L
    assert os.path.exists(L.LIB_PATH)


def test_gemm_against_torch(dbg_lib):
    lib = dbg_lib
    import ctypes as C
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    # the last four are tall enough (> 148 m-tiles) for the weight-resident mode of the short-K GEMMs
    for (M, N, K) in [(128, 256, 64), (300, 256, 256), (9600, 256, 832), (2400, 1024, 256),
                      (5000, 128, 256), (18944, 256, 1280), (40000, 768, 256), (38001, 256, 256),
                      (45000, 128, 256), (41000, 128, 128)]:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.1).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        ref = torch.relu(a.float() @ w.float().t() + bias)
        out = torch.full((M, N), float("nan"), device=dev,
                         dtype=torch.float32 if N == 256 else torch.bfloat16)
        rc = lib.fvtg_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K,
                               1, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.fvtg_last_error()
        torch.cuda.synchronize()
        tol = (1e-3 if N == 256 else 2e-2) * ref.abs().max().item()
        assert (out.float() - ref).abs().max().item() <= tol, (M, N, K)
    del C


def test_first_projection_kernel_against_torch(dbg_lib):
    lib = dbg_lib
    """inproj_kernel alone (LayerNorm(raw dim) folded into the GEMM, ReLU, LayerNorm(256)) against an fp64
    torch restatement: ragged feature dims (tail k-block), row counts off the tile size, features with a
    large common offset (the per-row shift that keeps the folded LayerNorm's cancellation benign)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    for rows, dim, offset in [(300, 770, 0.0), (129, 66, 0.0), (1000, 4096, 0.0), (77, 2818, 0.0),
                              (20000, 770, 0.0), (513, 4096, 40.0), (256, 514, -25.0)]:
        dim_pad = (dim + 63) // 64 * 64
        x = (torch.randn(rows, dim, generator=g) * (1.0 + torch.rand(rows, 1, generator=g)) + offset).to(dev)
        W = torch.randn(256, dim, generator=g) / dim ** 0.5
        gamma = 1.0 + 0.2 * torch.randn(dim, generator=g)
        beta = 0.1 * torch.randn(dim, generator=g)
        b = 0.1 * torch.randn(256, generator=g)
        g1 = 1.0 + 0.2 * torch.randn(256, generator=g)
        b1 = 0.1 * torch.randn(256, generator=g)
        wg = torch.zeros(256, dim_pad)
        wg[:, :dim] = W * gamma
        wg = wg.to(torch.bfloat16)
        wsum = wg.float().sum(1)
        cfold = W @ beta + b
        out = torch.full((rows, 256), float("nan"), device=dev, dtype=torch.bfloat16)
        t = [v.to(dev).contiguous() for v in (wg, wsum, cfold, g1, b1)]
        rc = lib.fvtg_dbg_inproj(x.data_ptr(), rows, dim, dim_pad, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                                 t[3].data_ptr(), t[4].data_ptr(), out.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.fvtg_last_error()
        torch.cuda.synchronize()
        xd = x.double().cpu()
        ln = (xd - xd.mean(1, keepdim=True)) / torch.sqrt(xd.var(1, unbiased=False, keepdim=True) + 1e-5)
        y = torch.relu((ln * gamma.double() + beta.double()) @ W.double().t() + b.double())
        y = (y - y.mean(1, keepdim=True)) / torch.sqrt(y.var(1, unbiased=False, keepdim=True) + 1e-5)
        ref = y * g1.double() + b1.double()
        err = (out.double().cpu() - ref).abs().max().item()
        assert err <= 3e-2 * max(ref.abs().max().item(), 1.0), (rows, dim, offset, err)   # bf16 operands + output
    # odd feature dims are rejected, not mis-read
    bad = lib.fvtg_dbg_inproj(x.data_ptr(), 4, 65, 128, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                              t[3].data_ptr(), t[4].data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert bad != 0


@pytest.mark.parametrize("entry", load_forward_index(), ids=lambda e: e["file"][:-4])
def test_forward_matches_oracle_and_golden(entry):
    from oracle import forward as O
    cfg, sd, batch, gold = regen_case(entry)
    _, r = _run(cfg, sd, batch)
    outs = O.forward_batch(sd, cfg, batch)
    x = float(sd["x"])
    for b, o in enumerate(outs):
        lv = int(batch["vid_len"][b])
        n = o["logit"].shape[0]
        got_logit = x * r.cls_logit[b, :n].cpu() + (1 - x) * r.conf_logit[b, :n].cpu()
        checks = {
            "video_emb": (r.video_emb[b, :lv].cpu(), o["video_emb"], TOL),
            "saliency": (r.saliency[b, :lv].cpu(), o["saliency"], TOL),
            "t2vattn": (r.t2vattn[b, :lv].cpu(), o["t2vattn"], TOL),
            "dummy_tokens": (r.dummy_tokens[b].cpu(), o["dummy_tokens"], TOL),
            # the contract: scores (sigmoid) and spans within 1e-2
            "score": (torch.sigmoid(got_logit), o["score"], TOL),
            "coord": (r.coord[b, :n].cpu(), o["coord"], TOL),
            # pre-sigmoid logits: informational looser bound (spread-init scales the last MLP layer
            # x64; tests/emulation.py shows 1-2e-2 is inherent to bf16 operands there)
            "cls": (r.cls_logit[b, :n].cpu(), o["cls"], TOL_LOGIT),
            "conf": (r.conf_logit[b, :n].cpu(), o["conf"], TOL_LOGIT),
            "logit": (got_logit, o["logit"], TOL_LOGIT),
        }
        for name, (got, want, tol) in checks.items():
            assert torch.isfinite(got).all(), name
            e = max_rel(got.numpy(), want.numpy())
            assert e < tol, f"{entry['file']} video {b} {name}: max-norm rel err {e:.3e}"
        # spans of every point (decode applied to our coord) vs the oracle's
        pt, ps = o["point_t"], o["point_s"]
        cg = r.coord[b, :n].cpu()
        spans = torch.stack([(pt - cg[:, 0] * ps), (pt + cg[:, 1] * ps)], 1) * cfg.clip_length
        assert max_rel(spans.numpy(), o["spans"].numpy()) < TOL
        # against the unmodified reference's outputs (golden fixtures)
        assert max_rel(r.saliency[b, :lv].cpu().numpy(), gold[f"saliency_{b}"]) < TOL
        assert max_rel(r.t2vattn[b, :lv].cpu().numpy(), gold[f"t2vattn_{b}"]) < TOL
        assert max_rel(torch.sigmoid(got_logit).numpy(),
                       torch.sigmoid(torch.from_numpy(gold[f"logit_{b}"])).numpy()) < TOL
        assert max_rel(r.coord[b, :n].cpu().numpy(), gold[f"coord_{b}"]) < TOL
        # rows beyond the true length are zero
        assert float(r.saliency[b, lv:].abs().sum()) == 0.0
        # ranked spans: compare as a set keyed by the point each row came from (ranking of bf16
        # scores may legitimately differ from fp32 between near-ties)
        cnt = int(r.count[b])
        assert cnt == min(n, cfg.max_num_moment) == gold[f"boundary_{b}"].shape[0]
        sc = r.boundary[b, :cnt, 2].cpu().numpy()
        assert np.all(np.diff(sc) <= 0), "boundary not sorted by score"
        spans_all = o["spans"].numpy()
        scale = np.abs(spans_all).max()
        gs = torch.sigmoid(got_logit).numpy()
        for row in r.boundary[b, :cnt].cpu().numpy():
            # the point(s) whose score this row carries: saturated sigmoids tie exactly (spread-init fixtures),
            # so the row must match the span of ONE of the points with that score
            cand = np.nonzero(np.abs(gs - row[2]) < 1e-6)[0]
            assert cand.size > 0
            assert min(np.abs(spans_all[j] - row[:2]).max() for j in cand) < TOL * scale


@pytest.mark.parametrize("entry", load_forward_index(), ids=lambda e: e["file"][:-4])
def test_forward_matches_bf16_emulation(entry):
    """Sharper than the fp32 bound: tests/emulation.py rounds to bf16 exactly where the kernels
    do, so the GPU must agree with it far inside the 1e-2 budget (a wrong mask, tap, residual or
    LayerNorm shows up here even when it hides under 1e-2 against fp32)."""
    import emulation as E
    cfg, sd, batch, gold = regen_case(entry)
    _, r = _run(cfg, sd, batch)
    outs = E.forward_batch(sd, cfg, batch)
    report = {}
    for b, o in enumerate(outs):
        lv = int(batch["vid_len"][b])
        n = o["logit"].shape[0]
        checks = {
            "video_emb": (r.video_emb[b, :lv], o["video_emb"]),
            "saliency": (r.saliency[b, :lv], o["saliency"]),
            "t2vattn": (r.t2vattn[b, :lv], o["t2vattn"]),
            "dummy_tokens": (r.dummy_tokens[b], o["dummy_tokens"]),
            "cls": (r.cls_logit[b, :n], o["cls"]),
            "conf": (r.conf_logit[b, :n], o["conf"]),
            "coord": (r.coord[b, :n], o["coord"]),
        }
        for name, (got, want) in checks.items():
            report[f"{b}.{name}"] = max_rel(got.cpu().numpy(), want.numpy())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_emulation_{entry['file'][:-4]}.json"), "w") as f:
        json.dump(report, f, indent=1)
    for k, e in report.items():
        # the x64-scaled score-head logits amplify single bf16 rounding flips: looser bound there
        tol = TOL_EMU_LOGIT if k.endswith((".cls", ".conf")) else TOL_EMU
        assert e < tol, f"{entry['file']} {k}: vs bf16 emulation {e:.3e}"


@pytest.mark.parametrize("entry", load_forward_index()[:3], ids=lambda e: e["file"][:-4])
def test_decode_topk_postproc_nms_bit_exact_on_identical_scores(entry):
    """Kernel group C fed the ORACLE's fp32 head outputs: ranking, compose, PostProcessorDETR and
    NMS must reproduce the oracle bit for bit."""
    from flashvtg_b200.config import postprocessor_preset
    from oracle import forward as O
    from oracle import postproc as P
    cfg, sd, batch, gold = regen_case(entry)
    outs = O.forward_batch(sd, cfg, batch)
    dev = torch.device("cuda:0")
    B, Lv = batch["src_vid"].shape[:2]
    n_max = cfg.num_points(Lv)
    cls = torch.zeros(B, n_max)
    conf = torch.zeros(B, n_max)
    coord = torch.zeros(B, n_max, 2)
    for b, o in enumerate(outs):
        n = o["cls"].shape[0]
        cls[b, :n], conf[b, :n], coord[b, :n] = o["cls"], o["conf"], o["coord"]
    m = _model(cfg, sd)
    for mode in ("normal", "linear"):
        bnd, win, nms_w, nms_o, count, nms_c = m.decode(cls.to(dev), conf.to(dev), coord.to(dev),
                                                       batch["vid_len"].to(dev), Lv,
                                                       duration=batch["duration"].to(dev), nms=mode)
        torch.cuda.synchronize()
        clip_ts, mn, mx, rnd = postprocessor_preset(cfg)
        for b, o in enumerate(outs):
            k = int(count[b])
            want = o["boundary"].numpy()
            assert k == want.shape[0]
            got = bnd[b, :k].cpu().numpy()
            # spans bit-exact (same fp32 op order); sigmoid may differ from torch's by an ulp
            np.testing.assert_array_equal(got[:, :2], want[:, :2])
            np.testing.assert_allclose(got[:, 2], want[:, 2], rtol=0, atol=2e-7)
            # compose + PostProcessorDETR on OUR ranked rows == oracle arithmetic, bit-exact
            w64 = P.post_process(P.compose_windows(got, float(batch["duration"][b])),
                                 cfg.clip_length, clip_ts, mn, mx, rnd)
            np.testing.assert_array_equal(win[b, :k].cpu().numpy(), w64.astype(np.float32))
            # NMS on identical windows: selection order and zero mask bit-exact
            out, order, _ = P.nms_reference_order(w64, cfg.nms_thd, mode)
            np.testing.assert_array_equal(nms_o[b, :k].cpu().numpy(), order)
            np.testing.assert_array_equal(nms_w[b, :k].cpu().numpy(), out)
            assert int(nms_c[b]) == k


def _nms_cases():
    with open(os.path.join(GOLDEN, "nms_cases.json")) as f:
        return json.load(f)


def test_temporal_nms_matches_reference_golden_bit_exact():
    from flashvtg_b200.postprocessing import temporal_nms
    cases = [denan(c["windows"]) for c in _nms_cases()]
    raw = _nms_cases()
    dev = torch.device("cuda:0")
    M = max(len(w) for w in cases)
    buf = np.zeros((len(cases), M, 3), np.float32)
    cnt = np.zeros(len(cases), np.int32)
    for i, w in enumerate(cases):
        buf[i, : len(w)] = np.asarray(w, np.float64).astype(np.float32)
        cnt[i] = len(w)
    win = torch.from_numpy(buf).to(dev)
    count = torch.from_numpy(cnt).to(dev)
    for mode in ("normal", "linear"):
        for thd in (0.7, 0.5):
            out, order, oc = temporal_nms(win, count, thd, mode)
            out, order = out.cpu().numpy(), order.cpu().numpy()
            for i, w in enumerate(cases):
                n = len(w)
                cout, corder, _ = c_nms_f32(w, thd, mode)
                np.testing.assert_array_equal(order[i, :n], corder)
                assert np.array_equal(out[i, :n], cout, equal_nan=True)
                assert (order[i, n:] == -1).all()
                gold = np.array(denan(raw[i][f"{mode}_{thd}"]), np.float64).reshape(-1, 3)
                gs = gold[:, 2]
                distinct = np.array([np.sum(gs == s) == 1 for s in gs])
                assert np.array_equal(gold[distinct], out[i, :n].astype(np.float64)[distinct],
                                      equal_nan=True)
    for thd, mx in ((0.7, 100), (0.5, 5)):
        out, order, oc = temporal_nms(win, count, thd, "hull", mx)
        out, order, oc = out.cpu().numpy(), order.cpu().numpy(), oc.cpu().numpy()
        for i, w in enumerate(cases):
            w32 = np.asarray(w, np.float64).astype(np.float32).astype(np.float64)
            cout, csrc = c_nms_hull(w32, thd, mx)
            assert oc[i] == len(csrc)
            np.testing.assert_array_equal(order[i, : oc[i]], csrc)
            np.testing.assert_array_equal(out[i, : oc[i]].astype(np.float64), cout)


def test_post_processing_mr_nms_dropin_list_api():
    from flashvtg_b200.postprocessing import post_processing_mr_nms, temporal_nms_list
    raw = _nms_cases()
    sub = [dict(qid=i, pred_relevant_windows=[list(r) for r in denan(c["windows"])])
           for i, c in enumerate(raw[10:40])]
    res = post_processing_mr_nms(sub, nms_thd=0.7, max_before_nms=1000, max_after_nms=10,
                                 nms_type="normal")
    for e, c in zip(res, raw[10:40]):
        gold = np.array(denan(c["normal_0.7"]), np.float64).reshape(-1, 3)
        got = np.array(e["pred_relevant_windows"], np.float64).reshape(-1, 3)
        assert got.shape == gold.shape
        np.testing.assert_array_equal(np.sort(got[:, 2]), np.sort(gold[:, 2]))
    with pytest.raises(ValueError):
        post_processing_mr_nms(sub, 0.7, 1000, 10, "bogus")
    w = [[0.0, 20.0, 0.9], [0.0, 14.0, 0.8], [40.0, 60.0, 0.7]]
    assert temporal_nms_list(w, 0.5) == [w[0], w[2]]
    assert temporal_nms_list(w[:1], 0.5) == w[:1]


def test_forward_dropin_signature_bs1():
    """The reference call `model(**model_inputs, targets=targets)` (inference.py:255) at bs=1."""
    entry = load_forward_index()[1]
    cfg, sd, batch, gold = regen_case(entry)
    dev = torch.device("cuda:0")
    m = _model(cfg, sd)
    out = m(src_txt=batch["src_txt"].to(dev), src_txt_mask=batch["src_txt_mask"].to(dev),
            src_vid=batch["src_vid"].to(dev), src_vid_mask=batch["src_vid_mask"].to(dev),
            vid=None, qid=None, targets={})
    assert set(out) >= {"_avg_factor", "saliency_scores", "t2vattnvalues", "_out",
                        "saliency_scores_neg", "t2vattnvalues_neg", "real_neg_mask", "dummy_tokens"}
    assert out["_out"]["boundary"].shape == gold["boundary_0"].shape
    assert out["_out"]["saliency"].shape == (75,)
    assert out["_out"]["video_msk"].dtype == torch.int32
    assert max_rel(out["saliency_scores"][0].cpu().numpy(), gold["saliency_0"]) < TOL
    with pytest.raises(RuntimeError):
        m(src_txt=batch["src_txt"], src_txt_mask=batch["src_txt_mask"], src_vid=batch["src_vid"],
          src_vid_mask=batch["src_vid_mask"], vid=None, qid=None, targets={})


def test_full_size_properties_b1024():
    """BASELINE config #2 size (B=1024, QVH-IV2): chunk/batch independence (a video's result does
    not depend on which batch or chunk it rides in - bit-exact), sortedness, NMS idempotence on
    the keep set, finite outputs."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.postprocessing import temporal_nms
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    B = 1024
    small = synth.make_inputs(cfg, 8, 75, 32, seed=77, ragged=True, min_lv=20)
    dev = torch.device("cuda:0")
    rep = B // 8
    big = {k: v.repeat(rep, *([1] * (v.dim() - 1))) for k, v in small.items()}
    m = _model(cfg, sd)

    def run(bt):
        r = m.infer(bt["src_vid"].to(dev), bt["vid_len"].to(dev), bt["src_txt"].to(dev),
                    bt["txt_len"].to(dev), duration=bt["duration"].to(dev), want_heads=True)
        torch.cuda.synchronize()
        return r
    rs, rb = run(small), run(big)
    for name in ("saliency", "t2vattn", "cls_logit", "conf_logit", "coord", "boundary", "windows",
                 "nms_windows", "nms_order", "count"):
        a, b = getattr(rs, name), getattr(rb, name)
        assert torch.isfinite(b.float()).all(), name
        for k in (0, 1, 37, rep - 1):
            assert torch.equal(a, b[k * 8:(k + 1) * 8]), f"{name}: batch position changes the result"
    sc = rb.boundary[:, :, 2]
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    # idempotence: NMS of the NMS output changes nothing
    out2, order2, _ = temporal_nms(rb.nms_windows, rb.nms_count, cfg.nms_thd, "normal")
    assert torch.equal(out2, rb.nms_windows)
    assert rb.launches > 0
    # ... and the B = 1024 run itself against the oracle (not only against the small run): the 8 seed videos,
    # read from three different chunk positions of the big batch, within the north_star tolerance
    from oracle import forward as O
    outs = O.forward_batch(sd, cfg, small)
    x = float(sd["x"])
    for k in (0, 63, rep - 1):
        for b, o in enumerate(outs):
            g = k * 8 + b
            lv, n = int(small["vid_len"][b]), o["logit"].shape[0]
            logit = x * rb.cls_logit[g, :n].cpu() + (1 - x) * rb.conf_logit[g, :n].cpu()
            for name, got, want in (("saliency", rb.saliency[g, :lv].cpu(), o["saliency"]),
                                    ("t2vattn", rb.t2vattn[g, :lv].cpu(), o["t2vattn"]),
                                    ("score", torch.sigmoid(logit), o["score"]),
                                    ("coord", rb.coord[g, :n].cpu(), o["coord"])):
                e = max_rel(got.numpy(), want.numpy())
                assert e < TOL, f"B=1024 video {g} {name}: max-norm rel err {e:.3e} vs the oracle"


def test_infer_host_pipeline_equals_device_call():
    """The host-buffer entry (chunked H2D overlapped with compute on a second stream, the call
    bench.py times as e2e) returns bit-identical results to one device-resident infer()."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    batch = synth.make_inputs(cfg, 11, 75, 32, seed=9, ragged=True, min_lv=9)
    m, r = _run(cfg, sd, batch)
    host = {k: v.pin_memory() for k, v in batch.items()}
    for chunk in (3, 4, 64):
        o = m.infer_host(host["src_vid"], host["vid_len"], host["src_txt"], host["txt_len"],
                         duration=host["duration"], nms="normal", chunk_videos=chunk)
        for name in ("boundary", "windows", "nms_windows", "nms_order", "count", "nms_count",
                     "saliency", "t2vattn"):
            assert torch.equal(o[name], getattr(r, name).cpu()), f"{name} differs at chunk {chunk}"
        assert o["launches"] > 0
    # second call reusing the pinned result buffers
    o2 = m.infer_host(host["src_vid"], host["vid_len"], host["src_txt"], host["txt_len"],
                      duration=host["duration"], nms="normal", chunk_videos=4, out=o)
    assert torch.equal(o2["nms_windows"], r.nms_windows.cpu())


@pytest.mark.parametrize("preset,B,Lv,Lt,ragged", [
    ("tacos", 2, 389, 9, False),          # BASELINE config #4: TACoS test-set max length
    ("tacos_deep", 1, 701, 13, False),    # train max with the MR_32 deep pyramid (strides 1..32)
    ("charades_vgg", 3, 184, 10, True),   # BASELINE config #3: Charades-STA VGG, ragged lengths
    ("charades_vgg", 1, 432, 6, False),   # Charades max length at 6 fps
    ("qvh_iv2", 4, 75, 40, True),         # QVH with max_q_l = 40 text tokens
    ("qvh_sfclip", 3, 5, 3, True),        # tiny videos: most pyramid levels vanish (blocks.py:56)
    ("qvh_iv2", 2, 1, 1, False),          # a single clip and a single token
])
def test_forward_matches_oracle_at_baseline_shapes(preset, B, Lv, Lt, ragged):
    """Shapes of BASELINE.json's configs that the golden fixtures do not hold (long videos, deep
    pyramids, degenerate lengths): CUDA path vs the fp32 oracle on the same seeded inputs,
    1e-2 max-norm relative on the index-aligned tensors."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from oracle import forward as O
    cfg = PRESETS[preset]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    batch = synth.make_inputs(cfg, B, Lv, Lt, seed=4242, ragged=ragged, min_lv=max(1, Lv // 3),
                              min_lt=min(2, Lt))
    _, r = _run(cfg, sd, batch)
    outs = O.forward_batch(sd, cfg, batch)
    x = float(sd["x"])
    for b, o in enumerate(outs):
        lv = int(batch["vid_len"][b])
        n = o["logit"].shape[0]
        got_logit = x * r.cls_logit[b, :n].cpu() + (1 - x) * r.conf_logit[b, :n].cpu()
        for name, got, want in (("saliency", r.saliency[b, :lv].cpu(), o["saliency"]),
                                ("t2vattn", r.t2vattn[b, :lv].cpu(), o["t2vattn"]),
                                ("video_emb", r.video_emb[b, :lv].cpu(), o["video_emb"]),
                                ("score", torch.sigmoid(got_logit), o["score"]),
                                ("coord", r.coord[b, :n].cpu(), o["coord"])):
            assert torch.isfinite(got).all(), name
            e = max_rel(got.numpy(), want.numpy())
            # a one-clip video has a ONE-element saliency tensor: the max-norm metric degenerates to the
            # relative error of a single bf16-operand quadratic form (no larger element to normalise by)
            tol = 2 * TOL if (name == "saliency" and lv == 1) else TOL
            assert e < tol, f"{preset} Lv={Lv} video {b} {name}: max-norm rel err {e:.3e}"
        cnt = int(r.count[b])
        assert cnt == min(n, cfg.max_num_moment)
        sc = r.boundary[b, :cnt, 2].cpu().numpy()
        assert np.all(np.diff(sc) <= 0)


def test_tcgen05_attention_forced_for_short_shapes_matches_default():
    """FVTG_ATTN_TC=1 routes the short QVHighlights shapes through the tcgen05 attention kernel too
    (persistent, TMEM-resident softmax; by default it only serves sequences of more than 160 keys): same
    forward as the default mma.sync kernel within the bf16 noise floor of the path."""
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from flashvtg_b200 import synth
from flashvtg_b200.config import PRESETS
from flashvtg_b200.model import FlashVTGB200
cfg = PRESETS["qvh_iv2"]
m = FlashVTGB200(cfg).eval(); m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
b = synth.make_inputs(cfg, 9, 75, 32, seed=11, ragged=True, min_lv=7)
dev = torch.device("cuda:0")
r = m.infer(b["src_vid"].to(dev), b["vid_len"].to(dev), b["src_txt"].to(dev), b["txt_len"].to(dev),
            duration=b["duration"].to(dev), want_heads=True)
torch.save({"sal": r.saliency.cpu(), "cls": r.cls_logit.cpu(), "coord": r.coord.cpu(), "t2v": r.t2vattn.cpu()}, sys.argv[1])
''' % ROOT
    import tempfile
    outs = []
    for tc in ("-1", "1"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            env = dict(os.environ, FVTG_ATTN_TC=tc)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            outs.append(torch.load(f.name))
    for k in outs[0]:
        e = max_rel(outs[1][k].numpy(), outs[0][k].numpy())
        assert e < TOL, f"{k}: tcgen05 attention vs default max-norm rel err {e:.3e}"


def test_tile_handoff_is_bit_identical_to_the_grid_wide_dependency():
    """The per-tile release/acquire between a T2V layer kernel and the next attention (FVTG_TILE_HANDOFF, default on)
    only changes WHEN an attention CTA may start: every output must equal the grid-wide griddepcontrol.wait chain bit
    for bit - at a size where every layer-kernel CTA owns several tiles (600 tiles on 148 SMs), twice in a row on the
    same workspace (stale flags of the previous forward must not satisfy the next one), ragged lengths."""
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from flashvtg_b200 import synth
from flashvtg_b200.config import PRESETS
from flashvtg_b200.model import FlashVTGB200
cfg = PRESETS["qvh_iv2"]
m = FlashVTGB200(cfg).eval(); m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
b = synth.make_inputs(cfg, 64, 75, 32, seed=12, ragged=True, min_lv=9)
dev = torch.device("cuda:0")
d = {k: v.repeat(16, *([1] * (v.dim() - 1))).contiguous().to(dev) for k, v in b.items()}
out = []
for _ in range(2):
    r = m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], duration=d["duration"], want_heads=True)
    torch.cuda.synchronize()
    out.append({"sal": r.saliency.cpu(), "cls": r.cls_logit.cpu(), "coord": r.coord.cpu(), "t2v": r.t2vattn.cpu(),
                "win": r.windows.cpu()})
for k in out[0]:
    assert torch.equal(out[0][k], out[1][k]), k
torch.save(out[1], sys.argv[1])
''' % ROOT
    import tempfile
    outs = []
    for on in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            env = dict(os.environ, FVTG_TILE_HANDOFF=on)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            outs.append(torch.load(f.name))
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), f"{k}: tile hand-off changed the result"


def test_uniform_length_hint_changes_nothing():
    """FvtgBatch.uniform_vid_len (compact sine table for chunks whose videos share one length) is a
    pure layout optimisation: bit-identical outputs with and without the hint, full and short length."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    dev = torch.device("cuda:0")
    m = _model(cfg, sd)
    for lv_true in (75, 41):
        batch = synth.make_inputs(cfg, 5, 75, 32, seed=21, ragged=False)
        if lv_true != 75:   # shorter than the padded length, still uniform
            batch["vid_len"][:] = lv_true
            batch["src_vid"][:, lv_true:] = 0
            batch["duration"][:] = lv_true * cfg.clip_length
        outs = []
        for hint in (False, True):
            r = m.infer(batch["src_vid"].to(dev), batch["vid_len"].to(dev), batch["src_txt"].to(dev),
                        batch["txt_len"].to(dev), duration=batch["duration"].to(dev), want_heads=True,
                        uniform_len=hint)
            torch.cuda.synchronize()
            outs.append(r)
        for name in ("saliency", "t2vattn", "cls_logit", "conf_logit", "coord", "boundary", "nms_windows"):
            assert torch.equal(getattr(outs[0], name), getattr(outs[1], name)), (lv_true, name)


def test_workspace_reuse_across_lengths_leaves_no_stale_rows():
    """The pyramid row spaces (H1 / H2) are no longer cleared per forward: the level-0 kernel zeroes exactly the
    rows no producer writes.  A model whose workspace just held full-length videos must give, on a ragged
    batch, bit-identical results to a fresh model (stale rows of the longer videos would leak into the conv
    taps), and the grouped pyramid launches must equal the one-launch-per-step schedule."""
    import subprocess
    import sys
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.model import FlashVTGB200
    cfg = PRESETS["qvh_iv2"]
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    full = synth.make_inputs(cfg, 12, 75, 32, seed=5)
    rag = synth.make_inputs(cfg, 12, 75, 32, seed=6, ragged=True, min_lv=3)

    def run(m, b):
        r = m.infer(b["src_vid"].to(dev), b["vid_len"].to(dev), b["src_txt"].to(dev), b["txt_len"].to(dev),
                    duration=b["duration"].to(dev), want_heads=True)
        torch.cuda.synchronize()
        return {k: getattr(r, k).clone() for k in ("saliency", "cls_logit", "conf_logit", "coord", "boundary", "count")}

    used = FlashVTGB200(cfg).eval()
    used.load_state_dict(sd)
    run(used, full)
    run(used, full)
    got = run(used, rag)
    fresh = FlashVTGB200(cfg).eval()
    fresh.load_state_dict(sd)
    ref = run(fresh, rag)
    for k in got:
        assert torch.equal(got[k], ref[k]), k
    # grouped vs per-step pyramid launches (separate process: the switch is read once)
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from flashvtg_b200 import synth
from flashvtg_b200.config import PRESETS
from flashvtg_b200.model import FlashVTGB200
cfg = PRESETS["qvh_iv2"]
m = FlashVTGB200(cfg).eval(); m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
b = synth.make_inputs(cfg, 12, 75, 32, seed=6, ragged=True, min_lv=3)
dev = torch.device("cuda:0")
r = m.infer(b["src_vid"].to(dev), b["vid_len"].to(dev), b["src_txt"].to(dev), b["txt_len"].to(dev),
            duration=b["duration"].to(dev), want_heads=True)
torch.save({"cls": r.cls_logit.cpu(), "conf": r.conf_logit.cpu(), "coord": r.coord.cpu()}, sys.argv[1])
''' % ROOT
    import tempfile
    outs = []
    for grp in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=dict(os.environ, FVTG_PYR_GROUP=grp),
                           timeout=300)
            outs.append(torch.load(f.name))
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert torch.equal(outs[1]["cls"], ref["cls_logit"].cpu())


def test_torch_ops_and_standalone_stage_chain_equal_the_fused_forward():
    """Every torch.ops.fvtg.* operator (flashvtg_b200/ops.py) is callable, and the standalone
    C-ABI chain fvtg_fusion_fwd -> fvtg_pyramid_heads_fwd -> fvtg_decode_nms (the three kernel groups as separate
    calls) equals fvtg_forward bit for bit: same kernels, same launch parameters, only the entry point differs."""
    import flashvtg_b200.ops as ops
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    dev = torch.device("cuda:0")
    m = _model(cfg, sd)
    h = ops.register_model(m)
    b = synth.make_inputs(cfg, 5, 75, 32, seed=21, ragged=True, min_lv=9)
    d = {k: v.to(dev) for k, v in b.items()}
    r = m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], duration=d["duration"], nms="normal",
                want_heads=True, want_emb=True, want_dummy=True)
    emb, sal, t2v, dummy = torch.ops.fvtg.fusion_fwd(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], h)
    assert torch.equal(sal, r.saliency) and torch.equal(t2v, r.t2vattn)
    assert torch.equal(emb, r.video_emb) and torch.equal(dummy, r.dummy_tokens)
    cls, conf, coord = torch.ops.fvtg.pyramid_heads_fwd(emb, d["vid_len"], h)
    assert torch.equal(cls, r.cls_logit) and torch.equal(conf, r.conf_logit) and torch.equal(coord, r.coord)
    bnd, win, nms_w, nms_o, cnt, nms_c = torch.ops.fvtg.decode_nms(cls, conf, coord, d["vid_len"], d["duration"],
                                                                  75, h, 0, cfg.nms_thd)
    assert torch.equal(cnt, r.count) and torch.equal(nms_c, r.nms_count)
    for bi in range(5):
        n = int(cnt[bi])
        assert torch.equal(bnd[bi, :n], r.boundary[bi, :n]) and torch.equal(win[bi, :n], r.windows[bi, :n])
        assert torch.equal(nms_w[bi, :n], r.nms_windows[bi, :n]) and torch.equal(nms_o[bi, :n], r.nms_order[bi, :n])
    out = torch.ops.fvtg.forward(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], d["duration"], h, 0,
                                 cfg.nms_thd)
    assert torch.equal(out[0], r.saliency) and torch.equal(out[4], r.count)
    for bi in range(5):
        n = int(cnt[bi])
        assert torch.equal(out[2][bi, :n], r.boundary[bi, :n]) and torch.equal(out[5][bi, :n], r.nms_windows[bi, :n])
    w2, o2, c2 = torch.ops.fvtg.temporal_nms(r.windows, r.count, cfg.nms_thd, 0, 100)
    for bi in range(5):
        n = int(cnt[bi])
        assert torch.equal(w2[bi, :n], r.nms_windows[bi, :n]) and torch.equal(o2[bi, :n], r.nms_order[bi, :n])
    with pytest.raises(NotImplementedError):
        torch.ops.fvtg.temporal_nms(r.windows.cpu(), r.count.cpu(), 0.7, 0, 100)   # no CPU implementation


def test_compute_mr_results_matches_the_oracle_post_processing():
    """flashvtg_b200.postprocessing.compute_mr_results (forward + compose + PostProcessorDETR
    [+ NMS] in one device pass) against the oracle's restatement of inference.py:232-355 applied to the SAME
    device boundaries: every submission row identical."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS, postprocessor_preset
    from flashvtg_b200.postprocessing import compute_mr_results
    from oracle import postproc as P
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 2025, spread=True)
    dev = torch.device("cuda:0")
    m = _model(cfg, sd)
    b = synth.make_inputs(cfg, 6, 75, 32, seed=31, ragged=True, min_lv=5)
    inp = {k: b[k].to(dev) for k in ("src_vid", "src_vid_mask", "src_txt", "src_txt_mask")}
    metas = [dict(qid=i, query="q", vid=f"v{i}", duration=float(b["duration"][i])) for i in range(6)]
    clip_ts, mn, mx, rnd = postprocessor_preset(cfg)
    r = m.infer(inp["src_vid"], b["vid_len"].to(dev), inp["src_txt"], b["txt_len"].to(dev),
                duration=b["duration"].to(dev), nms=None)
    for nms in (None, "normal"):
        got = compute_mr_results(m, [(metas, inp)], nms=nms)
        assert [g["qid"] for g in got] == list(range(6))
        for i, g in enumerate(got):
            n = int(r.count[i])
            w, out, _ = P.full_postproc(r.boundary[i, :n].cpu().numpy(), metas[i]["duration"], cfg.clip_length,
                                        clip_ts, mn, mx, rnd, cfg.nms_thd, "normal")
            want = np.asarray(out if nms else w, np.float64)
            have = np.asarray(g["pred_relevant_windows"], np.float64)
            assert have.shape == want.shape
            if nms:
                assert np.array_equal(have.astype(np.float32), want.astype(np.float32))
            else:
                assert np.array_equal(have, want)
            lv = int(b["vid_len"][i])
            sal = np.rint(r.saliency[i, :lv].cpu().numpy().astype(np.float64) * 1e4) / 1e4
            assert np.array_equal(np.asarray(g["pred_saliency_scores"]), sal)
