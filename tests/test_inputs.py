"""Device-resident input pipeline (flashvtg_b200/inputs.py, fvtg_prepare_inputs) against the oracle
restatement of the reference loader (oracle/inputs.py), and the oracle against the unmodified
reference functions when /root/reference is present."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _raw_case(seed=3, B=5, Lv=23, Lt=9, dims=(48, 20), Dt=40, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    vlen = rng.integers(1, Lv + 1, size=B); vlen[0] = Lv
    tlen = rng.integers(1, Lt + 1, size=B); tlen[0] = Lt
    groups = [(rng.standard_normal((B, Lv, d)) * (1 + g)).astype(dtype) for g, d in enumerate(dims)]
    groups[0][1, 0] = 0                       # an all-zero row: 0 / (0 + eps) = 0
    txt = (rng.standard_normal((B, Lt, Dt)) * 3).astype(dtype)
    return groups, txt, vlen.astype(np.int32), tlen.astype(np.int32)


def _oracle_batch(groups, txt, vlen, tlen, **kw):
    from oracle import inputs as OI
    vs, qs = [], []
    for b in range(len(vlen)):
        v, q = OI.prepare_item([g[b, :vlen[b]] for g in groups], txt[b, :tlen[b]], **kw)
        vs.append(v); qs.append(q)
    return OI.collate(vs), OI.collate(qs)


def test_oracle_matches_reference_loader_functions():
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "utils", "basic_utils.py")):
        pytest.skip("reference tree not present (GPU box)")
    sys.path.insert(0, ref)
    try:
        from utils.basic_utils import l2_normalize_np_array
        from utils.tensor_utils import pad_sequences_1d
    finally:
        sys.path.remove(ref)
    from oracle import inputs as OI
    groups, txt, vlen, tlen = _raw_case()
    a = groups[0][0]
    assert np.array_equal(OI.l2_normalize(a), l2_normalize_np_array(a))
    items = [torch.from_numpy(groups[1][b, :vlen[b]]) for b in range(len(vlen))]
    pad, mask = pad_sequences_1d(items, dtype=torch.float32, fixed_length=None)
    opad, omask = OI.collate([x.numpy() for x in items])
    assert np.array_equal(pad.numpy(), opad) and np.array_equal(mask.numpy(), omask)
    # TEF: the expression of start_end_dataset.py:175-177 evaluated verbatim
    for L in (1, 7, 75):
        st = torch.arange(0, L, 1.0) / L
        assert np.array_equal(OI.tef(L), torch.stack([st, st + 1.0 / L], dim=1).numpy())


def test_oracle_matches_golden_inputs_case():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "inputs_case.npz"))
    groups, txt, vlen, tlen = _raw_case()
    (v, vm), (q, qm) = _oracle_batch(groups, txt, vlen, tlen)
    for name, got in (("src_vid", v), ("vid_mask", vm), ("src_txt", q), ("txt_mask", qm)):
        assert np.array_equal(got, gold[name]), name


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("flags", [(True, True, True), (False, True, False), (True, False, True)])
def test_prepare_inputs_matches_oracle(dtype, flags):
    from flashvtg_b200.inputs import prepare_inputs
    nv, nt, tef = flags
    groups, txt, vlen, tlen = _raw_case()
    dev = torch.device("cuda:0")
    tg = [torch.from_numpy(g).to(dtype) for g in groups]
    tt = torch.from_numpy(txt).to(dtype)
    # the oracle sees exactly the (possibly rounded) raw values the device sees
    (v, vm), (q, qm) = _oracle_batch([g.float().numpy() for g in tg], tt.float().numpy(), vlen, tlen,
                                     normalize_v=nv, normalize_t=nt, use_tef=tef)
    sv, svm, st, stm = prepare_inputs([g.to(dev) for g in tg], torch.from_numpy(vlen).to(dev), tt.to(dev),
                                      torch.from_numpy(tlen).to(dev), normalize_v=nv, normalize_t=nt,
                                      use_tef=tef)
    torch.cuda.synchronize()
    assert np.array_equal(svm.cpu().numpy(), vm) and np.array_equal(stm.cpu().numpy(), qm)
    sv, st = sv.cpu().numpy(), st.cpu().numpy()
    # normalised values: 1e-6 relative (numpy's pairwise sum vs a warp-shuffle sum); TEF bit-exact
    assert np.abs(sv - v).max() <= 1e-6 * max(np.abs(v).max(), 1e-12) + 1e-9
    assert np.abs(st - q).max() <= 1e-6 * max(np.abs(q).max(), 1e-12) + 1e-9
    if tef:
        assert np.array_equal(sv[:, :, -2:], v[:, :, -2:])
    for b in range(len(vlen)):
        assert not sv[b, vlen[b]:].any() and not st[b, tlen[b]:].any()


@pytest.mark.gpu
def test_raw_pipeline_feeds_the_forward():
    """raw fp16 features -> prepare_inputs -> infer equals infer on host-prepared fp32 inputs (1e-6 on the
    inputs, so the forward outputs agree to bf16 noise)."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.inputs import prepare_inputs
    from flashvtg_b200.model import FlashVTGB200
    from helpers import max_rel
    cfg = PRESETS["qvh_iv2"]
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
    rng = np.random.Generator(np.random.PCG64(5))
    B, Lv, Lt = 4, 75, 32
    raw_v = torch.from_numpy(rng.standard_normal((B, Lv, 768)).astype(np.float32)).half()
    raw_t = torch.from_numpy(rng.standard_normal((B, Lt, 4096)).astype(np.float32)).half()
    vlen = torch.tensor([75, 60, 75, 33], dtype=torch.int32)
    tlen = torch.tensor([32, 12, 25, 32], dtype=torch.int32)
    (v, _), (q, _) = _oracle_batch([raw_v.float().numpy()], raw_t.float().numpy(), vlen.numpy(), tlen.numpy())
    sv, _, st, _ = prepare_inputs([raw_v.to(dev)], vlen.to(dev), raw_t.to(dev), tlen.to(dev))
    r1 = m.infer(sv, vlen.to(dev), st, tlen.to(dev))
    r2 = m.infer(torch.from_numpy(v).to(dev), vlen.to(dev), torch.from_numpy(q).to(dev), tlen.to(dev))
    torch.cuda.synchronize()
    assert max_rel(r1.saliency.cpu().numpy(), r2.saliency.cpu().numpy()) < 2e-3
    assert torch.equal(r1.count, r2.count)


@pytest.mark.gpu
def test_infer_raw_host_equals_prepare_then_infer():
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.inputs import prepare_inputs
    from flashvtg_b200.model import FlashVTGB200
    cfg = PRESETS["qvh_iv2"]
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
    g = torch.Generator().manual_seed(8)
    B, Lv, Lt = 7, 75, 32
    raw_v = torch.randn(B, Lv, 768, generator=g).half().pin_memory()
    raw_t = torch.randn(B, Lt, 4096, generator=g).half().pin_memory()
    vlen = torch.tensor([75, 60, 75, 33, 75, 75, 9], dtype=torch.int32)
    tlen = torch.tensor([32, 12, 25, 32, 4, 32, 17], dtype=torch.int32)
    sv, _, st, _ = prepare_inputs([raw_v.to(dev)], vlen.to(dev), raw_t.to(dev), tlen.to(dev))
    r = m.infer(sv, vlen.to(dev), st, tlen.to(dev))
    torch.cuda.synchronize()
    for chunk in (3, 64):
        o = m.infer_raw_host([raw_v], vlen, raw_t, tlen, chunk_videos=chunk)
        for name in ("boundary", "nms_windows", "count", "saliency"):
            assert torch.equal(o[name], getattr(r, name).cpu()), (name, chunk)
