"""CPU, build container only: the oracle against the UNMODIFIED reference imported from
/root/reference (skipped where the tree is absent, e.g. on the GPU box - the committed golden
fixtures cover that case)."""
import numpy as np
import pytest
import torch

from helpers import max_rel
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")


@pytest.mark.parametrize("name,lv,lt", [("qvh_iv2", 75, 32), ("charades_iv2", 27, 9), ("tacos_deep", 70, 6)])
def test_forward_oracle_equals_reference_bs1(name, lv, lt):
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from oracle import forward as O
    cfg = PRESETS[name]
    sd = synth.make_state_dict(cfg, 77, spread=True)
    batch = synth.make_inputs(cfg, 1, lv, lt, seed=78)
    model = ref_loader.build_reference_model(cfg, sd)
    ref = ref_loader.reference_forward_bs1(model, batch["src_txt"], batch["src_txt_mask"],
                                           batch["src_vid"], batch["src_vid_mask"])
    o = O.forward_batch(sd, cfg, batch)[0]
    assert max_rel(o["saliency"], ref["saliency_scores"][0]) < 2e-5
    assert max_rel(o["t2vattn"], ref["t2vattnvalues"][0]) < 2e-5
    assert max_rel(o["dummy_tokens"], ref["dummy_tokens"][0]) < 2e-5
    rb = ref["_out"]["boundary"].numpy()
    assert o["boundary"].shape == rb.shape
    assert max_rel(o["boundary"], rb) < 2e-5


def test_nms_and_postproc_oracles_equal_reference_functions():
    from oracle import postproc as P
    nms = ref_loader.reference_nms()
    hull = ref_loader.reference_temporal_nms()
    rng = np.random.Generator(np.random.PCG64(5))
    for n in (1, 2, 9, 50):
        st = np.round(rng.uniform(0, 140, size=n) / 2) * 2
        w = np.stack([st, st + np.round(rng.uniform(0, 40, size=n) / 2) * 2,
                      np.round(rng.uniform(0, 1, size=n), 4)], 1).tolist()
        for mode in ("normal", "linear"):
            res = nms([dict(pred_relevant_windows=[list(r) for r in w])], nms_thd=0.7,
                      max_before_nms=1000, max_after_nms=10, nms_type=mode)[0]["pred_relevant_windows"]
            out, _, _ = P.nms_reference_order(w, 0.7, mode)
            got, want = np.asarray(out, np.float64), np.asarray(res, np.float64)
            np.testing.assert_array_equal(np.sort(got[:, 2]), np.sort(want[:, 2]))
            distinct = np.array([np.sum(want[:, 2] == s) == 1 for s in want[:, 2]])
            np.testing.assert_array_equal(got[distinct], want[distinct])
        kept, _ = P.temporal_nms_hull(w, 0.5, 7)
        assert kept == hull([list(r) for r in w], 0.5, max_after_nms=7)
