"""CPU tests of the host side: C-ABI surface, struct layout, weight ABI, config logic, error
behaviour.  No compute entry point is called (there is no GPU here and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flashvtg_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fvtg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from flashvtg_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/flashvtg_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in flashvtg_b200/_lib.py"
    assert lib.fvtg_abi_version() == _lib.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True)
    exported = set(re.findall(r"\b(fvtg_[a-z0-9_]+)\b", nm.stdout))
    assert set(syms) <= exported
    # debug / test hooks and the probe micro-benchmarks live in the debug flavour only
    assert not [s for s in exported if s.startswith("fvtg_dbg_")], "debug hooks leaked into the product library"
    dbg_text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "flashvtg_b200_dbg.h")).read(), flags=re.S)
    dbg_syms = sorted(set(re.findall(r"\b(fvtg_dbg_[a-z0-9_]+)\s*\(", dbg_text)))
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.DBG_LIB_PATH)], capture_output=True, text=True)
    dbg_exported = set(re.findall(r"\b(fvtg_[a-z0-9_]+)\b", nm.stdout))
    assert set(dbg_syms) <= dbg_exported and set(syms) <= dbg_exported
    for s in dbg_syms:
        assert s in _lib.DBG_SIGNATURES, f"{s} has no ctypes signature in flashvtg_b200/_lib.py"


def test_ctypes_structs_match_the_c_header():
    """sizeof / offsetof of every boundary struct as gcc sees the header == the ctypes mirror."""
    from flashvtg_b200 import _lib
    structs = {"FvtgCfg": ["abi_version", "max_num_moment", "clip_len"],
               "FvtgLN": ["g", "b"], "FvtgLinear": ["w", "b"],
               "FvtgInProj": ["ln0", "fc0", "ln1", "fc1", "fc0_wsum"],
               "FvtgEncLayer": ["in_proj", "out_proj", "norm1", "ff1", "ff2", "norm2", "prelu"],
               "FvtgPyrConv": ["conv", "ln"],
               "FvtgScoreHead": ["conv", "mlp", "last_w", "last_b"],
               "FvtgWeights": ["vid", "txt", "dummy_tok", "dummy", "t2v", "enc", "sal_w1", "pyr",
                               "cls", "conf", "coord1", "coord2", "coef", "x"],
               "FvtgBatch": ["B", "Lv", "Lt", "uniform_vid_len", "vid", "txt", "vid_len", "txt_len"],
               "FvtgRawBatch": ["B", "Lv", "Lt", "n_groups", "group_dim", "t_dim", "dtype", "normalize_v",
                                "normalize_t", "use_tef", "vid", "txt", "vid_len", "txt_len"],
               "FvtgEvalBatch": ["n_queries", "max_pred", "max_gt", "max_sal", "max_clips", "pred_win", "pred_cnt",
                                 "gt_win", "gt_cnt", "pred_sal", "pred_sal_len", "gt_sal", "gt_clips"],
               "FvtgFusionOut": ["video_emb", "saliency", "t2v", "dummy_tokens"],
               "FvtgHeadsOut": ["n_max", "cls_logit", "conf_logit", "coord"],
               "FvtgDecodeParams": ["nms_thd", "x", "clip_len", "inv_clip_len", "min_ts", "max_ts",
                                    "topk", "num_levels", "clip_ts", "round_multiple", "nms_mode",
                                    "max_after_nms"],
               "FvtgDecodeOut": ["boundary", "windows", "nms_windows", "nms_order", "count", "nms_count"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for name, fields in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "abi.c")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c11", "-o", os.path.join(d, "abi"), src], check=True)
        out = subprocess.run([os.path.join(d, "abi")], capture_output=True, text=True, check=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for name, fields in structs.items():
        cls = getattr(_lib, name)
        assert int(got[name]) == C.sizeof(cls), name
        for f in fields:
            assert int(got[f"{name}.{f}"]) == getattr(cls, f).offset, f"{name}.{f}"


def test_compute_entry_points_fail_loudly_without_a_gpu(lib):
    """No CPU fallback: on this GPU-less box every compute entry point returns an error code and a
    message; on a GPU box the same calls with NULL arguments return FVTG_EINVAL."""
    from flashvtg_b200 import _lib
    rc = lib.fvtg_temporal_nms(None, None, 1, 4, 0.7, 0, 10, None, None, None, None)
    assert rc in (_lib.EINVAL, _lib.EARCH)
    assert lib.fvtg_last_error()
    cfg = _lib.FvtgCfg()
    assert lib.fvtg_workspace_bytes(C.byref(cfg), 1, 75, 32) == 0  # abi_version 0 -> rejected


def test_model_refuses_cpu_tensors_and_bad_checkpoints():
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.model import FlashVTGB200
    cfg = PRESETS["qvh_iv2"]
    m = FlashVTGB200(cfg).eval()
    with pytest.raises(RuntimeError, match="no weights"):
        m.state_dict()
    sd = synth.make_state_dict(cfg, 2024)
    m.load_state_dict(sd, strict=True)
    b = synth.make_inputs(cfg, 1, 75, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(src_txt=b["src_txt"], src_txt_mask=b["src_txt_mask"], src_vid=b["src_vid"],
          src_vid_mask=b["src_vid_mask"], vid=None, qid=None, targets={})
    bad = dict(sd)
    bad.pop("coef")
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict(bad, strict=True)
    bad = dict(sd)
    bad["extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(bad, strict=True)
    bad = dict(sd)
    bad["x"] = torch.zeros(2)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(bad, strict=True)
    m.train()
    with pytest.raises(RuntimeError, match="inference"):
        m(src_txt=b["src_txt"], src_txt_mask=b["src_txt_mask"], src_vid=b["src_vid"],
          src_vid_mask=b["src_vid_mask"], vid=None, qid=None, targets={})


@pytest.mark.parametrize("name", ["qvh_iv2", "qvh_sfclip", "charades_vgg", "charades_iv2", "tacos", "tacos_deep"])
def test_weight_abi_matches_synth_and_reference(name):
    """expected_shapes (the checkpoint ABI) == what synth emits == the unmodified reference's
    state_dict when /root/reference is importable (build container only)."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.weights import PackedWeights, expected_shapes
    cfg = PRESETS[name]
    exp = expected_shapes(cfg)
    sd = synth.make_state_dict(cfg, 1)
    assert set(exp) == set(sd)
    for k, shp in exp.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.build_reference_model(cfg).state_dict()
        assert set(ref) == set(exp)
        for k, v in ref.items():
            assert tuple(v.shape) == tuple(exp[k]), k
    # packing runs on CPU tensors too (it is pure layout work): spot-check the folded layouts
    W = PackedWeights(cfg, sd, torch.device("cpu"))
    assert W.struct.x == pytest.approx(float(sd["x"]))
    assert W.nbytes > 20e6


def test_packed_layouts_fold_taps_and_token_type():
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.weights import PackedWeights
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 3, spread=True)
    W = PackedWeights(cfg, sd, torch.device("cpu"))
    k = cfg.kernel_size
    conv = W.view(W.struct.cls.conv[0].w, (256, k * 256), torch.bfloat16).float()   # [256][k*256], tap-major K
    ref = sd["class_head.convs.0.weight"][:, :, 0, :].permute(0, 2, 1).reshape(256, -1)
    assert torch.allclose(conv, ref.to(torch.bfloat16).float())
    pyr = W.view(W.struct.pyr[2][1].conv.w, (256, 512), torch.bfloat16).float()
    ref = sd["pyramid.blocks.2.6.weight"].permute(0, 2, 1).reshape(256, 512)
    assert torch.allclose(pyr, ref.to(torch.bfloat16).float())
    b = W.view(W.struct.vid.fc1.b, (256,))
    assert torch.allclose(b, sd["input_vid_proj.1.net.1.bias"] + sd["token_type_embeddings.weight"][1])
    fc0 = W.view(W.struct.vid.fc0.w, (256, 832), torch.bfloat16)
    assert float(fc0[:, 770:].float().abs().sum()) == 0.0
    # LayerNorm over the raw dim is folded into the first projection (csrc/inproj.cu)
    w0, g0, b0 = (sd[f"input_vid_proj.0.{n}"] for n in ("net.1.weight", "LayerNorm.weight", "LayerNorm.bias"))
    assert torch.equal(fc0[:, :770].float(), (w0 * g0[None, :]).to(torch.bfloat16).float())
    assert torch.allclose(W.view(W.struct.vid.fc0.b, (256,)), w0 @ b0 + sd["input_vid_proj.0.net.1.bias"], atol=1e-6)
    assert torch.allclose(W.view(W.struct.vid.fc0_wsum, (256,)), fc0.float().sum(1), atol=1e-5)
    c2 = W.view(W.struct.coord2.w, (16, 768), torch.bfloat16).float()
    assert float(c2[2:].abs().sum()) == 0.0
    ref = sd["coord_head.module.3.weight"].permute(0, 2, 1).reshape(2, 768)
    assert torch.equal(c2[:2], ref.to(torch.bfloat16).float())
    assert torch.equal(W.view(W.struct.sal_w2t, (256, 256)), sd["saliency_proj2.weight"].t().contiguous())
    assert W.struct.t2v[0].in_proj.w is None and W.struct.enc[0].in_proj.w is not None


def test_pack_weights_is_strict_like_torch():
    """fvtg_pack_weights (the C-ABI checkpoint loader) rejects a missing key, an unexpected key and a wrong
    element count, naming the key - the strict=True behaviour of inference.py:471."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    from flashvtg_b200.weights import PackedWeights
    cfg = PRESETS["qvh_iv2"]
    sd = synth.make_state_dict(cfg, 3)
    bad = dict(sd)
    bad.pop("class_head.convs.0.bias")
    with pytest.raises(RuntimeError, match="Missing key.*class_head.convs.0.bias"):
        PackedWeights(cfg, bad, torch.device("cpu"))
    bad = dict(sd)
    bad["extra.weight"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="Unexpected key.*extra.weight"):
        PackedWeights(cfg, bad, torch.device("cpu"))
    bad = dict(sd)
    bad["coef"] = torch.ones(3)
    with pytest.raises(RuntimeError, match="size mismatch for coef"):
        PackedWeights(cfg, bad, torch.device("cpu"))


def test_config_presets_and_point_counts():
    from flashvtg_b200.config import PRESETS, gemm_flops_per_video, postprocessor_preset
    q = PRESETS["qvh_iv2"]
    assert q.level_lengths(75) == [75, 37, 18, 9, 4] and q.num_points(75) == 143
    assert PRESETS["qvh_sfclip"].num_points(75) == 139
    assert PRESETS["tacos_deep"].level_lengths(20) == [20, 10, 5, 2, 1]     # level 32 skipped (blocks.py:56)
    assert postprocessor_preset(q) == (True, 0.0, 150.0, True)
    assert postprocessor_preset(PRESETS["tacos"]) == (False, 0.0, 50000.0, True)
    # the reference tests v_feat_dim == 4096 AFTER TEF added 2: the 360 s branch never fires for 4098
    assert postprocessor_preset(PRESETS["charades_vgg"])[2] == 150.0
    assert 1.5e9 < gemm_flops_per_video(q, 75, 32) < 1.65e9
    with pytest.raises(ValueError):
        q.with_(strides=(1, 2, 3))
    with pytest.raises(ValueError):
        q.with_(hidden_dim=512)


def test_from_opt_reads_the_reference_option_names():
    from types import SimpleNamespace as NS
    from flashvtg_b200.config import ModelConfig

    class Cfg(dict):
        __getattr__ = dict.__getitem__
    model = Cfg(strides=(1, 2, 4, 8), buffer_size=1024, max_num_moment=50,
                coord_head_cfg=dict(type="ConvHead", kernal_size=3))
    opt = NS(v_feat_dim=2818, t_feat_dim=512, num_dummies=35, dummy_layers=3, t2v_layers=8, enc_layers=3,
             kernel_size=5, num_conv_layers=2, num_mlp_layers=5, clip_length=2, hidden_dim=256, nheads=8,
             dim_feedforward=1024, n_input_proj=2, max_q_l=-1, dset_name="tacos", nms_thd=0.7,
             cfg=NS(model=model))
    cfg = ModelConfig.from_opt(opt)
    assert cfg.t2v_layers == 8 and cfg.max_q_l == 100 and cfg.num_levels == 4


def test_synthetic_inputs_follow_the_loader_contract():
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    cfg = PRESETS["qvh_sfclip"]
    b = synth.make_inputs(cfg, 4, 75, 32, seed=9, ragged=True)
    v = b["src_vid"]
    for i in range(4):
        lv = int(b["vid_len"][i])
        assert float(v[i, lv:].abs().sum()) == 0.0                      # zero padding
        np.testing.assert_allclose(v[i, :lv, :2304].norm(dim=-1).numpy(), 1.0, atol=1e-3)  # per-group L2
        np.testing.assert_allclose(v[i, :lv, 2304:2816].norm(dim=-1).numpy(), 1.0, atol=1e-3)
        np.testing.assert_allclose(v[i, :lv, -2].numpy(), np.arange(lv) / lv, atol=1e-6)    # TEF
        assert float(b["duration"][i]) == lv * cfg.clip_length
    assert b["src_vid_mask"].sum(1).int().tolist() == b["vid_len"].tolist()


def test_product_never_imports_or_calls_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    import pathlib
    import re
    pkg = pathlib.Path(__file__).resolve().parent.parent / "flashvtg_b200"
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|oracle[./]", re.M)
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert not pat.search(p.read_text()), f"{p} references the oracle"
