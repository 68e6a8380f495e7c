"""TEST INFRASTRUCTURE ONLY - imports the UNMODIFIED reference.

The reference tree is looked up at $FLASHVTG_REFERENCE, /root/reference (the build container) and then
baseline/_ref/ - the git-ignored copy __graft_entry__.build() installs next to the repo so that it travels
to the GPU box for `bench.py --impl reference` (the reference is pure Python with no setup.py, so "install"
is a file copy of its packages; oracle/shim supplies the two imports that cannot be installed offline).
Used by tests/golden/make_golden.py to generate the committed fixtures, by the CPU tests that cross-check
oracle/forward.py against the real thing when the tree is present, and by bench.py's reference arm.
Nothing in the product (flashvtg_b200/) imports this module.
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace
from pathlib import Path

SHIM = Path(__file__).resolve().parent / "shim"
INSTALLED = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
REF_PACKAGES = ("FlashVTG", "blocks", "utils", "standalone_eval")


def _find_root() -> Path:
    env = os.environ.get("FLASHVTG_REFERENCE")
    for cand in ([Path(env)] if env else []) + [Path("/root/reference"), INSTALLED]:
        if (cand / "FlashVTG" / "model.py").exists():
            return cand
    return Path("/root/reference")


REF_ROOT = _find_root()


def install(src: Path = Path("/root/reference"), dst: Path = INSTALLED) -> bool:
    """Copy the reference's Python packages (unmodified) into baseline/_ref/ - the counterpart of
    `pip install --target baseline/_ref` for a tree without a build system.  Returns False when `src` is absent."""
    import shutil
    if not (src / "FlashVTG" / "model.py").exists():
        return False
    for pkg in REF_PACKAGES:
        if (src / pkg).is_dir():
            shutil.copytree(src / pkg, dst / pkg, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.jsonl", "*.json", "*.md"))
    for name in ("LISCENSE", "LICENSE"):
        if (src / name).exists():
            shutil.copy2(src / name, dst / name)
    return True


def available() -> bool:
    return (REF_ROOT / "FlashVTG" / "model.py").exists()


def _ensure_path():
    for p in (str(SHIM), str(REF_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:  # noqa: BLE001  (imported-but-unused in inference.py:12)
            sys.modules["wandb"] = types.ModuleType("wandb")


def reference_args(cfg) -> Namespace:
    """The argparse namespace build_model1 / FlashVTG read, from a ModelConfig."""
    _ensure_path()
    import nncore
    model_cfg = nncore.Config(dict(model=dict(
        strides=tuple(cfg.strides), buffer_size=cfg.buffer_size, max_num_moment=cfg.max_num_moment,
        pyramid_cfg=dict(type="ConvPyramid"), pooling_cfg=dict(type="AdaPooling"),
        class_head_cfg=dict(type="ConvHead", kernal_size=3),
        coord_head_cfg=dict(type="ConvHead", kernal_size=cfg.coord_kernel), loss_cfg=None)))
    return Namespace(
        device="cpu", hidden_dim=cfg.hidden_dim, dropout=0.1, nheads=cfg.nheads,
        dim_feedforward=cfg.dim_feedforward, enc_layers=cfg.enc_layers, pre_norm=False,
        t2v_layers=cfg.t2v_layers, dummy_layers=cfg.dummy_layers, num_dummies=cfg.num_dummies,
        position_embedding="sine", max_q_l=cfg.max_q_l, input_dropout=0.5,
        t_feat_dim=cfg.t_feat_dim, v_feat_dim=cfg.v_feat_dim, n_input_proj=cfg.n_input_proj,
        kernel_size=cfg.kernel_size, num_conv_layers=cfg.num_conv_layers,
        num_mlp_layers=cfg.num_mlp_layers, label_loss_coef=4, lw_saliency=1.0, lw_reg=1.0,
        lw_cls=5.0, lw_sal=0.1, eos_coef=0.1, saliency_margin=0.2, dset_name=cfg.dset_name,
        clip_length=cfg.clip_length, use_neg=False, cfg=model_cfg, span_loss_type="l1",
        contrastive_align_loss=False, aux_loss=False)


def build_reference_model(cfg, state_dict=None):
    """build_model1(args) (FlashVTG/model.py:792) in eval mode, optionally loading `state_dict`
    with strict=True (the reference's checkpoint ABI, inference.py:471)."""
    _ensure_path()
    from FlashVTG.model import build_model1
    model, _ = build_model1(reference_args(cfg))
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model


def reference_forward_bs1(model, src_txt, src_txt_mask, src_vid, src_vid_mask):
    """One unmodified bs=1 eval forward (FlashVTG/model.py:138)."""
    import torch
    with torch.no_grad():
        return model(src_txt=src_txt, src_txt_mask=src_txt_mask, src_vid=src_vid,
                     src_vid_mask=src_vid_mask, vid=None, qid=None, targets={})


def reference_nms():
    """post_processing_mr_nms (FlashVTG/inference.py:36) without importing the dataset stack."""
    _ensure_path()
    import importlib.util
    src = (REF_ROOT / "FlashVTG" / "inference.py").read_text()
    start = src.index("def post_processing_mr_nms")
    end = src.index("def eval_epoch_post_processing")
    import nncore
    import torch
    from nncore.ops import temporal_iou
    ns = dict(torch=torch, nncore=nncore, temporal_iou=temporal_iou)
    exec(compile(src[start:end], str(REF_ROOT / "FlashVTG" / "inference.py"), "exec"), ns)  # noqa: S102
    del importlib
    return ns["post_processing_mr_nms"]


def reference_temporal_nms():
    _ensure_path()
    from utils.temporal_nms import temporal_nms
    return temporal_nms


def reference_postprocessor():
    _ensure_path()
    src = (REF_ROOT / "FlashVTG" / "postprocessing.py").read_text()
    start = src.index("class PostProcessorDETR")
    import torch

    def tqdm(x, **kw):
        return x
    ns = dict(torch=torch, tqdm=tqdm)
    exec(compile(src[start:], str(REF_ROOT / "FlashVTG" / "postprocessing.py"), "exec"), ns)  # noqa: S102
    return ns["PostProcessorDETR"]


def reference_eval_functions():
    """The reference's own eval loop pieces, imported unmodified: (compute_mr_results, post_processing_mr_nms)
    from FlashVTG/inference.py:232,36."""
    _ensure_path()
    import importlib
    inf = importlib.import_module("FlashVTG.inference")
    return inf.compute_mr_results, inf.post_processing_mr_nms


def reference_eval_opt(cfg) -> Namespace:
    """The option fields compute_mr_results / eval_epoch_post_processing read (inference.py:246-352,79-87)."""
    return Namespace(device="cpu", pin_memory=False, clip_length=cfg.clip_length, span_loss_type="l1",
                     dset_name=cfg.dset_name, v_feat_dim=cfg.v_feat_dim, nms_thd=cfg.nms_thd,
                     max_before_nms=50, max_after_nms=10, nms_type=cfg.nms_type)


def bs1_loader(batch, n):
    """What DataLoader(StartEndDataset, collate_fn=start_end_collate, batch_size=1) yields for n synthetic videos:
    (query_meta, batched_model_inputs) with (tensor, mask) pairs cut to the true lengths (start_end_collate pads
    to the batch maximum, which at bs=1 is the video's own length)."""
    out = []
    for b in range(n):
        lv, lt = int(batch["vid_len"][b]), int(batch["txt_len"][b])
        meta = [dict(qid=b, query="synthetic", vid=f"v{b}", duration=float(batch["duration"][b]))]
        inputs = dict(query_feat=(batch["src_txt"][b:b + 1, :lt], batch["src_txt_mask"][b:b + 1, :lt]),
                      video_feat=(batch["src_vid"][b:b + 1, :lv], batch["src_vid_mask"][b:b + 1, :lv]),
                      vid=[f"v{b}"], qid=[b])
        out.append((meta, inputs))
    return out
