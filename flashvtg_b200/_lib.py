"""ctypes binding of include/flashvtg_b200.h.

The shared object is the product: there is no Python / PyTorch fallback.  If it
is missing (or was built for another source revision) loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libflashvtg_b200.so"

ABI_VERSION = 2
MAX_LAYERS = 8
MAX_LEVELS = 8
MAX_CONVS = 4
MAX_MLP = 8
MAX_TOPK = 64
PROF_CLASSES = ("gemm", "attention", "inproj", "decode_nms", "other", "layer")

NMS_NONE, NMS_NORMAL, NMS_LINEAR, NMS_HULL = -1, 0, 1, 2

OK, EINVAL, EARCH, ELAUNCH, EWORKSPACE = 0, -1, -2, -3, -4

i32, f32, vp = C.c_int32, C.c_float, C.c_void_p


class FvtgCfg(C.Structure):
    _fields_ = [(n, i32) for n in (
        "abi_version", "v_dim", "t_dim", "v_dim_pad", "t_dim_pad", "num_dummies", "dummy_layers",
        "t2v_layers", "enc_layers", "num_levels", "head_k", "num_conv_layers", "num_mlp_layers",
        "coord_k", "max_num_moment")] + [("clip_len", f32)]


class FvtgLN(C.Structure):
    _fields_ = [("g", vp), ("b", vp)]


class FvtgLinear(C.Structure):
    _fields_ = [("w", vp), ("b", vp)]


class FvtgInProj(C.Structure):
    _fields_ = [("ln0", FvtgLN), ("fc0", FvtgLinear), ("ln1", FvtgLN), ("fc1", FvtgLinear),
                ("fc0_wsum", vp)]


class FvtgEncLayer(C.Structure):
    _fields_ = [("in_proj", FvtgLinear), ("out_proj", FvtgLinear), ("norm1", FvtgLN),
                ("ff1", FvtgLinear), ("ff2", FvtgLinear), ("norm2", FvtgLN),
                ("prelu", f32), ("_pad", i32)]


class FvtgPyrConv(C.Structure):
    _fields_ = [("conv", FvtgLinear), ("ln", FvtgLN)]


class FvtgScoreHead(C.Structure):
    _fields_ = [("conv", FvtgLinear * MAX_CONVS), ("mlp", FvtgLinear * MAX_MLP),
                ("last_w", vp), ("last_b", f32), ("_pad", i32)]


class FvtgWeights(C.Structure):
    _fields_ = [("vid", FvtgInProj), ("txt", FvtgInProj),
                ("dummy_tok", vp), ("dummy_pos", vp),
                ("dummy", FvtgEncLayer * MAX_LAYERS),
                ("t2v", FvtgEncLayer * MAX_LAYERS),
                ("enc", FvtgEncLayer * MAX_LAYERS),
                ("sal_w1", vp), ("sal_b1", vp), ("sal_w2t", vp), ("sal_b2", vp),
                ("pyr", (FvtgPyrConv * MAX_LEVELS) * MAX_LEVELS),
                ("cls", FvtgScoreHead), ("conf", FvtgScoreHead),
                ("coord1", FvtgLinear), ("coord2", FvtgLinear),
                ("coef", f32 * MAX_LEVELS), ("x", f32), ("_pad", i32)]


class FvtgParam(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", vp), ("numel", C.c_int64)]


class FvtgBatch(C.Structure):
    _fields_ = [("B", i32), ("Lv", i32), ("Lt", i32), ("uniform_vid_len", i32),
                ("vid", vp), ("txt", vp), ("vid_len", vp), ("txt_len", vp)]


class FvtgFusionOut(C.Structure):
    _fields_ = [("video_emb", vp), ("saliency", vp), ("t2v", vp), ("dummy_tokens", vp)]


class FvtgHeadsOut(C.Structure):
    _fields_ = [("n_max", i32), ("_pad", i32), ("cls_logit", vp), ("conf_logit", vp),
                ("coord", vp)]


class FvtgDecodeParams(C.Structure):
    _fields_ = [("nms_thd", C.c_double), ("x", f32), ("clip_len", f32), ("inv_clip_len", f32),
                ("min_ts", f32), ("max_ts", f32), ("topk", i32), ("num_levels", i32),
                ("clip_ts", i32), ("round_multiple", i32), ("nms_mode", i32),
                ("max_after_nms", i32), ("_pad", i32)]


class FvtgDecodeOut(C.Structure):
    _fields_ = [("boundary", vp), ("windows", vp), ("nms_windows", vp), ("nms_order", vp),
                ("count", vp), ("nms_count", vp)]


RAW_MAX_GROUPS = 4
RAW_F32, RAW_F16, RAW_BF16 = 0, 1, 2


class FvtgRawBatch(C.Structure):
    _fields_ = [("B", i32), ("Lv", i32), ("Lt", i32), ("n_groups", i32),
                ("group_dim", i32 * RAW_MAX_GROUPS), ("t_dim", i32), ("dtype", i32),
                ("normalize_v", i32), ("normalize_t", i32), ("use_tef", i32), ("_pad", i32),
                ("vid", vp * RAW_MAX_GROUPS), ("txt", vp), ("vid_len", vp), ("txt_len", vp)]


class FvtgEvalBatch(C.Structure):
    _fields_ = [("n_queries", i32), ("max_pred", i32), ("max_gt", i32), ("max_sal", i32), ("max_clips", i32),
                ("_pad", i32), ("pred_win", vp), ("pred_cnt", vp), ("gt_win", vp), ("gt_cnt", vp),
                ("pred_sal", vp), ("pred_sal_len", vp), ("gt_sal", vp), ("gt_clips", vp)]


# name -> (restype, argtypes); also the list the ABI test checks against the header.
SIGNATURES = {
    "fvtg_workspace_bytes": (C.c_size_t, [C.POINTER(FvtgCfg), i32, i32, i32]),
    "fvtg_chunk_videos": (i32, [C.POINTER(FvtgCfg), i32, i32]),
    "fvtg_packed_weights_bytes": (C.c_size_t, [C.POINTER(FvtgCfg)]),
    "fvtg_pack_weights": (i32, [C.POINTER(FvtgCfg), C.POINTER(FvtgParam), i32, vp, C.c_size_t,
                                C.POINTER(FvtgWeights), vp]),
    "fvtg_pack_weights_host": (i32, [C.POINTER(FvtgCfg), C.POINTER(FvtgParam), i32, vp, C.c_size_t, vp,
                                     C.POINTER(FvtgWeights)]),
    "fvtg_fusion_fwd": (i32, [C.POINTER(FvtgCfg), C.POINTER(FvtgWeights), C.POINTER(FvtgBatch),
                              C.POINTER(FvtgFusionOut), vp, C.c_size_t, vp]),
    "fvtg_pyramid_heads_fwd": (i32, [C.POINTER(FvtgCfg), C.POINTER(FvtgWeights), i32, i32, vp, vp,
                                     C.POINTER(FvtgHeadsOut), vp, C.c_size_t, vp]),
    "fvtg_decode_nms": (i32, [C.POINTER(FvtgDecodeParams), i32, i32, i32, vp, vp, vp, vp, vp,
                              C.POINTER(FvtgDecodeOut), vp]),
    "fvtg_temporal_nms": (i32, [vp, vp, i32, i32, C.c_double, i32, i32, vp, vp, vp, vp]),
    "fvtg_temporal_nms_hull_f64": (i32, [vp, vp, i32, i32, C.c_double, i32, vp, vp, vp]),
    "fvtg_forward": (i32, [C.POINTER(FvtgCfg), C.POINTER(FvtgWeights), C.POINTER(FvtgBatch), vp,
                           C.POINTER(FvtgDecodeParams), C.POINTER(FvtgFusionOut),
                           C.POINTER(FvtgHeadsOut), C.POINTER(FvtgDecodeOut), vp, C.c_size_t, vp]),
    "fvtg_prepare_inputs": (i32, [C.POINTER(FvtgRawBatch), vp, vp, vp, vp, vp]),
    "fvtg_eval_submission": (i32, [C.POINTER(FvtgEvalBatch), i32, vp, vp, vp, vp, vp, vp]),
    "fvtg_last_launch_count": (C.c_int64, []),
    "fvtg_last_error": (C.c_char_p, []),
    "fvtg_abi_version": (i32, []),
    "fvtg_prof_enable": (None, [i32]),
    "fvtg_prof_collect": (i32, [vp, vp, i32]),
}
# include/flashvtg_b200_dbg.h: only in libflashvtg_b200_dbg.so (product sources + -DFVTG_DEBUG_HOOKS + probe.cu)
DBG_SIGNATURES = {
    "fvtg_dbg_set_trace": (None, [vp]),
    "fvtg_dbg_gemm": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "fvtg_dbg_stream_probe": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "fvtg_dbg_inproj": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "fvtg_dbg_mufu_probe": (i32, [i32, i32, vp, vp]),
    "fvtg_dbg_store_probe": (i32, [i32, i32, i32, vp, vp, vp]),
}
DBG_LIB_PATH = PKG_DIR / "libflashvtg_b200_dbg.so"

_lib = None
_dbg = None


def load() -> C.CDLL:
    """Load libflashvtg_b200.so (built in-tree by flashvtg_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("FVTG_DEBUG_LIB", "0") not in ("", "0"):
        _lib = load_debug()   # tools/trace_*.py: the whole path through the hook-carrying build
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m flashvtg_b200._build` (needs nvcc). "
            "flashvtg_b200 has no CPU or PyTorch fallback.")
    import torch  # noqa: F401  (loads libcudart.so.12, which the library links against)
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError:
        C.CDLL("/usr/local/cuda/lib64/libcudart.so.12", mode=C.RTLD_GLOBAL)
        lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == symbol missing from the build
        fn.restype = res
        fn.argtypes = args
    if lib.fvtg_abi_version() != ABI_VERSION:
        raise RuntimeError("libflashvtg_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def load_debug() -> C.CDLL:
    """libflashvtg_b200_dbg.so: every product symbol plus the fvtg_dbg_* hooks (kernel-level unit tests, tools/)."""
    global _dbg
    if _dbg is not None:
        return _dbg
    if not DBG_LIB_PATH.exists():
        raise RuntimeError(f"{DBG_LIB_PATH} is missing: run `python -m flashvtg_b200._build` (needs nvcc)")
    import torch  # noqa: F401
    lib = C.CDLL(str(DBG_LIB_PATH))
    for name, (res, args) in {**SIGNATURES, **DBG_SIGNATURES}.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _dbg = lib
    return lib


def check(rc: int, what: str, lib=None) -> None:
    if rc != OK:
        msg = (lib or load()).fvtg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
