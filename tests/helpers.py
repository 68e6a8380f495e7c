"""Shared helpers of the test-suite (oracle access lives here: tests may import oracle/)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_forward_index():
    with open(os.path.join(GOLDEN, "forward_index.json")) as f:
        return json.load(f)


def regen_case(entry):
    """Weights + inputs of a golden forward case, regenerated from its seeds and pinned by the
    checksums stored with the fixture."""
    from flashvtg_b200 import synth
    from flashvtg_b200.config import PRESETS
    cfg = PRESETS[entry["preset"]]
    sd = synth.make_state_dict(cfg, entry["weight_seed"], spread=entry["spread"])
    batch = synth.make_inputs(cfg, entry["B"], entry["Lv"], entry["Lt"], seed=entry["input_seed"],
                              ragged=entry["ragged"])
    assert abs(synth.state_dict_checksum(sd) - entry["weights_checksum"]) <= 1e-6 * max(
        1.0, abs(entry["weights_checksum"])), "weight RNG drifted from the golden fixture"
    chk = float(batch["src_vid"].double().sum() + batch["src_txt"].double().sum())
    assert abs(chk - entry["inputs_checksum"]) <= 1e-6 * max(1.0, abs(chk)), "input RNG drifted"
    assert batch["vid_len"].tolist() == entry["vid_len"]
    assert batch["txt_len"].tolist() == entry["txt_len"]
    gold = np.load(os.path.join(GOLDEN, entry["file"]))
    return cfg, sd, batch, gold


def max_rel(a, b):
    """Per-tensor max-norm relative error max|a-b| / max|b| (the parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


_oracle_lib = None


def oracle_c():
    """The C restatement of the NMS arithmetic (oracle/nms_c.c), built on demand with gcc."""
    global _oracle_lib
    if _oracle_lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        src = os.path.join(ROOT, "oracle", "nms_c.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True,
                           capture_output=True)
        lib = C.CDLL(path)
        lib.oracle_nms_f32.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                       C.c_void_p]
        lib.oracle_nms_f32.restype = None
        lib.oracle_nms_hull.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                        C.c_void_p]
        lib.oracle_nms_hull.restype = C.c_int
        _oracle_lib = lib
    return _oracle_lib


def c_nms_f32(windows, thd, mode):
    rows = np.array(windows, dtype=np.float64).astype(np.float32).reshape(-1, 3).copy()
    n = rows.shape[0]
    order = np.zeros(n, np.int32)
    sel = np.zeros(n, np.int32)
    oracle_c().oracle_nms_f32(rows.ctypes.data, n, thd, 0 if mode == "normal" else 1,
                              order.ctypes.data, sel.ctypes.data)
    return rows, order, sel


def c_nms_hull(windows, thd, max_after):
    rows = np.array(windows, dtype=np.float64).reshape(-1, 3).copy()
    n = rows.shape[0]
    out = np.zeros((n, 3), np.float64)
    src = np.zeros(n, np.int32)
    cnt = oracle_c().oracle_nms_hull(rows.ctypes.data, n, thd, max_after, out.ctypes.data,
                                     src.ctypes.data)
    return out[:cnt], src[:cnt]


def denan(o):
    if isinstance(o, str) and o == "nan":
        return float("nan")
    if isinstance(o, list):
        return [denan(x) for x in o]
    return o


def to_torch_batch(batch, device):
    return {k: v.to(device) for k, v in batch.items()}


__all__ = ["load_forward_index", "regen_case", "max_rel", "c_nms_f32", "c_nms_hull", "denan",
           "to_torch_batch", "torch"]
