"""Developer loop: device-resident step time + per-kernel-class CUDA-event times of one preset (not a bench line).
usage: python tools/quick_bench.py [preset] [steps]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from flashvtg_b200 import _lib, synth  # noqa: E402
from flashvtg_b200.config import PRESETS  # noqa: E402
from flashvtg_b200.model import FlashVTGB200  # noqa: E402


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "qvh_iv2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cfg = PRESETS[preset]
    B, LV, LT, _ = WORKLOADS[preset]
    dev = torch.device("cuda:0")
    model = FlashVTGB200(cfg).eval()
    model.load_state_dict(synth.make_state_dict(cfg, 2024), strict=True)
    lib = _lib.load()
    nb = min(64, B)
    base = synth.make_inputs(cfg, nb, LV, LT, seed=1234)
    d = {k: v.repeat(B // nb, *([1] * (v.dim() - 1))).contiguous().to(dev) for k, v in base.items()}
    out = model.alloc_outputs(B, LV, dev, "normal")

    def step():
        return model.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], duration=d["duration"],
                           nms="normal", uniform_len=True, out=out)
    for _ in range(5):
        r = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.fvtg_prof_enable(1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = len(_lib.PROF_CLASSES)
    ms_c, ln_c = (C.c_double * n)(), (C.c_int64 * n)()
    _lib.check(lib.fvtg_prof_collect(ms_c, ln_c, n), "prof")
    lib.fvtg_prof_enable(0)
    print(json.dumps({"preset": preset, "ms_per_step": round(ms, 4), "videos_per_s": round(B / ms * 1e3),
                      "launches": r.launches,
                      "class_ms": {c: round(ms_c[i] / 3, 4) for i, c in enumerate(_lib.PROF_CLASSES)},
                      "class_launches": {c: int(ln_c[i] // 3) for i, c in enumerate(_lib.PROF_CLASSES)}}))


if __name__ == "__main__":
    main()
