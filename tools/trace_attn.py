"""Debug: clock64 phases of the per-video attention kernel (CTA 0 = first wave, CTA 1000 = a later wave)."""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib, synth  # noqa: E402
from flashvtg_b200.config import PRESETS  # noqa: E402
from flashvtg_b200.model import FlashVTGB200  # noqa: E402

cfg = PRESETS["qvh_iv2"]
dev = torch.device("cuda:0")
m = FlashVTGB200(cfg).eval()
m.load_state_dict(synth.make_state_dict(cfg, 2024))
base = synth.make_inputs(cfg, 64, 75, 32, seed=1)
d = {k: v.repeat(16, *([1] * (v.dim() - 1))).contiguous().to(dev) for k, v in base.items()}
lib = _lib.load()
for _ in range(2):
    m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], uniform_len=True)
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.fvtg_dbg_set_trace(buf.data_ptr())
m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"], uniform_len=True)
torch.cuda.synchronize()
lib.fvtg_dbg_set_trace(None)
t = buf.cpu()[3072:3088].tolist()
for name, o in (("CTA 0", 0), ("CTA 1000", 8)):
    v = t[o:o + 5]
    print(f"{name}: stage {v[1]-v[0]} cycles, compute {v[2]-v[1]}, sync {v[3]-v[2]}, store {v[4]-v[3]}, total {v[4]-v[0]}")
