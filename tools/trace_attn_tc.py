"""Debug: clock64 phases of the persistent tcgen05 attention kernel (CTA 0, first 8 work items)."""
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
t = buf.cpu()[3072:3072 + 256].view(2, 8, 16)
t0 = int(t[0, 0, 0])
names_c = ["item", "kv_full", "S0 issued", "pre p_ready0", "p_ready0", "PV0 issued", "S1 issued", "pre p_ready1",
           "p_ready1", "PV1 issued"]
names_w = ["h0 begin", "s_full", "max done", "P stored", "o_full", "h0 end", "-", "-", "h1 begin", "s_full",
           "max done", "P stored", "o_full", "h1 end"]
for it in range(8):
    c = [int(x) - t0 for x in t[0, it, :10]]
    w = [int(x) - t0 for x in t[1, it, :14]]
    print(f"item {it} ctrl:", " ".join(f"{n}={v}" for n, v in zip(names_c, c)))
    print(f"item {it} wg  :", " ".join(f"{n}={v}" for n, v in zip(names_w, w) if n != "-"))
