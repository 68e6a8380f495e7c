"""Debug: device time of fvtg_decode_nms by NMS mode (run on a B200)."""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import synth  # noqa: E402
from flashvtg_b200.config import PRESETS  # noqa: E402
from flashvtg_b200.model import FlashVTGB200  # noqa: E402

cfg = PRESETS["qvh_iv2"]
dev = torch.device("cuda:0")
m = FlashVTGB200(cfg).eval()
m.load_state_dict(synth.make_state_dict(cfg, 2025, spread=True))
B, N = 1024, cfg.num_points(75)
g = torch.Generator(device="cpu").manual_seed(1)
cls = torch.randn(B, N, generator=g).to(dev)
conf = torch.randn(B, N, generator=g).to(dev)
coord = torch.rand(B, N, 2, generator=g).to(dev) * 3
vlen = torch.full((B,), 75, dtype=torch.int32, device=dev)
for mode in (None, "normal", "linear", "hull"):
    for _ in range(3):
        m.decode(cls, conf, coord, vlen, 75, nms=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.decode(cls, conf, coord, vlen, 75, nms=mode)
    e1.record()
    torch.cuda.synchronize()
    print(mode, f"{e0.elapsed_time(e1) * 100:.1f} us per call (incl. output allocation)")

from flashvtg_b200 import _lib
lib = _lib.load()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.fvtg_dbg_set_trace(buf.data_ptr())
m.decode(cls, conf, coord, vlen, 75, nms="normal")
torch.cuda.synchronize()
lib.fvtg_dbg_set_trace(None)
t = buf.cpu()[2048:2053].tolist()
print("decode CTA 0 cycles: keys %d, sort %d, decode+compose %d, nms %d" % (t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3]))
